# round 2, GPU call 32: the index builder on its own radix sort / scans / selection (no CUB): SA equality, full-size indices through the scale tests, build time
set -x
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_index_dir.py tests/test_gpu_scale.py -m gpu -x -q --durations=4 2>&1 | tail -10
python scripts/probe.py --genes 60000 --pairs 2000000 --batch 2000000 --iters 1 > gpurun_out/r2ae_probe.txt 2>&1
grep "index\|stage 0" gpurun_out/r2ae_probe.txt
