# round 2, GPU call 39: radix sort with the tile staged in sorted order in shared memory (contiguous runs out): primitives' tests, index tests, build time
set -x
timeout 900 python -m pytest tests/test_gpu_sortscan.py tests/test_gpu_pipeline.py -m gpu -x -q -k "sort or scan or index" 2>&1 | tail -4
python scripts/probe.py --genes 60000 --pairs 2000000 --batch 2000000 --iters 1 > gpurun_out/r2ak_probe.txt 2>&1
grep "index\|stage 0" gpurun_out/r2ak_probe.txt | cut -c1-120
