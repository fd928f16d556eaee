# round 2, GPU call 26: ONE two-strand pass per stage (class tail + the lean instantiation's other-strand declines) against two
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
for v in _head "" _head ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2y_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2y_probe$v.txt)"
done
