# round 2, GPU call 11: persistent-lane lockstep kernels (lanes draw reads from a counter; one loop: begin / step / finish) against HEAD's
# one-read-per-thread kernels, register cap and filter-stage geometry variants; parity tests first
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5
for v in _head "" _m4 _d3 _d3t _d3m4 _d2; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2k_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2k_probe$v.txt)"
done
