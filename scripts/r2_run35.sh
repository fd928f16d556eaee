# round 2, GPU call 35: lean collector with its intervals written straight to the hand-over slot (20 words of shared memory per read instead of 44); blocks per SM again
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in "" _m5 _m7 _m8 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2ah_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2ah_probe$v.txt)"
done
