# round 2, GPU call 16: resident blocks per SM of the lockstep kernels now that they wait on latency, not on DRAM bandwidth (4 / 5 / 6 / 7 / 8 x 128 threads)
set -x
for v in "" _m4 _m6 _m7 _m8 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2p_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2p_probe$v.txt)"
done
