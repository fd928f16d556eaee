# round 2, GPU call 37: resident blocks per SM of the warp kernels now that they draw their reads from a counter (general 3 / 4 / 5, fast 4 / 5 / 6)
set -x
for v in "" _g4 _g5 _w5 _w6 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2ai_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2ai_probe$v.txt)"
done
