# round 2, GPU call 40: table look-ups per walk round (2 / 4 / 8) and lcp bytes per extension chunk (2 / 4 / 8) at the final structure
set -x
for v in "" _w2 _w8 _x2 _x8 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2al_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2al_probe$v.txt)"
done
