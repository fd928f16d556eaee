# round 2, GPU call 34: shared-memory footprint of the collector (intervals kept per strand 8 / 6 / 5), 7 blocks per SM with the smaller footprint, 32-thread blocks
set -x
for v in "" _fi6 _fi6m7 _fi5 _t32 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2ag_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2ag_probe$v.txt)"
done
