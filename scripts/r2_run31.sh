# round 2, GPU call 31: phase profile of the warp kernels (cycles of lane 0 per phase; reads by log2 cycles), wt only
set -x
CG_LIB_PATH=circall_b200/libcircall_gpu_prof.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > gpurun_out/r2ad_prof.txt 2>&1
grep "WPROF\|stage" gpurun_out/r2ad_prof.txt
