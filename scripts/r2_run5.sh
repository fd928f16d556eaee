# round 2, GPU call 5: full GPU test suite incl. the BASELINE-scale parity tests and the production-tier interval dump; phase profile of the warp kernels
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -25
CG_LIB_PATH=circall_b200/libcircall_gpu_prof.so python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > gpurun_out/r2e_prof.txt 2>&1
grep "WPROF\|stage" gpurun_out/r2e_prof.txt
