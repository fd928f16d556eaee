# round 2, GPU call 20: lockstep tier as collector kernel + finishing kernel (instruction-cache footprint) against the single kernel
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
for v in _head "" _head ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2t_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2t_probe$v.txt)"
done
ncu --set full --clock-control none --import-source on -k regex:"k_thr_col|k_wt_fin|k_bsj_fin" -s 4 -c 8 -o gpurun_out/r2t_full -f python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > gpurun_out/r2t_ncu.log 2>&1
