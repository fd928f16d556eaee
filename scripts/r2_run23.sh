# round 2, GPU call 23: warp kernels drawing their reads from a counter, 4096-item compaction tiles; variants: text prefetch in the comparisons, two-strand collector at 4 / 5 blocks per SM
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in "" _pf _v4 _v5 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2w_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2w_probe$v.txt)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file gpurun_out/r2w_kt.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2w_kt.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep "k_wt\|k_bsj\|k_pack\|k_thr\|k_flag_scan"
