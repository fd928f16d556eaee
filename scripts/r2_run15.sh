# round 2, GPU call 15: lcp bytes in the extension (one text comparison per interval instead of one per suffix), with / without
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  if [ $v = 1 ]; then export CG_NO_LCP=1; fi
  python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2o_probe_nolcp$v.txt 2>&1
  echo "no_lcp=$v: $(grep 'index' gpurun_out/r2o_probe_nolcp$v.txt | tr '\n' ' ') $(grep 'stage 0 iter 2' gpurun_out/r2o_probe_nolcp$v.txt)"
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum --clock-control none -k regex:"k_wt_thr|k_bsj_thr|k_wt_pre|k_bsj_cls|k_wt_warp|k_bsj_warp" -s 5 -c 10 --csv --log-file gpurun_out/r2o_kt$v.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
done
