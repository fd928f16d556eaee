# round 2, GPU call 45: SA / transcript id / lcp of a suffix-array index as one 8-byte entry for the thread tiers (IndexView::sat2), with / without
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export CG_NO_SAT2=1; else unset CG_NO_SAT2; fi
  python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2aq_probe$v.txt 2>&1
  echo "no_sat2=$v: $(grep 'bsj index' gpurun_out/r2aq_probe$v.txt) $(grep 'stage 0 iter 2' gpurun_out/r2aq_probe$v.txt)"
done
unset CG_NO_SAT2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file gpurun_out/r2aq_kt.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2aq_kt.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep "k_wt\|k_bsj\|k_thr" | cut -c1-60
