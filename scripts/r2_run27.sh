# round 2, GPU call 27: second sort key of the lockstep tier's work (which of 5 evenly spaced 16-mers of the read are absent from the text), with / without
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export CG_NO_SUBKEY=1; else unset CG_NO_SUBKEY; fi
  python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2z_probe$v.txt 2>&1
  echo "no_subkey=$v: $(grep 'stage 0 iter 2' gpurun_out/r2z_probe$v.txt)"
done
unset CG_NO_SUBKEY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_" -c 400 --csv --log-file gpurun_out/r2z_kt.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2z_kt.csv',errors='replace')) if len(r)>10]
h=rows[0]; ki,mi,vi=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
d={}
for r in rows[1:]:
    d.setdefault((r[h.index('ID')],r[ki].split('(')[0][-22:]),{})[r[mi].split('.')[0][-12:]]=r[vi]
for (i,k),x in d.items():
    if any(t in k for t in ('k_thr','k_wt','k_bsj','k_cls','k_flag_scan')): print(k,x)
P
