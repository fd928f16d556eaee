# round 2, GPU call 44: two-level filter stage (probes 15 and 30 first, the other four only when one of those is in the text), with / without
set -x
CG_LIB_PATH=circall_b200/libcircall_gpu_f2l.so timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in "" _f2l "" _f2l; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2ap_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2ap_probe$v.txt)"
done
for v in "" _f2l; do
CG_LIB_PATH=circall_b200/libcircall_gpu$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"k_thr_col" -s 3 -c 3 --csv --log-file gpurun_out/r2ap_kt$v.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python - "$v" <<'P'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/r2ap_kt%s.csv'%sys.argv[1],errors='replace')) if len(r)>10]
h=rows[0]; ki,mi,vi=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
d={}
for r in rows[1:]: d.setdefault((r[h.index('ID')],r[ki].split('(')[0][-16:]),{})[r[mi].split('.')[0][-14:]]=r[vi]
print(sys.argv[1] or 'default', d)
P
done
