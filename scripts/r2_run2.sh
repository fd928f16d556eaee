# round 2, GPU call 2: what one missing 32-byte sector costs, by load instruction (gather microbenchmark, 16 GB buffer)
set -x
CG_DEBUG=1 CG_L2_FETCH=32 python - <<'P' 2>&1 | tee gpurun_out/r2b_gather.txt
import sys
sys.path.insert(0, '.')
import circall_b200 as cg
for mode in range(10):
    for gb in (16,):
        v = cg.bench_gather(int(gb * (1 << 30)), 64, 3, mode=mode)
        print("mode %d gather %d GB: %.0f GB/s useful = %.1f G sectors/s" % (mode, gb, v, v / 32), flush=True)
P
python - <<'P' 2>&1 | tee -a gpurun_out/r2b_gather.txt
import sys
sys.path.insert(0, '.')
import circall_b200 as cg
for mode in range(10):
    v = cg.bench_gather(16 << 30, 64, 3, mode=mode)
    print("default limit: mode %d gather 16 GB: %.0f GB/s useful = %.1f G sectors/s" % (mode, v, v / 32), flush=True)
P
cat > /tmp/g.py <<'P'
import sys
sys.path.insert(0, '.')
import circall_b200 as cg
for mode in range(10):
    cg.bench_gather(16 << 30, 16, 1, mode=mode)
P
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_sectors.sum,lts__t_requests_srcunit_tex_op_read.sum --clock-control none -k regex:k_gather --csv --log-file gpurun_out/r2b_ncu.csv python /tmp/g.py > /dev/null 2>&1
python - <<'P' | tee -a gpurun_out/r2b_gather.txt
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2b_ncu.csv', errors='replace')) if len(r) > 10]
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split('(')[0]), {})[r[mi]] = float(r[vi].replace(',', ''))
for k in sorted(d):
    m = d[k]
    print(k, " ".join("%s=%.4g" % (a.split('__')[-1], b) for a, b in m.items()))
P
