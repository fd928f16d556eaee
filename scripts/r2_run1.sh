# round 2, GPU call 1: L2 fetch granularity (default / 32 / 64 / 128) on the gather ceiling and on the pipeline; per-kernel fixed costs at small batches
set -x
for g in 0 32 64 128; do
  export CG_L2_FETCH=$g
  python - <<'P' 2>&1 | tee -a gpurun_out/r2a_gather.txt
import os, sys
sys.path.insert(0, '.')
import circall_b200 as cg
for gb in (0.0625, 4, 32):
    v = cg.bench_gather(int(gb * (1 << 30)), 64, 3)
    print("CG_L2_FETCH=%s gather %.3f GB: %.0f GB/s useful = %.1f G sectors/s" % (os.environ.get("CG_L2_FETCH"), gb, v, v / 32), flush=True)
P
  python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2a_probe_$g.txt 2>&1
  grep "stage" gpurun_out/r2a_probe_$g.txt
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:"k_wt_pre|k_wt_thr|k_bsj_thr|k_bsj_cls|k_gather" -c 12 --csv --log-file gpurun_out/r2a_ncu_$g.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
done
unset CG_L2_FETCH
# per-kernel times of a 1 Mi-pair batch against a 4 M-pair batch (the per-launch tails)
for b in 1048576 4194304; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wt|k_bsj|k_flag|k_cls|k_pack|k_exp|k_cand" -c 400 --csv --log-file gpurun_out/r2a_kt_$b.csv python scripts/probe.py --genes 60000 --pairs 4194304 --batch $b --iters 1 > /dev/null 2>&1
  python scripts/ktimes.py gpurun_out/r2a_kt_$b.csv | tee gpurun_out/r2a_kt_$b.txt
done
