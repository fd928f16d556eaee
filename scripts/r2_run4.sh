# round 2, GPU call 4: seed filter (16-mer presence bit set) in the lockstep tier's walks: tests, timing with and without it, DRAM sectors
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for fm in 16 0; do
  CG_FM=$fm python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2d_probe_fm$fm.txt 2>&1
  grep "index\|stage" gpurun_out/r2d_probe_fm$fm.txt | tail -4
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors.sum --clock-control none -k regex:"k_wt|k_bsj" -c 40 --csv --log-file gpurun_out/r2d_ncu.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2d_ncu.csv | tr ' ' '\n' | grep -v "flag\|cls_\|commit\|exp_split\|cand" 
