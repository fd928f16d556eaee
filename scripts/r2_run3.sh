# round 2, GPU call 3: index loads with .L2::64B (default build) against plain loads (_ltc0); GPU tests after the state-machine tier's removal
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "" _ltc0; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2c_probe$v.txt 2>&1
  grep "stage" gpurun_out/r2c_probe$v.txt | tail -2
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors.sum --clock-control none -k regex:"k_wt|k_bsj" -c 40 --csv --log-file gpurun_out/r2c_ncu$v.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
done
