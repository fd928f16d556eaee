# round 2, GPU call 13: walk as ONE loop (a filter stage and a table round per trip) and denser filter stages, against HEAD
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in _head "" _d3 _d3t _d2; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2m_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2m_probe$v.txt)"
done
for v in "" _d3; do
CG_LIB_PATH=circall_b200/libcircall_gpu$v.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_wt_thr|k_bsj_thr" -s 2 -c 4 --csv --log-file gpurun_out/r2m_kt$v.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
done
