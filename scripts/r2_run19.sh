# round 2, GPU call 19: source-level profile of the step kernels at the current state (full set, import source)
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack" -s 36 -c 12 -o gpurun_out/r2s_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2s_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2s_ncu1.log 2>&1
ls -la gpurun_out | grep r2s
