# round 2, GPU call 25: full ncu capture of one whole batch of the step (20 kernels of the stages + k_pack), for profiles/ and traffic.json
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack|k_thr_" -s 60 -c 20 -o gpurun_out/r2x_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2x_ncu2.log 2>&1
ls -la gpurun_out/ | grep r2x_full
