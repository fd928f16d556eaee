# round 2, GPU call 47: full ncu capture of the last batch of the step (16 kernels: k_pack + the two stages), for profiles/ and traffic.json
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack|k_thr_" -s 48 -c 16 -o gpurun_out/r2fin_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2fin_ncu2.log 2>&1
ls -la gpurun_out/ | grep r2fin_full
