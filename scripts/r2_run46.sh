# round 2, GPU call 46: final profiles — launch list and full ncu capture of one whole batch of the step (18 kernels of the stages + k_pack)
set -x
python bench.py --pairs 8000000 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2fin_plain.json 2> gpurun_out/r2fin_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2fin_launches.csv python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2fin_ncu1.log 2>&1
python scripts/ktimes.py gpurun_out/r2fin_launches.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep -c "k_wt\|k_bsj\|k_pack\|k_thr"
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack|k_thr_" -s 51 -c 17 -o gpurun_out/r2fin_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2fin_ncu2.log 2>&1
ls -la gpurun_out/ | grep r2fin
