# round 2, GPU call 21: resident blocks / block size of the collector kernel after the split; per-kernel times of the step
set -x
for v in "" _m5 _m7 _m8 _t128; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2u_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2u_probe$v.txt)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file gpurun_out/r2u_kt.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2u_kt.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep "k_wt\|k_bsj\|k_pack\|k_thr\|k_flag_scan"
