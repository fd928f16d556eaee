# round 2, GPU call 12: why the persistent-lane variant is slower: ncu of k_bsj_thr / k_wt_thr in both builds (lanes per instruction, instructions, stalls)
set -x
for v in "" _pers; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so ncu --set full --clock-control none --import-source on -k regex:"k_wt_thr|k_bsj_thr" -s 2 -c 4 -o gpurun_out/r2l_full$v -f python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > gpurun_out/r2l_ncu$v.log 2>&1
done
ls -la gpurun_out | grep r2l
