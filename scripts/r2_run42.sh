# round 2, GPU call 42: general warp kernels with the hit enumeration's candidate list (sort + binary searches) in shared memory
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for v in _head "" _head ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2an_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2an_probe$v.txt)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wt_warp|k_bsj_warp" -c 8 --csv --log-file gpurun_out/r2an_kt.csv python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2an_kt.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep "k_wt\|k_bsj"
