# round 2, GPU call 6: 32-ary extension search in the warp kernels + per-suffix transcript array: tests, step time, per-kernel times
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_cli.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4
python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2f_probe.txt 2>&1
grep "index\|stage" gpurun_out/r2f_probe.txt | tail -5
CG_NO_SATID=1 python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 2 > gpurun_out/r2f_probe_nosatid.txt 2>&1
grep "stage 0" gpurun_out/r2f_probe_nosatid.txt | tail -1
for b in 4194304 1048576; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wt|k_bsj|k_pack" -c 200 --csv --log-file gpurun_out/r2f_kt_$b.csv python scripts/probe.py --genes 60000 --pairs 4194304 --batch $b --iters 1 > /dev/null 2>&1
python scripts/ktimes.py gpurun_out/r2f_kt_$b.csv | tr ' ' '\n' | grep -v "flag\|cls_\|commit\|exp_split\|cand"
done
