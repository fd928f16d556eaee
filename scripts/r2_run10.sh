# round 2, GPU call 10 (re-entry): state of HEAD — full GPU tests, default bench line, launch list, full ncu capture of the step kernels
set -x
nvidia-smi -L | head -2
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) 2>&1 | tail -20
python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r2j_bench_n1.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['roofline'], d['cpu_baseline'])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_ncu1.log 2>&1
python scripts/ktimes.py gpurun_out/r2j_launches.csv | tr ' ' '\n' | grep "k_wt\|k_bsj\|k_pack" 
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack" -s 36 -c 12 -o gpurun_out/r2j_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_ncu2.log 2>&1
ls -la gpurun_out/
