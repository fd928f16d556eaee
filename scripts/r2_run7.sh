# round 2, GPU call 7: full GPU suite, default bench line, reference arm (its own suffix sort), 1 Mi-batch device-resident rate
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
( time python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2g_bench.err
python - <<'P'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('e2e', d['e2e'])
print('roofline', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'traffic', 'avg_launch_ms', 'algorithmic_bytes_per_pair')})
print('cpu', d['cpu_baseline'])
P
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err ) 2>&1 | tail -3
cat gpurun_out/r2g_ref.json | cut -c 1-900
python bench.py --no-cpu --no-e2e --batch 1048576 --steps 3 > gpurun_out/r2g_bench_1mi.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_1mi.json')); print('1Mi batches resident', d['value'], d['ms_per_step'])"
