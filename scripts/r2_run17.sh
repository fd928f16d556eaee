# round 2, GPU call 17: full GPU suite (index directory in the reference's layout, gzip inputs), default bench line
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) 2>&1 | tail -16
python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r2q_bench_n1.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['cpu_baseline']['value'], d['cpu_baseline']['parity_on_sample'])
P
