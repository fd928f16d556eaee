# round 2, GPU call 36: final state — full GPU tests, smoke, default bench line, launch list, full ncu capture of one batch
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 ) 2>&1 | tail -12
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r2z_bench_n1.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('frac_of_h2d_ceiling'), d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['cpu_baseline']['value'], d['cpu_baseline']['parity_on_sample'], d['details']['index_build_s'])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2z_ncu1.log 2>&1
python scripts/ktimes.py gpurun_out/r2z_launches.csv | sed 's/ \([a-z:_A-Z]*k_\)/\n\1/g' | grep "k_wt\|k_bsj\|k_pack\|k_thr\|k_list" | cut -c1-80
