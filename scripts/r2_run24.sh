# round 2, GPU call 24: state of HEAD — full GPU tests, default bench line, launch list, full ncu capture of the step kernels
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 ) 2>&1 | tail -12
python bench.py > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r2x_bench_n1.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('frac_of_h2d_ceiling'), d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['cpu_baseline']['value'], d['cpu_baseline']['parity_on_sample'])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2x_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wt_|k_bsj_|k_pack|k_thr_" -s 66 -c 22 -o gpurun_out/r2x_full -f python bench.py --pairs 8000000 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2x_ncu2.log 2>&1
ls -la gpurun_out/ | grep r2x
