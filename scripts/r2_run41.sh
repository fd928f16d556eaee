# round 2, GPU call 41: full GPU suite at HEAD (two look-ups per walk round)
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r2am_bench_n1.json 2> gpurun_out/r2am_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r2am_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
