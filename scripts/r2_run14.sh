# round 2, GPU call 14: device-resident step dealt to 1 / 2 / 3 / 4 sessions (tail kernels of one batch under the thread tiers of another)
set -x
for rs in 1 2 3 4; do
python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --resident-sessions $rs > gpurun_out/r2n_bench_rs$rs.json 2> gpurun_out/r2n_bench_rs$rs.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_rs$rs.json')); print('rs=$rs', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
python bench.py --no-cpu --steps 4 --warmup 3 --resident-sessions 2 --batch 5000000 > gpurun_out/r2n_bench_rs2b.json 2> gpurun_out/r2n_bench_rs2b.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_rs2b.json')); print('rs=2 batch 5M', d['value'], d['ms_per_step'], d['e2e']['value'])"
