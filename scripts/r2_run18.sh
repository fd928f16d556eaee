# round 2, GPU call 18 (N GPUs): the concurrent host-to-device ceiling of the box and the bench at N (the line carries e2e.frac_of_h2d_ceiling)
N=${1:-8}
set -x
nvidia-smi -L | wc -l
nvidia-smi topo -m 2>/dev/null | head -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/h2d_concurrent.py 2>&1 | grep H2D | tee gpurun_out/r2r_h2d_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r2r_bench_n$N.json 2> gpurun_out/r2r_bench_n$N.err
echo "rc=$? lines=$(wc -l < gpurun_out/r2r_bench_n$N.json)"
python - <<P
import json
d = json.load(open('gpurun_out/r2r_bench_n$N.json'))
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e'])
P
