# round 2, GPU call 8 (N GPUs): multi-GPU product tests, concurrent H2D ceiling, bench at N
N=${1:-2}
set -x
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or allreduce or interleave or failed_batch or intervals" 2>&1 | tail -40
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/h2d_concurrent.py 2>&1 | grep H2D | tee gpurun_out/r2h_h2d_n$N.txt
python scripts/h2d_concurrent.py 2>&1 | grep H2D | tee -a gpurun_out/r2h_h2d_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
echo "rc=$? lines=$(wc -l < gpurun_out/r2h_bench_n$N.json)"
python - <<P
import json
d = json.load(open('gpurun_out/r2h_bench_n$N.json'))
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
P
