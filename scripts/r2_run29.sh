# round 2, GPU calls 29 and 50 (N GPUs): bench at N with the final kernels (r2ab: mid-round series, r2av: last commit)
N=${1:-2}
set -x
if [ $N = 1 ]; then
python bench.py --steps 5 --warmup 3 > gpurun_out/r2av_bench_n1.json 2> gpurun_out/r2av_bench_n1.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2av_bench_n$N.json 2> gpurun_out/r2av_bench_n$N.err
fi
echo "rc=$? lines=$(wc -l < gpurun_out/r2av_bench_n$N.json)"
python - <<P
import json
d = json.load(open('gpurun_out/r2av_bench_n$N.json'))
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['h2d_ceiling_gbs'], d['e2e']['frac_of_h2d_ceiling'], d['clocks'])
P
