# round 2, GPU call 28: L2 prefetches along the collector's dependent chain (variant _pfc); 8 M-pair samples of configs 3 / 4 / 5
set -x
for v in "" _pfc "" _pfc; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2aa_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2aa_probe$v.txt)"
done
for c in 4 5 3; do
  python bench.py --config $c --pairs 8000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2aa_bench_cfg$c.json 2> gpurun_out/r2aa_bench_cfg$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r2aa_bench_cfg$c.json')); print('cfg$c', round(d['value']/1e6,1), 'M pairs/s resident', round(d['e2e']['value']/1e6,1), 'e2e', d['details']['index_device_gb'], 'GB index')"
done
