# round 2, GPU call 43: filter-stage geometry again at the final structure (the collectors sit at the random-access ceiling of the memory system:
# every probe is a DRAM access) — 6 probes 5 apart (default) / 4 probes 7 apart / 3 probes 15 apart / 5 probes 5 apart / 5 probes 7 apart
set -x
for v in "" _f24 _f13 _f35 _f25 ""; do
  CG_LIB_PATH=circall_b200/libcircall_gpu$v.so python scripts/probe.py --genes 60000 --pairs 8000000 --batch 8000000 --iters 3 > gpurun_out/r2ao_probe$v.txt 2>&1
  echo "variant '$v': $(grep 'stage 0 iter 2' gpurun_out/r2ao_probe$v.txt)"
done
