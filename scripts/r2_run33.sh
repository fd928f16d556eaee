# round 2, GPU call 33: the reference arm (CPU restatement on the box's host cores, nothing of the product loaded)
set -x
nproc
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2af_bench_reference_arm.json 2> gpurun_out/r2af_bench_reference_arm.err ) 2>&1 | tail -4
cat gpurun_out/r2af_bench_reference_arm.json | cut -c1-600
grep -c libcircall /proc/self/maps
