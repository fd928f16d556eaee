import sys
sys.path.insert(0,'/root/repo')
import circall_b200 as cg
for gb in (0.125, 0.5, 1, 2, 4, 8, 16, 32, 64):
    print("gather %.3f GB: %.0f GB/s useful (32 B sectors) = %.1f G sectors/s" % (gb, cg.bench_gather(int(gb * (1 << 30)), 64, 3), cg.bench_gather(int(gb * (1 << 30)), 64, 3)/32), flush=True)
