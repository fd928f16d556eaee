# usage: _exp2.sh "<kernel regex>" <skip> <count> variant:filterlog2 ...
RX=$1; SK=$2; CN=$3; shift 3
for vf in "$@"; do
  v=${vf%%:*}; f=${vf##*:}
  CG_FILTER_LOG2=$f CG_LIB_PATH=circall_b200/libcircall_gpu$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"$RX" -s $SK -c $CN --csv --log-file gpurun_out/kt$v-f$f.csv python scripts/probe.py --genes 60000 --pairs 4000000 --batch 4000000 --iters 1 > /dev/null 2>&1
  python scripts/ktimes.py gpurun_out/kt$v-f$f.csv
done
