#!/bin/bash
# ncu --set full of the radix-sort and index-sort kernels (they are launched hundreds of times per step: capture a window)
set -u
TAG=${1:-r1s}; WL=${2:-C2}
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --target-processes application-only --set full --clock-control none -k 'regex:k_sort_scatter|k_sort_hist|k_sort_rowscan|k_index_sort2' -s 40 -c 60 -o $O/prof_sort_${WL}_$TAG -f $CMD > $O/ncu_sort_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i $O/prof_sort_${WL}_$TAG.ncu-rep --page raw --csv > $O/prof_sort_${WL}_${TAG}_raw.csv 2> /dev/null
rm -f $O/prof_sort_${WL}_$TAG.ncu-rep
ls -la $O | tail -5
