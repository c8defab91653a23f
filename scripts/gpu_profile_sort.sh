#!/bin/bash
# ncu of the radix-sort kernels: the 18 launches of the round-1 tuple sort, then the 8 launches of the contig-table sort
# (they follow the ~70 sort launches of Stage 1 in a step).  Metric subset instead of --set full: fewer replays.
set -u
TAG=${1:-r1s}; WL=${2:-C2}
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --target-processes application-only --metrics $M --clock-control none -k 'regex:k_sort_scatter|k_sort_hist' -c 18 --csv --log-file $O/sort_main_${WL}_$TAG.csv $CMD > $O/ncu_sort_a_$TAG.log 2>&1
echo "ncu main rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 &&
ncu --target-processes application-only --metrics $M --clock-control none -k 'regex:k_sort_scatter<unsigned long long|k_sort_hist<unsigned long long|k_index_sort2' -c 15 --csv --log-file $O/sort_kmer_${WL}_$TAG.csv $CMD > $O/ncu_sort_b_$TAG.log 2>&1
echo "ncu kmer rc=$?"
ls -la $O | tail -6
