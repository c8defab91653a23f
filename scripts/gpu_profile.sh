#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, the default bench, then the ncu launch list and full captures.
# usage: scripts/gpu_profile.sh <tag> [workload]
set -u
TAG=${1:-r1}; WL=${2:-C2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
python bench.py --workload $WL --steps 3 --warmup 3 > $O/bench_${WL}_$TAG.json 2> $O/bench_${WL}_$TAG.err; echo "bench rc=$?"
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_${WL}_$TAG.csv $CMD > $O/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 &&
ncu --target-processes application-only --set full --clock-control none --import-source on -k 'regex:k_s2_probe|k_index_sort|k_consensus|k_pack_classify|k_sketch_lh|k_s2_singles' -c 22 -o $O/prof_${WL}_$TAG -f $CMD > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
$CMD > $O/plain3_$TAG.log 2>&1 &&
ncu --target-processes application-only --set full --clock-control none --import-source on -k 'regex:k_sort_scatter|k_sort_hist|k_scan_apply' -c 6 -o $O/prof_sort_${WL}_$TAG -f $CMD > $O/ncu_full_sort_$TAG.log 2>&1
echo "ncu sort rc=$?"
ls -la $O
