#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, the default bench, then the ncu launch list and full captures.
# gpurun copies back at most 64 MiB: reports are exported to CSV on the box and large .ncu-rep files are dropped.
# usage: scripts/gpu_profile.sh <tag> [workload] [steps: tests,bench,launches,full]
set -u
TAG=${1:-r1}; WL=${2:-C2}; STEPS=${3:-tests,bench,launches,full}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
if [[ $STEPS == *tests* ]]; then
  T0=$(date +%s); python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)" | tee -a $O/pytest_gpu_$TAG.log; tail -5 $O/pytest_gpu_$TAG.log
fi
if [[ $STEPS == *bench* ]]; then
  python bench.py --workload $WL --steps 3 --warmup 3 > $O/bench_${WL}_$TAG.json 2> $O/bench_${WL}_$TAG.err; echo "bench rc=$?"
fi
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
if [[ $STEPS == *launches* ]]; then
  $CMD > $O/plain_$TAG.log 2>&1 &&
  ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_${WL}_$TAG.csv $CMD > $O/ncu_launch_$TAG.log 2>&1
  echo "ncu launches rc=$?"
fi
if [[ $STEPS == *full* ]]; then
  KREGEX=${KREGEX:-'regex:k_s2_|k_ix_|k_index_sort|k_consensus|k_pack_classify|k_sketch_lh|k_resketch|k_sort_|k_bucket_local|k_cb_match|k_cb_consensus'}
  { [[ $STEPS == *launches* ]] || $CMD > $O/plain2_$TAG.log 2>&1; } &&
  ncu --target-processes application-only --set full --clock-control none -k "$KREGEX" -c ${KCOUNT:-40} -o $O/prof_${WL}_$TAG -f $CMD > $O/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?"
  ncu -i $O/prof_${WL}_$TAG.ncu-rep --page raw --csv > $O/prof_${WL}_${TAG}_raw.csv 2> /dev/null
  ncu -i $O/prof_${WL}_$TAG.ncu-rep --page details --csv > $O/prof_${WL}_${TAG}_details.csv 2> /dev/null
  ncu -i $O/prof_${WL}_$TAG.ncu-rep --page source --csv > $O/prof_${WL}_${TAG}_source.csv 2> /dev/null
  gzip -f $O/prof_${WL}_${TAG}_source.csv
  SZ=$(stat -c %s $O/prof_${WL}_$TAG.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -gt 30000000 ]; then rm -f $O/prof_${WL}_$TAG.ncu-rep; echo "dropped .ncu-rep ($SZ bytes), CSV exports kept"; fi
fi
du -sh $O; ls -la $O
