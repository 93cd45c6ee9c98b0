#!/bin/bash
# ncu only: launch list of one bench step, then `--set full` captures of the longest launch of each named kernel.
#   gpurun -- 'bash tools/gpu_ncu.sh <tag> "<kernel regexes>" ["<bench args>"]'
TAG=${1:-n}; KERNELS=$2; ARGS=$3
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --profile-run --steps 1 --warmup 0 $ARGS > $OUT/ncu_list_$TAG.log 2>&1
python tools/pick_launch.py $OUT/launches_$TAG.csv --summary > $OUT/launch_summary_$TAG.md; cat $OUT/launch_summary_$TAG.md
for K in $KERNELS; do
  IDX=$(python tools/pick_launch.py $OUT/launches_$TAG.csv $K)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $IDX -c 1 -f -o $OUT/prof_${K}_$TAG \
      python bench.py --profile-run --steps 1 --warmup 0 $ARGS > $OUT/ncu_full_${K}_$TAG.log 2>&1
  echo "$K idx $IDX: $(tail -1 $OUT/ncu_full_${K}_$TAG.log)"
done
