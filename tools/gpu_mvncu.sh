#!/bin/bash
# per-launch (ncu, cold cache) times of the upwards / solve kernels under several tuning configurations
#   gpurun -- 'bash tools/gpu_mvncu.sh <tag> "<mv_bench args>"'
TAG=${1:-mv}; ARGS=$2
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'upwards|solve_split|hdiff|leaf_solve' -c 4000 --csv --log-file $OUT/mvlaunches_$TAG.csv \
    python tools/mv_bench.py --ncu --reps 1 $ARGS > $OUT/mvncu_$TAG.log 2>&1
tail -3 $OUT/mvncu_$TAG.log
