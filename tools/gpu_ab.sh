#!/bin/bash
# A/B call: GPU parity tests, then bench lines for the argument sets given (one per argument).
#   gpurun -- 'bash tools/gpu_ab.sh <tag> "" "--no-symmetry"'
TAG=${1:-ab}; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_$TAG.log
i=0
for ARGS in "$@"; do
  timeout 900 python bench.py $ARGS > $OUT/bench_${TAG}_$i.json 2> $OUT/bench_${TAG}_$i.err; echo "bench [$ARGS] exit $?"
  cat $OUT/bench_${TAG}_$i.json; tail -5 $OUT/bench_${TAG}_$i.err
  i=$((i+1))
done
