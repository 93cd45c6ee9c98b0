#!/bin/bash
# Ad-hoc experiment call: GPU parity tests, then whatever experiments the arguments name.
#   gpurun -- 'bash tools/gpu_exp.sh <tag> gemm "7 18 19"'      GEMM tuning variants (tools/gemm_bench.py)
#   gpurun -- 'bash tools/gpu_exp.sh <tag> bench "<bench args>"' one bench line
#   gpurun -- 'bash tools/gpu_exp.sh <tag> mv "--level 8 --nx 16 --configs 1,0,0,0;1,0,0,1"'  upwards/solve A/B (tools/mv_bench.py)
TAG=${1:-x}; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
while [ $# -gt 0 ]; do
  case "$1" in
    gemm) timeout 600 python tools/gemm_bench.py $2 > $OUT/gemm_$TAG.log 2>&1; cat $OUT/gemm_$TAG.log; shift 2;;
    bench) timeout 900 python bench.py $2 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err; shift 2;;
    mv) timeout 900 python tools/mv_bench.py $2 2>&1 | tee -a $OUT/mv_$TAG.log; shift 2;;
    env) export $2; shift 2;;
    *) shift;;
  esac
done
