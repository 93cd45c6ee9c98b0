#!/bin/bash
# Round 2, one GPU: fast pivot reciprocal in the 128 x 128 base case (tuning key 10): test + bench A/B
TAG=${1:-r2v}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_edge.py -q -x -k "reciprocal or blocked" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_$TAG.log | cut -c1-300
for T in 1 0; do
  F=$OUT/bench_${TAG}_t10$T
  timeout 900 python bench.py --no-cpu-baseline --tuning 10=$T > $F.json 2> $F.err; echo "bench tuning 10=$T exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['kernel_ms_per_step']['invert_small'], d['kernel_ms_per_step']['gemm_Xinv'])"
done
