#!/bin/bash
# Round 2, one GPU: look-ahead in the blocked base case (tuning 7: 0 look-ahead, 2 without, 1 round-1 kernel)
TAG=${1:-r2i}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_refscale.py -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_$TAG.log | cut -c1-300
for T in 0 2; do
  F=$OUT/bench_${TAG}_t7$T
  timeout 900 python bench.py --no-cpu-baseline --tuning 7=$T > $F.json 2> $F.err; echo "bench tuning 7=$T exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['kernel_ms_per_step'])"
done
