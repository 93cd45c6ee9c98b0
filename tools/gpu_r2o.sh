#!/bin/bash
# Round 2, one GPU: TMA operand staging inside the merge products (tuning key 8): tests, bench A/B, launch list + ncu of the root T product
TAG=${1:-r2o}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_refscale.py tests/test_gpu_graphs.py tests/test_gpu_rebuild.py tests/test_gpu_staged.py -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_$TAG.log | cut -c1-400
for T in 1 0; do
  F=$OUT/bench_${TAG}_t8$T
  timeout 900 python bench.py --no-cpu-baseline --tuning 8=$T > $F.json 2> $F.err; echo "bench tuning 8=$T exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['roofline']['achieved'], d['kernel_ms_per_step'])"
done
if [ -n "$NCU" ]; then bash tools/gpu_ncu.sh $TAG "bgemm_tma_kernel"; fi
