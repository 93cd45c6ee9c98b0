#!/bin/bash
# Round 2, one GPU: cluster-cooperative 256 x 256 base case (tuning key 9): tests, bench A/B, optional ncu of the kernel
TAG=${1:-r2t}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_refscale.py tests/test_gpu_graphs.py tests/test_gpu_staged.py -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_$TAG.log | cut -c1-300
for T in 1 0; do
  F=$OUT/bench_${TAG}_t9$T
  timeout 900 python bench.py --no-cpu-baseline --tuning 9=$T > $F.json 2> $F.err; echo "bench tuning 9=$T exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['kernel_ms_per_step'], d['gpu_launches'])"
done
if [ -n "$NCU" ]; then bash tools/gpu_ncu.sh $TAG "invert_cluster_kernel"; fi
