#!/bin/bash
# Round 2, one GPU: blocked tensor-core base case of the block inversion (A/B against the round-1 register kernel), adaptive
# re-build, VTU writer; then the whole GPU suite.
TAG=${1:-r2g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_rebuild.py tests/test_gpu_sampling.py -q -s > $OUT/pytest_${TAG}_new.log 2>&1; echo "new tests exit $?"; tail -15 $OUT/pytest_${TAG}_new.log | cut -c1-400
for T in 0 1; do
  F=$OUT/bench_${TAG}_t7$T
  timeout 900 python bench.py --no-cpu-baseline --tuning 7=$T > $F.json 2> $F.err; echo "bench tuning 7=$T exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['kernel_ms_per_step'])"
done
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log | cut -c1-300
bash tools/gpu_ncu.sh ${TAG} "invert_blk128_kernel" ""
