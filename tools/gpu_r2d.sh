#!/bin/bash
# Round 2, one GPU: variable-coefficient leaf kernels (warp-level factor, tensor-core DtN): parity tests, configs[3] bench A/B
# against the round-1 kernels (tuning key 6), ncu of the new kernels.
TAG=${1:-r2d}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_sampling.py tests/test_gpu_staged.py -q -x -k "variable or varcoef or coefficient or golden or reference_dump" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_$TAG.log
for T in 0 1; do
  timeout 900 python bench.py --no-cpu-baseline --adaptive 4 9 --threshold 1.6 --problem varcoef --tuning 6=$T > $OUT/bench_${TAG}_c3_t6$T.json 2> $OUT/bench_${TAG}_c3_t6$T.err; echo "bench c3 tuning 6=$T exit $?"
  python -c "import json; d=json.loads(open('$OUT/bench_${TAG}_c3_t6$T.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['e2e']['ms_per_step'], d['e2e_device_sampling'], d['kernel_ms_per_step'])"
done
bash tools/gpu_ncu.sh ${TAG}_c3 "leaf_var_factor_warp_kernel leaf_var_dtn_mma_kernel leaf_var_solve_warp_kernel" "--adaptive 4 9 --threshold 1.6 --problem varcoef"
