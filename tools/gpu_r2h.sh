#!/bin/bash
# Round 2, one GPU: transposes fused into the GEMM epilogues (second, transposed destination): full GPU suite, bench, ncu launch
# list + full capture of the longest bgemm launch (root T) for roofline.traffic.
TAG=${1:-r2h}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log | cut -c1-300
F=$OUT/bench_$TAG
timeout 900 python bench.py --no-cpu-baseline > $F.json 2> $F.err; echo "bench exit $?"; tail -2 $F.err
python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['gpu_launches'], d['kernel_ms_per_step'])"
git rev-parse HEAD > $OUT/head_$TAG.txt 2>/dev/null
bash tools/gpu_ncu.sh ${TAG} "bgemm_kernel" ""
