#!/bin/bash
# Round 2, one GPU: TMA-staged GEMM A/B (tests, TFLOP/s table, one ncu --set full capture of each operand path on 4 x 4096^3)
TAG=${1:-r2n}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tma or dgemm" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_$TAG.log | cut -c1-400
timeout 600 python tools/gemm_tma_ab.py --iters 10 --out $OUT/gemm_tma_ab_$TAG.json > $OUT/gemm_tma_ab_$TAG.log 2>&1; echo "ab exit $?"; tail -12 $OUT/gemm_tma_ab_$TAG.log | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bgemm_kernel|dgemm_tma_kernel" -c 2 -f -o $OUT/prof_gemm_ab_$TAG \
    python tools/gemm_tma_ab.py --ncu 1 > $OUT/ncu_gemm_ab_$TAG.log 2>&1; echo "ncu exit $?"; tail -3 $OUT/ncu_gemm_ab_$TAG.log
