#!/bin/bash
# Round 2, one GPU: stream lanes + CUDA graphs: all GPU tests, then the three workloads with and without them.
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -rs -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_$TAG.log | cut -c1-300
run() {  # name, env, args
  F=$OUT/bench_${TAG}_$1
  env $2 timeout 900 python bench.py --no-cpu-baseline $3 > $F.json 2> $F.err; echo "bench $1 exit $?"; tail -2 $F.err
  python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['stages'], d['kernel_ms_per_step'])"
}
run c0 "X=1" "--adaptive 0 7"
run c0_plain "EFGPU_GRAPHS=0 EFGPU_LANES=0" "--adaptive 0 7"
run c0_graphs_only "EFGPU_LANES=0" "--adaptive 0 7"
run c3 "X=1" "--adaptive 4 9 --threshold 1.6 --problem varcoef"
run c3_plain "EFGPU_GRAPHS=0 EFGPU_LANES=0" "--adaptive 4 9 --threshold 1.6 --problem varcoef"
run c1 "X=1" ""
run c1_plain "EFGPU_GRAPHS=0 EFGPU_LANES=0" ""
