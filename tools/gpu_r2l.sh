#!/bin/bash
# Round 2, multi-GPU call after the staged transposed stores (tuning key 4), device-side barrier epochs and graph replay of
# peer-mapped trees:  gpurun --gpus N -- 'bash tools/gpu_r2l.sh r2l N [single]'
#   sharded tests on the box's GPUs, bench at N ranks with graphs on / off, a per-launch trace of one profiled pass (EFGPU_TRACE),
#   optionally (third argument) the new single-GPU test and the single-GPU A/B of tuning key 4.
TAG=${1:-r2l}; N=${2:-2}; SINGLE=$3
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_$TAG.txt
if [ -n "$SINGLE" ]; then
  timeout 600 python -m pytest tests/test_gpu_edge.py -q -x > $OUT/pytest_edge_$TAG.log 2>&1; echo "pytest edge exit $?"; tail -3 $OUT/pytest_edge_$TAG.log | cut -c1-300
  for T in 0 1; do
    F=$OUT/bench_${TAG}_t4$T
    timeout 600 python bench.py --no-cpu-baseline --tuning 4=$T > $F.json 2> $F.err; echo "bench tuning 4=$T exit $?"; tail -2 $F.err
    python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['kernel_ms_per_step'])"
  done
fi
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -rs ${TESTK:+-k "$TESTK"} > $OUT/pytest_sharded_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_sharded_$TAG.log
fi
run() {   # name, env..., -- bench args
  local NAME=$1; shift
  local F=$OUT/bench_${TAG}_n${N}_$NAME
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 $BARGS > $F.json 2> $F.err
  echo "bench N=$N $NAME exit $?"; tail -3 $F.err
  [ -s $F.json ] && python -c "import json; d=json.loads([l for l in open('$F.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['stages']['build_ms'], d['config']['parity'], d['kernel_ms_per_step'])"
}
BARGS=""
run graphs EFGPU_TRACE=$OUT/trace_${TAG}_n${N}
[ -z "$SKIP_PLAIN" ] && run plain EFGPU_GRAPHS_PEER=0
if [ -n "$ROWSPLIT" ]; then BARGS="--tuning 11=0"; run rowsplit EFGPU_X=0; BARGS=""; fi
if [ -n "$BIG" ]; then BARGS="$BIG"; run big EFGPU_X=0; fi
ls $OUT | grep trace_${TAG} | head -40
