#!/bin/bash
# Round 2, multi-GPU call: sharded tests on the box's GPUs, then bench lines at the given rank counts, peer mode (default) and
# round 1's NCCL-callback path (EFGPU_P2P=0), optionally a larger tree.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2b.sh r2b "2" [extra bench args]'
TAG=${1:-r2b}; NS=${2:-"2"}; EXTRA=$3
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_$TAG.txt; nvidia-smi topo -m >> $OUT/smi_$TAG.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -rs ${TESTK:+-k "$TESTK"} > $OUT/pytest_sharded_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_sharded_$TAG.log
fi
for N in $NS; do
  for P2P in 1 0; do
    [ "$P2P" = "0" ] && [ -n "$SKIP_NCCL" ] && continue
    F=$OUT/bench_${TAG}_n${N}_p2p$P2P
    EFGPU_P2P=$P2P timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$P2P bench.py --gpus $N --steps 5 --warmup 3 $EXTRA > $F.json 2> $F.err
    echo "bench N=$N p2p=$P2P exit $?"; tail -3 $F.err
    [ -s $F.json ] && python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stages'], d['config']['parity'], d['kernel_ms_per_step'])"
  done
done
if [ -n "$BIG" ]; then   # a tree that does not fit one GPU (e.g. BIG="--level 9"): peer mode only
  for N in $NS; do
    F=$OUT/bench_${TAG}_n${N}_big
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 $BIG > $F.json 2> $F.err
    echo "bench N=$N $BIG exit $?"; tail -3 $F.err
    [ -s $F.json ] && python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stages'], d['config']['parity'], d['kernel_ms_per_step'], d['config']['l2'])"
  done
fi
if [ -n "$SPLITS" ]; then   # A/B of the row-split threshold of the inversion's products
  for N in $NS; do for SM in $SPLITS; do
    F=$OUT/bench_${TAG}_n${N}_split$SM
    EFGPU_SPLIT_MIN_ROWS=$SM timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > $F.json 2> $F.err
    echo "bench N=$N split_min=$SM exit $?"; tail -3 $F.err
    [ -s $F.json ] && python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms_per_step'])"
  done; done
fi
if [ -n "$ADAPT" ]; then   # sharded adaptive / variable-coefficient workloads (BASELINE configs[0], configs[3])
  for N in $NS; do
    F=$OUT/bench_${TAG}_n${N}_c0
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --adaptive 2 7 > $F.json 2> $F.err
    echo "bench N=$N adaptive 2-7 exit $?"; tail -3 $F.err
    [ -s $F.json ] && python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['stages'], d['config']['sharding'][:120])"
    F=$OUT/bench_${TAG}_n${N}_c3
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 --adaptive 4 9 --threshold 1.6 --problem varcoef > $F.json 2> $F.err
    echo "bench N=$N adaptive 4-9 varcoef exit $?"; tail -3 $F.err
    [ -s $F.json ] && python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['stages'], d['kernel_ms_per_step'])"
  done
fi
