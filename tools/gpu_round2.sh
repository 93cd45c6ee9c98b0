#!/bin/bash
# First GPU call of round 2: everything written at the end of round 1 that has not been timed yet, in one call.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r2a'            (1 GPU)
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_round2.sh r2a multi "1 2 4 8"'
TAG=${1:-r2a}; MODE=${2:-single}; NS=${3:-"2 4 8"}
OUT=gpurun_out; mkdir -p $OUT
export EFGPU_TEST_STAGED=1     # tests/test_gpu_staged.py: the switches below, against the default path
if [ "$MODE" = "single" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
  # A/B: symmetric diagonal blocks of T as block triangles (327 -> 315 n^3 per merge; CPU-emulated at 1 / 4 / 8 ranks)
  for T in "" "--tuning 5=1" "--lazy-root-dtn" "--tuning 5=1 --lazy-root-dtn"; do
    N=$(echo "$T" | tr -dc '0-9a-z'); timeout 600 python bench.py --no-cpu-baseline $T > $OUT/bench_${TAG}_t${N:-0}.json 2> $OUT/bench_${TAG}_t${N:-0}.err
    echo "bench [$T] exit $?"; python -c "import json,sys; d=json.load(open('$OUT/bench_${TAG}_t${N:-0}.json')); print(d['ms_per_step'], d['roofline']['frac'], d['kernel_ms_per_step'])"
  done
  timeout 300 python tools/sample_bench.py > $OUT/sample_bench_$TAG.log 2>&1; tail -2 $OUT/sample_bench_$TAG.log
  timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json
else
  timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_sharded_$TAG.log
  for N in 2 4; do   # cut at level 1: four subtrees, level-1 merges whole on their owners (model: -5 % / -11 % at 2 / 4 GPUs)
    F=$OUT/bench_${TAG}_n${N}_cut1
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --cut 1 > $F.json 2> $F.err
    echo "bench N=$N cut=1 exit $?"; [ -s $F.json ] && python -c "import json; d=json.load(open('$F.json')); print(d['ms_per_step'], d['stages'], d['config']['sharding'])"
  done
  for N in 4 8; do   # three tiers: level-1 merges inside rank groups (model: -4.5 of 52 ms at 8 GPUs)
    [ $(nvidia-smi -L | wc -l) -ge $N ] || continue
    F=$OUT/bench_${TAG}_n${N}_grouped
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --grouped > $F.json 2> $F.err
    echo "bench N=$N grouped exit $?"; [ -s $F.json ] && python -c "import json; d=json.load(open('$F.json')); print(d['ms_per_step'], d['stages'], d['config']['sharding'])"
  done
  for N in $NS; do
    for AG in 0 1; do
      F=$OUT/bench_${TAG}_n${N}_ag$AG
      if [ "$N" = "1" ]; then [ $AG = 0 ] && timeout 900 python bench.py --no-cpu-baseline > $F.json 2> $F.err
      else EFGPU_SHARE_ALLGATHER=$AG timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $F.json 2> $F.err; fi
      echo "bench N=$N allgather=$AG exit $?"; [ -s $F.json ] && python -c "import json; d=json.load(open('$F.json')); print(d['ms_per_step'], d['stages'], d['config']['sharding'])"
    done
  done
fi
