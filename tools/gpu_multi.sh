#!/bin/bash
# multi-GPU round: sharded tests + bench at N = 1, 2, ... (usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> "<N list>"')
TAG=${1:-r1}
NS=${2:-"1 2"}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > $OUT/pytest_sharded_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_sharded_$TAG.log
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
  fi
  echo "bench N=$N exit $?"; cat $OUT/bench_${TAG}_n$N.json; tail -8 $OUT/bench_${TAG}_n$N.err
done
