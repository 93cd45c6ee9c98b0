#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench command and
# full captures of the dominant kernels.  Usage: gpurun -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
if [ "$2" != "noncu" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --profile-run --steps 1 --warmup 0 > $OUT/ncu_list_$TAG.log 2>&1
python tools/pick_launch.py $OUT/launches_$TAG.csv --summary > $OUT/launch_summary_$TAG.md; cat $OUT/launch_summary_$TAG.md
for K in ${KERNELS:-bgemm_kernel rb_upwards_w_kernel rb_upwards_h_kernel rb_solve_split_kernel solve_split_kernel leaf_solve_const_mma_kernel invert_reg_kernel assemble_X_kernel}; do
  IDX=$(python tools/pick_launch.py $OUT/launches_$TAG.csv $K)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $IDX -c 1 -f -o $OUT/prof_${K}_$TAG \
      python bench.py --profile-run --steps 1 --warmup 0 > $OUT/ncu_full_${K}_$TAG.log 2>&1
  echo "$K idx $IDX: $(tail -1 $OUT/ncu_full_${K}_$TAG.log)"
done
fi
