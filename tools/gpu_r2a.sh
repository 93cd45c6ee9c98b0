#!/bin/bash
# Round 2, GPU call 1 (one GPU): every regular GPU test (the formerly staged ones and the bench-scale parity against the compiled
# reference included), the default bench line, the A/B of tuning key 5, the adaptive / variable-coefficient workloads, the
# reference arm, and ncu captures of the variable-coefficient leaf kernels and the coarsening stencil.
TAG=${1:-r2a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc_$TAG.txt; free -g >> $OUT/nproc_$TAG.txt
timeout 1800 python -m pytest tests -m gpu -q -rs -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_$TAG.log
grep -E "passed|failed|error" $OUT/pytest_$TAG.log | tail -3
grep -E "^lambda|^level|FAILED|Error" $OUT/pytest_$TAG.log | head -20
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 600 python bench.py --no-cpu-baseline --tuning 5=1 > $OUT/bench_${TAG}_t5.json 2> $OUT/bench_${TAG}_t5.err; echo "bench t5 exit $?"
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_t5.json')); print('t5', d['ms_per_step'], d['roofline']['frac'], d['kernel_ms_per_step'])"
timeout 600 python bench.py --no-cpu-baseline --adaptive 0 7 > $OUT/bench_${TAG}_c0.json 2> $OUT/bench_${TAG}_c0.err; echo "bench c0 exit $?"
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_c0.json')); print('c0', d['ms_per_step'], d['e2e'], d['stages'], d['kernel_ms_per_step'])"
timeout 900 python bench.py --no-cpu-baseline --adaptive 4 9 --threshold 1.6 --problem varcoef > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 exit $?"
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_c3.json')); print('c3', d['ms_per_step'], d['e2e'], d['e2e_device_sampling'], d['stages'], d['kernel_ms_per_step'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_${TAG}_ref.json 2> $OUT/bench_${TAG}_ref.err; echo "ref arm exit $?"; cat $OUT/bench_${TAG}_ref.json
bash tools/gpu_ncu.sh ${TAG}_c3 "leaf_var_factor_kernel leaf_var_solve_kernel coarsen_T_kernel" "--adaptive 4 9 --threshold 1.6 --problem varcoef"
