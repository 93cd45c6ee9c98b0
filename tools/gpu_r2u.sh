#!/bin/bash
# Round 2, one GPU: the level-9 tree of the multi-GPU runs on ONE GPU under the lean-T memory policy (the N = 1 anchor of that
# scaling series), and the adaptive / variable-coefficient configs on the current build
TAG=${1:-r2u}
OUT=gpurun_out; mkdir -p $OUT
F=$OUT/bench_${TAG}_l9_lean
timeout 900 python bench.py --no-cpu-baseline --level 9 --lean-T --steps 3 --warmup 3 > $F.json 2> $F.err; echo "bench L9 lean exit $?"; tail -3 $F.err
python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['roofline']['frac'], d['kernel_ms_per_step'], d['stages'])"
F=$OUT/bench_${TAG}_c0
timeout 600 python bench.py --no-cpu-baseline --adaptive 0 7 --steps 20 --warmup 5 > $F.json 2> $F.err; echo "bench c0 exit $?"; tail -2 $F.err
python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['stages'], d['kernel_ms_per_step'])"
F=$OUT/bench_${TAG}_c3
timeout 600 python bench.py --no-cpu-baseline --adaptive 4 9 --threshold 1.6 --problem varcoef --steps 10 --warmup 5 > $F.json 2> $F.err; echo "bench c3 exit $?"; tail -2 $F.err
python -c "import json; d=json.loads(open('$F.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['linf_error_vs_exact'], d['stages'], d['kernel_ms_per_step'], d.get('e2e'), d.get('e2e_device_sampling'))"
