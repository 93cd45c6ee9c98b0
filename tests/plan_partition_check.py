"""Helper of tests/test_merge_plan.py, run as a subprocess because EFGPU_SPLIT_MIN_ROWS is read once per process: emulates the
merge plan of one batch at `nranks` ranks (row-partitioned products of the block inversion, S and T rows, all-gathers) on
synthetic signed-symmetric children and compares X^-1, S, T with the oracle's merge4to1.
usage: plan_partition_check.py n nranks split_min_rows [plans, default "1,0"] [tuning KEY=VALUE]"""
import os, sys, time
os.environ["EFGPU_SPLIT_MIN_ROWS"] = sys.argv[3] if len(sys.argv) > 3 else "256"
HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE):
    sys.path.insert(0, p)
import numpy as np
import test_merge_plan as tm
n, nranks = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 5:      # "KEY=VALUE": efgpu_set_tuning before the plans are made
    from ellipticforest_b200 import _lib
    assert _lib.load().efgpu_set_tuning(*[int(v) for v in sys.argv[5].split("=")]) == 0
rng = np.random.default_rng(1)
d = np.ones(4 * n); d[:n] = -1; d[2 * n:3 * n] = -1      # W, E, S, N: d = -1 on W and S
A = rng.standard_normal((4 * n, 4 * n)) / np.sqrt(4 * n)
M = A @ A.T * 0.3 + np.eye(4 * n)
T = d[:, None] * M
Tc = [T, T, T, T]
t0 = time.time(); ref = tm.oracle_merge(Tc, n); print("oracle %.1fs" % (time.time() - t0))
for sym in [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "1,0").split(",")]:
    for level in (1, 0):
        t0 = time.time()
        states, flops = tm.emulate(n, level, nranks, sym, Tc, ref.X)
        e = [max(tm.rel(tm.view(s.ops[tm.OP_XINV], 0, 4 * n, 4 * n, 4 * n), np.linalg.inv(ref.X)), tm.rel(s.ops[tm.OP_S].reshape(4 * n, 8 * n), ref.S),
                 tm.rel(s.ops[tm.OP_T].reshape(8 * n, 8 * n), ref.T)) for s in states]
        plans = tm.get_plan(n, level, 0, nranks, sym)
        ngather = sum(1 for st in plans[0] if int(st[6]))
        print("n", n, "nranks", nranks, "sym", sym, "level", level, "max err %.2e" % max(e), "gathers", ngather, "flops/n^3 %.1f" % (flops / n ** 3 / nranks), "%.1fs" % (time.time() - t0))
        assert max(e) < 1e-10
