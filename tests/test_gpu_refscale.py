"""Parity at benchmark scale against the UNMODIFIED reference, run on this box.

The per-node golden fixtures stop at child side n = 128; the benchmarked tree (BASELINE configs[1]) reaches n = 2048 and six
levels of unpivoted block inversion.  Here oracle/_ref/ref_driver (the reference compiled from its own sources by oracle/Makefile;
the binary travels with the snapshot, /root/reference does not) builds and solves a uniform level-6 tree of 16x16 patches
(1.05 M DOFs, root child side n = 512, X of order 2048, four recursion levels of the blocked inversion - everything the level-8 tree
does except two more levels of the same recursion) and a level-5 tree of 32x32 patches with Helmholtz lambda = -1, dumps the
root's DtN map T, solution operator S, h, w and every leaf's u (--dump-root-only), and the CUDA path is compared entry by entry:
relative max-norm 1e-10 (north_star).  Reference: src/HPSAlgorithm.hpp:870-968 (mergeX_/mergeS_/mergeT_ is what the dump pins).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import ellipticforest_b200 as ef
import hps_oracle as O
from conftest import ROOT
from refdump import read_dump

pytestmark = pytest.mark.gpu
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
TOL = 1e-10


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("problem,level,nx", [("poisson", 6, 16), ("helmholtz", 5, 32)])
def test_bench_scale_tree_against_compiled_reference(problem, level, nx, tmp_path):
    if not os.path.exists(REF_DRIVER):
        pytest.skip("oracle/_ref/ref_driver not built (needs /root/reference at build time)")
    threads = str(max(1, min(len(os.sched_getaffinity(0)), 64)))
    dump = str(tmp_path / "ref.bin")
    cmd = [REF_DRIVER, "--problem", problem, "--solver", "fishpack", "--min-level", str(level), "--max-level", str(level), "--nx", str(nx),
           "--domain", "0", repr(np.pi), "0", repr(np.pi), "--dump", dump, "--dump-root-only", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True, env=dict(os.environ, OPENBLAS_NUM_THREADS=threads)).stdout
    assert "REF_RESULT" in out
    D = read_dump(dump)
    os.remove(dump)

    P = O.problem(problem)
    mesh = ef.Mesh().refineByFunction(None, 0.0, level, level, ef.FiniteVolumeGrid(nx, 0.0, np.pi, nx, 0.0, np.pi))
    # traversal order of the real p4est run, bit-exact
    assert "".join(mesh.path(i) + ";" for i in range(mesh.n_nodes)) == D["order/pre"]
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90"
    s.lambda_function = P["lam"]
    hps = ef.HPSAlgorithm(mesh, s)
    hps.buildStage()
    hps.upwardsStage(P["f"])
    u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0))
    assert hps.is_symmetric()
    err = {"T": relerr(hps.operator(0, "T"), D["build/0/T"]), "S": relerr(hps.operator(0, "S"), D["build/0/S"]),
           "h": relerr(hps.vector(0, "h"), D["up/0/h"]), "w": relerr(hps.vector(0, "w"), D["up/0/w"])}
    u_ref = np.stack([D["solve/%s/u" % mesh.path(int(i))].reshape(nx, nx) for i in mesh.leaf_nodes])
    err["u"] = relerr(u, u_ref)
    st = hps.stats()
    print("level %d nx %d %s: %s; pivots min %.3e max %.3e block ratio %.3e negative %d" % (
        level, nx, problem, {k: "%.1e" % v for k, v in err.items()}, st["min_pivot"], st["max_pivot"], st["pivot_ratio_min"], st["negative_pivots"]))
    for k, v in err.items():
        assert v < TOL, (k, v)
    assert st["negative_pivots"] == 0        # lambda <= 0: every merge matrix is positive definite
    # the general plan (what adaptive trees take) at the same size
    del hps
    gen = ef.HPSAlgorithm(mesh, s)
    gen.no_symmetry = True
    gen.buildStage()
    gen.upwardsStage(P["f"])
    u2 = gen.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0))
    assert not gen.is_symmetric()
    assert relerr(gen.operator(0, "T"), D["build/0/T"]) < TOL and relerr(gen.operator(0, "S"), D["build/0/S"]) < TOL and relerr(u2, u_ref) < TOL
