"""CPU tests that pin the oracle (oracle/hps_oracle.py) to the reference.

* golden fixtures = full dumps of the unmodified reference (tests/golden/make_golden.py);
* the reference's own known-answer tests for the primitives on this path;
* the 9-digit transcript in examples/patch-solver/README.md:17-25.
"""
import numpy as np
import pytest

import hps_oracle as O
from conftest import GOLDEN_CASES, golden_case_args, load_golden

TOL = 1e-11  # relative max-norm, oracle vs compiled reference (both LAPACK partial pivoting)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_dump(case):
    gold = load_golden(case)
    hps = O.run(**golden_case_args(gold))
    nodes = hps.nodes
    # ordering contract: p4est DFS, post-order for merge, pre-order for split (bit-exact)
    post = "".join(nodes[i].path + ("L" if nodes[i].leaf else "P") + ";" for i in O.post_order(nodes))
    pre = "".join(nodes[i].path + ";" for i in O.pre_order(nodes))
    assert post == gold["order/post"]
    assert pre == gold["order/pre"]
    seen = set()
    for nd in nodes:
        box = gold["grid0/" + nd.path]
        meta = gold["build/%s/meta" % nd.path]
        assert int(meta[0]) == nd.n_coarsens
        assert int(box[5]) == nd.level
        for stage, names in (("build", "TSXH"), ("up", "hwf"), ("solve", "gu")):
            for nm in names:
                key = "%s/%s/%s" % (stage, nd.path, nm)
                if key in gold:
                    mine = getattr(nd, nm)
                    assert mine is not None and mine.shape == gold[key].shape, key
                    assert relerr(mine, gold[key]) < TOL, key
                    seen.add(nm)
    assert seen >= (set("TSXHhwfgu") if any(not nd.leaf for nd in nodes) else set("Thfgu"))   # (a single-patch tree has no merge)
    # child boxes are produced by midpoint splitting: bit-exact
    for nd in nodes:
        if nd.leaf:
            box = gold["grid0/" + nd.path]
            assert (nd.grid.xl, nd.grid.xu, nd.grid.yl, nd.grid.yu) == tuple(box[:4])


def test_tag2_fixture_really_has_double_coarsening():
    gold = load_golden("adaptive_tag2_m8_helmholtz_rect")
    ncs = [int(v[0]) for k, v in gold.items() if k.endswith("/meta")]
    assert max(ncs) == 2


def test_linear_solve_known_answer():
    # test/test_matrix.cpp:155-192
    A = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 10.0]])
    x = O.sla.solve(A, np.ones(3))
    assert np.max(np.abs(x - np.array([-1, 1, 0.0]))) < 1e-14
    B = np.tile(np.array([1, 2, 3, 4.0]), (3, 1))
    X = O.sla.solve(A, B)
    assert np.max(np.abs(X - np.array([[-1, -2, -3, -4], [1, 2, 3, 4], [0, 0, 0, 0.0]]))) < 1e-13


def test_block_permute_known_answer():
    # test/test_matrix.cpp:270-300 with equal block sizes (the only form the HPS path uses)
    m = np.arange(6 * 4, dtype=np.float64).reshape(6, 4)
    out = O.block_permute_rows(m, [2, 0, 1], 2)
    assert np.array_equal(out, np.concatenate([m[4:6], m[0:2], m[2:4]]))
    # WESN permutation used by reorderOperators_ (HPSAlgorithm.hpp:984)
    v = np.repeat(np.arange(8.0), 3)
    assert np.array_equal(O.block_permute_rows(v, O.PI_WESN, 3)[::3], np.array([0, 4, 2, 6, 1, 3, 5, 7.0]))


def test_grid_points_known_answer():
    # test/test_finite_volume_grid.cpp:30-49
    g = O.Grid(4, 0.0, 1.0, 1.0, 3.0)
    assert np.allclose(g.x(np.arange(4)), [0.125, 0.375, 0.625, 0.875], atol=1e-16, rtol=0)
    assert np.allclose(g.y(np.arange(4)), [1.25, 1.75, 2.25, 2.75], atol=1e-16, rtol=0)


@pytest.mark.parametrize("n,expected", [(8, 6.62291775e-03), (16, 1.65048768e-03), (32, 4.12493989e-04)])
def test_patch_solver_transcript(n, expected):
    # examples/patch-solver/README.md:17-25 (FivePointStencil, [-1,1]^2, u = sin x + sin y)
    g = O.Grid(n, -1.0, 1.0, -1.0, 1.0)
    u = lambda x, y: np.sin(x) + np.sin(y)
    s = O.Solver(kind="fivepoint")
    ys = g.y(np.arange(n)); xs = g.x(np.arange(n))
    gd = np.concatenate([u(-1.0, ys), u(1.0, ys), u(xs, -1.0), u(xs, 1.0)])
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    sol = s.solve(g, gd, (-u(X, Y)).reshape(-1))
    err = np.max(np.abs(sol - u(X, Y).reshape(-1)))
    assert abs(err - expected) < 5e-9 * 1  # transcript prints 9 significant digits
    # the FISHPACK-type system is the same linear system for alpha = beta = 1
    sol2 = O.Solver(kind="fishpack").solve(g, gd, (-u(X, Y)).reshape(-1))
    assert relerr(sol2, sol) < 1e-12


def test_interpolation_stencils():
    # src/SpecialMatrices.hpp:93-173
    L = O.L12(8)
    assert np.array_equal(L[0, :3], [1.40625, -0.5625, 0.15625])
    assert np.array_equal(L[7, 1:], [0.15625, -0.5625, 1.40625])
    assert np.array_equal(L[1, :2], [0.75, 0.25]) and np.array_equal(L[2, :2], [0.25, 0.75])
    assert np.allclose(L.sum(axis=1), 1.0)
    assert np.array_equal(O.L21(4)[1], [0, 0, 0.5, 0.5, 0, 0, 0, 0])
