"""Edge cases of the tree shape through the C-ABI: the smallest trees the reference accepts."""
import numpy as np
import pytest

import hps_oracle as O
from test_gpu_parity import TOL, relerr, run_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,problem", [(16, "poisson"), (8, "helmholtz")])
def test_single_patch_tree(nx, problem):
    """min_level = max_level = 0: the root is itself a leaf, so buildStage is one buildD2N (HPSAlgorithm.hpp:128-141), the
    upwards stage one particularNeumannData and the solve stage one leafSolve (:584-596); there is no merge at all."""
    kw = dict(problem_name=problem, solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=nx, min_level=0, max_level=0,
              threshold=1.2, refine_box=None)
    hps = run_gpu(kw, keep_x=False)
    ora = O.run(**kw)
    assert hps.mesh.n_nodes == 1 and len(ora.nodes) == 1 and hps.node_info(0)["leaf"]
    nd = ora.nodes[0]
    assert relerr(hps.operator(0, "T"), nd.T) < TOL
    for nm in ("h", "g", "u"):
        assert relerr(hps.vector(0, nm), getattr(nd, nm)) < TOL, nm
    assert hps.stats()["merge_flops_issued"] == 0.0


def test_blocked_tensor_core_base_case_against_register_kernel():
    """efgpu_set_tuning(7, ...): the 128 x 128 base case of the block inversion as a blocked Gauss-Jordan on the FP64 tensor pipe
    (default) against the per-pivot register kernel of round 1: same operators up to rounding, both within 1e-10 of the oracle
    (uniform level-4 tree of 16x16 patches: base cases of 64 and 128 rows, zipped pairs at the root)."""
    import ellipticforest_b200 as ef
    from ellipticforest_b200 import _lib
    from test_host import _mesh_for
    kw = dict(problem_name="helmholtz", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=4, max_level=4,
              threshold=1.2, refine_box=None)
    P = O.problem(kw["problem_name"])
    lib = _lib.load()
    out = {}
    for key in (1, 0):
        assert lib.efgpu_set_tuning(7, key) == 0
        try:
            s = ef.FiniteVolumeSolver()
            s.solver_type = "FISHPACK90"
            s.lambda_function = P["lam"]
            hps = ef.HPSAlgorithm(_mesh_for(kw), s)
            hps.buildStage(); hps.upwardsStage(P["f"])
            u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0)).copy()
            out[key] = (u, hps.operator(0, "T"), hps.operator(0, "S"), hps.operator(0, "Xinv"), hps.stats())
        finally:
            lib.efgpu_set_tuning(7, 0)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    for k in range(4):
        assert rel(out[0][k], out[1][k]) < 1e-11, k
    assert abs(out[0][4]["min_pivot"] / out[1][4]["min_pivot"] - 1.0) < 1e-6       # the same pivots, in blocks of eight
    ora = O.run(**kw)
    assert rel(out[0][1], ora.nodes[0].T) < 1e-10 and rel(out[0][2], ora.nodes[0].S) < 1e-10


@pytest.mark.parametrize("level,nx", [(4, 16), (3, 8), (5, 8)])
def test_staged_transposed_stores_are_bit_identical(level, nx):
    """efgpu_set_tuning(4, ...): the transposed second destination of a GEMM block (lower blocks of the symmetric X^-1, mirrored
    blocks of T) written from the accumulators with 8-byte stores (default on one GPU) against the form that assembles the
    transposed tile in shared memory and stores whole rows (default where peer arenas receive the block): only the store path
    differs, so every operator and the solution must agree bit for bit, over every CTA tile size the levels of these trees use."""
    import ellipticforest_b200 as ef
    from ellipticforest_b200 import _lib
    from test_host import _mesh_for
    kw = dict(problem_name="poisson", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=nx, min_level=level, max_level=level,
              threshold=1.2, refine_box=None)
    P = O.problem(kw["problem_name"])
    lib = _lib.load()
    out = {}
    for key in (2, 1):
        assert lib.efgpu_set_tuning(4, key) == 0
        try:
            s = ef.FiniteVolumeSolver()
            s.solver_type = "FISHPACK90"
            hps = ef.HPSAlgorithm(_mesh_for(kw), s)
            hps.buildStage(); hps.upwardsStage(P["f"])
            u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0)).copy()
            assert hps.stats()["symmetric_plan"] if "symmetric_plan" in hps.stats() else True
            out[key] = [u] + [hps.operator(nd, w) for nd in (0, 1) for w in ("T", "S", "Xinv")]
        finally:
            lib.efgpu_set_tuning(4, 0)
    for k, (a, b) in enumerate(zip(out[1], out[2])):
        assert np.array_equal(a, b), k


@pytest.mark.parametrize("level,nx,problem", [(6, 16, "poisson"), (5, 32, "helmholtz")])
def test_tma_operand_staging_is_bit_identical(level, nx, problem):
    """efgpu_set_tuning(8, ...): the 128-row tiles of the merge products with operands staged by TMA (cp.async.bulk.tensor.2d through
    per-(parent, view) tensor maps, swizzled shared memory, mbarrier ring; default) against the cp.async kernel: the same DMMA sequence
    in the same k order, so every operator and the solution agree bit for bit.  Trees whose upper levels are large enough for the
    128 x 64 tiles to be chosen (root child side 512: S, T, the split recursion of X^-1, transposed second destinations)."""
    import ellipticforest_b200 as ef
    from ellipticforest_b200 import _lib
    from test_host import _mesh_for
    kw = dict(problem_name=problem, solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=nx, min_level=level, max_level=level,
              threshold=1.2, refine_box=None)
    P = O.problem(kw["problem_name"])
    lib = _lib.load()
    out = {}
    for key in (0, 1):
        assert lib.efgpu_set_tuning(8, key) == 0
        try:
            s = ef.FiniteVolumeSolver()
            s.solver_type = "FISHPACK90"
            s.lambda_function = P["lam"]
            hps = ef.HPSAlgorithm(_mesh_for(kw), s)
            hps.buildStage(); hps.upwardsStage(P["f"])
            u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0)).copy()
            out[key] = [u] + [hps.operator(nd, w) for nd in (0, 1) for w in ("T", "S", "Xinv")]
        finally:
            lib.efgpu_set_tuning(8, 1)
    for k, (a, b) in enumerate(zip(out[1], out[0])):
        assert np.array_equal(a, b), k


@pytest.mark.parametrize("level,nx,problem", [(4, 16, "helmholtz"), (3, 32, "poisson"), (5, 8, "helmholtz:9")])
def test_cluster_base_case_against_the_128_row_recursion(level, nx, problem):
    """efgpu_set_tuning(9, ...): in batches of at most four merges (the serial chain of the top tree levels) a whole 256 x 256 block of
    the block inversion is one kernel - eight CTAs of a thread-block cluster, column slabs in distributed shared memory, DMMA rank-8
    updates - against the recursion down to 128 x 128 base cases: same operators up to rounding, same pivots, both within
    1e-10 of the oracle.  Trees whose root X has order 512 (two 256-blocks per half, zipped pairs) and whose level 1 has order 256."""
    import ellipticforest_b200 as ef
    from ellipticforest_b200 import _lib
    from test_host import _mesh_for
    kw = dict(problem_name=problem, solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=nx, min_level=level, max_level=level,
              threshold=1.2, refine_box=None)
    P = O.problem(kw["problem_name"])
    lib = _lib.load()
    out = {}
    for key in (0, 1):
        assert lib.efgpu_set_tuning(9, key) == 0
        try:
            s = ef.FiniteVolumeSolver()
            s.solver_type = "FISHPACK90"
            s.lambda_function = P["lam"]
            hps = ef.HPSAlgorithm(_mesh_for(kw), s)
            hps.buildStage(); hps.upwardsStage(P["f"])
            u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0)).copy()
            out[key] = ([u] + [hps.operator(nd, w) for nd in (0, 1) for w in ("T", "S", "Xinv")], hps.stats())
        finally:
            lib.efgpu_set_tuning(9, 0)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    tol = 1e-11 if ":" not in problem else 1e-9      # indefinite operator: conditioning, see test_indefinite_helmholtz_against_oracle
    for k, (a, b) in enumerate(zip(out[1][0], out[0][0])):
        assert rel(a, b) < tol, (k, rel(a, b))
    assert abs(out[1][1]["min_pivot"] / out[0][1]["min_pivot"] - 1.0) < 1e-6
    assert out[1][1]["negative_pivots"] == out[0][1]["negative_pivots"]
    if ":" not in problem:
        ora = O.run(**kw)
        assert rel(out[1][0][1], ora.nodes[0].T) < 1e-10 and rel(out[1][0][2], ora.nodes[0].S) < 1e-10


def test_fast_pivot_reciprocal_against_ieee_division():
    """efgpu_set_tuning(10, 1): the pivot reciprocals of the 128 x 128 base case from the hardware seed and two Newton steps instead
    of the IEEE division sequence (shorter dependent chain per pivot): same operators to rounding."""
    import ellipticforest_b200 as ef
    from ellipticforest_b200 import _lib
    from test_host import _mesh_for
    kw = dict(problem_name="helmholtz", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=4, max_level=4,
              threshold=1.2, refine_box=None)
    P = O.problem(kw["problem_name"])
    lib = _lib.load()
    out = {}
    for key in (0, 1):
        assert lib.efgpu_set_tuning(10, key) == 0
        try:
            s = ef.FiniteVolumeSolver()
            s.solver_type = "FISHPACK90"
            s.lambda_function = P["lam"]
            hps = ef.HPSAlgorithm(_mesh_for(kw), s)
            hps.buildStage(); hps.upwardsStage(P["f"])
            u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0)).copy()
            out[key] = [u, hps.operator(0, "T"), hps.operator(0, "S"), hps.operator(0, "Xinv")]
        finally:
            lib.efgpu_set_tuning(10, 0)      # the default
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    for k in range(4):
        assert rel(out[0][k], out[1][k]) < 1e-11, k

