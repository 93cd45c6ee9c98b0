"""SURVEY 8(f) rank 1 - the callers either side of the path: sampling coordinates written by the device and the error norms of
the reference's drivers reduced on the device (coefficient arrays handed over in device memory: tests/test_gpu_staged.py).

Oracle: the host formulas of the Python mirror (pinned to the compiled reference through the golden dumps) and a numpy
restatement of the driver loop examples/elliptic-multiple/main.cpp:346-371.  Coordinates: bit-exact.  Norms: linf exact,
l1 / l2 to 1e-12 relative (the device reduces in a tree, the reference loop adds sequentially)."""
import numpy as np
import pytest

import ellipticforest_b200 as ef
import hps_oracle as O
from test_host import _mesh_for

pytestmark = pytest.mark.gpu

CASES = {
    "uniform": dict(problem_name="poisson", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=2, max_level=2,
                    threshold=1.2, refine_box=None),
    "adaptive_rect": dict(problem_name="helmholtz", solver_kind="fishpack", box=(-10.0, 10.0, -3.0, 7.0), nx=8, min_level=1, max_level=4,
                          threshold=1.2, refine_box=None),
}


def _solver(kw):
    P = O.problem(kw["problem_name"])
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90" if kw["solver_kind"] == "fishpack" else "FivePointStencil"
    s.alpha_function, s.beta_function, s.lambda_function = P["alpha"], P["beta"], P["lam"]
    return P, s


@pytest.mark.parametrize("case", list(CASES))
def test_leaf_points_are_bit_identical_to_the_host_mirror(case):
    kw = CASES[case]
    m = _mesh_for(kw)
    P, s = _solver(kw)
    hps = ef.HPSAlgorithm(m, s)           # before any build: only the two small tables are uploaded
    X, Y = m.leaf_cell_centres()
    b = m.box[m.leaf_nodes]
    hx = ((b[:, 1] - b[:, 0]) / m.nx)[:, None, None] / 2.0
    hy = ((b[:, 3] - b[:, 2]) / m.nx)[:, None, None] / 2.0
    want = {"centre": (X, Y), "W": (X - hx, Y), "E": (X + hx, Y), "S": (X, Y - hy), "N": (X, Y + hy)}
    for which, (wx, wy) in want.items():
        x, y = hps.leafPoints(which)
        assert np.array_equal(x, np.broadcast_to(wx, x.shape)), which
        assert np.array_equal(y, np.broadcast_to(wy, y.shape)), which
    assert hps.stats()["device_bytes"] < 1e6 + 16 * X.size     # no operator storage was allocated for this


@pytest.mark.parametrize("case", list(CASES))
def test_error_norms_match_the_reference_driver_loop(case):
    kw = CASES[case]
    m = _mesh_for(kw)
    P, s = _solver(kw)
    hps = ef.HPSAlgorithm(m, s)
    with pytest.raises(ef.EfgpuError):
        hps.errorNorms(P["u"])            # no solution yet
    hps.buildStage()
    hps.upwardsStage(P["f"])
    u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0))
    X, Y = m.leaf_cell_centres()
    exact = P["u"](X, Y)
    b = m.box[m.leaf_nodes]
    dxdy = ((b[:, 1] - b[:, 0]) / m.nx) * ((b[:, 3] - b[:, 2]) / m.nx)
    diff = np.abs(u - exact)
    area = (kw["box"][1] - kw["box"][0]) * (kw["box"][3] - kw["box"][2])
    l1 = float(np.sum(dxdy[:, None, None] * diff)) / area
    l2 = float(np.sqrt(np.sum(dxdy[:, None, None] * diff ** 2) / area))
    li = float(diff.max())
    g1, g2, gi = hps.errorNorms(P["u"])
    assert gi == li
    assert abs(g1 - l1) <= 1e-12 * l1 and abs(g2 - l2) <= 1e-12 * l2, (g1, l1, g2, l2)
    assert hps.errorNorms(exact) == (g1, g2, gi)      # array form, and deterministic
    assert np.isfinite(gi) and 0.0 < g1 <= g2 <= gi   # area-weighted means: l1 <= l2 <= linf
