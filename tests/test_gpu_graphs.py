"""Stream lanes (independent batches of one tree level on parallel streams) and CUDA-graph replay of the three stages: the second
issue of a stage with the same key captures its launch sequence, later ones replay it.  Neither changes any arithmetic, so every
repetition must reproduce, bit for bit, what a handle with both switched off (EFGPU_GRAPHS=0, EFGPU_LANES=0) computes."""
import numpy as np
import pytest

import ellipticforest_b200 as ef
import hps_oracle as O
from test_host import _mesh_for

pytestmark = pytest.mark.gpu

CASES = {
    "uniform_l3_m16": dict(problem_name="poisson", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=3, max_level=3,
                           threshold=1.2, refine_box=None),
    "adaptive_l0_5_m8": dict(problem_name="helmholtz", solver_kind="fishpack", box=(-10.0, 10.0, -10.0, 10.0), nx=8, min_level=0, max_level=5,
                             threshold=1.2, refine_box=None),
    "adaptive_l1_4_m16_varcoef": dict(problem_name="varcoef", solver_kind="fivepoint", box=(-10.0, 10.0, -10.0, 10.0), nx=16, min_level=1, max_level=4,
                                      threshold=1.2, refine_box=None),
}


def _make(kw):
    P = O.problem(kw["problem_name"])
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90" if kw["solver_kind"] == "fishpack" else "FivePointStencil"
    s.alpha_function, s.beta_function, s.lambda_function = P["alpha"], P["beta"], P["lam"]
    return P, ef.HPSAlgorithm(_mesh_for(kw), s)


@pytest.mark.parametrize("case", list(CASES))
def test_graph_replay_and_lanes_reproduce_the_plain_launch_sequence(case, monkeypatch):
    kw = CASES[case]
    monkeypatch.setenv("EFGPU_GRAPHS", "0")
    monkeypatch.setenv("EFGPU_LANES", "0")
    P, plain = _make(kw)
    plain.buildStage(); plain.upwardsStage(P["f"])
    bc = lambda side, x, y: (P["u"](x, y), 1.0, 0.0)
    u0 = plain.solveStage(bc).copy()
    T0, S0 = plain.operator(0, "T"), plain.operator(0, "S")
    plain.upwardsStage(P["f"], 1.5)
    u0s = plain.solveStage(lambda side, x, y: (1.5 * P["u"](x, y), 1.0, 0.0)).copy()
    launches0 = sum(v[1] for v in plain.profile().values())
    monkeypatch.delenv("EFGPU_GRAPHS")
    monkeypatch.delenv("EFGPU_LANES")
    P, hps = _make(kw)
    for rep in range(4):           # 0: eager, 1: capture + launch, 2, 3: replay
        hps.buildStage()
        hps.upwardsStage(P["f"])
        assert np.array_equal(hps.solveStage(bc), u0), rep
        assert np.array_equal(hps.operator(0, "T"), T0) and np.array_equal(hps.operator(0, "S"), S0), rep
        hps.upwardsStage(P["f"], 1.5)          # another load on the resident operators: same graphs, new data
        assert np.array_equal(hps.solveStage(lambda side, x, y: (1.5 * P["u"](x, y), 1.0, 0.0)), u0s), rep
    # launch accounting survives the replay (bench.py's gpu_launches): four repetitions of what the plain handle launched once
    assert sum(v[1] for v in hps.profile().values()) == 4 * launches0
    # a different key (homogeneous right-hand side changes the solve sweep's kernels) does not reuse the captured graph
    hps.options["homogeneous-rhs"] = True
    plain.options["homogeneous-rhs"] = True
    plain.upwardsStage(P["f"]); hps.upwardsStage(P["f"])
    assert np.array_equal(hps.solveStage(bc), plain.solveStage(bc))
    st = hps.stats()
    assert st["build_ms"] > 0 and st["solve_ms"] > 0
