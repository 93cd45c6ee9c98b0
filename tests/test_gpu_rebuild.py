"""Adaptive re-build: after a local change of the mesh only the dirty ancestor chains are merged again,
the operators of unchanged subtrees are taken over from the previous build.  Reference: the capability paper.md:44 advertises;
src/HPSAlgorithm.hpp:50-55 (isBuilt, never read) is all the reference has of it, so the parity statement is the one the paper
implies: the re-build must equal a build from scratch of the new mesh - here bit for bit, node by node."""
import numpy as np
import pytest

import ellipticforest_b200 as ef
import hps_oracle as O
from test_host import _mesh_for

pytestmark = pytest.mark.gpu


def _kw(problem, solver, nx, lo, hi, box):
    return dict(problem_name=problem, solver_kind=solver, box=(-10.0, 10.0, -10.0, 10.0), nx=nx, min_level=lo, max_level=hi,
                threshold=1.2, refine_box=box)


def _hps(kw):
    P = O.problem(kw["problem_name"])
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90" if kw["solver_kind"] == "fishpack" else "FivePointStencil"
    s.alpha_function, s.beta_function, s.lambda_function = P["alpha"], P["beta"], P["lam"]
    h = ef.HPSAlgorithm(_mesh_for(kw), s)
    h.keep_x = True
    return P, h


@pytest.mark.parametrize("name,problem,solver,nx,lo,hi,box_a,box_b", [
    ("refine", "helmholtz", "fishpack", 8, 2, 5, (-10.0, -5.0, -10.0, -5.0), (-10.0, -3.5, -10.0, -5.0)),        # the refined corner grows
    ("coarsen", "poisson", "fishpack", 16, 1, 4, (2.0, 10.0, -3.0, 10.0), (4.0, 10.0, 0.0, 10.0)),               # ... and shrinks
    ("varcoef", "varcoef", "fivepoint", 8, 2, 4, (-10.0, -5.0, -10.0, -5.0), (-10.0, -5.0, -10.0, -2.0)),
])
def test_rebuild_equals_build_from_scratch(name, problem, solver, nx, lo, hi, box_a, box_b):
    P, old = _hps(_kw(problem, solver, nx, lo, hi, box_a))
    old.buildStage()
    P, fresh = _hps(_kw(problem, solver, nx, lo, hi, box_b))
    fresh.buildStage()
    P, re = _hps(_kw(problem, solver, nx, lo, hi, box_b))
    reused, rebuilt = re.rebuildStage(old)
    m = re.mesh
    n_parents = m.n_nodes - m.n_leaves
    assert reused + rebuilt == n_parents and reused > 0 and 0 < rebuilt < n_parents, (reused, rebuilt)
    assert m.n_nodes != old.mesh.n_nodes                      # the mesh really changed
    for i in range(m.n_nodes):
        info = re.node_info(i)
        assert info == fresh.node_info(i)
        for nm in (["T"] if info["leaf"] else ["T", "S", "X", "H", "Xinv"]):
            assert np.array_equal(re.operator(i, nm), fresh.operator(i, nm)), (m.path(i), nm)
    f_re, f_fresh = re.stats()["merge_flops_issued"], fresh.stats()["merge_flops_issued"]
    assert 0 < f_re < f_fresh
    bc = lambda side, x, y: (P["u"](x, y), 1.0, 0.0)
    re.upwardsStage(P["f"]); fresh.upwardsStage(P["f"])
    assert np.array_equal(re.solveStage(bc), fresh.solveStage(bc))
    # the old handle is untouched and the re-built one builds from scratch like any other
    old.upwardsStage(P["f"]); old.solveStage(bc)
    re.buildStage(); re.upwardsStage(P["f"])
    assert np.array_equal(re.solveStage(bc), fresh.u_leaves)
    print(name, "merges reused %d, rebuilt %d of %d; flops %.3g of %.3g" % (reused, rebuilt, n_parents, f_re, f_fresh))


def test_rebuild_rejects_what_it_cannot_reuse():
    P, old = _hps(_kw("poisson", "fishpack", 8, 1, 3, (-10.0, 0.5, -10.0, 0.5)))
    P, new = _hps(_kw("helmholtz", "fishpack", 8, 1, 3, (-10.0, 0.5, -10.0, 0.5)))
    with pytest.raises(ef.EfgpuError):
        new.rebuildStage(old)                 # old has not been built
    old.buildStage()
    with pytest.raises(ef.EfgpuError):
        new.rebuildStage(old)                 # lambda changed: every operator is dirty
