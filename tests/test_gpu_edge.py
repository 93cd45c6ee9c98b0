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
