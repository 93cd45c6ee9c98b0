"""CPU tests of the host logic above the C-ABI: the library loads, exports every declared symbol,
fails loudly without a GPU, and the stand-alone mesh builder reproduces p4est's result."""
import ctypes
import os
import re

import numpy as np
import pytest

import ellipticforest_b200 as ef
from ellipticforest_b200 import _lib
from conftest import GOLDEN_CASES, ROOT, golden_case_args, load_golden
import hps_oracle as O


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "efgpu.h")).read()
    declared = set(re.findall(r"\b(efgpu_[a-z_]+)\s*\(", header)) - {"efgpu_refine_fn"}
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_builtin_refine_indicator_matches_the_callback():
    """efgpu_refine_elliptic_single (used for large adaptive meshes built from Python) = the per-point callback."""
    g = ef.FiniteVolumeGrid(16, -10.0, 10.0, 16, -10.0, 10.0)
    for thr, lo, hi in [(1.2, 0, 5), (1.6, 2, 5)]:
        a = ef.Mesh().refineByFunction("elliptic-single", thr, lo, hi, g)
        b = ef.Mesh().refineByFunction(lambda x, y: abs(-(np.sin(x) + np.sin(y))) > thr, thr, lo, hi, g)
        assert np.array_equal(a.level, b.level) and np.array_equal(a.child, b.child) and np.array_equal(a.box, b.box)
        assert np.array_equal(a.leaf_nodes, b.leaf_nodes)


def _mesh_for(kw):
    ind = O.refine_box_indicator(kw["refine_box"]) if kw["refine_box"] is not None else O.refine_indicator(kw["threshold"])
    g = ef.FiniteVolumeGrid(kw["nx"], kw["box"][0], kw["box"][1], kw["nx"], kw["box"][2], kw["box"][3])
    return ef.Mesh().refineByFunction(lambda x, y: bool(ind(x, y)), kw["threshold"], kw["min_level"], kw["max_level"], g)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_mesh_builder_matches_p4est(case):
    """Leaf/level indexing and traversal order bit-exact against the reference's real p4est run."""
    gold = load_golden(case)
    m = _mesh_for(golden_case_args(gold))
    leaf = set(int(i) for i in m.leaf_nodes)
    pre = "".join(m.path(i) + ";" for i in range(m.n_nodes))
    post = "".join(m.path(i) + ("L" if i in leaf else "P") + ";" for i in m.post_order())
    assert pre == gold["order/pre"]
    assert post == gold["order/post"]
    for i in range(m.n_nodes):
        box = gold["grid0/" + m.path(i)]
        assert int(box[5]) == m.level[i]
        if i in leaf:  # child boxes by midpoint splitting: bit-exact
            assert tuple(m.box[i]) == tuple(box[:4])


def test_uniform_mesh_counts():
    g = ef.FiniteVolumeGrid(16, 0.0, np.pi, 16, 0.0, np.pi)
    m = ef.Mesh().refineByFunction(None, 0.0, 5, 5, g)
    assert m.n_leaves == 4 ** 5 and m.n_nodes == (4 ** 6 - 1) // 3
    X, Y = m.leaf_cell_centres()
    assert X.shape == (1024, 16, 16)
    # first leaf is the lower-left corner patch, second is its right neighbour (Morton order)
    assert X[0, 0, 0] < X[1, 0, 0] and Y[0, 0, 0] == Y[1, 0, 0]


def test_create_fails_loudly_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    g = ef.FiniteVolumeGrid(8, 0.0, 1.0, 8, 0.0, 1.0)
    m = ef.Mesh().refineByFunction(None, 0.0, 1, 1, g)
    with pytest.raises(ef.EfgpuError) as e:
        ef.HPSAlgorithm(m, ef.FiniteVolumeSolver())
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_rejected():
    lib = _lib.load()
    out = ctypes.c_void_p()
    assert lib.efgpu_mesh_create(0.0, 1.0, 0.0, 1.0, 8, 3, 2, _lib.REFINE_FN(), None, ctypes.byref(out)) == 2
    assert lib.efgpu_mesh_create(1.0, 0.0, 0.0, 1.0, 8, 0, 1, _lib.REFINE_FN(), None, ctypes.byref(out)) == 2
    assert lib.efgpu_build(None, 0) == 2
    d = ctypes.c_double()
    assert lib.efgpu_leaf_points(None, 0, None, None) == 2
    assert lib.efgpu_leaf_points_device(None, 9, None, None, 1) == 2
    assert lib.efgpu_error_norms(None, None, ctypes.byref(d), ctypes.byref(d), ctypes.byref(d)) == 2
    assert lib.efgpu_error_norms_device(None, None, None, None, None, None) == 2
    assert lib.efgpu_set_leaf_variable_device(None, None, None, None, None, None, None) == 2


@pytest.mark.parametrize("threads", [1, 4])
def test_cpp_binding_samples_like_the_reference_and_has_no_cpu_fallback(threads):
    """include/EllipticForestB200.hpp compiled against the unmodified reference (oracle/_ref/dropin_driver), without a
    device: setupStage and upwardsStage must throw (nothing falls back to the reference's CPU stages), and the load the
    binding samples into every leaf's vectorF - serially or on `sampling_threads` host threads - is bit-identical to what
    the reference's own upwardsStage stored (HPSAlgorithm.hpp:241-249)."""
    import json
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/dropin_driver not built (needs /root/reference at build time)")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([driver, "--problem", "poisson", "--solver", "fishpack", "--min-level", "1", "--max-level", "4", "--nx", "8",
                        "--domain", "-10", "10", "-10", "10", "--sampling-only", "--threads", str(threads)],
                       capture_output=True, text=True, timeout=300, env=env)
    line = [l for l in p.stdout.splitlines() if l.startswith("DROPIN_SAMPLING")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(line[-1][len("DROPIN_SAMPLING "):])
    assert res["same"] and res["threw"] == 2 and res["threads"] == threads and res["cells"] == res["leaves"] * 64 > 0, res
