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


def _read_vtu(path):
    """Minimal reader of the raw-appended .vtu efgpu_write_vtu produces: {array name: numpy array}."""
    import re
    raw = open(path, "rb").read()
    cut = raw.index(b'<AppendedData encoding="raw">')
    head = raw[:cut].decode()
    blob = raw[raw.index(b"_", cut) + 1:]
    assert 'header_type="UInt64"' in head and 'byte_order="LittleEndian"' in head
    out = {"n_points": int(re.search(r'NumberOfPoints="(\d+)"', head).group(1)), "n_cells": int(re.search(r'NumberOfCells="(\d+)"', head).group(1))}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*offset="(\d+)"', head):
        typ, name, off = m.group(1), m.group(2), int(m.group(3))
        nbytes = int(np.frombuffer(blob, dtype="<u8", count=1, offset=off)[0])
        dt = {"Float64": "<f8", "Int64": "<i8", "UInt8": "u1"}[typ]
        out[name] = np.frombuffer(blob, dtype=dt, count=nbytes // np.dtype(dt).itemsize, offset=off + 8)
    return out


def test_binary_vtu_from_device_buffers(tmp_path):
    """the mesh of Mesh::setMeshFromQuadtree (src/Mesh.hpp:186-267: four own corner points per leaf cell,
    leaves in traversePreOrder, cells i-slow j-fast, VTK_QUAD) and the solution as CellData, written as raw appended binary."""
    kw = dict(problem_name="poisson", solver_kind="fishpack", box=(-10.0, 10.0, -10.0, 10.0), nx=8, min_level=1, max_level=3,
              threshold=1.2, refine_box=(-10.0, 0.5, -10.0, 0.5))
    m = _mesh_for(kw)
    P = O.problem("poisson")
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90"
    hps = ef.HPSAlgorithm(m, s)
    hps.buildStage(); hps.upwardsStage(P["f"])
    u = hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0))
    path = tmp_path / "mesh.vtu"
    hps.toVTK(path, {"u": None})
    V = _read_vtu(path)
    M, nl = m.nx, m.n_leaves
    ncell = nl * M * M
    assert V["n_cells"] == ncell and V["n_points"] == 4 * ncell
    # the reference's corner formulas (Mesh.hpp:211-227), leaf by leaf
    b = m.box[m.leaf_nodes]
    dx, dy = (b[:, 1] - b[:, 0]) / M, (b[:, 3] - b[:, 2]) / M
    i = np.arange(M)
    x0 = b[:, 0][:, None] + i[None, :] * dx[:, None]; x1 = b[:, 0][:, None] + (i[None, :] + 1) * dx[:, None]
    y0 = b[:, 2][:, None] + i[None, :] * dy[:, None]; y1 = b[:, 2][:, None] + (i[None, :] + 1) * dy[:, None]
    pts = np.zeros((nl, M, M, 4, 3))
    pts[..., 0, 0] = x0[:, :, None]; pts[..., 1, 0] = x1[:, :, None]; pts[..., 2, 0] = x1[:, :, None]; pts[..., 3, 0] = x0[:, :, None]
    pts[..., 0, 1] = y0[:, None, :]; pts[..., 1, 1] = y0[:, None, :]; pts[..., 2, 1] = y1[:, None, :]; pts[..., 3, 1] = y1[:, None, :]
    assert np.array_equal(V["points"].reshape(nl, M, M, 4, 3), pts)                 # bit-exact
    assert np.array_equal(V["connectivity"], np.arange(4 * ncell))
    assert np.array_equal(V["offsets"], 4 * (np.arange(ncell) + 1))
    assert np.all(V["types"] == 9) and V["types"].size == ncell
    assert np.array_equal(V["u"].reshape(nl, M, M), u)
