"""Host-side mirror of the reference's operator/plugin interface for the HPS path, over the C-ABI.

Names, argument meaning and call order follow the reference so that the parity tests read like its
own drivers (examples/elliptic-single/main.cpp:104-217):

    grid   = FiniteVolumeGrid(nx, x_lower, x_upper, ny, y_lower, y_upper)
    mesh   = Mesh(); mesh.refineByFunction(fn, threshold, min_level, max_level, grid)
    solver = FiniteVolumeSolver(); solver.solver_type = "FISHPACK90"; solver.lambda_function = ...
    hps    = HPSAlgorithm(mesh, solver)
    hps.setupStage(); hps.buildStage(); hps.upwardsStage(f); hps.solveStage(bc)

The reference is C++; this Python layer exists for the tests and the benchmark harness (the C++
drop-in is include/EllipticForestB200.hpp).  All arithmetic of the path runs
in the CUDA library; numpy is used only to sample user callbacks at cell centres, as the
reference's host code does (src/HPSAlgorithm.hpp:241-249, :375-400).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import Stats, TreeDesc, check

CACHE_OPERATORS, HOMOGENEOUS_RHS, KEEP_X, LEAN_T, NO_SYMMETRY, LAZY_ROOT_DTN = 1, 2, 4, 8, 16, 32
OP = dict(T=0, S=1, X=2, H=3, Xinv=4, T_uncoarsened=5)
VEC = dict(h=0, w=1, g=2, u=3, f=4)


@dataclass
class FiniteVolumeGrid:
    """src/Patches/FiniteVolume/FiniteVolumeGrid.hpp:21-183 (cell-centred grid of one patch)."""
    nx: int
    x_lower: float
    x_upper: float
    ny: int
    y_lower: float
    y_upper: float

    @property
    def dx(self):
        return (self.x_upper - self.x_lower) / self.nx

    @property
    def dy(self):
        return (self.y_upper - self.y_lower) / self.ny

    def point(self, dim, index):  # FiniteVolumeGrid.cpp:21-36
        if dim == 0:
            return (self.x_lower + self.dx / 2) + np.asarray(index) * self.dx
        if dim == 1:
            return (self.y_lower + self.dy / 2) + np.asarray(index) * self.dy
        raise ValueError("Invalid `dim` argument")

    __call__ = point


class FiniteVolumeSolver:
    """src/Patches/FiniteVolume/FiniteVolumeSolver.hpp:53-162: the patch-solver plugin's public fields."""

    def __init__(self):
        self.solver_type = "FivePointStencil"  # FiniteVolumeSolver.hpp:64 default
        self.alpha_function: Callable = lambda x, y: 1.0 + 0 * x
        self.beta_function: Callable = lambda x, y: 1.0 + 0 * x
        self.lambda_function: Callable = lambda x, y: 0.0 * x

    def name(self):
        return "FiniteVolumeSolver"


class Mesh:
    """src/Mesh.hpp: quadtree mesh; the node table is in p4est pre-order."""

    def __init__(self):
        self._m = None
        self.root_grid: Optional[FiniteVolumeGrid] = None

    def __del__(self):
        if getattr(self, "_m", None):
            try:
                _lib.load().efgpu_mesh_destroy(self._m)
            except TypeError:      # interpreter shutdown: the module globals are already gone, the process frees the mesh
                pass
            self._m = None

    def refineByFunction(self, fn: Optional[Callable[[float, float], bool]], threshold, min_level, max_level,
                         root_grid: FiniteVolumeGrid):
        """src/Mesh.hpp:111-180.  `fn(x, y) -> bool` is evaluated at the nx*ny cell centres of a quadrant; the string
        "elliptic-single" selects the library's built-in |-(sin x + sin y)| > threshold (examples/elliptic-single)."""
        lib = _lib.load()
        if root_grid.nx != root_grid.ny:
            raise ValueError("square patches only (nx == ny), as the reference's merge assumes")
        self.root_grid = root_grid
        user = None
        if fn == "elliptic-single":
            thr = C.c_double(float(threshold))
            cb = C.cast(lib.efgpu_refine_elliptic_single, _lib.REFINE_FN)
            user = C.cast(C.pointer(thr), C.c_void_p)
        else:
            cb = _lib.REFINE_FN(lambda x, y, _u: 1 if fn(x, y) else 0) if fn is not None else _lib.REFINE_FN()
        out = C.c_void_p()
        check(lib.efgpu_mesh_create(root_grid.x_lower, root_grid.x_upper, root_grid.y_lower, root_grid.y_upper,
                                    root_grid.nx, min_level, max_level, cb, user, C.byref(out)))
        self._m = out
        d = TreeDesc()
        check(lib.efgpu_mesh_desc(self._m, C.byref(d)))
        self.desc = d
        n = d.n_nodes
        self.nx = d.nx
        self.level = np.ctypeslib.as_array(d.level, shape=(n,)).copy()
        self.child = np.ctypeslib.as_array(d.child, shape=(n, 4)).copy()
        self.box = np.ctypeslib.as_array(d.box, shape=(n, 4)).copy()
        nl = lib.efgpu_mesh_n_leaves(self._m)
        self.leaf_nodes = np.ctypeslib.as_array(lib.efgpu_mesh_leaf_nodes(self._m), shape=(nl,)).copy()
        return self

    @property
    def n_nodes(self):
        return int(self.desc.n_nodes)

    @property
    def n_leaves(self):
        return int(len(self.leaf_nodes))

    def path(self, node):
        buf = C.create_string_buffer(64)
        check(_lib.load().efgpu_mesh_path(self._m, int(node), buf, 64))
        return buf.value.decode()

    def post_order(self):
        """Order of Quadtree::merge callbacks (src/Quadtree.hpp:322-337)."""
        out = []
        stack = [(0, False)]
        while stack:
            i, done = stack.pop()
            if done or self.child[i, 0] < 0:
                out.append(i)
            else:
                stack.append((i, True))
                for c in range(3, -1, -1):
                    stack.append((int(self.child[i, c]), False))
        return out

    def leaf_cell_centres(self):
        """x, y of every leaf cell, shape (n_leaves, nx, ny), index [leaf, i, j] (HPSAlgorithm.hpp:241-249)."""
        b = self.box[self.leaf_nodes]
        M = self.nx
        dx = (b[:, 1] - b[:, 0]) / M
        dy = (b[:, 3] - b[:, 2]) / M
        k = np.arange(M)
        xs = (b[:, 0] + dx / 2)[:, None] + k[None, :] * dx[:, None]
        ys = (b[:, 2] + dy / 2)[:, None] + k[None, :] * dy[:, None]
        X = np.broadcast_to(xs[:, :, None], (len(b), M, M))
        Y = np.broadcast_to(ys[:, None, :], (len(b), M, M))
        return X, Y


def sample_leaf_coefficients(solver, leaf_boxes, nx, threads=1):
    """alpha, beta_w, beta_e, beta_s, beta_n, lambda of every leaf in `leaf_boxes` (rows x_lower, x_upper, y_lower, y_upper), each
    (n_leaves, nx, nx): alpha and lambda at the cell centres, beta at the four face midpoints (FiniteVolumeSolver.cpp:63-79).
    The callbacks are vectorised numpy functions; blocks of leaves are evaluated on `threads` host threads (numpy releases the
    GIL inside its kernels) - 1 for callbacks that are not thread-safe.  The reference calls them point by point, serially,
    inside every leaf solve."""
    b = np.asarray(leaf_boxes, dtype=np.float64).reshape(-1, 4)
    nl, M = len(b), int(nx)
    dx1, dy1 = (b[:, 1] - b[:, 0]) / M, (b[:, 3] - b[:, 2]) / M
    k = np.arange(M)
    xs = (b[:, 0] + dx1 / 2)[:, None] + k[None, :] * dx1[:, None]
    ys = (b[:, 2] + dy1 / 2)[:, None] + k[None, :] * dy1[:, None]
    X = np.broadcast_to(xs[:, :, None], (nl, M, M))
    Y = np.broadcast_to(ys[:, None, :], (nl, M, M))
    dx, dy = dx1[:, None, None], dy1[:, None, None]
    out = [np.empty((nl, M, M), dtype=np.float64) for _ in range(6)]

    def block(lo, hi):
        x, y, hx, hy = X[lo:hi], Y[lo:hi], dx[lo:hi] / 2.0, dy[lo:hi] / 2.0
        vals = (solver.alpha_function(x, y), solver.beta_function(x - hx, y), solver.beta_function(x + hx, y),
                solver.beta_function(x, y - hy), solver.beta_function(x, y + hy), solver.lambda_function(x, y))
        for o, v in zip(out, vals):
            o[lo:hi] = v          # broadcasts scalar / lower-dimensional results

    nthreads = max(1, min(int(threads), nl))
    if nthreads == 1:
        block(0, nl)
    else:
        from concurrent.futures import ThreadPoolExecutor
        step = max(1, -(-nl // (4 * nthreads)))
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(lambda lo: block(lo, min(nl, lo + step)), range(0, nl, step)))
    return out


class HPSAlgorithm:
    """src/HPSAlgorithm.hpp:26-596 for <FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double>."""

    def __init__(self, mesh: Mesh, patch_solver: FiniteVolumeSolver, device: int = 0, options: Optional[dict] = None):
        self.mesh = mesh
        self.patch_solver = patch_solver
        # app.options keys read by the hot path (HPSAlgorithm.hpp:134,532,587,1202)
        self.options = {"cache-operators": False, "homogeneous-rhs": False}
        if options:
            self.options.update(options)
        self.keep_x = False
        self.lean_T = False   # EFGPU_LEAN_T: DtN maps of interior nodes are transient (memory policy)
        self.no_symmetry = False   # EFGPU_NO_SYMMETRY: general merge plan even where X and diag(d) T are symmetric
        # EFGPU_LAZY_ROOT_DTN: the DtN map of the whole domain is formed by its first reader (Robin solve, operator(0, "T")) instead
        # of by buildStage - nothing on the Dirichlet path reads it
        self.lazy_root_dtn = False
        # pivoting policy (efgpu_set_refine_inverse): None = automatic - indefinite operators (lambda > 0) get one Newton-Schulz
        # step on every X^-1 after the unpivoted block inversion; True / False force it on / off
        self.refine_inverse = None
        # FivePointStencil leaves: the reference evaluates alpha/beta/lambda inside every leaf solve; here they are sampled on
        # the host once per buildStage.  False keeps the coefficient arrays already resident in HBM (same functions).
        self.resample_coefficients = True
        self._coefficients_set = False
        self.sampling_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.efgpu_create(C.byref(mesh.desc), device, C.byref(self._h)))
        self.isBuilt = False

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.efgpu_destroy(self._h)
            self._h = None

    # -- stages ----------------------------------------------------------------------------------
    def setupStage(self):  # HPSAlgorithm.hpp:91-106: nothing is done in the reference either
        return None

    def _flags(self):
        return ((CACHE_OPERATORS if self.options["cache-operators"] else 0)
                | (HOMOGENEOUS_RHS if self.options["homogeneous-rhs"] else 0) | (KEEP_X if self.keep_x else 0)
                | (LEAN_T if self.lean_T else 0) | (NO_SYMMETRY if self.no_symmetry else 0)
                | (LAZY_ROOT_DTN if self.lazy_root_dtn else 0))

    def is_symmetric(self):
        """True when the root's DtN map was built by the symmetric merge plan (efgpu_is_symmetric)."""
        return bool(self._lib.efgpu_is_symmetric(self._h))

    def buildStage(self):  # HPSAlgorithm.hpp:120-161
        s = self.patch_solver
        if s.solver_type == "FISHPACK90":
            lam = float(s.lambda_function(np.float64(0.0), np.float64(0.0)))  # FiniteVolumeSolver.cpp:254
            check(self._lib.efgpu_set_leaf_constant(self._h, lam), self._h)
        elif s.solver_type == "FivePointStencil":
            if self.resample_coefficients or not self._coefficients_set:
                self._set_variable_coefficients()
                self._coefficients_set = True
        else:
            raise ValueError("unknown solver_type")
        check(self._lib.efgpu_set_refine_inverse(self._h, -1 if self.refine_inverse is None else int(bool(self.refine_inverse))), self._h)
        check(self._lib.efgpu_build(self._h, self._flags()), self._h)
        self.isBuilt = True

    def rebuildStage(self, old: "HPSAlgorithm"):
        """buildStage of a CHANGED mesh that re-uses `old`'s operators for every unchanged subtree (efgpu_rebuild_from): what
        paper.md:44 describes and the reference leaves as a TODO (isBuilt, HPSAlgorithm.hpp:50-55).  Returns (merges copied,
        merges computed); the result is bit-identical to buildStage()."""
        s = self.patch_solver
        if s.solver_type == "FISHPACK90":
            check(self._lib.efgpu_set_leaf_constant(self._h, float(s.lambda_function(np.float64(0.0), np.float64(0.0)))), self._h)
        elif self.resample_coefficients or not self._coefficients_set:
            self._set_variable_coefficients()
            self._coefficients_set = True
        check(self._lib.efgpu_set_refine_inverse(self._h, -1 if self.refine_inverse is None else int(bool(self.refine_inverse))), self._h)
        reused, rebuilt = C.c_double(), C.c_double()
        check(self._lib.efgpu_rebuild_from(self._h, old._h, self._flags(), C.byref(reused), C.byref(rebuilt)), self._h)
        self.isBuilt = True
        return int(reused.value), int(rebuilt.value)

    def _set_variable_coefficients(self):
        """Sample alpha, beta, lambda where FiniteVolumeSolver.cpp:63-79 samples them (sample_leaf_coefficients) and hand the
        six leaf-major arrays to the library."""
        m = self.mesh
        out = sample_leaf_coefficients(self.patch_solver, m.box[m.leaf_nodes], m.nx, self.sampling_threads)
        check(self._lib.efgpu_set_leaf_variable(self._h, *[a.ctypes.data for a in out]), self._h)

    def upwardsStage(self, rhs, scale: float = 1.0):
        """HPSAlgorithm.hpp:178-272.  `rhs` is f(x, y) (sampled at leaf cell centres) or a ready
        (n_leaves, nx, ny) array holding vectorF of every leaf."""
        if callable(rhs):
            X, Y = self.mesh.leaf_cell_centres()
            rhs = rhs(X, Y)
        f = np.ascontiguousarray(rhs, dtype=np.float64).reshape(-1)
        if f.size != self.mesh.n_leaves * self.mesh.nx ** 2:
            raise ValueError("load vector has the wrong size")
        self._f_host = f
        check(self._lib.efgpu_upwards(self._h, f.ctypes.data, float(scale), self._flags()), self._h)

    def root_boundary_points(self):
        """Sampling points of solveStage (HPSAlgorithm.hpp:375-400): sides W, E, S, N of the merged root grid."""
        size = self.node_info(0)["size"]
        xl, xu, yl, yu = self.mesh.box[0]
        g = FiniteVolumeGrid(size, xl, xu, size, yl, yu)
        k = np.arange(size)
        xs, ys = g.point(0, k), g.point(1, k)
        x = np.concatenate([np.full(size, xl), np.full(size, xu), xs, xs])
        y = np.concatenate([ys, ys, np.full(size, yl), np.full(size, yu)])
        side = np.repeat(np.arange(4), size)
        return side, x, y

    def solveStage(self, boundary):
        """HPSAlgorithm.hpp:291-445.  `boundary` is either a ready root vectorG (array, WESN) or a
        function (side, x, y) -> (r, a, b) of the general condition a u + b du/dn = r."""
        u = np.empty(self.mesh.n_leaves * self.mesh.nx ** 2)
        if callable(boundary):
            side, x, y = self.root_boundary_points()
            r, a, b = boundary(side, x, y)
            r, a, b = (np.ascontiguousarray(np.broadcast_to(v, x.shape), dtype=np.float64) for v in (r, a, b))
            check(self._lib.efgpu_solve_robin(self._h, a.ctypes.data, b.ctypes.data, r.ctypes.data, self._flags(), u.ctypes.data), self._h)
        else:
            g = np.ascontiguousarray(boundary, dtype=np.float64).reshape(-1)
            if g.size != 4 * self.node_info(0)["size"]:
                raise ValueError("root Dirichlet vector has the wrong size")
            check(self._lib.efgpu_solve_dirichlet(self._h, g.ctypes.data, self._flags(), u.ctypes.data), self._h)
        self.u_leaves = u.reshape(self.mesh.n_leaves, self.mesh.nx, self.mesh.nx)
        return self.u_leaves

    # -- parity accessors ------------------------------------------------------------------------
    def node_info(self, node):
        v = [C.c_int() for _ in range(4)]
        check(self._lib.efgpu_node_info(self._h, int(node), *[C.byref(x) for x in v]), self._h)
        return dict(size=v[0].value, n_coarsens=v[1].value, leaf=bool(v[2].value), leaf_index=v[3].value)

    def operator(self, node, which):
        r, c = C.c_int(), C.c_int()
        check(self._lib.efgpu_operator_shape(self._h, int(node), OP[which], C.byref(r), C.byref(c)), self._h)
        out = np.empty((r.value, c.value))
        check(self._lib.efgpu_get_operator(self._h, int(node), OP[which], out.ctypes.data, out.size), self._h)
        return out

    def vector(self, node, which):
        n = C.c_int()
        check(self._lib.efgpu_vector_length(self._h, int(node), VEC[which], C.byref(n)), self._h)
        out = np.empty(n.value)
        check(self._lib.efgpu_get_vector(self._h, int(node), VEC[which], out.ctypes.data, out.size), self._h)
        return out

    # -- device-resident variants (inputs already in HBM; used by the benchmark's `value` leg) ---
    def upwardsStageDevice(self, f_dev_ptr: int, scale: float = 1.0, sync: bool = True):
        check(self._lib.efgpu_upwards_device(self._h, C.c_void_p(f_dev_ptr), float(scale), self._flags(), int(sync)), self._h)

    def solveStageDevice(self, g_dev_ptr: int, u_dev_ptr: int = 0, sync: bool = True):
        check(self._lib.efgpu_solve_dirichlet_device(self._h, C.c_void_p(g_dev_ptr), self._flags(),
                                                     C.c_void_p(u_dev_ptr) if u_dev_ptr else None, int(sync)), self._h)

    # -- callers either side of the path: sampling coordinates and error norms on the device ----
    POINTS = {"centre": 0, "W": 1, "E": 2, "S": 3, "N": 4}

    def leafPoints(self, which="centre"):
        """(x, y) of the sampling points of every leaf, shape (n_leaves, nx, ny), computed on the device: the cell centres
        (load, alpha, lambda) or the W / E / S / N face midpoints (beta) - bit-identical to Mesh.leaf_cell_centres()."""
        m = self.mesh
        x, y = (np.empty((m.n_leaves, m.nx, m.nx)) for _ in range(2))
        check(self._lib.efgpu_leaf_points(self._h, self.POINTS[which], x.ctypes.data, y.ctypes.data), self._h)
        return x, y

    def leafPointsDevice(self, which, x_dev_ptr: int, y_dev_ptr: int, sync: bool = True):
        check(self._lib.efgpu_leaf_points_device(self._h, self.POINTS[which], C.c_void_p(x_dev_ptr) if x_dev_ptr else None,
                                                 C.c_void_p(y_dev_ptr) if y_dev_ptr else None, int(sync)), self._h)

    def setVariableCoefficientsDevice(self, alpha, beta_w, beta_e, beta_s, beta_n, lam):
        """FivePointStencil leaves from six device arrays (pointers) the caller evaluated on leafPointsDevice coordinates;
        the following buildStage keeps them (no host sampling)."""
        check(self._lib.efgpu_set_leaf_variable_device(self._h, *[C.c_void_p(int(p)) for p in (alpha, beta_w, beta_e, beta_s, beta_n, lam)]), self._h)
        self._coefficients_set = True
        self.resample_coefficients = False

    def errorNorms(self, exact):
        """(l1, l2, linf) of the last solveStage against `exact` (a function of (x, y) sampled at the cell centres, or a
        ready (n_leaves, nx, ny) array), as the reference's drivers compute them (examples/elliptic-multiple/main.cpp:346-371)
        but reduced on the device."""
        if callable(exact):
            X, Y = self.mesh.leaf_cell_centres()
            exact = exact(X, Y)
        e = np.ascontiguousarray(np.broadcast_to(exact, (self.mesh.n_leaves, self.mesh.nx, self.mesh.nx)), dtype=np.float64)
        out = [C.c_double() for _ in range(3)]
        check(self._lib.efgpu_error_norms(self._h, e.ctypes.data, *[C.byref(v) for v in out]), self._h)
        return tuple(v.value for v in out)

    def errorNormsDevice(self, exact_dev_ptr: int, u_dev_ptr: int = 0):
        out = [C.c_double() for _ in range(3)]
        check(self._lib.efgpu_error_norms_device(self._h, C.c_void_p(u_dev_ptr) if u_dev_ptr else None, C.c_void_p(exact_dev_ptr),
                                                 *[C.byref(v) for v in out]), self._h)
        return tuple(v.value for v in out)

    def toVTK(self, filename, fields=None):
        """Mesh + cell fields as one binary .vtu written from device buffers (src/Mesh.hpp:186-267 + src/VTK.cpp:237-300 write
        ASCII from the host).  fields: {name: device pointer or None}; None = the solution of the last solveStage."""
        fields = {"u": None} if fields is None else fields
        names = (C.c_char_p * len(fields))(*[k.encode() for k in fields])
        ptrs = (C.c_void_p * len(fields))(*[C.c_void_p(int(v)) if v else None for v in fields.values()])
        check(self._lib.efgpu_write_vtu(self._h, str(filename).encode(), len(fields), names, ptrs), self._h)

    def sync(self):
        check(self._lib.efgpu_sync(self._h), self._h)

    def stream(self) -> int:
        return int(self._lib.efgpu_stream(self._h) or 0)

    def set_profiling(self, on: bool = True):
        check(self._lib.efgpu_set_profiling(self._h, int(on)), self._h)

    def profile(self):
        """{class name: (device ms, launches)} accumulated since set_profiling()."""
        out = {}
        for cls in range(64):
            name = self._lib.efgpu_profile_class_name(cls)
            if not name:
                break
            ms, n = C.c_double(), C.c_double()
            check(self._lib.efgpu_get_profile(self._h, cls, C.byref(ms), C.byref(n)), self._h)
            out[name.decode()] = (ms.value, n.value)
        return out

    def stats(self):
        s = Stats()
        check(self._lib.efgpu_get_stats(self._h, C.byref(s)), self._h)
        return {k: getattr(s, k) for k, _ in Stats._fields_}
