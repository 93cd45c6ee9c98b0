#!/usr/bin/env python
"""Bandwidth of the SURVEY 8(f) rank-1 kernels (csrc/sample.cu): sampling coordinates of every leaf (16 B written per point)
and the device reduction of the drivers' error norms (16 B read per cell), CUDA events on the library's stream, next to
the host loops they replace (numpy on one thread: Mesh.leaf_cell_centres and the norm formulas)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ellipticforest_b200 as ef

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=8)
ap.add_argument("--nx", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()

PI = 3.141592653589793
mesh = ef.Mesh().refineByFunction(None, 0.0, a.level, a.level, ef.FiniteVolumeGrid(a.nx, 0.0, PI, a.nx, 0.0, PI))
solver = ef.FiniteVolumeSolver(); solver.solver_type = "FISHPACK90"
hps = ef.HPSAlgorithm(mesh, solver)
cells = mesh.n_leaves * a.nx * a.nx
x, y, u, v = (torch.empty(cells, dtype=torch.float64, device="cuda") for _ in range(4))
s = torch.cuda.ExternalStream(hps.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(a.reps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


t_pts = timed(lambda: hps.leafPointsDevice("centre", x.data_ptr(), y.data_ptr(), sync=False))
u.copy_(torch.sin(x) + torch.sin(y)); v.copy_(u + 1e-3 * torch.cos(x))
torch.cuda.synchronize()
t_err = timed(lambda: hps.errorNormsDevice(v.data_ptr(), u.data_ptr()))      # includes the 24-byte read-back and a stream sync
t0 = time.perf_counter(); X, Y = mesh.leaf_cell_centres(); X = np.ascontiguousarray(X); Y = np.ascontiguousarray(Y); t_host_pts = (time.perf_counter() - t0) * 1e3
assert np.array_equal(x.cpu().numpy().reshape(X.shape), X) and np.array_equal(y.cpu().numpy().reshape(Y.shape), Y)
uh, vh = u.cpu().numpy(), v.cpu().numpy()
t0 = time.perf_counter(); d = np.abs(uh - vh); ref = (d.sum() / cells, np.sqrt((d * d).sum() / cells), d.max()); t_host_err = (time.perf_counter() - t0) * 1e3
got = hps.errorNormsDevice(v.data_ptr(), u.data_ptr())
print("cells %d | points: %.3f ms = %.0f GB/s (host numpy %.1f ms) | error norms: %.3f ms = %.0f GB/s (host numpy %.1f ms) | l1 %.6e l2 %.6e linf %.6e (host %.6e %.6e %.6e)" % (
    cells, t_pts, 16.0 * cells / t_pts / 1e6, t_host_pts, t_err, 16.0 * cells / t_err / 1e6, t_host_err, *got, *ref))
