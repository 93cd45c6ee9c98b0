#!/usr/bin/env python
"""A/B measurement of the bandwidth-bound stages (upwards / solve) under the efgpu_set_tuning knobs.
Builds the tree once, then times `reps` upwards + solve passes per configuration (CUDA events per kernel class inside
the library) and checks that every configuration reproduces the first one's solution."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ellipticforest_b200 as ef
from ellipticforest_b200 import dist as efdist
from ellipticforest_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=8)
ap.add_argument("--nx", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--configs", default="2;0;2,0,0,1", help="semicolon-separated efgpu_set_tuning values for keys 0, 1, 2, ...")
ap.add_argument("--ncu", action="store_true", help="one pass per configuration, no warm-up (for an ncu launch list)")
a = ap.parse_args()

PI = 3.141592653589793
lib = _lib.load()
grid = ef.FiniteVolumeGrid(a.nx, 0.0, PI, a.nx, 0.0, PI)
mesh = ef.Mesh().refineByFunction(None, 0.0, a.level, a.level, grid)
solver = ef.FiniteVolumeSolver(); solver.solver_type = "FISHPACK90"
hps = efdist.make_hps(mesh, solver)
u_exact = lambda x, y: np.sin(x) + np.sin(y)
f, g = hps.sample_inputs(lambda x, y: -u_exact(x, y), u_exact)
f_dev, g_dev = torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()
u_dev = torch.empty_like(f_dev)
hps.buildStage()
st = hps.stats()
ref = None
for cfg in a.configs.split(";"):
    k = [int(v) for v in cfg.split(",")]
    for key, val in enumerate(k):
        assert lib.efgpu_set_tuning(key, val) == 0
    for _ in range(0 if a.ncu else 3):
        hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=False); hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True)
    hps.set_profiling(True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    s = torch.cuda.ExternalStream(hps.stream())
    tu = ts = 0.0
    for _ in range(a.reps):
        e[0].record(s); hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=False)
        e[1].record(s); hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True)
        e[2].record(s); torch.cuda.synchronize()
        tu += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2])
    p = hps.profile(); hps.set_profiling(False)
    u = u_dev.cpu().numpy()
    if ref is None:
        ref = u
    tu /= a.reps; ts /= a.reps
    print("cfg %-8s upwards %.3f ms %.0f GB/s | solve %.3f ms %.0f GB/s | kernels: up_mv %.3f so_mv %.3f leaf %.3f | max|du| %.2e err %.2e" % (
        cfg, tu, st["upwards_bytes"] / tu / 1e6, ts, st["solve_bytes"] / ts / 1e6,
        p["upwards_matvec"][0] / a.reps, p["solve_matvec"][0] / a.reps, p["leaf_solve"][0] / a.reps,
        float(np.max(np.abs(u - ref))), float(np.max(np.abs(u.reshape(-1) - u_exact(*mesh.leaf_cell_centres()).reshape(-1))))), flush=True)
