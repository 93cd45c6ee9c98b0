"""Benchmark/driver-facing wrappers of the HPS path: one GPU, or the quadtree sharded by level-2
subtree over the ranks of one NVSwitch domain (reference: Morton-curve partition of
p4est, src/Mesh.hpp:169-170, rank-shared upper tree src/Quadtree.hpp:146-151,464-507).

torch / torch.distributed are plumbing only (device buffers, NCCL); every kernel of the path lives
in libefgpu.so.
"""
from __future__ import annotations

import numpy as np

from .hps import HPSAlgorithm


class SingleGpuHPS(HPSAlgorithm):
    """HPSAlgorithm plus the host/device step helpers bench.py uses."""

    def sharding(self):
        return "single GPU (whole tree on cuda:%d)" % self._device

    def __init__(self, mesh, solver, device=0, options=None):
        super().__init__(mesh, solver, device=device, options=options)
        self._device = device

    def sample_inputs(self, f_fn, u_fn):
        """Host sampling of the load at leaf cell centres (HPSAlgorithm.hpp:241-249) and of the Dirichlet
        data at the root boundary (HPSAlgorithm.hpp:375-400, a = 1, b = 0)."""
        X, Y = self.mesh.leaf_cell_centres()
        f = np.ascontiguousarray(f_fn(X, Y), dtype=np.float64).reshape(-1)
        _side, x, y = self.root_boundary_points()
        g = np.ascontiguousarray(u_fn(x, y), dtype=np.float64).reshape(-1)
        return f, g

    def upwardsStageHost(self, f_host):
        self.upwardsStage(f_host.reshape(self.mesh.n_leaves, self.mesh.nx, self.mesh.nx))

    def solveStageHost(self, g_host, u_host):
        import ctypes as C
        from ._lib import check
        check(self._lib.efgpu_solve_dirichlet(self._h, g_host.ctypes.data, self._flags(), u_host.ctypes.data), self._h)

    def total_issued_flops(self):
        return self.stats()["merge_flops_issued"]

    def max_error(self, u_dev, u_fn):
        X, Y = self.mesh.leaf_cell_centres()
        u = u_dev.cpu().numpy().reshape(X.shape)
        return float(np.max(np.abs(u - u_fn(X, Y))))


def make_hps(mesh, solver, device=0, rank=0, world=1, options=None, cut=2, grouped=False, balance="count"):
    """cut: tree level whose 4^cut subtrees are dealt to the ranks (2: the 16 subtrees of SURVEY 8(e); 1: four subtrees, for 2 or 4
    ranks - the level-1 merges then run whole on their owners and only the root merge is row-partitioned)."""
    if world == 1:
        return SingleGpuHPS(mesh, solver, device=device, options=options)
    if grouped:     # three tiers: forests, level-1 merges inside rank groups, root merge over all ranks (4, 8 or 16 ranks)
        from .sharded import GroupedShardedHPS
        return GroupedShardedHPS(mesh, solver, device=device, rank=rank, world=world, options=options)
    from .sharded import ShardedHPS
    return ShardedHPS(mesh, solver, device=device, rank=rank, world=world, options=options, cut=cut, balance=balance)
