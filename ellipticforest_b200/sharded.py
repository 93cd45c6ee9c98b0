"""Quadtree sharded by subtree over the GPUs of one NVSwitch domain.

Reference mechanism being replaced: p4est_partition gives every MPI rank a contiguous Morton range
of leaves (src/Mesh.hpp:169-170); nodes above are replicated on every rank that shares them and
their children arrive by MPI::broadcast of the WHOLE child patch (src/Quadtree.hpp:464-507,
src/QuadNode.hpp:191-199, FiniteVolumePatch.cpp:118-134), after which every sharing rank repeats
the same merge.  Here ownership is static: the tree is cut at `cut` (level 2: 16 subtrees, Morton
order); rank r owns subtrees [16 r / P, 16 (r+1) / P) - exactly p4est's partition of a uniform
tree - builds them as one batched forest on its GPU, and only what a parent needs crosses NVLink:
the subtree roots' DtN maps (build), particular Neumann data h (upwards) and Dirichlet data g
(solve).  The upper tree (levels < cut) is merged once, on rank 0, never redundantly.

ShardPlan is pure host logic (numpy) and is what the world_size-2 gloo tests exercise; the engines
(forest / upper-tree handles of libefgpu.so on the GPU; the numpy oracle in the CPU tests) plug in
underneath through the same small interface.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import TreeDesc, check
from .hps import CACHE_OPERATORS, HOMOGENEOUS_RHS, LAZY_ROOT_DTN, NO_SYMMETRY, OP, VEC, FiniteVolumeGrid, sample_leaf_coefficients


class ShardPlan:
    """Static ownership of the subtrees below `cut` and the node tables of every piece."""

    def __init__(self, level, child, box, nx, world, cut=2, balance="count"):
        """balance: how the 4^cut subtrees (Morton order) are dealt to the ranks, always as contiguous blocks of whole subtrees:
        "count" - equally many per rank (p4est_partition on a uniform tree; needs world | 4^cut);
        "leaves" - blocks with the smallest possible maximum number of leaves (p4est_partition weights leaves uniformly,
                   src/Mesh.hpp:169-170; here whole subtrees stay together);
        "work"  - the same with the merge work below the cut as the weight (sum of n^3 over the subtree's merges, n = child
                   side, plus nx^3 per leaf for its DtN map): what balances the build stage of an adaptive tree."""
        self.level = np.asarray(level, dtype=np.int32)
        self.child = np.asarray(child, dtype=np.int32).reshape(-1, 4)
        self.box = np.asarray(box, dtype=np.float64).reshape(-1, 4)
        self.nx, self.world, self.cut = int(nx), int(world), int(cut)
        nn = len(self.level)
        leaf = self.child[:, 0] < 0
        if np.any(leaf & (self.level < cut)):
            raise ValueError("tree too shallow to shard: a leaf lies above the cut level %d" % cut)
        self.cut_nodes = np.nonzero(self.level == cut)[0]            # pre-order = Morton order
        ncut = len(self.cut_nodes)
        if balance not in ("count", "leaves", "work"):
            raise ValueError("balance must be 'count', 'leaves' or 'work'")
        if ncut != 4 ** cut or world < 1 or world > ncut or (balance == "count" and ncut % world):
            raise ValueError("%d subtrees cannot be dealt to %d ranks" % (ncut, world))
        self.balance = balance
        # subtree extents: the table is in pre-order, so a subtree is a contiguous id range
        self.sub_end = np.empty(ncut, dtype=np.int64)
        for k, r in enumerate(self.cut_nodes):
            e = r + 1
            while e < nn and self.level[e] > cut:
                e += 1
            self.sub_end[k] = e
        # node sizes bottom-up: leaf = nx, parent = 2 * min(child sizes) (mergePatch_, HPSAlgorithm.hpp:1004-1009)
        self.size = np.zeros(nn, dtype=np.int64)
        for i in range(nn - 1, -1, -1):
            self.size[i] = self.nx if leaf[i] else 2 * min(self.size[c] for c in self.child[i])
        self.leaf = leaf
        self.leaf_index = np.cumsum(leaf) - 1                        # global leaf numbering (pre-order)
        # weights of the subtrees and their owners
        self.weight = np.ones(ncut)
        if balance != "count":
            for k, r in enumerate(self.cut_nodes):
                ids = np.arange(r, self.sub_end[k])
                if balance == "leaves":
                    self.weight[k] = float(np.sum(leaf[ids]))
                else:
                    half = self.size[ids] / 2.0
                    self.weight[k] = float(np.sum(np.where(leaf[ids], float(self.nx) ** 3, half ** 3)))
        # "count": contiguous Morton blocks of equal length, as p4est_partition on a uniform tree
        self.owner = self._contiguous_blocks(self.weight, world) if balance != "count" else (np.arange(ncut) * world) // ncut

    @staticmethod
    def _contiguous_blocks(w, parts):
        """Owner of each item when the sequence is cut into `parts` non-empty contiguous blocks with the smallest possible
        maximum block weight (dynamic programme over the prefix sums; ties resolved towards earlier cuts: deterministic)."""
        n = len(w)
        pre = np.concatenate([[0.0], np.cumsum(w)])
        best = np.full((parts + 1, n + 1), np.inf)
        arg = np.zeros((parts + 1, n + 1), dtype=np.int64)
        best[0, 0] = 0.0
        for p in range(1, parts + 1):
            for i in range(p, n - (parts - p) + 1):
                for j in range(p - 1, i):
                    v = max(best[p - 1, j], pre[i] - pre[j])
                    if v < best[p, i]:
                        best[p, i], arg[p, i] = v, j
        owner = np.zeros(n, dtype=np.int64)
        i = n
        for p in range(parts, 0, -1):
            j = arg[p, i]
            owner[j:i] = p - 1
            i = j
        return owner

    # -- pieces ------------------------------------------------------------------------------
    def subtrees_of(self, rank) -> List[int]:
        return [k for k in range(len(self.cut_nodes)) if self.owner[k] == rank]

    def local_table(self, rank):
        """Forest of the subtrees rank owns: (global ids, level, child (local ids), box, local id of each subtree root)."""
        ids = np.concatenate([np.arange(self.cut_nodes[k], self.sub_end[k]) for k in self.subtrees_of(rank)])
        remap = -np.ones(len(self.level), dtype=np.int64)
        remap[ids] = np.arange(len(ids))
        child = self.child[ids].copy()
        m = child >= 0
        child[m] = remap[child[m]]
        roots = remap[[self.cut_nodes[k] for k in self.subtrees_of(rank)]]
        return ids, self.level[ids].copy(), child.astype(np.int32), self.box[ids].copy(), roots.astype(np.int64)

    def local_leaf_range(self, rank):
        """Global leaf indices [lo, hi) owned by rank (contiguous: Morton order)."""
        ks = self.subtrees_of(rank)
        lo_node, hi_node = self.cut_nodes[ks[0]], self.sub_end[ks[-1]]
        lo = int(np.sum(self.leaf[:lo_node]))
        return lo, lo + int(np.sum(self.leaf[lo_node:hi_node]))

    def top_table(self):
        """Upper tree (levels <= cut) whose leaves are the subtree roots: (global ids, level, child (local), box, ext sizes)."""
        ids = np.nonzero(self.level <= self.cut)[0]
        remap = -np.ones(len(self.level), dtype=np.int64)
        remap[ids] = np.arange(len(ids))
        child = self.child[ids].copy()
        child[self.level[ids] == self.cut] = -1
        m = child >= 0
        child[m] = remap[child[m]]
        ext = self.size[self.cut_nodes].astype(np.int32)             # one per top-tree leaf, pre-order
        return ids, self.level[ids].copy(), child.astype(np.int32), self.box[ids].copy(), ext

    def root_size(self):
        return int(self.size[0])


# ---------------------------------------------------------------------------------------------
def _dev_tensor(ptr: int, n: int, device: int):
    """torch view of `n` doubles of device memory owned by libefgpu.so on GPU `device` (no copy).  The device is explicit:
    the memory belongs to the handle's GPU whatever torch's current device is."""
    import torch

    class _View:
        pass

    v = _View()
    v.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(v, device=torch.device("cuda", int(device)))


def global_symmetry(dist, local_flag: bool, world: int, device=None, group=None) -> bool:
    """The symmetric merge plan of a shared upper tree is a GLOBAL decision: efgpu_is_symmetric() only looks at the subtree
    roots one rank built, and a rank whose subtrees are uniform would otherwise run the mirrored plan on maps another rank
    built with coarsening (not signed-symmetric) - wrong operators, or mismatched collectives.  MIN over every rank that
    contributes leaves to the handle (the world, or `group`)."""
    if world <= 1:
        return bool(local_flag)
    import torch
    dev = "cpu" if device is None else torch.device("cuda", int(device))
    t = torch.tensor([1 if local_flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t[0]))


class GpuEngine:
    """One libefgpu handle: a forest of owned subtrees (leaves = FV patches) or the upper tree (external leaves)."""

    def __init__(self, level, child, box, nx, device, ext_sizes=None, stream=None):
        self._lib = _lib.load()
        self.level = np.ascontiguousarray(level, dtype=np.int32)
        self.child = np.ascontiguousarray(child, dtype=np.int32)
        self.box = np.ascontiguousarray(box, dtype=np.float64)
        d = TreeDesc()
        d.n_nodes, d.nx = len(self.level), int(nx)
        d.level = self.level.ctypes.data_as(C.POINTER(C.c_int32))
        d.child = self.child.ctypes.data_as(C.POINTER(C.c_int32))
        d.box = self.box.ctypes.data_as(C.POINTER(C.c_double))
        self._h = C.c_void_p()
        self.device = int(device)
        self.peer = False
        ext = None
        if ext_sizes is not None:
            self._ext = np.ascontiguousarray(ext_sizes, dtype=np.int32)
            ext = self._ext.ctypes.data_as(C.POINTER(C.c_int32))
        check(self._lib.efgpu_create_ex(C.byref(d), device, ext, C.byref(self._h)))
        if stream:
            check(self._lib.efgpu_set_stream(self._h, C.c_void_p(stream)), self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.efgpu_destroy(self._h)
            self._h = None

    def stream(self):
        return int(self._lib.efgpu_stream(self._h) or 0)

    def set_leaf_constant(self, lam):
        check(self._lib.efgpu_set_leaf_constant(self._h, float(lam)), self._h)

    def set_leaf_variable(self, arrays):
        """FivePointStencil leaves: alpha, beta_w, beta_e, beta_s, beta_n, lambda of this handle's leaves (host arrays, leaf-major)."""
        arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
        check(self._lib.efgpu_set_leaf_variable(self._h, *[a.ctypes.data for a in arrays]), self._h)

    def operator_view(self, node, which):
        p, r, c = C.c_void_p(), C.c_int(), C.c_int()
        check(self._lib.efgpu_operator_device(self._h, int(node), OP[which], C.byref(p), C.byref(r), C.byref(c)), self._h)
        return _dev_tensor(p.value, r.value * c.value, self.device)

    def vector_view(self, node, which):
        p, n = C.c_void_p(), C.c_int()
        code = {"h0": 5}.get(which, VEC.get(which))
        check(self._lib.efgpu_vector_device(self._h, int(node), code, C.byref(p), C.byref(n)), self._h)
        return _dev_tensor(p.value, n.value, self.device)

    def build(self, flags):
        check(self._lib.efgpu_build(self._h, flags), self._h)

    def is_symmetric(self):
        return bool(self._lib.efgpu_is_symmetric(self._h))

    def set_symmetric_leaves(self, on):
        check(self._lib.efgpu_set_symmetric_leaves(self._h, int(bool(on))), self._h)

    def set_refine_inverse(self, mode):
        check(self._lib.efgpu_set_refine_inverse(self._h, int(mode)), self._h)

    def set_partition(self, rank, nranks, allgather=None):
        """Row partition of a replicated tree; `allgather(tensor)` must all-gather the tensor's equal slices in place."""
        check(self._lib.efgpu_set_partition(self._h, int(rank), int(nranks)), self._h)
        if allgather is not None:
            def _cb(buf, nbytes, _user):
                try:
                    allgather(_dev_tensor(buf, nbytes * nranks // 8, self.device))
                    return 0
                except Exception as e:  # never unwind through the C frames
                    import traceback
                    traceback.print_exc()
                    return 1
            self._ag_cb = _lib.ALLGATHER_FN(_cb)      # keep the trampoline alive as long as the handle
            check(self._lib.efgpu_set_allgather(self._h, self._ag_cb, None), self._h)

    def peer_setup(self, dist, world, group=None):
        """Peer mode (include/efgpu.h: efgpu_peer_export / efgpu_peer_attach): map every rank's shared operator arena into this
        process.  The 64-byte CUDA IPC handles travel through torch.distributed (the one host-side collective of the set-up);
        afterwards the library exchanges row slices by itself - GEMM epilogue stores into the peers' arenas and flag barriers -
        and no collective callback is installed."""
        import torch
        buf = (C.c_ubyte * 64)()
        check(self._lib.efgpu_peer_export(self._h, buf), self._h)
        dev = torch.device("cuda", self.device)
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        raw = bytes(bytearray(allh.cpu().numpy().tobytes()))
        check(self._lib.efgpu_peer_attach(self._h, raw, int(world)), self._h)
        self.peer = True

    def peer_barrier(self):
        check(self._lib.efgpu_peer_barrier(self._h), self._h)

    def peer_broadcast(self, t):
        """t: a device view of this handle (inside the shared arena): its bytes go to the same place on every other rank."""
        check(self._lib.efgpu_peer_broadcast(self._h, C.c_void_p(t.data_ptr()), t.numel() * t.element_size()), self._h)

    def complete_root_T(self):
        """Collective: gather the row slices of the root's DtN map and mirror the blocks of the symmetric plan."""
        check(self._lib.efgpu_complete_root_dtn(self._h), self._h)

    def build_begin(self, flags):
        check(self._lib.efgpu_build_begin(self._h, flags), self._h)

    def build_level(self, level, phase):
        check(self._lib.efgpu_build_level(self._h, int(level), int(phase)), self._h)

    def build_end(self):
        check(self._lib.efgpu_build_end(self._h), self._h)

    def upwards(self, f_dev_ptr, scale, flags, sync=False):
        check(self._lib.efgpu_upwards_device(self._h, C.c_void_p(f_dev_ptr) if f_dev_ptr else None, float(scale), flags, int(sync)), self._h)

    def solve_from_roots(self, u_dev_ptr, flags, sync=False):
        check(self._lib.efgpu_solve_from_roots_device(self._h, flags, C.c_void_p(u_dev_ptr) if u_dev_ptr else None, int(sync)), self._h)

    def sync(self):
        check(self._lib.efgpu_sync(self._h), self._h)

    def set_profiling(self, on):
        check(self._lib.efgpu_set_profiling(self._h, int(on)), self._h)

    def profile(self):
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
        s = _lib.Stats()
        check(self._lib.efgpu_get_stats(self._h, C.byref(s)), self._h)
        return {k: getattr(s, k) for k, _ in _lib.Stats._fields_}


class ShardedExchange:
    """The three exchanges of a sharded run, written against torch.distributed point-to-point calls so
    that the same code runs over NCCL (GPU) and gloo (CPU tests).  `local` / `top` expose
    root_T(k) / root_h(k) / root_g(k) for the k-th owned subtree and leaf_T(j) / leaf_h(j) / leaf_g(j)
    for the j-th leaf of the upper tree (tensors that alias the engines' storage)."""

    def __init__(self, plan: ShardPlan, rank: int, dist, root_rank=0):
        self.plan, self.rank, self.dist, self.root = plan, rank, dist, root_rank

    def _gather(self, send_of, recv_of):
        """subtree k: owner -> root (root's own subtrees are copied)."""
        ops, copies = [], []
        for k in range(len(self.plan.cut_nodes)):
            o = int(self.plan.owner[k])
            if self.rank == self.root:
                if o == self.root:
                    copies.append((recv_of(k), send_of(k)))
                else:
                    ops.append(self.dist.P2POp(self.dist.irecv, recv_of(k), o))
            elif o == self.rank:
                ops.append(self.dist.P2POp(self.dist.isend, send_of(k), self.root))
        for dst, src in copies:
            dst.copy_(src)
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def _scatter(self, send_of, recv_of):
        """subtree k: root -> owner."""
        ops, copies = [], []
        for k in range(len(self.plan.cut_nodes)):
            o = int(self.plan.owner[k])
            if self.rank == self.root:
                if o == self.root:
                    copies.append((recv_of(k), send_of(k)))
                else:
                    ops.append(self.dist.P2POp(self.dist.isend, send_of(k), o))
            elif o == self.rank:
                ops.append(self.dist.P2POp(self.dist.irecv, recv_of(k), self.root))
        for dst, src in copies:
            dst.copy_(src)
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def share(self, local_view, top_view, flat=None):
        """Replicated upper tree: every rank ends up with every subtree root's operator / vector in its own
        upper-tree leaf buffers (owner copies, then broadcasts).

        `flat(first, total)`: optional - returns ONE tensor of `total` elements aliasing the destination buffers from
        `first` on.  When the K buffers lie back to back in subtree order with equal sizes (libefgpu keeps the leaf DtN maps
        of the upper tree that way) and every rank owns an equal Morton block, the K broadcasts collapse into one in-place
        all-gather.  Opt-in (EFGPU_SHARE_ALLGATHER=1) until it has been timed over NCCL."""
        world = self.dist.get_world_size() if self.dist.is_initialized() else 1
        K = len(self.plan.cut_nodes)
        dsts = [top_view(k) for k in range(K)]
        for k in range(K):
            if int(self.plan.owner[k]) == self.rank:
                dsts[k].copy_(local_view(k))
        if world == 1:
            return
        if os.environ.get("EFGPU_SHARE_ALLGATHER") == "1":
            full = self._back_to_back(dsts, flat, world)
            if full is not None:
                cnt = full.numel() // world
                self.dist.all_gather_into_tensor(full, full[self.rank * cnt:(self.rank + 1) * cnt])
                return
            if self._equal_blocks(dsts, world):
                # scattered destinations of equal size (the subtree roots' h: 16 small vectors): pack this rank's share, ONE
                # all-gather into a staging tensor, one multi-tensor copy into the leaf buffers - 3 launches instead of K
                import torch
                n = dsts[0].numel()
                mine = torch.cat([dsts[k].reshape(-1) for k in range(K) if int(self.plan.owner[k]) == self.rank])
                staged = torch.empty(K * n, dtype=mine.dtype, device=mine.device)
                self.dist.all_gather_into_tensor(staged, mine)
                torch._foreach_copy_([d.reshape(-1) for d in dsts], list(staged.split(n)))
                return
        for k in range(K):
            self.dist.broadcast(dsts[k], src=int(self.plan.owner[k]))

    def _equal_blocks(self, dsts, world):
        K = len(dsts)
        if K % world:
            return False
        per, n = K // world, dsts[0].numel()
        return all(int(self.plan.owner[k]) == k // per and dsts[k].numel() == n for k in range(K))

    def _back_to_back(self, dsts, flat, world):
        """The flat tensor over `dsts` when one all-gather can replace the broadcasts, else None (same answer on every rank:
        it depends on the plan and on the library's deterministic layout only)."""
        K = len(dsts)
        if flat is None or K % world:
            return None
        per = K // world
        n = dsts[0].numel()
        for k in range(K):
            if int(self.plan.owner[k]) != k // per or dsts[k].numel() != n:
                return None
            if dsts[k].data_ptr() != dsts[0].data_ptr() + k * n * dsts[0].element_size():
                return None
        return flat(dsts[0], K * n)

    def allgather_rows(self, full):
        """In-place all-gather of the contiguous, equally sized row slices of `full` (slice r was computed by rank r)."""
        world = self.dist.get_world_size() if self.dist.is_initialized() else 1
        if world == 1:
            return
        cnt = full.numel() // world
        self.dist.all_gather_into_tensor(full, full[self.rank * cnt:(self.rank + 1) * cnt])

    def gather_T(self, local, top):
        self._gather(lambda k: local.root_T(k), lambda k: top.leaf_T(k) if top is not None else None)

    def gather_h(self, local, top):
        self._gather(lambda k: local.root_h(k), lambda k: top.leaf_h(k) if top is not None else None)

    def scatter_g(self, top, local):
        self._scatter(lambda k: top.leaf_g(k) if top is not None else None, lambda k: local.root_g(k))


class _LocalGpu:
    def __init__(self, eng: GpuEngine, roots, my_subtrees):
        self.eng = eng
        self.idx = {k: int(r) for k, r in zip(my_subtrees, roots)}

    def root_T(self, k):
        return self.eng.operator_view(self.idx[k], "T_uncoarsened")

    def root_h(self, k):
        return self.eng.vector_view(self.idx[k], "h0")

    def root_g(self, k):
        return self.eng.vector_view(self.idx[k], "g")


class _TopGpu:
    def __init__(self, eng: GpuEngine, leaf_nodes):
        self.eng, self.leaf_nodes = eng, leaf_nodes

    def leaf_T(self, j):
        return self.eng.operator_view(self.leaf_nodes[j], "T_uncoarsened")

    def leaf_h(self, j):
        return self.eng.vector_view(self.leaf_nodes[j], "h0")

    def leaf_g(self, j):
        return self.eng.vector_view(self.leaf_nodes[j], "g")


class ShardedHPS:
    """Benchmark/driver-facing sharded HPS (same stage names as HPSAlgorithm; inputs/outputs are this rank's share)."""

    def __init__(self, mesh, solver, device=0, rank=0, world=1, options=None, cut=2, top_mode="replicated", balance="count"):
        import torch
        self.top_mode = top_mode
        self.p2p = False
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.mesh, self.patch_solver, self.rank, self.world, self._device = mesh, solver, rank, world, device
        self.options = {"cache-operators": False, "homogeneous-rhs": False}
        if options:
            self.options.update(options)
        if solver.solver_type not in ("FISHPACK90", "FivePointStencil"):
            raise ValueError("unknown solver_type")
        self.plan = ShardPlan(mesh.level, mesh.child, mesh.box, mesh.nx, world, cut, balance=balance)
        if self.options["cache-operators"] and len(set(int(v) for v in np.asarray(mesh.level)[mesh.leaf_nodes])) != 1:
            # quirk q1 (HPSAlgorithm.hpp:134-139) reuses GLOBAL leaf 0's map for every leaf; a forest handle would reuse its own
            # first leaf's, which differs as soon as leaves have different sizes
            raise NotImplementedError("cache-operators on a sharded tree needs leaves of one size (the option is only valid on uniform meshes)")
        ids, lev, ch, box, roots = self.plan.local_table(rank)
        self.local = GpuEngine(lev, ch, box, mesh.nx, device)
        self._stream = torch.cuda.ExternalStream(self.local.stream(), device=torch.device("cuda", int(device)))
        self.local_if = _LocalGpu(self.local, roots, self.plan.subtrees_of(rank))
        self.top = self.top_if = None
        if rank == 0 or top_mode == "replicated":
            tids, tlev, tch, tbox, ext = self.plan.top_table()
            self.top = GpuEngine(tlev, tch, tbox, mesh.nx, device, ext_sizes=ext, stream=self.local.stream())
            if top_mode == "replicated":
                def _ag(t):
                    with torch.cuda.stream(self._stream):
                        self.xchg.allgather_rows(t)
                # peer mode (default on one NVSwitch domain): the library exchanges row slices itself through peer-mapped arenas;
                # EFGPU_P2P=0 keeps round 1's path (ncclAllGather through torch.distributed from a host callback)
                self.p2p = world > 1 and world <= 8 and os.environ.get("EFGPU_P2P", "1") != "0"
                self.top.set_partition(rank, world, _ag if (world > 1 and not self.p2p) else None)     # must precede the first device view
                if self.p2p:
                    self.top.peer_setup(dist, world)
            self.top_if = _TopGpu(self.top, [int(i) for i in np.nonzero(tch[:, 0] < 0)[0]])
            self._top_level = tlev
            self._top_interior = [int(i) for i in np.nonzero(tch[:, 0] >= 0)[0]]
        self.xchg = ShardedExchange(self.plan, rank, dist)
        self.leaf_lo, self.leaf_hi = self.plan.local_leaf_range(rank)
        if solver.solver_type == "FISHPACK90":
            lam = float(solver.lambda_function(np.float64(0.0), np.float64(0.0)))      # FiniteVolumeSolver.cpp:254
            self.local.set_leaf_constant(lam)
        else:   # this rank's leaves only: sampled where FiniteVolumeSolver.cpp:63-79 samples (general merge plan above them)
            boxes = mesh.box[mesh.leaf_nodes[self.leaf_lo:self.leaf_hi]]
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            self.local.set_leaf_variable(sample_leaf_coefficients(solver, boxes, mesh.nx, max(1, threads // max(1, world))))

    def __del__(self):
        self.top_if = self.top = None      # the upper-tree handle borrows the forest handle's stream: release it first
        self.local_if = self.local = None

    # -- helpers -------------------------------------------------------------------------------
    no_symmetry = False
    lazy_root_dtn = False      # EFGPU_LAZY_ROOT_DTN on the upper tree (the forests have no level-0 merge)

    def _flags(self):
        return ((CACHE_OPERATORS if self.options["cache-operators"] else 0) | (HOMOGENEOUS_RHS if self.options["homogeneous-rhs"] else 0)
                | (NO_SYMMETRY if self.no_symmetry else 0) | (LAZY_ROOT_DTN if self.lazy_root_dtn else 0))

    def sharding(self):
        K = len(self.plan.cut_nodes)
        per = ("%d per GPU" % (K // self.world) if self.plan.balance == "count" else
               "%s per GPU, balanced by %s" % ("/".join(str(len(self.plan.subtrees_of(r))) for r in range(self.world)), self.plan.balance))
        if self.top_mode == "replicated":
            if self.p2p:
                return ("level-%d subtrees in Morton blocks over %d GPUs (%s); upper tree replicated and row-partitioned %d ways: every rank maps every "
                        "rank's operator arena (CUDA IPC over NVLink), the merge GEMMs store their row slices of the X^-1 products, S and T into all "
                        "arenas from the epilogue between device-side flag barriers (no NCCL on the data path), subtree-root T / h by peer stores"
                        % (self.plan.cut, self.world, per, self.world))
            how = "all-gathered in place (one collective)" if os.environ.get("EFGPU_SHARE_ALLGATHER") == "1" else "broadcast"
            return ("level-%d subtrees in Morton blocks over %d GPUs (%s); upper tree replicated: subtree-root T %s over NCCL, "
                    "X^-1 on every rank, rows of S and T split %d ways and all-gathered" % (self.plan.cut, self.world, per, how, self.world))
        return "level-%d subtrees in Morton blocks over %d GPUs (%s); subtree-root T/h gathered to rank 0 over NCCL, g scattered back" % (
            self.plan.cut, self.world, per)

    def stream(self):
        return self.local.stream()

    def setupStage(self):
        return None

    def sample_inputs(self, f_fn, u_fn):
        """This rank's share of f (its leaves, global pre-order) and the root Dirichlet data (rank 0 uses it)."""
        m = self.mesh
        b = m.box[m.leaf_nodes[self.leaf_lo:self.leaf_hi]]
        M = m.nx
        dx = (b[:, 1] - b[:, 0]) / M
        dy = (b[:, 3] - b[:, 2]) / M
        k = np.arange(M)
        xs = (b[:, 0] + dx / 2)[:, None] + k[None, :] * dx[:, None]
        ys = (b[:, 2] + dy / 2)[:, None] + k[None, :] * dy[:, None]
        X = np.broadcast_to(xs[:, :, None], (len(b), M, M))
        Y = np.broadcast_to(ys[:, None, :], (len(b), M, M))
        self._XY = (X, Y)
        f = np.ascontiguousarray(f_fn(X, Y), dtype=np.float64).reshape(-1)
        size = self.plan.root_size()
        xl, xu, yl, yu = m.box[0]
        g = FiniteVolumeGrid(size, xl, xu, size, yl, yu)
        kk = np.arange(size)
        xs, ys = g.point(0, kk), g.point(1, kk)
        x = np.concatenate([np.full(size, xl), np.full(size, xu), xs, xs])
        y = np.concatenate([ys, ys, np.full(size, yl), np.full(size, yu)])
        gr = np.ascontiguousarray(u_fn(x, y), dtype=np.float64).reshape(-1)
        return f, gr

    # -- stages --------------------------------------------------------------------------------
    def buildStage(self):
        fl = self._flags()
        self.local.build(fl)
        # the subtree roots' DtN maps are signed-symmetric when EVERY rank's forest used the symmetric plan (collective)
        sym = global_symmetry(self.dist, self.local.is_symmetric(), self.world, self._device)
        # ... and the upper tree refines its X^-1 (indefinite operator, lambda > 0) when ANY forest did: not (MIN of the negation)
        refine = not global_symmetry(self.dist, self.local.stats()["inverse_residual"] < 0.0, self.world, self._device)
        if self.top is not None:
            self.top.set_symmetric_leaves(sym)
            self.top.set_refine_inverse(1 if refine else 0)
        if self.top_mode != "replicated":
            with self.torch.cuda.stream(self._stream):
                self.xchg.gather_T(self.local_if, self.top_if)
            if self.top is not None:
                self.top.build(fl)
            return
        with self.torch.cuda.stream(self._stream):
            if self.p2p:
                self._share_p2p(self.local_if.root_T, self.top_if.leaf_T)
            else:
                self.xchg.share(self.local_if.root_T, self.top_if.leaf_T, flat=lambda first, total: _dev_tensor(first.data_ptr(), total, self._device))
        # levels above the cut: X^-1 products, S and T are computed in row slices and all-gathered by the library
        # through the callback above (the root's DtN map stays row-distributed)
        self.top.build(fl)

    def _share_p2p(self, local_view, top_view):
        """Every rank ends up with every subtree root's operator / vector in its upper-tree leaf buffers: the owner copies its own
        into place and stores them into the same place of every other rank's arena (NVLink), between two flag barriers (the first:
        nobody still reads the buffers from the previous step; the second: every rank's stores have landed)."""
        self.top.peer_barrier()
        for k in self.plan.subtrees_of(self.rank):
            dst = top_view(k)
            dst.copy_(local_view(k))
            self.top.peer_broadcast(dst)
        self.top.peer_barrier()

    def gather_root_T(self):
        """Parity/debug: the root DtN map assembled from its row slices (every rank gets the whole matrix)."""
        if self.top_mode == "replicated":
            self.top.complete_root_T()
        return self.top.operator_view(0, "T_uncoarsened")

    def upwardsStageDevice(self, f_dev_ptr, scale=1.0, sync=True):
        fl = self._flags()
        self.local.upwards(f_dev_ptr, scale, fl)
        if not (fl & HOMOGENEOUS_RHS):
            with self.torch.cuda.stream(self._stream):
                if self.top_mode == "replicated" and self.p2p:
                    self._share_p2p(self.local_if.root_h, self.top_if.leaf_h)
                elif self.top_mode == "replicated":
                    self.xchg.share(self.local_if.root_h, self.top_if.leaf_h)
                else:
                    self.xchg.gather_h(self.local_if, self.top_if)
            if self.top is not None:
                self.top.upwards(0, 1.0, fl)
        if sync:
            self.local.sync()

    def solveStageDevice(self, g_dev_ptr, u_dev_ptr=0, sync=True):
        fl = self._flags()
        if self.top is not None:
            groot = self.top.vector_view(0, "g")
            with self.torch.cuda.stream(self._stream):
                groot.copy_(_dev_tensor(g_dev_ptr, groot.numel(), self._device))
            self.top.solve_from_roots(0, fl)
        with self.torch.cuda.stream(self._stream):
            if self.top_mode == "replicated":   # every rank walked the upper tree itself: its subtree roots' g are local
                for k in self.plan.subtrees_of(self.rank):
                    self.local_if.root_g(k).copy_(self.top_if.leaf_g(k))
            else:
                self.xchg.scatter_g(self.top_if, self.local_if)
        self.local.solve_from_roots(u_dev_ptr, fl, sync=sync)

    def upwardsStageHost(self, f_host):
        t = self.torch
        if getattr(self, "_f_dev", None) is None:
            self._f_dev = t.empty(f_host.size, dtype=t.float64, device=t.device("cuda", self._device))
        with t.cuda.stream(self._stream):
            self._f_dev.copy_(t.from_numpy(f_host), non_blocking=True)
        self.upwardsStageDevice(self._f_dev.data_ptr(), 1.0, sync=True)

    def solveStageHost(self, g_host, u_host):
        t = self.torch
        if getattr(self, "_u_dev", None) is None:
            self._u_dev = t.empty(u_host.size, dtype=t.float64, device=t.device("cuda", self._device))
            self._g_dev = t.empty(g_host.size, dtype=t.float64, device=t.device("cuda", self._device))
        with t.cuda.stream(self._stream):
            self._g_dev.copy_(t.from_numpy(g_host), non_blocking=True)
        self.solveStageDevice(self._g_dev.data_ptr(), self._u_dev.data_ptr(), sync=False)
        with t.cuda.stream(self._stream):
            t.from_numpy(u_host).copy_(self._u_dev, non_blocking=True)
        self.local.sync()

    def sync(self):
        self.local.sync()

    # -- reporting -----------------------------------------------------------------------------
    def set_profiling(self, on=True):
        self.local.set_profiling(on)
        if self.top is not None:
            self.top.set_profiling(on)

    def profile(self):
        """Per-class (ms, launches) of this rank: its forest plus, on rank 0, the upper tree."""
        p = self.local.profile()
        if self.top is not None:
            for k, (ms, n) in self.top.profile().items():
                p[k] = (p[k][0] + ms, p[k][1] + n)
        return p

    def stats(self):
        s = self.local.stats()
        if self.top is not None:
            ts = self.top.stats()
            for k in ("merge_flops_canonical", "merge_flops_issued", "upwards_bytes", "solve_bytes", "device_bytes"):
                s[k] += ts[k]
            for k in ("build_ms", "upwards_ms", "solve_ms"):
                s[k] += ts[k]
        return s

    def total_issued_flops(self):
        """Whole-tree issued flops (sum over ranks)."""
        t = self.torch.tensor([self.stats()["merge_flops_issued"]], dtype=self.torch.float64, device=self.torch.device("cuda", self._device))
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t[0])

    def max_error(self, u_dev, u_fn):
        X, Y = self._XY
        u = u_dev.cpu().numpy().reshape(X.shape)
        e = self.torch.tensor([float(np.max(np.abs(u - u_fn(X, Y))))], dtype=self.torch.float64, device=self.torch.device("cuda", self._device))
        if self.world > 1:
            self.dist.all_reduce(e, op=self.dist.ReduceOp.MAX)
        return float(e[0])


class GroupedExchange:
    """The exchanges of GroupedShardedHPS on plain tensors (device views over NCCL on the GPUs; CPU tensors over gloo in
    tests/test_sharded_gloo.py): world = 4 gs ranks, group k = ranks [k gs, (k + 1) gs) owns the four level-2 subtrees
    4 k .. 4 k + 3 under level-1 node k, rank q of a group the subtrees [4 q / gs, 4 (q + 1) / gs) of those four."""

    def __init__(self, plan: ShardPlan, rank: int, world: int, dist):
        if world % 4 or world // 4 not in (1, 2, 4):
            raise ValueError("grouped sharding needs 4, 8 or 16 ranks")
        self.plan, self.rank, self.world, self.dist = plan, rank, world, dist
        self.gs = gs = world // 4
        self.group_id, self.group_rank = rank // gs, rank % gs
        self.group = None
        for k in range(4):      # every rank creates every group (new_group is collective); ranks keep their own
            g = dist.new_group(list(range(k * gs, (k + 1) * gs))) if gs > 1 else None
            if k == self.group_id:
                self.group = g

    def local_index(self, k):
        """Leaf slot (0..3) of global subtree k under this rank's level-1 node."""
        j = k - 4 * self.group_id
        if not 0 <= j < 4:
            raise ValueError("subtree %d does not belong to group %d" % (k, self.group_id))
        return j

    def allgather_group(self, t):
        """In place: equal contiguous chunks of t, chunk q from rank q of this rank's group."""
        if self.gs > 1:
            cnt = t.numel() // self.gs
            self.dist.all_gather_into_tensor(t, t[self.group_rank * cnt:(self.group_rank + 1) * cnt], group=self.group)

    def allgather_world(self, t):
        cnt = t.numel() // self.world
        self.dist.all_gather_into_tensor(t, t[self.rank * cnt:(self.rank + 1) * cnt])

    def subtree_T_to_group(self, slab, n, root_T):
        """slab: the four leaf DtN maps (n doubles each, back to back) of the level-1 handle; root_T(k): this rank's subtree roots."""
        for k in self.plan.subtrees_of(self.rank):
            j = self.local_index(k)
            slab[j * n:(j + 1) * n].copy_(root_T(k))
        self.allgather_group(slab)      # rank q owns leaves [4 q / gs, 4 (q + 1) / gs): equal contiguous chunks

    def level1_T_to_world(self, Tk, slab):
        """Tk: this group's level-1 DtN map (whole on every rank of the group); slab: the four leaf maps of the root handle.
        Chunk r of the slab is part r % gs of leaf r // gs: every rank sends 1 / gs of its group's map."""
        cnt = slab.numel() // self.world
        slab[self.rank * cnt:(self.rank + 1) * cnt].copy_(Tk[self.group_rank * cnt:(self.group_rank + 1) * cnt])
        self.allgather_world(slab)

    def subtree_h_to_group(self, leaf_h, root_h):
        """leaf_h(j): leaf vector j (0..3) of the level-1 handle; root_h(k): this rank's subtree roots."""
        per = 4 // self.gs
        for k in self.plan.subtrees_of(self.rank):
            leaf_h(self.local_index(k)).copy_(root_h(k))
        if self.gs > 1:
            for j in range(4):
                self.dist.broadcast(leaf_h(j), src=self.group_id * self.gs + j // per, group=self.group)

    def level1_h_to_world(self, own_h, leaf_h):
        """own_h: h of this group's level-1 node (every rank of the group holds it: its first rank sends); leaf_h(k) of the root handle."""
        leaf_h(self.group_id).copy_(own_h)
        for k in range(4):
            self.dist.broadcast(leaf_h(k), src=k * self.gs)


class GroupedShardedHPS(ShardedHPS):
    """Three tiers for world = 4 * gs ranks (gs = 1, 2 or 4; eight GPUs: gs = 2) on a tree whose 16 level-2 subtrees have equal
    roots.  The four subtrees under level-1 node k belong to the ranks of group k = [k gs, (k + 1) gs) (Morton blocks), and in
    the solve a rank only descends through its own level-1 parent.  So

      local  the forest of this rank's subtrees (no communication);
      mid    level-1 node k alone, its four subtree roots as external leaves, row-partitioned over the gs ranks of the group:
             the subtree roots' DtN maps and every all-gather of that merge stay inside the group;
      top    the root merge alone, the four level-1 nodes as external leaves, row-partitioned over all ranks: only the level-1
             DtN maps are exchanged between groups (one in-place all-gather, each rank sending 1 / gs of its group's map).

    Against the replicated upper tree of ShardedHPS this removes the 16 world-wide broadcasts of the subtree roots' maps, gathers
    the level-1 S and the products of its inversion over gs instead of 4 gs ranks, and no rank repeats another group's small
    products (model estimate -4.5 of 52 ms at eight GPUs).  Not yet run on GPUs."""

    def __init__(self, mesh, solver, device=0, rank=0, world=4, options=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.top_mode = "grouped"
        self.mesh, self.patch_solver, self.rank, self.world, self._device = mesh, solver, rank, world, device
        self.options = {"cache-operators": False, "homogeneous-rhs": False}
        if options:
            self.options.update(options)
        if world % 4 or world // 4 not in (1, 2, 4):
            raise ValueError("grouped sharding needs 4, 8 or 16 ranks")
        if solver.solver_type != "FISHPACK90":
            raise NotImplementedError("grouped sharding: constant-coefficient leaves only")
        self.plan = plan = ShardPlan(mesh.level, mesh.child, mesh.box, mesh.nx, world, 2)
        l1 = np.nonzero(plan.level == 1)[0]                              # the four level-1 nodes, Morton order
        if len(set(int(v) for v in plan.size[plan.cut_nodes])) != 1 or len(set(int(v) for v in plan.size[l1])) != 1:
            raise ValueError("grouped sharding needs equal subtree roots (uniform upper levels)")
        ids, lev, ch, box, roots = plan.local_table(rank)
        self.local = GpuEngine(lev, ch, box, mesh.nx, device)
        self._stream = torch.cuda.ExternalStream(self.local.stream(), device=torch.device("cuda", int(device)))
        self.local_if = _LocalGpu(self.local, roots, plan.subtrees_of(rank))
        self.xchg = ShardedExchange(plan, rank, dist)
        self.gx = GroupedExchange(plan, rank, world, dist)
        self.gs, self.group_id, self.group_rank, self.group = self.gx.gs, self.gx.group_id, self.gx.group_rank, self.gx.group
        gs = self.gs
        self.leaf_lo, self.leaf_hi = plan.local_leaf_range(rank)
        self.local.set_leaf_constant(float(solver.lambda_function(np.float64(0.0), np.float64(0.0))))
        star = np.array([[1, 2, 3, 4]] + [[-1] * 4] * 4, dtype=np.int32)   # one merge, four external leaves
        me1 = int(l1[self.group_id])
        kids = [int(c) for c in plan.child[me1]]
        assert kids == [int(plan.cut_nodes[4 * self.group_id + j]) for j in range(4)]
        self.mid = GpuEngine(np.array([1, 2, 2, 2, 2], dtype=np.int32), star, plan.box[[me1] + kids], mesh.nx, device,
                             ext_sizes=plan.size[kids].astype(np.int32), stream=self.local.stream())

        def _ag_group(t):
            with torch.cuda.stream(self._stream):
                self.gx.allgather_group(t)

        def _ag_world(t):
            with torch.cuda.stream(self._stream):
                self.gx.allgather_world(t)
        if gs > 1:
            self.mid.set_partition(self.group_rank, gs, _ag_group)
        self.top = GpuEngine(np.array([0, 1, 1, 1, 1], dtype=np.int32), star, plan.box[[0] + [int(i) for i in l1]], mesh.nx, device,
                             ext_sizes=plan.size[l1].astype(np.int32), stream=self.local.stream())
        self.top.set_partition(rank, world, _ag_world)
        self.top_if = None

    def __del__(self):
        self.top = self.mid = None         # both borrow the forest handle's stream: release them first
        self.local_if = self.local = None

    def sharding(self):
        return ("level-2 subtrees in Morton blocks over %d GPUs (%d per GPU); level-1 merges row-split inside groups of %d GPUs, root merge over "
                "all %d; between groups only the level-1 DtN maps travel (one in-place all-gather)" % (self.world, 16 // self.world, self.gs, self.world))

    def _slab(self, eng, which):
        """One tensor over the four leaf buffers of a star handle (libefgpu lays the leaf DtN maps out back to back)."""
        v = [eng.operator_view(1 + j, which) for j in range(4)]
        n = v[0].numel()
        if any(v[j].data_ptr() != v[0].data_ptr() + 8 * j * n or v[j].numel() != n for j in range(4)):
            raise RuntimeError("leaf DtN maps are not contiguous")
        return _dev_tensor(v[0].data_ptr(), 4 * n, self._device), n

    def buildStage(self):
        t, dist, fl = self.torch, self.dist, self._flags()
        self.local.build(fl)
        # symmetric plans are decided by every rank that feeds the handle: the group for `mid`, the world for `top`
        self.mid.set_symmetric_leaves(global_symmetry(dist, self.local.is_symmetric(), self.gs, self._device, group=self.group))
        refine = 1 if self.local.stats()["inverse_residual"] >= 0.0 else 0      # constant lambda: the same answer on every rank
        self.mid.set_refine_inverse(refine)
        self.top.set_refine_inverse(refine)
        with t.cuda.stream(self._stream):
            slab, n = self._slab(self.mid, "T_uncoarsened")
            self.gx.subtree_T_to_group(slab, n, self.local_if.root_T)
        self.mid.build(fl & ~LAZY_ROOT_DTN)      # its root is a level-1 node: gathered and mirrored inside the group by the library
        self.top.set_symmetric_leaves(global_symmetry(dist, self.mid.is_symmetric(), self.world, self._device))
        with t.cuda.stream(self._stream):
            slab, n = self._slab(self.top, "T_uncoarsened")
            self.gx.level1_T_to_world(self.mid.operator_view(0, "T_uncoarsened"), slab)
        self.top.build(fl)

    def gather_root_T(self):
        self.top.complete_root_T()
        return self.top.operator_view(0, "T_uncoarsened")

    def upwardsStageDevice(self, f_dev_ptr, scale=1.0, sync=True):
        t, dist, fl = self.torch, self.dist, self._flags()
        self.local.upwards(f_dev_ptr, scale, fl)
        if not (fl & HOMOGENEOUS_RHS):
            with t.cuda.stream(self._stream):
                self.gx.subtree_h_to_group(lambda j: self.mid.vector_view(1 + j, "h0"), self.local_if.root_h)
            self.mid.upwards(0, 1.0, fl)
            with t.cuda.stream(self._stream):
                self.gx.level1_h_to_world(self.mid.vector_view(0, "h0"), lambda k: self.top.vector_view(1 + k, "h0"))
            self.top.upwards(0, 1.0, fl)
        if sync:
            self.local.sync()

    def solveStageDevice(self, g_dev_ptr, u_dev_ptr=0, sync=True):
        t, fl = self.torch, self._flags()
        groot = self.top.vector_view(0, "g")
        with t.cuda.stream(self._stream):
            groot.copy_(_dev_tensor(g_dev_ptr, groot.numel(), self._device))
        self.top.solve_from_roots(0, fl)
        with t.cuda.stream(self._stream):
            self.mid.vector_view(0, "g").copy_(self.top.vector_view(1 + self.group_id, "g"))
        self.mid.solve_from_roots(0, fl)
        with t.cuda.stream(self._stream):
            for k in self.plan.subtrees_of(self.rank):
                self.local_if.root_g(k).copy_(self.mid.vector_view(1 + self.gx.local_index(k), "g"))
        self.local.solve_from_roots(u_dev_ptr, fl, sync=sync)

    def set_profiling(self, on=True):
        for e in (self.local, self.mid, self.top):
            e.set_profiling(on)

    def profile(self):
        p = self.local.profile()
        for e in (self.mid, self.top):
            for k, (ms, n) in e.profile().items():
                p[k] = (p[k][0] + ms, p[k][1] + n)
        return p

    def stats(self):
        s = self.local.stats()
        for e in (self.mid, self.top):
            ts = e.stats()
            for k in ("merge_flops_canonical", "merge_flops_issued", "upwards_bytes", "solve_bytes", "device_bytes", "build_ms", "upwards_ms", "solve_ms"):
                s[k] += ts[k]
        return s
