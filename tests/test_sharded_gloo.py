"""World-size-2 gloo tests (CPU) of the sharded path's host logic: the ShardPlan node tables, the
Morton-block ownership and the three exchanges (gather T, gather h, scatter g) of
ellipticforest_b200.sharded.ShardedExchange.  The arithmetic underneath is the numpy oracle (test
infrastructure); on the GPU box the same plan/exchange code drives libefgpu handles over NCCL
(tests/test_gpu_sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hps_oracle as O
from ellipticforest_b200.sharded import GroupedExchange, ShardPlan, ShardedExchange

HERE = os.path.dirname(os.path.abspath(__file__))


def _tables(nodes):
    level = np.array([n.level for n in nodes], dtype=np.int32)
    child = np.array([n.children if n.children else [-1] * 4 for n in nodes], dtype=np.int32)
    box = np.array([[n.grid.xl, n.grid.xu, n.grid.yl, n.grid.yu] for n in nodes])
    return level, child, box


class OracleEngine:
    """Forest (FV leaves) or upper tree (external leaves) driven by the numpy oracle's per-node functions."""

    def __init__(self, level, child, box, nx, solver, ext_sizes=None):
        self.nodes, self.ext = [], ext_sizes is not None
        li = 0
        for i in range(len(level)):
            leaf = child[i][0] < 0
            n = int(ext_sizes[li]) if (leaf and self.ext) else nx
            li += int(leaf)
            self.nodes.append(O.Node(path=str(i), level=int(level[i]), grid=O.Grid(n, *box[i]), leaf=bool(leaf),
                                     children=[] if leaf else [int(c) for c in child[i]]))
        for i, nd in enumerate(self.nodes):
            for c in nd.children:
                self.nodes[c].parent = i
        self.roots = [i for i, nd in enumerate(self.nodes) if nd.parent < 0]
        self.hps = O.HPS(self.nodes, solver)
        self.buf = {}

    def _post(self):
        return [i for r in self.roots for i in O.post_order(self.nodes, r)]

    def tensor(self, node, name, n):
        """storage that outlives the exchange (the engines' 'device views')"""
        key = (node, name)
        if key not in self.buf:
            self.buf[key] = torch.zeros(n, dtype=torch.float64)
        return self.buf[key]

    def build(self):
        for i in self._post():
            nd = self.nodes[i]
            if nd.leaf:
                if self.ext:
                    nd.T = self.buf[(i, "T")].numpy().reshape(4 * nd.grid.nx, 4 * nd.grid.nx).copy()
                else:
                    nd.T = self.hps.solver.buildD2N(nd.grid)
            else:
                self.hps.merge4to1(nd, *[self.nodes[c] for c in nd.children])

    def upwards(self, f_fn):
        for i in self._post():
            nd = self.nodes[i]
            if nd.leaf:
                if self.ext:
                    nd.h = self.buf[(i, "h")].numpy().copy()
                else:
                    g = nd.grid
                    X, Y = np.meshgrid(g.x(np.arange(g.nx)), g.y(np.arange(g.nx)), indexing="ij")
                    nd.f = np.asarray(f_fn(X, Y), dtype=np.float64).reshape(-1)
                    nd.h = self.hps.solver.particularNeumannData(g, nd.f)
            else:
                self.hps.upwards4to1(nd, *[self.nodes[c] for c in nd.children])

    def solve(self):
        order = []
        def pre(i):
            order.append(i)
            for c in self.nodes[i].children:
                pre(c)
        for r in self.roots:
            pre(r)
        for i in order:
            nd = self.nodes[i]
            if nd.leaf:
                if not self.ext:
                    nd.u = self.hps.solver.solve(nd.grid, nd.g, nd.f)
                else:  # a subtree root tagged by its parent receives coarsened data: uncoarsen_ (HPSAlgorithm.hpp:1165-1183)
                    for n in range(nd.n_coarsens):
                        nfine = nd.grid.nx // (2 ** (nd.n_coarsens - (n + 1)))
                        nd.g = O._blkdiag4(O.L12(nfine)) @ nd.g
            else:
                self.hps.split1to4(nd, *[self.nodes[c] for c in nd.children])


class LocalIf:
    def __init__(self, eng, roots, subtrees):
        self.eng, self.idx = eng, {k: int(r) for k, r in zip(subtrees, roots)}

    def root_T(self, k):
        nd = self.eng.nodes[self.idx[k]]
        return torch.from_numpy(np.ascontiguousarray(nd.T)).reshape(-1)

    def root_h(self, k):
        return torch.from_numpy(np.ascontiguousarray(self.eng.nodes[self.idx[k]].h))

    def root_g(self, k):
        nd = self.eng.nodes[self.idx[k]]
        return self.eng.tensor(self.idx[k], "g", 4 * nd.grid.nx)


class TopIf:
    def __init__(self, eng, back_to_back=False):
        self.eng = eng
        self.leaves = [i for i, nd in enumerate(eng.nodes) if nd.leaf]
        self.big = None
        if back_to_back:   # the layout libefgpu gives the leaf DtN maps of the upper tree: one slab, leaf order
            sizes = [(4 * eng.nodes[i].grid.nx) ** 2 for i in self.leaves]
            self.big = torch.zeros(sum(sizes), dtype=torch.float64)
            off = 0
            for i, n in zip(self.leaves, sizes):
                eng.buf[(i, "T")] = self.big[off:off + n]
                off += n

    def flat(self, first, total):
        assert first.data_ptr() == self.big.data_ptr()
        return self.big[:total]

    def leaf_T(self, j):
        i = self.leaves[j]
        return self.eng.tensor(i, "T", (4 * self.eng.nodes[i].grid.nx) ** 2)

    def leaf_h(self, j):
        i = self.leaves[j]
        return self.eng.tensor(i, "h", 4 * self.eng.nodes[i].grid.nx)

    def leaf_g(self, j):
        return torch.from_numpy(np.ascontiguousarray(self.eng.nodes[self.leaves[j]].g))


CASES = {
    "uniform": dict(problem_name="helmholtz", box=(0.0, np.pi, 0.0, np.pi), nx=8, min_level=3, max_level=3, refine_box=None),
    # adaptive below the cut, including a subtree root that is coarsened by its parent's merge
    "adaptive": dict(problem_name="poisson", box=(-10.0, 10.0, -10.0, 10.0), nx=8, min_level=2, max_level=4, refine_box=(-10.0, 0.5, -10.0, 0.5)),
}


def _worker_replicated(rank, world, port, case, out_dir, one_allgather=False, balance="count", cut=2):
    """Replicated upper tree: share() of the subtree roots, then per level the row slices of S and T are
    all-gathered in place (the oracle computes whole merges; rows a rank does not own are wiped first, so
    only the exchange can restore them)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kw = CASES[case]
        P = O.problem(kw["problem_name"])
        ind = O.refine_box_indicator(kw["refine_box"]) if kw["refine_box"] else O.refine_indicator(1.2)
        nodes = O.build_tree(ind, kw["box"], kw["nx"], kw["min_level"], kw["max_level"])
        solver = O.Solver(kind="fishpack", alpha=P["alpha"], beta=P["beta"], lam=P["lam"])
        plan = ShardPlan(*_tables(nodes), kw["nx"], world, cut=cut, balance=balance)
        ids, lev, ch, box, roots = plan.local_table(rank)
        local = OracleEngine(lev, ch, box, kw["nx"], solver)
        lif = LocalIf(local, roots, plan.subtrees_of(rank))
        tids, tlev, tch, tbox, ext = plan.top_table()
        top = OracleEngine(tlev, tch, tbox, kw["nx"], solver, ext_sizes=ext)
        tif = TopIf(top, back_to_back=one_allgather)
        x = ShardedExchange(plan, rank, dist)
        local.build()
        if one_allgather:
            os.environ["EFGPU_SHARE_ALLGATHER"] = "1"
            calls = []
            real = dist.broadcast
            dist.broadcast = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        x.share(lif.root_T, tif.leaf_T, flat=tif.flat if one_allgather else None)
        if one_allgather:
            dist.broadcast = real
            # equal subtree roots in equal Morton blocks travel as ONE all-gather; anything else falls back to broadcasts
            sizes = {tif.leaf_T(j).numel() for j in range(len(tif.leaves))}
            assert (len(calls) == 0) == (len(sizes) == 1), (len(calls), sizes)
        for i in top._post():     # leaves first (from the shared buffers), then merges level by level
            nd = top.nodes[i]
            if nd.leaf:
                nd.T = top.buf[(i, "T")].numpy().reshape(4 * nd.grid.nx, 4 * nd.grid.nx).copy()
        for level in range(cut - 1, -1, -1):
            for i, nd in enumerate(top.nodes):
                if nd.leaf or nd.level != level:
                    continue
                top.hps.merge4to1(nd, *[top.nodes[c] for c in nd.children])
                for name in ("S", "T"):
                    full = torch.from_numpy(np.ascontiguousarray(getattr(nd, name))).reshape(-1)
                    cnt = full.numel() // world
                    keep = full[rank * cnt:(rank + 1) * cnt].clone()
                    full.zero_()
                    full[rank * cnt:(rank + 1) * cnt] = keep
                    x.allgather_rows(full)
                    setattr(nd, name, full.numpy().reshape(getattr(nd, name).shape).copy())
        local.upwards(P["f"])
        if one_allgather:   # scattered equal-size vectors: packed into one all-gather as well (no broadcast), else the fallback
            calls = []
            dist.broadcast = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        x.share(lif.root_h, tif.leaf_h)
        if one_allgather:
            dist.broadcast = real
            assert (len(calls) == 0) == (len({tif.leaf_h(j).numel() for j in range(len(tif.leaves))}) == 1), len(calls)
        top.upwards(None)
        r, a, b = O.HPS.root_boundary(type("R", (), {"nodes": [top.nodes[0]]})(), lambda s_, xx, yy: (float(P["u"](xx, yy)), 1.0, 0.0))
        top.nodes[0].g = r / a
        top.solve()
        for k in plan.subtrees_of(rank):
            local.nodes[lif.idx[k]].g = tif.leaf_g(k).numpy().copy()
        local.solve()
        np.save(os.path.join(out_dir, "u_%d.npy" % rank), np.concatenate([nd.u for nd in local.nodes if nd.leaf]))
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array(plan.local_leaf_range(rank)))
        if rank == 0:
            np.save(os.path.join(out_dir, "T_root.npy"), top.nodes[0].T)
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kw = CASES[case]
        P = O.problem(kw["problem_name"])
        ind = O.refine_box_indicator(kw["refine_box"]) if kw["refine_box"] else O.refine_indicator(1.2)
        nodes = O.build_tree(ind, kw["box"], kw["nx"], kw["min_level"], kw["max_level"])
        solver = O.Solver(kind="fishpack", alpha=P["alpha"], beta=P["beta"], lam=P["lam"])
        plan = ShardPlan(*_tables(nodes), kw["nx"], world)
        ids, lev, ch, box, roots = plan.local_table(rank)
        local = OracleEngine(lev, ch, box, kw["nx"], solver)
        lif = LocalIf(local, roots, plan.subtrees_of(rank))
        top = tif = None
        if rank == 0:
            tids, tlev, tch, tbox, ext = plan.top_table()
            top = OracleEngine(tlev, tch, tbox, kw["nx"], solver, ext_sizes=ext)
            tif = TopIf(top)
        x = ShardedExchange(plan, rank, dist)
        # build
        local.build()
        x.gather_T(lif, tif)
        if top:
            top.build()
        # upwards
        local.upwards(P["f"])
        x.gather_h(lif, tif)
        if top:
            top.upwards(None)
        # solve
        if top:
            r, a, b = O.HPS(nodes, solver).root_boundary.__func__(type("R", (), {"nodes": [top.nodes[0]]})(), lambda s, xx, yy: (float(P["u"](xx, yy)), 1.0, 0.0))
            top.nodes[0].g = r / a
            top.solve()
        x.scatter_g(tif, lif)
        for k in plan.subtrees_of(rank):
            local.nodes[lif.idx[k]].g = lif.root_g(k).numpy().copy()
        local.solve()
        u = np.concatenate([nd.u for nd in local.nodes if nd.leaf])
        np.save(os.path.join(out_dir, "u_%d.npy" % rank), u)
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array(plan.local_leaf_range(rank)))
        if rank == 0:
            np.save(os.path.join(out_dir, "T_root.npy"), top.nodes[0].T)
    finally:
        dist.destroy_process_group()


def _worker_grouped(rank, world, port, case, out_dir):
    """GroupedShardedHPS's three tiers with the numpy oracle as the engine: forest -> level-1 star merge inside the rank group ->
    root star merge over all ranks, every exchange through GroupedExchange (the object the GPU class uses).  The row partition of a
    tier is emulated as in _worker_replicated: a rank keeps its row slice of S and T, wipes the rest, and the tier's in-place
    all-gather (group / world) has to restore it."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kw = CASES[case]
        P = O.problem(kw["problem_name"])
        nodes = O.build_tree(O.refine_indicator(1.2), kw["box"], kw["nx"], kw["min_level"], kw["max_level"])
        solver = O.Solver(kind="fishpack", alpha=P["alpha"], beta=P["beta"], lam=P["lam"])
        plan = ShardPlan(*_tables(nodes), kw["nx"], world, cut=2)
        gx = GroupedExchange(plan, rank, world, dist)
        assert (gx.gs, gx.group_id, gx.group_rank) == (world // 4, rank // (world // 4), rank % (world // 4))
        ids, lev, ch, box, roots = plan.local_table(rank)
        local = OracleEngine(lev, ch, box, kw["nx"], solver)
        lif = LocalIf(local, roots, plan.subtrees_of(rank))
        star = np.array([[1, 2, 3, 4]] + [[-1] * 4] * 4, dtype=np.int32)
        l1 = np.nonzero(plan.level == 1)[0]
        me1 = int(l1[gx.group_id])
        kids = [int(c) for c in plan.child[me1]]
        assert kids == [int(plan.cut_nodes[4 * gx.group_id + j]) for j in range(4)]
        mid = OracleEngine(np.array([1, 2, 2, 2, 2]), star, plan.box[[me1] + kids], kw["nx"], solver, ext_sizes=plan.size[kids])
        top = OracleEngine(np.array([0, 1, 1, 1, 1]), star, plan.box[[0] + [int(i) for i in l1]], kw["nx"], solver, ext_sizes=plan.size[l1])
        mif, tif = TopIf(mid, back_to_back=True), TopIf(top, back_to_back=True)

        def merge_partitioned(eng, part_rank, nranks, allgather, gather_T):
            nd = eng.nodes[0]
            for j in range(4):
                leaf = eng.nodes[1 + j]
                leaf.T = eng.buf[(1 + j, "T")].numpy().reshape(4 * leaf.grid.nx, 4 * leaf.grid.nx).copy()
            eng.hps.merge4to1(nd, *[eng.nodes[c] for c in nd.children])
            for name in ("S", "T"):
                full = torch.from_numpy(np.ascontiguousarray(getattr(nd, name))).reshape(-1)
                if nranks > 1 and (name == "S" or gather_T):
                    cnt = full.numel() // nranks
                    keep = full[part_rank * cnt:(part_rank + 1) * cnt].clone()
                    full.zero_()
                    full[part_rank * cnt:(part_rank + 1) * cnt] = keep
                    allgather(full)
                setattr(nd, name, full.numpy().reshape(getattr(nd, name).shape).copy())

        # build
        local.build()
        n = mif.leaf_T(0).numel()
        gx.subtree_T_to_group(mif.big, n, lif.root_T)
        merge_partitioned(mid, gx.group_rank, gx.gs, gx.allgather_group, True)
        Tk = torch.from_numpy(np.ascontiguousarray(mid.nodes[0].T)).reshape(-1)
        own = Tk.clone()
        cnt = Tk.numel() // gx.gs      # only this rank's part of the group's map may travel: wipe the rest of the source
        Tk.zero_()
        Tk[gx.group_rank * cnt:(gx.group_rank + 1) * cnt] = own[gx.group_rank * cnt:(gx.group_rank + 1) * cnt]
        gx.level1_T_to_world(Tk, tif.big)
        assert torch.equal(tif.leaf_T(gx.group_id), own)
        merge_partitioned(top, rank, world, gx.allgather_world, True)
        # upwards
        local.upwards(P["f"])
        gx.subtree_h_to_group(mif.leaf_h, lif.root_h)
        mid.upwards(None)
        gx.level1_h_to_world(torch.from_numpy(np.ascontiguousarray(mid.nodes[0].h)), tif.leaf_h)
        top.upwards(None)
        # solve: no exchange - a rank descends through its own level-1 node only
        r, a, b = O.HPS.root_boundary(type("R", (), {"nodes": [top.nodes[0]]})(), lambda s_, xx, yy: (float(P["u"](xx, yy)), 1.0, 0.0))
        top.nodes[0].g = r / a
        top.solve()
        mid.nodes[0].g = tif.leaf_g(gx.group_id).numpy().copy()
        mid.solve()
        for k in plan.subtrees_of(rank):
            local.nodes[lif.idx[k]].g = mif.leaf_g(gx.local_index(k)).numpy().copy()
        local.solve()
        np.save(os.path.join(out_dir, "u_%d.npy" % rank), np.concatenate([nd.u for nd in local.nodes if nd.leaf]))
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array(plan.local_leaf_range(rank)))
        if rank == 0:
            np.save(os.path.join(out_dir, "T_root.npy"), top.nodes[0].T)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [4, 8])
def test_grouped_three_tier_run_matches_single_process_oracle(world, tmp_path):
    """world = 4 gs: groups of gs ranks per level-1 merge (gs = 2: sub-groups, group all-gathers and broadcasts)."""
    kw = CASES["uniform"]
    mp.spawn(_worker_grouped, args=(world, _free_port(), "uniform", str(tmp_path)), nprocs=world, join=True)
    ref = O.run(solver_kind="fishpack", **kw)
    u_ref = ref.leaf_solution()
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    lo_seen = 0
    for r in range(world):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert lo == lo_seen
        lo_seen = hi
        assert rel(np.load(tmp_path / ("u_%d.npy" % r)), np.concatenate(u_ref[lo:hi])) < 1e-11
    assert lo_seen == len(u_ref)
    assert rel(np.load(tmp_path / "T_root.npy"), ref.nodes[0].T) < 1e-11


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["root", "replicated", "replicated-one-allgather", "replicated-balanced", "replicated-cut1"])
@pytest.mark.parametrize("case", list(CASES))
def test_two_rank_sharded_run_matches_single_process_oracle(case, mode, tmp_path):
    kw = CASES[case]
    if mode == "root":
        mp.spawn(_worker, args=(2, _free_port(), case, str(tmp_path)), nprocs=2, join=True)
    else:
        mp.spawn(_worker_replicated, args=(2, _free_port(), case, str(tmp_path), mode == "replicated-one-allgather",
                                           "leaves" if mode == "replicated-balanced" else "count", 1 if mode == "replicated-cut1" else 2),
                 nprocs=2, join=True)
    ref = O.run(solver_kind="fishpack", **kw)
    u_ref = ref.leaf_solution()
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    lo_seen = 0
    for r in range(2):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert lo == lo_seen
        lo_seen = hi
        u = np.load(tmp_path / ("u_%d.npy" % r))
        assert rel(u, np.concatenate(u_ref[lo:hi])) < 1e-11
    assert lo_seen == len(u_ref)
    assert rel(np.load(tmp_path / "T_root.npy"), ref.nodes[0].T) < 1e-11


def test_shard_plan_tables():
    nodes = O.build_tree(O.refine_indicator(1.2), (0.0, 1.0, 0.0, 1.0), 8, 3, 3)
    level, child, box = _tables(nodes)
    for world in (1, 2, 4, 8, 16):
        plan = ShardPlan(level, child, box, 8, world)
        assert len(plan.cut_nodes) == 16
        seen, leaves = [], 0
        for r in range(world):
            ids, lev, ch, bx, roots = plan.local_table(r)
            assert len(roots) == 16 // world and np.all(lev[roots] == 2)
            assert np.all(ch[ch >= 0] < len(ids)) and np.all(ch[ch >= 0] > 0)
            seen.extend(ids.tolist())
            lo, hi = plan.local_leaf_range(r)
            assert lo == leaves
            leaves = hi
        assert leaves == 64 and sorted(seen) == [i for i in range(len(nodes)) if level[i] >= 2]
        tids, tlev, tch, tbox, ext = plan.top_table()
        assert len(tids) == 21 and list(ext) == [16] * 16 and plan.root_size() == 64
    with pytest.raises(ValueError):
        ShardPlan(level, child, box, 8, 3)
    # weighted dealing (adaptive trees): contiguous Morton blocks of whole subtrees, smallest possible maximum weight
    assert ShardPlan._contiguous_blocks(np.array([1, 1, 1, 1, 10, 1, 1, 1.0]), 3).tolist() == [0, 0, 0, 0, 1, 2, 2, 2]
    assert ShardPlan(level, child, box, 8, 4, balance="leaves").owner.tolist() == ShardPlan(level, child, box, 8, 4).owner.tolist()
    lop = O.build_tree(O.refine_box_indicator((-10.0, 0.5, -10.0, 0.5)), (-10.0, 10.0, -10.0, 10.0), 8, 2, 5)
    for world in (2, 3, 4):
        count = ShardPlan(*_tables(lop), 8, world if world != 3 else 4)
        for bal in ("leaves", "work"):
            plan = ShardPlan(*_tables(lop), 8, world, balance=bal)
            assert list(plan.owner) == sorted(plan.owner) and set(plan.owner) == set(range(world))      # contiguous, nobody idle
            load = [sum(plan.weight[k] for k in plan.subtrees_of(r)) for r in range(world)]
            if world != 3:
                even = [sum(plan.weight[k] for k in count.subtrees_of(r)) for r in range(world)]
                assert max(load) < max(even)                                                         # and better than equal counts
            ranges = [plan.local_leaf_range(r) for r in range(world)]
            assert ranges[0][0] == 0 and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    shallow = O.build_tree(O.refine_indicator(1.2), (0.0, 1.0, 0.0, 1.0), 8, 1, 1)
    with pytest.raises(ValueError):
        ShardPlan(*_tables(shallow), 8, 2)


def test_sharding_description_strings():
    """ShardedHPS.sharding() feeds bench.py's config line on every multi-GPU run: formatted without a device here."""
    from ellipticforest_b200.sharded import ShardedHPS
    nodes = O.build_tree(O.refine_box_indicator((-10.0, 0.5, -10.0, 0.5)), (-10.0, 10.0, -10.0, 10.0), 8, 2, 4)

    class Fake:
        pass
    for world, balance in ((2, "count"), (4, "count"), (8, "count"), (4, "work"), (3, "leaves")):
        for top_mode, p2p in (("replicated", True), ("replicated", False), ("root", False)):
            f = Fake()
            f.plan, f.world, f.top_mode, f.p2p = ShardPlan(*_tables(nodes), 8, world, balance=balance), world, top_mode, p2p
            text = ShardedHPS.sharding(f)
            assert "over %d GPUs" % world in text and "%" not in text, text
            assert ("balanced by " + balance in text) == (balance != "count"), text


def _worker_symmetry(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ellipticforest_b200.sharded import global_symmetry
        # rank 1's forest holds an internally coarsened subtree (efgpu_is_symmetric() == 0 there), rank 0's is uniform
        mixed = global_symmetry(dist, rank == 0, world)
        agree = global_symmetry(dist, True, world)
        g = dist.new_group([0, 1])
        sub = global_symmetry(dist, rank != 1, 2, group=g) if rank < 2 else True
        np.save(os.path.join(out_dir, "sym_%d.npy" % rank), np.array([mixed, agree, sub]))
    finally:
        dist.destroy_process_group()


def test_symmetric_plan_is_a_global_decision(tmp_path):
    """one rank with a uniform forest and one with an internally coarsened subtree of the same root size must
    agree on the upper tree's merge plan - the answer is the MIN over the ranks that feed the handle (world or sub-group)."""
    mp.spawn(_worker_symmetry, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    got = [np.load(tmp_path / ("sym_%d.npy" % r)) for r in range(3)]
    for r in range(3):
        assert not got[r][0] and got[r][1]
    assert not got[0][2] and not got[1][2] and got[2][2]
