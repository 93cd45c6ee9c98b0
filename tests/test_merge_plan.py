"""CPU tests of the merge planner (host logic of libefgpu.so, no device): the step / descriptor lists that
efgpu_build would launch are serialised by efgpu_debug_merge_plan, interpreted here with numpy (one state
per rank, all-gathers included) and the resulting X^-1, S, T compared with the oracle's merge4to1
(reference src/HPSAlgorithm.hpp:497-518) on the same children."""
import ctypes as C

import numpy as np
import pytest

from ellipticforest_b200 import _lib
import hps_oracle as O

OP_TC0, OP_XINV, OP_S, OP_T, OP_W1, OP_W2, OP_W3 = 0, 4, 5, 6, 7, 8, 9
CLS_GEMM_S, CLS_GEMM_T, CLS_MIRROR_T = 5, 6, 14


@pytest.fixture
def whole_diagonal_blocks():
    """Plans with tuning key 5 off (every diagonal block of T as one product): what the block-counting tests below describe."""
    lib = _lib.load()
    assert lib.efgpu_set_tuning(5, 0) == 0
    yield
    assert lib.efgpu_set_tuning(5, 1) == 0


def get_plan(n, level, rank, nranks, sym, peer=0):
    lib = _lib.load()
    ns, nb, nt = C.c_int(), C.c_int(), C.c_int()
    ws = np.zeros(3, dtype=np.int64)
    args = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    assert lib.efgpu_debug_merge_plan_ex(n, level, rank, nranks, sym, peer, None, C.byref(ns), None, None, C.byref(nb), None, C.byref(nt), args(ws)) == 0
    steps = np.zeros((ns.value, 16), dtype=np.int64)
    blocks = np.zeros((nb.value, 16), dtype=np.int64)
    terms = np.zeros((nb.value, 2, 8), dtype=np.int64)
    trans = np.zeros((max(nt.value, 1), 16), dtype=np.int64)
    assert lib.efgpu_debug_merge_plan_ex(n, level, rank, nranks, sym, peer, args(steps), C.byref(ns), args(blocks), args(terms), C.byref(nb),
                                         args(trans), C.byref(nt), args(ws)) == 0
    return steps, blocks, terms, trans[:nt.value], ws


def view(buf, off, ld, rows, cols):
    """rows x cols window of a flat row-major buffer (as_strided: writable, no copy)."""
    return np.lib.stride_tricks.as_strided(buf[off:], shape=(rows, cols), strides=(ld * 8, 8))


class RankState:
    def __init__(self, n, Tc, X, ws):
        self.ops = {OP_TC0 + c: np.ascontiguousarray(Tc[c], dtype=np.float64).reshape(-1).copy() for c in range(4)}
        self.ops[OP_XINV] = np.ascontiguousarray(X).reshape(-1).copy()
        self.ops[OP_S] = np.full(32 * n * n, np.nan)
        self.ops[OP_T] = np.full(64 * n * n, np.nan)
        for k, op in enumerate((OP_W1, OP_W2, OP_W3)):
            self.ops[op] = np.full(int(ws[k]) + 1, np.nan)

    def gemm(self, blk, trm, targets=None):
        """targets: the states that receive the result (peer mode: the epilogue stores into every rank's arena); default: this one"""
        c_op, c0_op, ldc, ldc0, c_off, c0_off, rows, cols, nterms = (int(v) for v in blk[:9])
        acc = np.zeros((rows, cols))
        if c0_op >= 0:
            acc += view(self.ops[c0_op], c0_off, ldc0, rows, cols)
        for t in range(nterms):
            a_op, b_op, lda, ldb, a_off, b_off, K, neg = (int(v) for v in trm[t])
            p = view(self.ops[a_op], a_off, lda, rows, K) @ view(self.ops[b_op], b_off, ldb, K, cols)
            acc += -p if neg else p
        ct_op1, ldct, ct_off, ct_neg = (int(v) for v in blk[9:13])
        for st in (targets if targets is not None else [self]):
            view(st.ops[c_op], c_off, ldc, rows, cols)[...] = acc
            if ct_op1:               # second destination: the signed transpose of the block, from the same epilogue
                view(st.ops[ct_op1 - 1], ct_off, ldct, cols, rows)[...] = (-acc if ct_neg else acc).T

    def run_step(self, st, blocks, terms, trans, n, targets=None):
        kind, first, count, off, N = (int(v) for v in st[:5])
        if kind == 0:
            for o in [off] + ([int(st[12])] if int(st[12]) >= 0 else []):    # a launch may invert two independent blocks
                v = view(self.ops[OP_XINV], o, 4 * n, N, N)
                v[...] = np.linalg.inv(v.copy())
        elif kind == 1:
            results = []   # one launch: every block reads the state before the launch
            for k in range(first, first + count):
                self.gemm(blocks[k], terms[k], targets)
        else:
            for k in range(first, first + count):
                s_op, d_op, lds, ldd, s_off, d_off, rows, cols, neg = (int(v) for v in trans[k][:9])
                src = view(self.ops[s_op], s_off, lds, rows, cols).copy()
                view(self.ops[d_op], d_off, ldd, cols, rows)[...] = (-src if neg else src).T


def allgather(states, op, off, doubles):
    nr = len(states)
    per = doubles // nr
    full = np.concatenate([states[r].ops[op][off + r * per: off + (r + 1) * per] for r in range(nr)])
    for s in states:
        s.ops[op][off: off + doubles] = full


def emulate(n, level, nranks, sym, Tc, X, peer=0):
    """peer = 1: the plan of a peer-mapped tree - the GEMMs of split products (gk 3), of S and of the DtN maps below the root
    store their results (and transposed second destinations) into EVERY rank's state, nothing is gathered afterwards."""
    plans = [get_plan(n, level, r, nranks, sym, peer) for r in range(nranks)]
    states = [RankState(n, Tc, X, plans[r][4]) for r in range(nranks)]
    nsteps = len(plans[0][0])
    assert all(len(p[0]) == nsteps for p in plans)

    colsplit = bool(int(plans[0][0][0][13]))     # peer-mapped trees over 2 / 4 / 8 ranks: S and T partitioned by block columns

    def run(i):
        st = plans[0][0][i]
        scatter = peer and nranks > 1 and int(st[0]) == 1 and (int(st[6]) == 3 or int(st[5]) == CLS_GEMM_S or (int(st[5]) == CLS_GEMM_T and level > 0))
        if scatter and colsplit and int(st[5]) == CLS_GEMM_S:
            # column partition: no barrier between S and T - every rank sees only its OWN columns of S (the rest of its S is still
            # NaN here) until the barrier after T; a T block that read another rank's columns would turn into NaN
            scatter = False
        if scatter:      # all ranks read the state before the launch (flag barrier), then store into every arena
            frozen = [RankState.__new__(RankState) for _ in range(nranks)]
            for r in range(nranks):
                frozen[r].ops = {k: v.copy() for k, v in states[r].ops.items()}
            for r in range(nranks):
                steps, blocks, terms, trans, _ = plans[r]
                frozen[r].run_step(steps[i], blocks, terms, trans, n, targets=states)
            return
        for r in range(nranks):
            steps, blocks, terms, trans, _ = plans[r]
            states[r].run_step(steps[i], blocks, terms, trans, n)
        gk, g_op, g_rows, g_cols, g_ld, g_off = (int(v) for v in st[6:12])
        if gk == 1:
            allgather(states, g_op, g_off, g_rows * g_cols)
        elif gk == 2:
            allgather(states, OP_W3, 0, g_rows * g_cols)
            for s in states:
                view(s.ops[g_op], g_off, g_ld, g_rows, g_cols)[...] = s.ops[OP_W3][:g_rows * g_cols].reshape(g_rows, g_cols)

    cls = [int(plans[0][0][i][5]) for i in range(nsteps)]
    for i in range(nsteps):          # phase 0: everything but T
        if cls[i] not in (CLS_GEMM_T, CLS_MIRROR_T):
            run(i)
    if nranks > 1 and not peer:
        allgather(states, OP_S, 0, 32 * n * n)
    for i in range(nsteps):          # phase 1
        if cls[i] == CLS_GEMM_T:
            run(i)
    if nranks > 1 and peer and colsplit:          # the barrier after T completes S: every rank's block columns are everywhere now
        w = 8 * n // nranks
        full = np.concatenate([view(states[r].ops[OP_S], r * w, 8 * n, 4 * n, w) for r in range(nranks)], axis=1)
        for s in states:
            s.ops[OP_S][:] = full.reshape(-1)
    if nranks > 1 and peer and level == 0 and colsplit:   # efgpu_complete_root_dtn of a column partition: every rank's block columns to everybody
        w = 8 * n // nranks
        full = np.concatenate([view(states[r].ops[OP_T], r * w, 8 * n, 8 * n, w) for r in range(nranks)], axis=1)
        for s in states:
            s.ops[OP_T][:] = full.reshape(-1)
    elif nranks > 1 and (not peer or level == 0):   # level 0: efgpu_complete_root_dtn (the root's map is gathered and mirrored only on demand)
        allgather(states, OP_T, 0, 64 * n * n)
    for i in range(nsteps):
        if cls[i] == CLS_MIRROR_T:
            run(i)
    flops = 0
    for steps, blocks, terms, trans, _ in plans:
        for st in steps:
            if int(st[0]) == 1:
                for k in range(int(st[1]), int(st[1]) + int(st[2])):
                    for t in range(int(blocks[k][8])):
                        flops += 2 * int(blocks[k][6]) * int(blocks[k][7]) * int(terms[k][t][6])
    return states, flops


def oracle_merge(Tc, n):
    nodes = [O.Node(path="0", level=0, grid=O.Grid(2 * n, 0.0, 2.0, 0.0, 2.0), leaf=False)]
    for c in range(4):
        nd = O.Node(path="0%d" % c, level=1, grid=O.Grid(n, (c & 1) * 1.0, (c & 1) + 1.0, (c >> 1) * 1.0, (c >> 1) + 1.0), leaf=True)
        nd.T = Tc[c].copy()
        nodes.append(nd)
    hps = O.HPS(nodes, None)
    hps.merge4to1(*nodes)
    return nodes[0]


def uniform_children(n_leaf, depth, name="poisson"):
    """Four identical-size DtN maps of uniform subtrees (child side n = n_leaf * 2^depth) from the oracle."""
    r = O.run(problem_name=name, solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=n_leaf, min_level=depth + 1,
              max_level=depth + 1)
    root = r.nodes[0]
    return [r.nodes[c].T for c in root.children], root


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


_CHILDREN = {}


@pytest.mark.parametrize("depth", [2, 3])
@pytest.mark.parametrize("nranks", [1, 2])
@pytest.mark.parametrize("sym", [0, 1])
def test_plan_reproduces_oracle_merge_uniform(sym, nranks, depth):
    # n = 64: X is 256 x 256, blocked inversion with a structured top level and register-resident base cases;
    # n = 128: two recursion levels, and with 2 ranks the 256-row products are split by rows and all-gathered
    if depth not in _CHILDREN:
        _CHILDREN[depth] = uniform_children(16, depth)
    Tc, root = _CHILDREN[depth]
    n = 16 << depth
    for level in ((0, 1) if nranks > 1 else (0,)):
        states, flops = emulate(n, level, nranks, sym, Tc, root.X)
        for s in states:
            assert rel(view(s.ops[OP_XINV], 0, 4 * n, 4 * n, 4 * n), np.linalg.inv(root.X)) < 1e-11
            assert rel(s.ops[OP_S].reshape(4 * n, 8 * n), root.S) < 1e-11
        for s in states:
            assert rel(s.ops[OP_T].reshape(8 * n, 8 * n), root.T) < 1e-11
        n3 = float(n) ** 3
        if nranks == 1:                      # (small replicated products are issued by every rank of a partition)
            assert flops < 400 * n3 if sym else flops <= 484 * n3   # symmetric plan: well below the general 484 n^3


@pytest.mark.parametrize("sym,cluster", [(0, 0), (1, 0), (1, 1)])
def test_zipped_diagonal_subinversions(sym, cluster):
    """n = 256: X is 1024 x 1024, the two 256 x 256 diagonal blocks of its leading block recurse (products inside), and
    their step lists are zipped into launches that carry both blocks (second branch on its own workspace slots).  With tuning
    key 9 (off by default) a batch of at most four merges stops the recursion at 256 x 256 blocks (cluster-cooperative base case):
    the pair is then ONE paired base-case launch and the Schur complement two more 256-blocks."""
    if 4 not in _CHILDREN:
        _CHILDREN[4] = uniform_children(16, 4)
    Tc, root = _CHILDREN[4]
    n = 256
    lib = _lib.load()
    assert lib.efgpu_set_tuning(9, cluster) == 0
    try:
        steps, blocks, terms, trans, ws = get_plan(n, 0, 0, 1, sym)
        paired = sum(1 for st in steps if int(st[0]) == 0 and int(st[12]) >= 0)
        if cluster:
            assert paired == 1 and sum(1 for st in steps if int(st[0]) == 0) == 3
        else:
            assert paired == 2                                                                   # two paired base-case launches
            assert any(int(st[0]) == 1 and int(st[2]) == 2 and int(st[5]) == 4 for st in steps)  # paired products of the inversion
        states, flops = emulate(n, 0, 1, sym, Tc, root.X)
    finally:
        assert lib.efgpu_set_tuning(9, 0) == 0
    s = states[0]
    assert rel(view(s.ops[OP_XINV], 0, 4 * n, 4 * n, 4 * n), np.linalg.inv(root.X)) < 1e-11
    assert rel(s.ops[OP_S].reshape(4 * n, 8 * n), root.S) < 1e-11
    assert rel(s.ops[OP_T].reshape(8 * n, 8 * n), root.T) < 1e-11


@pytest.mark.parametrize("nranks", [1, 2])
def test_symmetric_diagonal_blocks_of_T(nranks):
    """Tuning key 5 (default on since round 2): the eight diagonal n x n blocks of the signed-symmetric T are symmetric themselves;
    with the knob on they are multiplied as upper-triangular sub-blocks (n = 256: 2 x 2 of 128) and completed by transposes in
    the mirror step."""
    from ellipticforest_b200 import _lib
    lib = _lib.load()
    if 4 not in _CHILDREN:
        _CHILDREN[4] = uniform_children(16, 4)
    Tc, root = _CHILDREN[4]
    n = 256
    assert lib.efgpu_set_tuning(5, 0) == 0
    base = get_plan(n, 1, 0, nranks, 1)
    assert lib.efgpu_set_tuning(5, 1) == 0
    try:
        split = get_plan(n, 1, 0, nranks, 1)
        # one mirrored sub-block per diagonal block: a transpose step on a partitioned tree, a second (transposed) destination
        # of the producing block otherwise
        fused = lambda plan: int(np.sum(plan[1][:, 9] != 0))
        if nranks > 1:
            assert len(split[3]) == len(base[3]) + 8
        else:
            assert len(split[3]) == len(base[3]) == 0 and fused(split) == fused(base) + 8
        for level in ((0, 1) if nranks > 1 else (0,)):
            states, flops = emulate(n, level, nranks, 1, Tc, root.X)
            for s in states:
                assert rel(s.ops[OP_T].reshape(8 * n, 8 * n), root.T) < 1e-11
                assert rel(s.ops[OP_S].reshape(4 * n, 8 * n), root.S) < 1e-11
    finally:
        lib.efgpu_set_tuning(5, 1)


@pytest.mark.parametrize("n,nranks,split_min,plans", [(256, 4, 256, "1,0"), (512, 8, 256, "1")])
def test_partitions_of_4_and_8_ranks_with_row_split_inversion(n, nranks, split_min, plans):
    """The shapes of the 4- and 8-GPU runs: products of the block inversion split by rows over the ranks (512 rows over 4,
    1024 rows over 8) with in-place and staged all-gathers, S / T row slices of n / 2 and n rows, tree levels 1 (T gathered
    after the merge) and 0 (root: gathered and mirrored on demand).  Own process: the split threshold is read once."""
    import os
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "plan_partition_check.py"),
                        str(n), str(nranks), str(split_min), plans], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("n ")]
    assert len(lines) == 2 * len(plans.split(",")), p.stdout
    for l in lines:
        assert int(l.split("gathers")[1].split()[0]) >= 3, l      # the inversion's products really were split


def test_general_plan_on_nonsymmetric_children():
    rng = np.random.default_rng(0)
    n = 40                                       # N = 160: h = 80, q = 40; odd sizes for the transposes / tiles
    Tc, _ = uniform_children(8, 0)
    scale = np.max(np.abs(Tc[0]))
    Tc = []
    for c in range(4):
        A = rng.standard_normal((4 * n, 4 * n)) * 0.05 * scale / np.sqrt(n)
        d = np.ones(4 * n); d[:n] = -1; d[2 * n:3 * n] = -1
        Tc.append(A + np.diag(d) * scale)       # diagonally dominant with the sign pattern of a DtN map
    ref = oracle_merge(Tc, n)
    states, flops = emulate(n, 0, 1, 0, Tc, ref.X)
    s = states[0]
    assert rel(view(s.ops[OP_XINV], 0, 4 * n, 4 * n, 4 * n), np.linalg.inv(ref.X)) < 1e-11
    assert rel(s.ops[OP_S].reshape(4 * n, 8 * n), ref.S) < 1e-11
    assert rel(s.ops[OP_T].reshape(8 * n, 8 * n), ref.T) < 1e-11
    assert flops <= 484 * n ** 3                 # 100 (structured inverse) + 128 (S) + 256 (T)


def test_symmetric_plan_flop_count_and_balance(whole_diagonal_blocks):
    n = 256
    steps, blocks, terms, trans, ws = get_plan(n, 1, 0, 1, 1)
    assert len(trans) == 0          # unpartitioned: no transpose launches at all, the mirrored blocks ride on the epilogues
    # every off-diagonal block of T is either computed or mirrored (second, transposed destination of its partner), never both
    tsteps = [st for st in steps if int(st[5]) == CLS_GEMM_T]
    computed, mirrored = set(), set()
    for st in tsteps:
        for k in range(int(st[1]), int(st[1]) + int(st[2])):
            off = int(blocks[k][4]); computed.add((off // (8 * n) // n, off % (8 * n) // n))
            if int(blocks[k][9]):
                assert int(blocks[k][9]) - 1 == OP_T
                toff = int(blocks[k][11]); mirrored.add((toff // (8 * n) // n, toff % (8 * n) // n))
    assert len(mirrored) == 28
    steps2, blocks2, terms2, trans2, ws2 = get_plan(n, 1, 0, 2, 1)      # a partitioned tree keeps the 28 transposes
    assert len([t for t in trans2 if int(t[0]) == OP_T]) == 28
    assert len(computed) == 36 and not (computed & mirrored) and len(computed | mirrored) == 64
    for (p, q) in mirrored:
        assert (q, p) in computed
    rows = [sum(1 for (p, q) in computed if p == r) for r in range(8)]
    assert max(rows) - min(rows) <= 1


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_T_products_are_balanced_over_the_ranks(nranks, whole_diagonal_blocks):
    """Symmetric plan: 36 block products of n^3 x 4 flops; halves / quarters of the rows carry 18 / 9 of them, and with one
    block row per rank (8 ranks) the opposite pairs are shared half and half: 4.5 each."""
    n = 1024
    per_rank = []
    for r in range(nranks):
        steps, blocks, terms, trans, ws = get_plan(n, 1, r, nranks, 1)
        f = 0
        for st in steps:
            if int(st[0]) == 1 and int(st[5]) == CLS_GEMM_T:
                for k in range(int(st[1]), int(st[1]) + int(st[2])):
                    f += sum(2 * int(blocks[k][6]) * int(blocks[k][7]) * int(terms[k][t][6]) for t in range(int(blocks[k][8])))
        per_rank.append(f / n ** 3)
    assert per_rank == [144.0 / nranks] * nranks, per_rank


@pytest.mark.parametrize("nranks,n,level,sym,cols", [(2, 256, 0, 1, 0), (4, 256, 1, 1, 0), (8, 256, 0, 1, 0), (8, 256, 1, 1, 0),
                                                     (2, 256, 1, 0, 1), (8, 256, 0, 0, 1), (4, 256, 0, 1, 1), (8, 256, 1, 1, 1)])
def test_peer_plan_reproduces_oracle_merge(nranks, n, level, sym, cols, monkeypatch):
    """The plan of a peer-mapped tree (efgpu_peer_export): products of the inversion split down to 64-row slices per rank (here
    forced with EFGPU_SPLIT_MIN_ROWS), stored into every rank's arena together with their fused transposes; 8 ranks: one block
    row of T per rank with the opposite pairs shared half and half.  Every rank must end with the oracle's X^-1, S and T."""
    depth = {256: 4, 512: 5}[n]
    if depth not in _CHILDREN:
        _CHILDREN[depth] = uniform_children(16, depth)
    Tc, root = _CHILDREN[depth]
    monkeypatch.setenv("EFGPU_SPLIT_MIN_ROWS", "256")        # read whenever a plan is made
    lib = _lib.load()
    assert lib.efgpu_set_tuning(11, cols) == 0              # cols = 1: S and T partitioned by block columns (off by default)
    try:
        states, flops = emulate(n, level, nranks, sym, Tc, root.X, peer=1)
        plan0 = get_plan(n, level, 0, nranks, sym, 1)[0]
    finally:
        assert lib.efgpu_set_tuning(11, 0) == 0
    assert int(plan0[0][13]) == cols
    split_steps = sum(1 for st in plan0 if int(st[6]) == 3)
    for s in states:
        assert rel(view(s.ops[OP_XINV], 0, 4 * n, 4 * n, 4 * n), np.linalg.inv(root.X)) < 1e-11
        assert rel(s.ops[OP_S].reshape(4 * n, 8 * n), root.S) < 1e-11
        assert rel(s.ops[OP_T].reshape(8 * n, 8 * n), root.T) < 1e-11
    assert split_steps > 0
    print("peer plan n=%d ranks=%d level=%d: %d split products" % (n, nranks, level, split_steps))


@pytest.mark.parametrize("n,level,nranks,peer", [(512, 0, 1, 0), (2048, 0, 1, 0), (1024, 1, 4, 1), (2048, 0, 8, 1), (128, 3, 1, 0), (64, 4, 1, 0)])
@pytest.mark.parametrize("sym", [0, 1])
def test_tma_views_address_the_same_operands(n, level, nranks, peer, sym):
    """plan_tma (operand staging by TMA, csrc/gemm_tma.cu): every operand block of a step marked for the TMA kernel must be reachable as
    (view origin) + row * ld + col = the descriptor's element offset, inside one row of its view (no wrap), with whole 128 x 64 tiles
    and K a multiple of 16; the W1 workspace slots of the recursion get their own origins; steps with any other block stay on cp.async."""
    lib = _lib.load()
    for rank in sorted({0, nranks - 1}):
        steps, blocks, terms, trans, ws = get_plan(n, level, rank, nranks, sym, peer)
        nv = C.c_int()
        assert lib.efgpu_debug_tma_plan(n, level, rank, nranks, sym, peer, None, C.byref(nv), None, None) == 0
        views = np.zeros((max(nv.value, 1), 3), dtype=np.int64)
        tb = np.full((len(blocks), 2, 6), -7, dtype=np.int64)
        flags = np.zeros(len(steps), dtype=np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.efgpu_debug_tma_plan(n, level, rank, nranks, sym, peer, p(views), C.byref(nv), p(tb), p(flags)) == 0
        views = views[:nv.value]
        n_tma = 0
        for st, fl in zip(steps, flags):
            if int(st[0]) != 1:
                assert fl == 0
                continue
            first, count = int(st[1]), int(st[2])
            ok_all = count > 0
            for k in range(first, first + count):
                rows, cols, nterms = int(blocks[k][6]), int(blocks[k][7]), int(blocks[k][8])
                valid = tb[k][0][0] >= 0
                if valid:
                    assert rows % 128 == 0 and cols % 64 == 0
                    for t in range(nterms):
                        a_op, b_op, lda, ldb, a_off, b_off, K, _ = (int(v) for v in terms[k][t])
                        av, bv, ar, ac, br, bc = (int(v) for v in tb[k][t])
                        assert K % 16 == 0
                        assert tuple(views[av][[0, 2]]) == (a_op, lda) and tuple(views[bv][[0, 2]]) == (b_op, ldb)
                        assert int(views[av][1]) + ar * lda + ac == a_off and ac + K <= lda
                        assert int(views[bv][1]) + br * ldb + bc == b_off and bc + cols <= ldb
                ok_all = ok_all and valid
            assert bool(fl) == ok_all
            n_tma += int(fl)
        if n >= 512:
            assert n_tma > 0            # the large products of such a merge do run on TMA-staged operands
        if n <= 64:
            assert nv.value == 0 or n_tma >= 0
        # W1 views: one origin per recursion depth in use, all inside the W1 workspace
        for op, origin, ld in views:
            assert origin >= 0 and (op != OP_W1 or origin < ws[0])
            assert op == OP_W1 or origin == 0
