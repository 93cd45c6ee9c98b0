"""GPU tests of the sharded path: forest handles (several subtree roots in one batched plan), the
upper-tree handle with external leaves, and the NCCL exchanges.  With world = 1 one GPU owns all 16
subtrees, which exercises every code path except the wire; the 2-rank NCCL test runs when the box has
two GPUs.  Results must equal the unsharded handle's (same kernels, same per-node arithmetic)."""
import os
import socket

import numpy as np
import pytest

import ellipticforest_b200 as ef
import hps_oracle as O
from ellipticforest_b200.sharded import ShardedHPS
from test_host import _mesh_for

pytestmark = pytest.mark.gpu

CASES = {
    "uniform_l3_m16": dict(problem_name="helmholtz", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=3, max_level=3,
                           threshold=1.2, refine_box=None),
    "adaptive_l2_4_m8": dict(problem_name="poisson", solver_kind="fishpack", box=(-10.0, 10.0, -10.0, 10.0), nx=8, min_level=2, max_level=4,
                             threshold=1.2, refine_box=(-10.0, 0.5, -10.0, 0.5)),
}


# two GPUs only (not run by the single-GPU parts of this file): variable-coefficient leaves below the cut - the subtree roots'
# DtN maps are not signed-symmetric, so the upper tree takes the general plan - on a lopsided tree dealt by merge work
EXTRA_CASES = {
    "adaptive_l2_4_m8_varcoef": dict(problem_name="varcoef", solver_kind="fivepoint", box=(-10.0, 10.0, -10.0, 10.0), nx=8, min_level=2, max_level=4,
                                     threshold=1.2, refine_box=(-10.0, 0.5, -10.0, 0.5)),
}
# two GPUs only: large enough (root child side n = 256, X of order 1024) that, with EFGPU_SPLIT_MIN_ROWS=256, the products of the
# block inversion are split by rows too - in peer mode their slices are stored into the other rank's arena by the GEMM epilogue
EXTRA_CASES["uniform_l4_m16_split"] = dict(problem_name="poisson", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=4,
                                           max_level=4, threshold=1.2, refine_box=None)
ALL_CASES = dict(CASES, **EXTRA_CASES)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _solver(P, kind="fishpack"):
    s = ef.FiniteVolumeSolver()
    s.solver_type = "FISHPACK90" if kind == "fishpack" else "FivePointStencil"
    s.alpha_function, s.beta_function, s.lambda_function = P["alpha"], P["beta"], P["lam"]
    return s


def _run_sharded(kw, rank, world, device, top_mode="replicated", balance="count", cut=2):
    import torch
    P = O.problem(kw["problem_name"])
    m = _mesh_for(kw)
    hps = ShardedHPS(m, _solver(P, kw["solver_kind"]), device=device, rank=rank, world=world, top_mode=top_mode, balance=balance, cut=cut)
    f, g = hps.sample_inputs(P["f"], P["u"])
    f_dev = torch.from_numpy(f).cuda(device)
    g_dev = torch.from_numpy(g).cuda(device)
    u_dev = torch.empty_like(f_dev)
    hps.buildStage()
    hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=True)
    hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True)
    rootT = hps.gather_root_T().cpu().numpy() if hps.top is not None else None
    return hps, u_dev.cpu().numpy(), rootT


def _run_single(kw):
    P = O.problem(kw["problem_name"])
    m = _mesh_for(kw)
    hps = ef.HPSAlgorithm(m, _solver(P, kw["solver_kind"]))
    hps.buildStage()
    hps.upwardsStage(P["f"])
    hps.solveStage(lambda side, x, y: (P["u"](x, y), 1.0, 0.0))
    return hps


@pytest.mark.parametrize("top_mode", ["replicated", "root"])
@pytest.mark.parametrize("case", list(CASES))
def test_forest_and_upper_tree_on_one_gpu(case, top_mode):
    kw = CASES[case]
    single = _run_single(kw)
    hps, u, rootT = _run_sharded(kw, 0, 1, 0, top_mode)
    assert relerr(u, single.u_leaves.reshape(-1)) < 1e-12
    assert relerr(rootT, single.operator(0, "T").reshape(-1)) < 1e-12
    # and against the oracle, like every other parity test
    ora = O.run(**kw)
    assert relerr(u, np.concatenate(ora.leaf_solution())) < 1e-10


def _worker(rank, world, port, case, out_dir, top_mode, balance="count", cut=2):
    import faulthandler
    faulthandler.enable()
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if case.endswith("_split"):
        os.environ["EFGPU_SPLIT_MIN_ROWS"] = "256"
    if top_mode == "replicated-nccl":      # round 1's exchange: ncclAllGather from a host callback instead of peer-mapped arenas
        os.environ["EFGPU_P2P"] = "0"
        top_mode = "replicated"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        hps, u, rootT = _run_sharded(ALL_CASES[case], rank, world, rank, top_mode, balance, cut)
        np.save(os.path.join(out_dir, "u_%d.npy" % rank), u)
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array([hps.leaf_lo, hps.leaf_hi]))
        if rootT is not None and rank == 0:
            np.save(os.path.join(out_dir, "T_root.npy"), rootT)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("top_mode", ["replicated", "replicated-nccl", "root"])
@pytest.mark.parametrize("case", list(CASES))
def test_two_gpus_over_nccl(case, top_mode, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path), top_mode), nprocs=2, join=True)
    single = _run_single(CASES[case])
    u_ref = single.u_leaves.reshape(single.mesh.n_leaves, -1)
    for r in range(2):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert relerr(np.load(tmp_path / ("u_%d.npy" % r)), u_ref[lo:hi].reshape(-1)) < 1e-12
    assert relerr(np.load(tmp_path / "T_root.npy"), single.operator(0, "T").reshape(-1)) < 1e-12


@pytest.mark.parametrize("top_mode", ["replicated", "replicated-nccl"])
def test_two_gpus_row_split_inversion_products(top_mode, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = "uniform_l4_m16_split"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path), top_mode), nprocs=2, join=True)
    single = _run_single(ALL_CASES[case])
    u_ref = single.u_leaves.reshape(single.mesh.n_leaves, -1)
    for r in range(2):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert relerr(np.load(tmp_path / ("u_%d.npy" % r)), u_ref[lo:hi].reshape(-1)) < 1e-12
    assert relerr(np.load(tmp_path / "T_root.npy"), single.operator(0, "T").reshape(-1)) < 1e-12


@pytest.mark.parametrize("top_mode", ["replicated", "root"])
def test_two_gpus_variable_coefficients_balanced_by_work(top_mode, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = "adaptive_l2_4_m8_varcoef"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path), top_mode, "work"), nprocs=2, join=True)
    single = _run_single(ALL_CASES[case])
    u_ref = single.u_leaves.reshape(single.mesh.n_leaves, -1)
    ranges = [tuple(np.load(tmp_path / ("range_%d.npy" % r))) for r in range(2)]
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == single.mesh.n_leaves
    for r, (lo, hi) in enumerate(ranges):
        assert relerr(np.load(tmp_path / ("u_%d.npy" % r)), u_ref[lo:hi].reshape(-1)) < 1e-10
    assert relerr(np.load(tmp_path / "T_root.npy"), single.operator(0, "T").reshape(-1)) < 1e-10


@pytest.mark.parametrize("top_mode", ["replicated", "root"])
@pytest.mark.parametrize("case", list(CASES))
def test_cut_at_level_1_on_one_gpu(case, top_mode):
    """cut = 1: four subtrees; the level-1 merges belong to the forests and the upper tree is the root merge alone."""
    kw = CASES[case]
    single = _run_single(kw)
    hps, u, rootT = _run_sharded(kw, 0, 1, 0, top_mode, cut=1)
    assert relerr(u, single.u_leaves.reshape(-1)) < 1e-12
    assert relerr(rootT, single.operator(0, "T").reshape(-1)) < 1e-12


@pytest.mark.parametrize("case", list(CASES))
def test_cut_at_level_1_on_two_gpus(case, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path), "replicated", "count", 1), nprocs=2, join=True)
    single = _run_single(CASES[case])
    u_ref = single.u_leaves.reshape(single.mesh.n_leaves, -1)
    for r in range(2):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert relerr(np.load(tmp_path / ("u_%d.npy" % r)), u_ref[lo:hi].reshape(-1)) < 1e-12
    assert relerr(np.load(tmp_path / "T_root.npy"), single.operator(0, "T").reshape(-1)) < 1e-12


def _worker_grouped(rank, world, port, case, out_dir):
    import torch
    import torch.distributed as dist
    from ellipticforest_b200.sharded import GroupedShardedHPS
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        kw = CASES[case]
        P = O.problem(kw["problem_name"])
        hps = GroupedShardedHPS(_mesh_for(kw), _solver(P, kw["solver_kind"]), device=rank, rank=rank, world=world)
        f, g = hps.sample_inputs(P["f"], P["u"])
        f_dev, g_dev = torch.from_numpy(f).cuda(rank), torch.from_numpy(g).cuda(rank)
        u_dev = torch.empty_like(f_dev)
        hps.buildStage()
        hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=True)
        hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True)
        rootT = hps.gather_root_T().cpu().numpy()
        np.save(os.path.join(out_dir, "u_%d.npy" % rank), u_dev.cpu().numpy())
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array([hps.leaf_lo, hps.leaf_hi]))
        if rank == 0:
            np.save(os.path.join(out_dir, "T_root.npy"), rootT)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [4, 8])
def test_grouped_level_1_merges(world, tmp_path):
    """GroupedShardedHPS: forests, level-1 merges inside rank groups (1 rank at 4 GPUs, 2 at 8), root merge over all ranks."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    case = "uniform_l3_m16"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_grouped, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    single = _run_single(CASES[case])
    u_ref = single.u_leaves.reshape(single.mesh.n_leaves, -1)
    for r in range(world):
        lo, hi = np.load(tmp_path / ("range_%d.npy" % r))
        assert relerr(np.load(tmp_path / ("u_%d.npy" % r)), u_ref[lo:hi].reshape(-1)) < 1e-12
    assert relerr(np.load(tmp_path / "T_root.npy"), single.operator(0, "T").reshape(-1)) < 1e-12
