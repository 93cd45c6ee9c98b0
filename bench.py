#!/usr/bin/env python
"""Benchmark of the HPS hot path (BASELINE.json metric: HPS build & solve DOFs/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--level L] [--nx M] [--problem poisson|helmholtz] [--no-cpu-baseline]

A "step" is one full pass of the hot path over the workload's quadtree: buildStage (leaf DtN +
every 4-to-1 merge), upwardsStage and solveStage for one right-hand side.  DOFs = leaves * nx * ny
(reference: examples/elliptic-multiple/main.cpp:374).  The default workload is BASELINE.json
configs[1]: uniform level-8 quadtree of 16x16 finite-volume patches, constant-coefficient Poisson
on [0, pi]^2, f = -(sin x + sin y), Dirichlet data from u = sin x + sin y.

Legs of the JSON line:
  value        DOFs/s with f and the root Dirichlet data already resident in HBM, device events.
  e2e          the same step through the C-ABI with HOST buffers (efgpu_build, efgpu_upwards,
               efgpu_solve_dirichlet): f and g copied H2D from pinned memory, u copied D2H, every step.
  roofline     the dominant kernel (the FP64 tensor-core batched GEMM of the merges) timed with
               CUDA events on the library's stream around every launch, over a second pass of the same
               K steps (the events cost 3-10 % in the launch-bound top levels); numerator = flops issued.
  cpu_baseline the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from /root/reference by
               oracle/Makefile) on this box's host cores, on a bounded sample of the same workload.
`--impl reference` times only that reference build (bounded sample per step) and prints the same line.

Multi-GPU (torchrun, one rank per GPU): the quadtree is sharded by level-2 subtree;
timing is barrier + synchronize on both sides and the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
PI = 3.141592653589793


# ---------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--nx", type=int, default=16)
    ap.add_argument("--problem", default="poisson", choices=["poisson", "helmholtz", "varcoef"],
                    help="varcoef: alpha = 1, beta = 1 + sin x cos y / 2, lambda = -(1 + cos x cos y / 2), FivePointStencil leaves (BASELINE configs[3])")
    ap.add_argument("--adaptive", nargs=2, type=int, metavar=("MIN", "MAX"), default=None,
                    help="adaptive mesh of examples/elliptic-single (refine where |sin x + sin y| > --threshold at a cell centre, 2:1 balanced) "
                         "on [-10,10]^2: BASELINE configs[0] is --adaptive 0 7, configs[3] --adaptive 4 9 --threshold 1.6 --problem varcoef")
    ap.add_argument("--threshold", type=float, default=1.2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-symmetry", action="store_true", help="EFGPU_NO_SYMMETRY: general merge plan (A/B against the symmetric one)")
    ap.add_argument("--profile-run", action="store_true", help="for ncu: exactly --warmup/--steps device steps, nothing else, no JSON")
    ap.add_argument("--lean-T", action="store_true", help="EFGPU_LEAN_T memory policy (single GPU): interior DtN maps in a transient arena")
    ap.add_argument("--n-solves", type=int, default=0, help="BASELINE configs[2] pattern: after the timed steps, K x (upwards + solve) on the "
                    "resident operators with f and the boundary data scaled by (1 + k/K); reported as `repeat_solves`")
    ap.add_argument("--cut", type=int, default=2, help="multi-GPU: tree level of the subtrees dealt to the ranks (2: 16 subtrees, the default of "
                    "SURVEY 8(e); 1: four subtrees for 2 or 4 GPUs, level-1 merges whole on their owners)")
    ap.add_argument("--grouped", action="store_true", help="multi-GPU (4 or 8): level-1 merges row-split inside rank groups, root merge over all ranks "
                    "(GroupedShardedHPS; not yet run on GPUs)")
    ap.add_argument("--lazy-root-dtn", action="store_true", help="EFGPU_LAZY_ROOT_DTN: the DtN map of the whole domain (read by nothing on the Dirichlet "
                    "path) is left to its first reader; recorded in config.root_dtn - not the reference's buildStage, which always forms it")
    ap.add_argument("--tuning", action="append", default=[], metavar="KEY=VALUE",
                    help="efgpu_set_tuning knob for A/B runs (include/efgpu.h), e.g. --tuning 5=1; recorded in config.tuning")
    ap.add_argument("--cpu-level", type=int, default=None, help="tree depth of the CPU sample (default 6; the reference arm lowers it until warmup + steps runs fit EFGPU_REF_BUDGET_S)")
    return ap.parse_args()


PROBLEM_NAMES = {"poisson": "constant-coefficient Poisson", "helmholtz": "constant-coefficient Helmholtz lambda=-1",
                 "varcoef": "variable-coefficient alpha/beta/lambda elliptic (FivePointStencil leaves: one block-tridiagonal LU per leaf)"}


def workload_name(a):
    if a.adaptive:
        return "adaptive quadtree levels %d-%d (|sin x + sin y| > %g, 2:1 balanced), %dx%d FV patches, %s on [-10,10]^2 (BASELINE configs[%d] shape)" % (
            a.adaptive[0], a.adaptive[1], a.threshold, a.nx, a.nx, PROBLEM_NAMES[a.problem], 3 if a.problem == "varcoef" else 0)
    shape = "BASELINE configs[1]" if (a.level, a.nx, a.problem) == (8, 16, "poisson") else "BASELINE configs[1] shape at another size"
    return "uniform level-%d quadtree, %dx%d FV patches, %s on [0,pi]^2 (%s)" % (a.level, a.nx, a.nx, PROBLEM_NAMES[a.problem], shape)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def run_reference_sample(level, nx, problem, threads, adaptive=None, threshold=1.2):
    """One run of the unmodified reference on a uniform level-`level` tree (or the adaptive mesh `adaptive` = (min, max));
    returns its REF_RESULT dict."""
    if not os.path.exists(REF_DRIVER):
        raise RuntimeError("oracle/_ref/ref_driver is missing (built by __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS=str(threads), OMP_NUM_THREADS=str(threads))
    lo, hi = adaptive if adaptive else (level, level)
    dom = ["-10", "10", "-10", "10"] if adaptive else ["0", repr(PI), "0", repr(PI)]
    cmd = [REF_DRIVER, "--problem", problem, "--solver", "fivepoint" if problem == "varcoef" else "fishpack", "--min-level", str(lo),
           "--max-level", str(hi), "--nx", str(nx), "--threshold", repr(threshold), "--domain"] + dom + ["--ops", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("REF_RESULT")][-1]
    return json.loads(line[len("REF_RESULT "):])


def run_binding_timing(a, threads, passes=2):
    """The same step through the reference-side C++ binding (include/EllipticForestB200.hpp: the reference's own Mesh / Quadtree /
    FiniteVolumeSolver objects, std::function callbacks sampled per cell, Vector copies of f and u per leaf), timed by the
    reference's stage timers: oracle/_ref/dropin_driver --time-only (compiled against the unmodified reference)."""
    drv = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")
    if not os.path.exists(drv):
        raise RuntimeError("oracle/_ref/dropin_driver is missing")
    lo, hi = a.adaptive if a.adaptive else (a.level, a.level)
    dom = ["-10", "10", "-10", "10"] if a.adaptive else ["0", repr(PI), "0", repr(PI)]
    cmd = [drv, "--problem", a.problem, "--solver", "fivepoint" if a.problem == "varcoef" else "fishpack", "--min-level", str(lo), "--max-level", str(hi),
           "--nx", str(a.nx), "--threshold", repr(a.threshold), "--domain"] + dom + ["--time-only", str(passes), "--threads", str(threads)]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("DROPIN_TIMING")][-1]
    return json.loads(line[len("DROPIN_TIMING "):])


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_sample(a, default_level):
    """Bounded CPU sample of the workload: (kwargs of run_reference_sample, description)."""
    if a.problem == "varcoef":
        default_level -= 1      # the reference factorises a dense M^2 x M^2 matrix in every leaf solve call
    if a.adaptive:
        hi = a.cpu_level if a.cpu_level is not None else min(a.adaptive[1], default_level)
        lo = min(a.adaptive[0], max(hi - 3, 0))
        return dict(level=hi, adaptive=(lo, hi), threshold=a.threshold), "adaptive levels %d-%d" % (lo, hi)
    lvl = a.cpu_level if a.cpu_level is not None else min(a.level, default_level)
    return dict(level=lvl), "uniform level-%d" % lvl


def extrapolate_reference(a, t_lo, t_hi, lvl_hi):
    """Estimate of the reference's step time on the full workload from two timed levels of the same uniform family (labelled as
    an extrapolation wherever it is printed): t(L) = t(lvl_hi) * r^(L - lvl_hi) with the MEASURED growth r = t(lvl_hi) / t(lvl_hi - 1)
    per level.  DOFs grow 4x per level and the dense root merge 8x (810.67 n^3), so r lies between 4 and 8 and grows
    with L: the estimate is optimistic for the reference."""
    if a.adaptive or t_lo is None or a.level <= lvl_hi:
        return None
    r = t_hi / t_lo
    t = t_hi * r ** (a.level - lvl_hi)
    dofs = float(4 ** a.level * a.nx * a.nx)
    return {"to": "uniform level-%d (the workload of the repo arm)" % a.level, "ms_per_step": 1e3 * t, "value": dofs / t, "unit": "DOFs/s",
            "method": "t(L) = t(%d) * r^(L-%d), r = t(%d)/t(%d) = %.2f measured in this run; not a measurement" % (lvl_hi, lvl_hi, lvl_hi, lvl_hi - 1, r)}


def reference_arm(a):
    """`--impl reference`: the reference's own CPU implementation (compiled, unmodified) on the host cores.  Every step is one
    full build + upwards + solve of a BOUNDED sample of the workload: the largest uniform level (6 by default) whose
    warmup + steps runs fit EFGPU_REF_BUDGET_S (600 s); the level actually timed is a top-level config field, the line says
    same_config false, and an extrapolation to the full workload is printed next to it, labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    budget = float(os.environ.get("EFGPU_REF_BUDGET_S", "600"))
    kw, what = cpu_sample(a, 6)
    runs = a.warmup + a.steps
    t_first = None
    while True:   # first warm-up run decides whether this level fits the budget
        res = run_reference_sample(nx=a.nx, problem=a.problem, threads=threads, **kw)
        t_first = res["build_s"] + res["upwards_s"] + res["solve_s"]
        if a.cpu_level is not None or t_first * runs <= budget or kw["level"] <= 3:
            break
        if a.adaptive:
            kw["level"] -= 1; kw["adaptive"] = (min(kw["adaptive"][0], max(kw["level"] - 3, 0)), kw["level"]); what = "adaptive levels %d-%d" % kw["adaptive"]
        else:
            kw["level"] -= 1; what = "uniform level-%d" % kw["level"]
    t_lower = None
    if not a.adaptive and kw["level"] >= 2:   # one run a level below: the measured growth per level for the extrapolation
        lo = run_reference_sample(nx=a.nx, problem=a.problem, threads=threads, **dict(kw, level=kw["level"] - 1))
        t_lower = lo["build_s"] + lo["upwards_s"] + lo["solve_s"]
    for _ in range(max(a.warmup - 1, 0)):
        run_reference_sample(nx=a.nx, problem=a.problem, threads=threads, **kw)
    ts = []
    for _ in range(a.steps):
        res = run_reference_sample(nx=a.nx, problem=a.problem, threads=threads, **kw)
        ts.append(res["build_s"] + res["upwards_s"] + res["solve_s"])
    t = sum(ts) / len(ts)
    v = res["dofs"] / t
    sample = "%s, %dx%d patches (%d DOFs): build %.3f s, upwards %.3f s, solve %.3f s per step" % (
        what, a.nx, a.nx, res["dofs"], res["build_s"], res["upwards_s"], res["solve_s"])
    full = (not a.adaptive) and kw["level"] == a.level
    print(json.dumps({
        "impl": "reference", "metric": "HPS build+upwards+solve DOFs/s (FP64)", "value": v, "unit": "DOFs/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "timed": what, "level_timed": kw["level"], "dofs_timed": res["dofs"],
                   "same_config": bool(full), "sample": sample,
                   "note": None if full else "the reference's DOFs/s falls as the tree grows (dense root merge): a ratio against this line is a "
                                             "cross-config figure that favours the reference; see `extrapolated`",
                   "extrapolated": None if full else extrapolate_reference(a, t_lower, t, kw["level"])},
        "cpu_baseline": {"value": v, "unit": "DOFs/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def own_arm(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ellipticforest_b200 as ef
    from ellipticforest_b200 import dist as efdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the HPS path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep stdout for the one JSON line (NCCL_DEBUG=VERSION prints a banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    for kv in a.tuning:
        k, v = kv.split("=")
        if ef.load().efgpu_set_tuning(int(k), int(v)) != 0:
            raise SystemExit("bad --tuning %s" % kv)

    # ---- workload (untimed set-up: mesh, plan, host sampling of f and the boundary data) ----
    if a.adaptive:
        grid = ef.FiniteVolumeGrid(a.nx, -10.0, 10.0, a.nx, -10.0, 10.0)
        mesh = ef.Mesh().refineByFunction("elliptic-single", a.threshold, a.adaptive[0], a.adaptive[1], grid)
    else:
        grid = ef.FiniteVolumeGrid(a.nx, 0.0, PI, a.nx, 0.0, PI)
        mesh = ef.Mesh().refineByFunction(None, 0.0, a.level, a.level, grid)
    solver = ef.FiniteVolumeSolver()
    u_exact = lambda x, y: np.sin(x) + np.sin(y)
    if a.problem == "varcoef":
        # manufactured solution u = sin x + sin y of alpha div(beta grad u) + lambda u = f
        solver.solver_type = "FivePointStencil"
        beta = lambda x, y: 1.0 + 0.5 * np.sin(x) * np.cos(y)
        lamf = lambda x, y: -(1.0 + 0.5 * np.cos(x) * np.cos(y))
        solver.beta_function, solver.lambda_function = beta, lamf
        f_fn = lambda x, y: (0.5 * np.cos(x) * np.cos(y) * np.cos(x) - 0.5 * np.sin(x) * np.sin(y) * np.cos(y)
                             - beta(x, y) * u_exact(x, y) + lamf(x, y) * u_exact(x, y))
    else:
        solver.solver_type = "FISHPACK90"
        lam = 0.0 if a.problem == "poisson" else -1.0
        solver.lambda_function = lambda x, y: lam + 0.0 * x
        f_fn = lambda x, y: (lam - 1.0) * u_exact(x, y)

    # sharded adaptive trees: whole level-2 subtrees per GPU, dealt as contiguous Morton blocks of (nearly) equal merge work
    hps = efdist.make_hps(mesh, solver, device=local, rank=rank, world=world, cut=a.cut, grouped=a.grouped,
                          balance="work" if (a.adaptive and world > 1) else "count")
    hps.no_symmetry = a.no_symmetry
    hps.lazy_root_dtn = a.lazy_root_dtn
    if a.lean_T:
        if world > 1:
            raise SystemExit("--lean-T: single GPU only")
        hps.lean_T = True
    hps.setupStage()
    f_host, g_host = hps.sample_inputs(f_fn, u_exact)          # numpy, this rank's share
    f_pin = torch.from_numpy(f_host).pin_memory()
    g_pin = torch.from_numpy(g_host).pin_memory()
    u_pin = torch.empty(f_host.size, dtype=torch.float64).pin_memory()
    f_dev = f_pin.cuda()
    g_dev = g_pin.cuda()
    u_dev = torch.empty_like(f_dev)
    torch.cuda.synchronize()
    dofs = float(mesh.n_leaves * a.nx * a.nx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():      # every input resident in HBM: f, g and (FivePointStencil leaves) the sampled coefficient arrays
        hps.resample_coefficients = False
        hps.buildStage()
        hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=False)
        hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True)

    def step_host():        # host buffers: alpha/beta/lambda sampled and copied by buildStage, f and g copied H2D, u copied D2H
        hps.resample_coefficients = True
        hps.buildStage()
        hps.upwardsStageHost(f_pin.numpy())
        hps.solveStageHost(g_pin.numpy(), u_pin.numpy())

    def timed(fn, steps):
        """barrier + synchronize on both sides; returns seconds (max over ranks) for `steps` steps."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(hps.stream())
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        barrier()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        t = torch.tensor([dev, wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    if a.profile_run:   # launched under ncu: numbers printed by such a run are never bench values
        for _ in range(a.warmup + a.steps):
            step_device()
        print("profile run done: %d steps" % (a.warmup + a.steps))
        return

    # ---- warm-up, then the device-resident leg under the clock sampler ----
    # short steps (adaptive trees: ms) also get the clocks up: at least 0.5 s of warm-up.  The number of extra steps is agreed
    # between the ranks (max over ranks of the measured step time): every step contains collective flag barriers
    t_w = time.perf_counter()
    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    t_step = torch.tensor([(time.perf_counter() - t_w) / max(a.warmup, 3)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    for _ in range(int(min(200, max(0.0, 0.5 - max(a.warmup, 3) * float(t_step[0])) / max(float(t_step[0]), 1e-6)))):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_s, wall_s = timed(step_device, a.steps)
    ms_per_step_hint = 1e3 * dev_s / a.steps
    # Per-kernel CUDA events (two cudaEventRecord per launch on the library's stream, ~900 launches per step) are taken
    # over a second, identical pass of the same K steps: in the launch-bound top tree levels the event records themselves
    # cost 3 % of a step at N = 1 and 10 % at N = 8, which must not sit in the headline time.
    hps.set_profiling(True)
    prof_dev_s, _ = timed(step_device, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    prof = hps.profile()
    stats = hps.stats()
    hps.set_profiling(False)

    # correctness of the timed computation: discretisation error against the manufactured solution
    err = hps.max_error(u_dev, u_exact)
    mesh_stats = None
    if a.adaptive:   # leaves per level and the histogram of coarsening tags (HPSAlgorithm.hpp:676-741) over all nodes
        lv = np.bincount(mesh.level[mesh.leaf_nodes])
        mesh_stats = {"leaves_per_level": {str(i): int(c) for i, c in enumerate(lv) if c}, "nodes": mesh.n_nodes}
        if world == 1:
            tags = np.bincount([hps.node_info(i)["n_coarsens"] for i in range(mesh.n_nodes)])
            mesh_stats["n_coarsens_histogram"] = {str(i): int(c) for i, c in enumerate(tags)}

    # ---- sharded runs: parity against a single-GPU build of the SAME tree, measured here so that every driver-run line carries it
    # (each rank rebuilds the whole tree on its own GPU when it fits next to its shard, and compares its leaves' u; rank 0
    # compares the root's DtN map after the collective that gathers its row slices).  Reference: the per-node parity tests tie the
    # single-GPU path to the reference (tests/test_gpu_parity.py, tests/test_gpu_refscale.py).
    parity = None
    if world > 1:
        parity = {"linf_error_vs_exact": err}
        free_b, _tot_b = torch.cuda.mem_get_info()
        single_bytes = dofs * (256.0 * a.level + 300.0)      # 1 KiB n^2 per merge, sum of n^2 per level = DOFs / 4; + leaf maps, vectors, workspace
        fits = torch.tensor([1 if single_bytes < 0.85 * free_b else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if int(fits[0]) and not a.adaptive and a.problem != "varcoef":
            rootT = hps.gather_root_T()                       # collective: completes the row-distributed root map
            one = ef.HPSAlgorithm(mesh, solver, device=local)
            one.no_symmetry = a.no_symmetry
            one.buildStage()
            Xc, Yc = mesh.leaf_cell_centres()
            one.upwardsStage(f_fn(Xc, Yc))
            _s, bx, by = one.root_boundary_points()
            u_one = one.solveStage(np.ascontiguousarray(u_exact(bx, by)))
            lo, hi = hps.leaf_lo, hps.leaf_hi
            mine = u_dev.cpu().numpy().reshape(hi - lo, a.nx, a.nx)
            du = torch.tensor([float(np.max(np.abs(mine - u_one[lo:hi]))), float(np.max(np.abs(u_one[lo:hi])))], dtype=torch.float64, device="cuda")
            dist.all_reduce(du, op=dist.ReduceOp.MAX)
            parity["u_vs_single_gpu_rel_max"] = float(du[0] / du[1])
            if rank == 0:
                T1 = one.operator(0, "T").reshape(-1)
                Tn = rootT.cpu().numpy()
                parity["root_T_vs_single_gpu_rel_max"] = float(np.max(np.abs(Tn - T1)) / np.max(np.abs(T1)))
                del T1, Tn
            del one, u_one, mine
        else:
            parity["note"] = ("adaptive / variable-coefficient tree: no single-GPU rebuild in the bench line (tests/test_gpu_sharded.py compares them); "
                              "error against the exact solution only" if (a.adaptive or a.problem == "varcoef") else
                              "tree does not fit one GPU next to the shard: no single-GPU rebuild; error against the exact solution only")

    # ---- e2e leg: host buffers through the C-ABI, copies inside the timed region ----
    step_host()
    e2e_dev_s, e2e_wall_s = timed(step_host, a.steps)
    err_e2e = float(np.max(np.abs(u_pin.numpy() - u_dev.cpu().numpy())))

    # ---- variable coefficients: the same end-to-end step through the device-functor entry points ----
    # The reference calls alpha / beta / lambda / f as std::function per point on the host (FiniteVolumeSolver.cpp:63-79,
    # HPSAlgorithm.hpp:241-249); the host-array e2e leg above pays for that sampling (numpy on all host threads) plus 7 arrays of
    # H2D.  Here the library writes the sampling coordinates on the device (efgpu_leaf_points_device), the caller's functions are
    # evaluated there by its own device code (torch elementwise kernels stand in for the user's functors) and handed over as
    # device pointers (efgpu_set_leaf_variable_device, efgpu_upwards_device); g still comes from the host, u still goes back.
    e2e_devsample = None
    if a.problem == "varcoef" and world == 1:
        st_lib = torch.cuda.ExternalStream(hps.stream())
        shape = (mesh.n_leaves, a.nx, a.nx)
        xd, yd = (torch.empty(shape, dtype=torch.float64, device="cuda") for _ in range(2))
        t_u = lambda x, y: torch.sin(x) + torch.sin(y)
        t_beta = lambda x, y: 1.0 + 0.5 * torch.sin(x) * torch.cos(y)
        t_lam = lambda x, y: -(1.0 + 0.5 * torch.cos(x) * torch.cos(y))
        t_f = lambda x, y: (0.5 * torch.cos(x) * torch.cos(y) * torch.cos(x) - 0.5 * torch.sin(x) * torch.sin(y) * torch.cos(y)
                            - t_beta(x, y) * t_u(x, y) + t_lam(x, y) * t_u(x, y))

        def step_devsample():
            with torch.cuda.stream(st_lib):
                arrs = []
                for which, fn in (("centre", None), ("W", t_beta), ("E", t_beta), ("S", t_beta), ("N", t_beta), ("centre", t_lam)):
                    hps.leafPointsDevice(which, xd.data_ptr(), yd.data_ptr(), sync=False)
                    arrs.append(torch.ones(shape, dtype=torch.float64, device="cuda") if fn is None else fn(xd, yd))
                f_t = t_f(xd, yd)                  # xd, yd hold the cell centres again (last request)
                hps.setVariableCoefficientsDevice(*[t.data_ptr() for t in arrs])
                hps.buildStage()
                hps.upwardsStageDevice(f_t.data_ptr(), 1.0, sync=False)
                g_dev.copy_(g_pin, non_blocking=True)
                hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=False)
                u_pin.copy_(u_dev, non_blocking=True)
            hps.sync()

        u_host_leg = u_pin.numpy().copy()
        step_devsample()
        ds_dev_s, ds_wall_s = timed(step_devsample, a.steps)
        e2e_devsample = {"value": dofs * a.steps / ds_wall_s, "unit": "DOFs/s", "ms_per_step": 1e3 * ds_wall_s / a.steps,
                         "h2d_bytes_per_step": int(g_host.nbytes), "d2h_bytes_per_step": int(f_host.nbytes),
                         "max_abs_diff_vs_host_sampled": float(np.max(np.abs(u_pin.numpy() - u_host_leg))),
                         "what": "coordinates written by efgpu_leaf_points_device, alpha/beta/lambda/f evaluated on the device by the caller's "
                                 "elementwise kernels, efgpu_set_leaf_variable_device + efgpu_upwards_device; g H2D and u D2H as in e2e"}
        hps.resample_coefficients = True
        hps._coefficients_set = False

    # ---- stage split (extra passes, not part of the headline): each stage between barriers, max over ranks ----
    hps.resample_coefficients = False
    med = lambda v: sorted(v)[len(v) // 2]
    reps = 5 if ms_per_step_hint < 400.0 else 3          # median of several samples per stage
    build_ms = 1e3 * med([timed(lambda: hps.buildStage(), 1)[0] for _ in range(reps)])
    up_ms = 1e3 * med([timed(lambda: hps.upwardsStageDevice(f_dev.data_ptr(), 1.0, sync=True), 1)[0] for _ in range(reps)])
    so_ms = 1e3 * med([timed(lambda: hps.solveStageDevice(g_dev.data_ptr(), u_dev.data_ptr(), sync=True), 1)[0] for _ in range(reps)])
    # ---- repeated solves on the resident operators (the `thermal` usage pattern) ----
    repeat = None
    if a.n_solves > 0:
        K = a.n_solves
        st = torch.cuda.ExternalStream(hps.stream())
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        checks = {}
        barrier()
        for k in range(K):
            sc = 1.0 + k / K
            with torch.cuda.stream(st):
                g_k = g_dev * sc
            ev[k][0].record(st)
            hps.upwardsStageDevice(f_dev.data_ptr(), sc, sync=False)
            hps.solveStageDevice(g_k.data_ptr(), u_dev.data_ptr(), sync=False)
            ev[k][1].record(st)
            if k in (0, K // 2, K - 1):          # every solve is checkable against (1 + k/K) u
                hps.sync()
                checks[k] = hps.max_error(u_dev, lambda x, y, sc=sc: sc * u_exact(x, y))
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) for e0, e1 in ev], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.cpu().numpy()
        repeat = {"n": K, "mean_ms": float(ms.mean()), "min_ms": float(ms.min()), "max_ms": float(ms.max()),
                  "linf_error_vs_scaled_exact": {str(k): v for k, v in checks.items()}}
    # whole-job totals of the per-rank work models (flops issued, algorithmic bytes, device memory)
    agg = torch.tensor([stats[k] for k in ("merge_flops_issued", "merge_flops_canonical", "upwards_bytes", "solve_bytes", "device_bytes")],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
    tot = dict(zip(("issued", "canonical", "upwards_bytes", "solve_bytes", "device_bytes"), [float(v) for v in agg]))

    # ---- measured FP64 GEMM ceiling on this box (cuBLAS DGEMM 8192^3), the tensor roofline denominator ----
    dgemm_tf = None
    if rank == 0:
        n = 8192
        A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            A @ B
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); A @ B; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        dgemm_tf = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        del A, B

    # ---- CPU baseline: the unmodified reference on a bounded sample, rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            kw, what = cpu_sample(a, 6)
            threads = host_threads()
            r = run_reference_sample(nx=a.nx, problem=a.problem, threads=threads, **kw)
            t = r["build_s"] + r["upwards_s"] + r["solve_s"]
            cpu = {"value": r["dofs"] / t, "unit": "DOFs/s", "cores": threads, "kind": "reference",
                   "sample": "oracle/_ref/ref_driver (unmodified reference, OpenBLAS x%d threads), %s %dx%d patches, %d DOFs: "
                             "build %.2f s, upwards %.2f s, solve %.2f s" % (threads, what, a.nx, a.nx, r["dofs"], r["build_s"], r["upwards_s"], r["solve_s"]),
                   "build_dofs_per_s": r["dofs"] / r["build_s"], "solve_dofs_per_s": r["dofs"] / (r["upwards_s"] + r["solve_s"])}
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            cpu = {"value": None, "unit": "DOFs/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}

    # ---- the step through the reference-side C++ binding (per-cell std::function sampling included), rank 0, N = 1 only ----
    binding = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            torch.cuda.synchronize()
            binding = {}
            for thr in (1, host_threads()):
                r = run_binding_timing(a, thr)
                t = r["build_s"] + r["upwards_s"] + r["solve_s"]
                binding["sampling_threads_%d" % thr] = {"value": r["dofs"] / t, "unit": "DOFs/s", "ms_per_step": 1e3 * t, "build_ms": 1e3 * r["build_s"],
                                                        "upwards_ms": 1e3 * r["upwards_s"], "solve_ms": 1e3 * r["solve_s"], "setup_ms": 1e3 * r["setup_s"],
                                                        "linf_error_vs_exact": r["linf_error"]}
            binding["what"] = ("oracle/_ref/dropin_driver --time-only: HPSAlgorithmB200 (include/EllipticForestB200.hpp) on the reference's own Mesh / "
                               "Quadtree / FiniteVolumePatch objects; stage times from the reference's timers, i.e. including the per-cell "
                               "std::function sampling of f (and alpha / beta / lambda) and the per-leaf Vector copies of f and u; "
                               "sampling_threads > 1 is the binding's opt-in for thread-safe callbacks")
        except Exception as e:
            binding = {"error": str(e)}

    def cleanup():
        # torch frees pinned buffers by recording events on the streams they were used on: release every tensor
        # while the library's stream still exists, then the handles, then the process group
        nonlocal f_pin, g_pin, u_pin, f_dev, g_dev, u_dev, hps
        import gc
        torch.cuda.synchronize()
        hps._f_dev = hps._g_dev = hps._u_dev = None
        f_pin = g_pin = u_pin = f_dev = g_dev = u_dev = None
        gc.collect()
        torch.cuda.synchronize()
        hps = None
        gc.collect()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        cleanup()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture (never measured under this run)
    traffic = traffic_bytes = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = {"bytes": tr["traffic_bytes_per_launch"], "algorithmic_bytes": tr["algorithmic_bytes_per_launch"],
                   "launch": tr["launch"], "source": tr["source"], "captured_at_commit": tr.get("commit")}
        traffic_bytes = tr["traffic_bytes_per_launch"]
    except Exception:
        pass

    gemm_ms = sum(prof[k][0] for k in ("gemm_Xinv", "gemm_S", "gemm_T"))
    gemm_launches = sum(prof[k][1] for k in ("gemm_Xinv", "gemm_S", "gemm_T"))
    total_prof_ms = sum(v[0] for k, v in prof.items())
    launches = sum(v[1] for v in prof.values())
    flops_issued = stats["merge_flops_issued"] * a.steps      # rank 0's own merges (its subtrees + the upper tree when sharded)
    gemm_tf = flops_issued / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    ms_per_step = 1e3 * dev_s / a.steps
    out = {
        "metric": "HPS build+upwards+solve DOFs/s (FP64)", "value": dofs * a.steps / dev_s, "unit": "DOFs/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "dofs": dofs, "leaves": mesh.n_leaves, "mesh": mesh_stats, "l2": "inputs larger than L2 (%.1f GB of operators streamed per step)" % (tot["device_bytes"] / 1e9),
                   "sharding": hps.sharding(), "parity": parity, "tuning": a.tuning, "root_dtn": "deferred to its first reader (EFGPU_LAZY_ROOT_DTN)" if a.lazy_root_dtn else "formed by the build",
                   "merge_plan": "general (EFGPU_NO_SYMMETRY)" if a.no_symmetry else "symmetric where the subtree is uniform with constant-coefficient leaves%s" % (
                       "" if (a.adaptive or a.problem == "varcoef") else " (every merge of this workload)")},
        "stages": {"build_ms": build_ms, "upwards_ms": up_ms, "solve_ms": so_ms, "samples": "median of %d runs per stage" % reps,
                   "build_dofs_per_s": dofs / (build_ms * 1e-3), "solve_dofs_per_s": dofs / ((up_ms + so_ms) * 1e-3),
                   "upwards_gbs": tot["upwards_bytes"] / (up_ms * 1e-3) / 1e9, "solve_gbs": tot["solve_bytes"] / (so_ms * 1e-3) / 1e9,
                   "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
                   "merge_tflops_issued": tot["issued"] / (build_ms * 1e-3) / 1e12,
                   "merge_tflops_canonical": tot["canonical"] / (build_ms * 1e-3) / 1e12},
        "linf_error_vs_exact": err, "e2e_vs_device_max_abs_diff": err_e2e,
        "e2e": {"value": dofs * a.steps / e2e_wall_s, "unit": "DOFs/s", "h2d_bytes_per_step": int(f_host.nbytes * (7 if (a.problem == "varcoef" and world == 1) else 1) + g_host.nbytes),
                "d2h_bytes_per_step": int(f_host.nbytes), "ms_per_step": 1e3 * e2e_wall_s / a.steps, "timer": "host wall clock between device synchronisations"},
        "e2e_device_sampling": e2e_devsample, "e2e_cpp_binding": binding,
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "bgemm_tma_kernel / bgemm_kernel (FP64 DMMA batched GEMM of the merges: X^-1 blocks, S, T; 128-row tiles on TMA-staged operands, smaller tiles on cp.async)%s" % (" on rank 0" if world > 1 else ""),
                     "achieved": gemm_tf, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": (gemm_tf / dgemm_tf) if (gemm_tf and dgemm_tf) else None,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure; NVIDIA nominal FP64 tensor 37-40 TFLOP/s)",
                     "flops": "issued to the tensor pipe (symmetric plan: ~327 n^3 per merge, general plan 484 n^3; the reference's dgesv+dgemm count is 810.67 n^3)",
                     "launches": int(gemm_launches), "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
                     "timing": "CUDA events around every launch on the library's stream, over a second pass of the same %d steps (%.2f ms per step with "
                               "the events, %.2f without)" % (a.steps, 1e3 * prof_dev_s / a.steps, ms_per_step),
                     "traffic": traffic_bytes, "traffic_detail": traffic},
        "kernel_ms_per_step": {k: v[0] / a.steps for k, v in prof.items() if v[1] > 0 or v[0] > 0},
        "repeat_solves": None if repeat is None else dict(repeat, dofs_per_s_mean=dofs / (repeat["mean_ms"] * 1e-3),
                                                           gbs_mean=(tot["upwards_bytes"] + tot["solve_bytes"]) / (repeat["mean_ms"] * 1e-3) / 1e9,
                                                           hbm_frac_mean=(tot["upwards_bytes"] + tot["solve_bytes"]) / (repeat["mean_ms"] * 1e-3) / 1e9 / hbm_peak),
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))
    sys.stdout.flush()
    cleanup()


def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        own_arm(a)


if __name__ == "__main__":
    main()
