#!/usr/bin/env python
"""Analytic model of one build step of a uniform tree from the library's own merge plans (efgpu_debug_merge_plan: host logic,
no device): per tree level and per rank the flops issued to the tensor pipe, the serial base-case inversions, the launches and
the bytes all-gathered, turned into milliseconds with rates measured in round 1 on B200 (profiles/r1q_*, r1n_*):

    merge GEMMs      31.3 TFLOP/s issued  (big S / T products 32.7, small products of the inversion less: see --gemm-small)
    base case        73 us per 128 x 128 Gauss-Jordan launch, 1 or 2 matrices (invert_reg_kernel; a launch is serial in the chain)
    launch overhead  4 us per dependent small launch of the top levels
    all-gather       NCCL over NVSwitch: 20 us latency + bytes / 600 GB/s received per rank
    subtree roots' T 16 broadcasts (or one all-gather): bytes / 600 GB/s

It is a planning tool: it reproduces the measured L = 8, M = 16 steps (208 / 129.5 / 81.8 / 63.7 ms at 1 / 2 / 4 / 8 GPUs) to
within ~10 % and is used to rank the switches that have not been timed yet (tuning key 5, EFGPU_LAZY_ROOT_DTN, the balanced T
plan).  Nothing here is a measurement."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellipticforest_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=8)
ap.add_argument("--nx", type=int, default=16)
ap.add_argument("--gpus", default="1,2,4,8")
ap.add_argument("--cut", type=int, default=2)
ap.add_argument("--tuning5", type=int, default=0)
ap.add_argument("--lazy-root-dtn", action="store_true")
ap.add_argument("--grouped", action="store_true", help="GroupedShardedHPS: level-1 merges inside groups of gpus / 4 ranks")
ap.add_argument("--gemm", type=float, default=32.5, help="TFLOP/s of the large products (S, T)")
ap.add_argument("--gemm-small", type=float, default=22.0, help="TFLOP/s of the products of the block inversion")
ap.add_argument("--base-us", type=float, default=73.0)
ap.add_argument("--launch-us", type=float, default=4.0)
ap.add_argument("--nvlink", type=float, default=600.0, help="GB/s received per rank in an all-gather")
ap.add_argument("--gather-us", type=float, default=20.0)
ap.add_argument("--other-ms", type=float, default=8.0, help="assembly, leaf DtN, mirrors, upwards + solve at 1 GPU (scaled by 1 / ranks)")
a = ap.parse_args()

lib = _lib.load()
assert lib.efgpu_set_tuning(5, a.tuning5) == 0
CLS_XINV, CLS_S, CLS_T = 4, 5, 6


def plan(n, level, rank, nranks, sym=1):
    ns, nb, nt = C.c_int(), C.c_int(), C.c_int()
    ws = np.zeros(3, dtype=np.int64)
    p = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
    assert lib.efgpu_debug_merge_plan(n, level, rank, nranks, sym, None, C.byref(ns), None, None, C.byref(nb), None, C.byref(nt), p(ws)) == 0
    steps = np.zeros((ns.value, 16), dtype=np.int64); blocks = np.zeros((nb.value, 16), dtype=np.int64)
    terms = np.zeros((nb.value, 2, 8), dtype=np.int64); trans = np.zeros((max(nt.value, 1), 16), dtype=np.int64)
    assert lib.efgpu_debug_merge_plan(n, level, rank, nranks, sym, p(steps), C.byref(ns), p(blocks), p(terms), C.byref(nb), p(trans), C.byref(nt), p(ws)) == 0
    return steps, blocks, terms


def merge_cost(n, level, nranks, count):
    """components (ms) of one batch of `count` merges with child side n on tree level `level`, for the slowest rank."""
    worst = None
    for r in range(nranks):
        steps, blocks, terms = plan(n, level, r, nranks)
        fl = {CLS_XINV: 0.0, CLS_S: 0.0, CLS_T: 0.0}
        base_ms = 0.0
        launches = 0
        gbytes = 0.0; ngather = 0
        for st in steps:
            kind, first, cnt, N, cls, gk = int(st[0]), int(st[1]), int(st[2]), int(st[4]), int(st[5]), int(st[6])
            if kind == 0:       # one CTA per matrix (two per entry in a paired launch), one CTA per SM; cost ~ N^2 per pivot sweep
                mats = count * (2 if int(st[12]) >= 0 else 1)
                base_ms += a.base_us * 1e-3 * (N / 128.0) ** 2 * max(1.0, mats / 148.0)
            elif kind == 1:
                if cls == CLS_T and level == 0 and a.lazy_root_dtn:
                    continue
                for k in range(first, first + cnt):
                    for t in range(int(blocks[k][8])):
                        fl[cls] += 2.0 * int(blocks[k][6]) * int(blocks[k][7]) * int(terms[k][t][6])
            launches += 1
            if gk:
                gbytes += 8.0 * int(st[8]) * int(st[9]) * (nranks - 1) / nranks; ngather += 1
        if nranks > 1:      # S after phase 0; T after phase 1 except at the root
            gbytes += 8.0 * 32 * n * n * (nranks - 1) / nranks; ngather += 1
            if level > 0:
                gbytes += 8.0 * 64 * n * n * (nranks - 1) / nranks; ngather += 1
        c = dict(gemm_ST=count * (fl[CLS_S] + fl[CLS_T]) / (a.gemm * 1e9), gemm_Xinv=count * fl[CLS_XINV] / (a.gemm_small * 1e9),
                 base=base_ms, launch=launches * a.launch_us * 1e-3 * (1.0 if count <= 16 else 0.0),
                 gather=count * (gbytes / (a.nvlink * 1e6)) + count * ngather * a.gather_us * 1e-3)
        if worst is None or sum(c.values()) > sum(worst.values()):
            worst = c
    return worst


def step(nranks):
    L, M = a.level, a.nx
    tot = dict(gemm_ST=0.0, gemm_Xinv=0.0, base=0.0, launch=0.0, gather=0.0)
    grouped = a.grouped and nranks >= 4
    gs = nranks // 4
    for lev in range(L - 1, -1, -1):
        n = M << (L - 1 - lev)
        merges = 4 ** lev
        if grouped and lev == 1:
            c = merge_cost(n, lev, gs, 1)            # one level-1 merge per group, row-split over its gs ranks
        else:
            c = merge_cost(n, lev, 1, merges // nranks) if (nranks > 1 and lev >= a.cut) else merge_cost(n, lev, nranks, merges)
        for k, v in c.items():
            tot[k] += v
    if grouped:      # subtree roots' T inside the group, then the four level-1 T's all-gathered over everybody
        n2, n1 = M << (L - 2), M << (L - 1)
        tot["gather"] += 8.0 * 4 * (4 * n2) ** 2 * (gs - 1) / gs / (a.nvlink * 1e6) + (a.gather_us * 1e-3 if gs > 1 else 0.0)
        tot["gather"] += 8.0 * 4 * (4 * n1) ** 2 * (nranks - 1) / nranks / (a.nvlink * 1e6) + a.gather_us * 1e-3
    elif nranks > 1:
        n2 = M << (L - a.cut)
        tot["gather"] += 8.0 * (4 ** a.cut) * (4 * n2) ** 2 * (nranks - 1) / nranks / (a.nvlink * 1e6) + 16 * a.gather_us * 1e-3
    tot["other"] = a.other_ms / nranks
    return sum(tot.values()), tot


print("uniform level-%d tree of %dx%d patches, tuning5=%d, lazy root DtN=%s" % (a.level, a.nx, a.nx, a.tuning5, a.lazy_root_dtn))
for g in [int(v) for v in a.gpus.split(",")]:
    t, tot = step(g)
    print("%d GPU(s): %6.1f ms   " % (g, t) + "  ".join("%s %.1f" % kv for kv in tot.items()))
lib.efgpu_set_tuning(5, 0)
