"""A/B of the operand path of the FP64 tensor-core GEMM on merge-like shapes (one GPU): the default LDGSTS kernel
(efgpu_dgemm_batched: 3-stage cp.async, padded shared memory, __syncthreads per k-tile) against the TMA-staged variants
(efgpu_dgemm_batched_tma: cp.async.bulk.tensor.2d + mbarrier ring, 128-byte swizzle) and cuBLAS (torch.bmm), TFLOP/s each, with a
bit-identity check of the two kernels.  `--ncu SHAPE_INDEX` runs exactly one launch of each kernel on that shape (for a capture).
    python tools/gemm_tma_ab.py [--iters 10]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(8192, 8192, 8192, 1), (4096, 4096, 4096, 4), (2048, 2048, 4096, 16), (1024, 1024, 2048, 64), (512, 512, 1024, 256), (2048, 2048, 2048, 1),
          (4096, 16384, 4096, 1)]
TMA_VARIANTS = [4, 6, 3, 64]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ncu", type=int, default=-1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    import ellipticforest_b200 as ef
    lib = ef.load()
    rows = []
    shapes = SHAPES if a.ncu < 0 else [SHAPES[a.ncu]]
    for (m, n, k, batch) in shapes:
        A = torch.randn(batch, m, k, dtype=torch.float64, device="cuda")
        B = torch.randn(batch, k, n, dtype=torch.float64, device="cuda")
        C0 = torch.zeros(batch, m, n, dtype=torch.float64, device="cuda")
        C1 = torch.zeros(batch, m, n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ms = C.c_float()
        fl = 2.0 * m * n * k * batch
        iters = 0 if a.ncu >= 0 else a.iters
        assert lib.efgpu_dgemm_batched(A.data_ptr(), B.data_ptr(), C0.data_ptr(), m, n, k, batch, 128, iters, C.byref(ms)) == 0
        row = {"shape": [m, n, k, batch], "ldgsts_tflops": fl / (ms.value * 1e-3) / 1e12 if iters else None}
        for v in (TMA_VARIANTS if a.ncu < 0 else [4]):
            C1.zero_()
            rc = lib.efgpu_dgemm_batched_tma(A.data_ptr(), B.data_ptr(), C1.data_ptr(), m, n, k, batch, v, iters, C.byref(ms))
            assert rc == 0, (rc, lib.efgpu_last_error(None))
            row["tma%d_tflops" % v] = fl / (ms.value * 1e-3) / 1e12 if iters else None
            row["tma%d_bit_identical" % v] = bool(torch.equal(C0, C1))
        ref = A[0] @ B[0]
        row["rel_err_vs_cublas"] = float((C1[0] - ref).abs().max() / ref.abs().max())
        if iters:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.bmm(A, B)
            e0.record()
            for _ in range(iters):
                torch.bmm(A, B)
            e1.record(); torch.cuda.synchronize()
            row["cublas_tflops"] = fl * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
        rows.append(row)
        print(json.dumps(row), flush=True)
        del A, B, C0, C1
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
