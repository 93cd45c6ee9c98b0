"""Stand-alone timing of the descriptor-driven FP64 GEMM (efgpu_dgemm_batched) for a few merge-like shapes,
one process per tuning variant (EFGPU_GEMM_VARIANT is read once per process).  Prints TFLOP/s; also times
cuBLAS (torch.matmul) on the same shapes.  Run on the GPU box: python tools/gemm_bench.py"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(2048, 2048, 4096, 16), (4096, 4096, 4096, 4), (1024, 1024, 2048, 64), (512, 512, 1024, 256), (2048, 2048, 2048, 1), (1024, 1024, 1024, 1)]


def child():
    import torch
    import ellipticforest_b200 as ef
    lib = ef.load()
    out = []
    for (m, n, k, batch) in SHAPES:
        A = torch.randn(batch, m, k, dtype=torch.float64, device="cuda")
        B = torch.randn(batch, k, n, dtype=torch.float64, device="cuda")
        Cm = torch.zeros(batch, m, n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ms = C.c_float()
        rc = lib.efgpu_dgemm_batched(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), m, n, k, batch, 128, 5, C.byref(ms))
        assert rc == 0
        tf = 2.0 * m * n * k * batch / (ms.value * 1e-3) / 1e12
        err = float((Cm[0] - A[0] @ B[0]).abs().max() / (A[0] @ B[0]).abs().max())
        out.append("%dx%dx%d b%d: %.2f TF/s (err %.1e)" % (m, n, k, batch, tf, err))
        if os.environ.get("EFGPU_GEMM_VARIANT", "0") == "0":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.bmm(A, B)
            e0.record()
            for _ in range(5):
                torch.bmm(A, B)
            e1.record(); torch.cuda.synchronize()
            out[-1] += "  | cuBLAS %.2f TF/s" % (2.0 * m * n * k * batch * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del A, B, Cm
    print("variant %s\n  " % os.environ.get("EFGPU_GEMM_VARIANT", "0") + "\n  ".join(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for v in (sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6"]):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, EFGPU_GEMM_VARIANT=v))
