#!/usr/bin/env python
"""Design study (CPU, numpy): structure of one uniform 4-to-1 merge, taken from the oracle.

Findings recorded in DESIGN.md:
  * X is symmetric positive definite on uniform constant-coefficient subtrees;
  * H = D_t S_RHS^T with D_t = -1 on the E and N exterior blocks (the coordinate-derivative sign convention), the interface
    signs are all +1; hence T = T_LHS + D_t (S_RHS^T X^-1 S_RHS);
  * a Cholesky route (X = L L^T, Y = L^-1 S_RHS, T = T_LHS + D_t Y^T Y, S = L^-T Y) is NOT cheaper than the explicit inverse:
    Y fills in, so the SYRK costs ~288 n^3 for the 36 block pairs against 144 n^3 for H S with the 50 %-sparse H.
Usage: python tools/merge_structure_study.py [depth]   (child side n = 16 * 2^depth, default 2)"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hps_oracle as O  # noqa: E402  (test infrastructure; this is a study script, not product code)


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = 16 << depth
    r = O.run(problem_name="poisson", solver_kind="fishpack", box=(0.0, np.pi, 0.0, np.pi), nx=16, min_level=depth + 1, max_level=depth + 1)
    root = r.nodes[0]
    X, H, S_perm = root.X, root.H, root.S
    S_pre = np.zeros_like(S_perm)
    for p, q in enumerate(O.PI_WESN):
        S_pre[:, q * n:(q + 1) * n] = S_perm[:, p * n:(p + 1) * n]
    S_RHS = X @ S_pre
    w = np.linalg.eigvalsh((X + X.T) / 2)
    print("n = %d: |X - X^T| / |X| = %.1e, eigenvalues of X in [%.3g, %.3g]" % (n, np.max(np.abs(X - X.T)) / np.max(np.abs(X)), w[0], w[-1]))
    best = None
    for st in itertools.product([1, -1], repeat=8):
        e = np.max(np.abs(H - np.repeat(st, n)[:, None] * S_RHS.T)) / np.max(np.abs(H))
        if best is None or e < best[0]:
            best = (e, st)
    print("H = D_t S_RHS^T with block signs %s (pre-permutation order aW aS bE bS gW gN oE oN): residual %.1e" % (best[1], best[0]))
    L = np.linalg.cholesky(X)
    Y = np.linalg.solve(L, S_RHS)
    nzY = [[bool(np.max(np.abs(Y[k * n:(k + 1) * n, q * n:(q + 1) * n])) > 1e-12) for q in range(8)] for k in range(4)]
    nzH = [[bool(np.max(np.abs(H[q * n:(q + 1) * n, k * n:(k + 1) * n])) > 0) for k in range(4)] for q in range(8)]
    syrk = sum(2 * sum(nzY[k][p] and nzY[k][q] for k in range(4)) for p in range(8) for q in range(p, 8))
    hs = sum(2 * sum(nzH[p][k] for k in range(4)) for p in range(8) for q in range(p, 8))
    print("block products (x n^3 flops) for the 36 block pairs of T: Y^T Y %d, H S %d" % (syrk, hs))


if __name__ == "__main__":
    main()
