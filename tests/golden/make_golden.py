"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   make -C oracle ref && python tests/golden/make_golden.py
Each fixture is the complete dump of oracle/_ref/ref_driver for one small case: p4est traversal
orders, every node's grid box, n_coarsens, T/S/X/H after buildStage, h/w/f after upwardsStage and
g/u after solveStage.  The GPU box has no /root/reference; tests only read the committed .npz.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from refdump import read_dump  # noqa: E402

PI = repr(np.pi)
CASES = {
    # name: ref_driver arguments
    "uniform_l2_m8_poisson": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "2", "--max-level", "2", "--nx", "8", "--domain", "0", PI, "0", PI],
    "uniform_l1_m16_helmholtz": ["--problem", "helmholtz", "--solver", "fishpack", "--min-level", "1", "--max-level", "1", "--nx", "16", "--domain", "0", PI, "0", PI],
    "adaptive_l1_3_m8_poisson": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "1", "--max-level", "3", "--nx", "8", "--domain", "-10", "10", "-10", "10", "--refine-box", "-10", "0.5", "-10", "0.5"],
    "adaptive_tag2_m8_helmholtz_rect": ["--problem", "helmholtz", "--solver", "fishpack", "--min-level", "1", "--max-level", "4", "--nx", "8", "--domain", "0", "2", "0", "1", "--refine-box", "1.0", "2.0", "0.5", "1.0"],
    # indefinite operator (lambda = +5 > first Dirichlet eigenvalue 2 of [0,pi]^2; hstcrt IERROR = 6, tolerated): the root merge matrix has negative eigenvalues
    "uniform_l2_m8_helmholtz_indefinite": ["--problem", "helmholtz", "--lambda", "5.0", "--solver", "fishpack", "--min-level", "2", "--max-level", "2", "--nx", "8", "--domain", "0", PI, "0", PI],
    # patch sizes of the reference's convergence driver / plots outside 8 ... 32 (examples/elliptic-multiple/main.cpp:444)
    "adaptive_l1_4_m4_helmholtz": ["--problem", "helmholtz", "--solver", "fishpack", "--min-level", "1", "--max-level", "4", "--nx", "4", "--domain", "-10", "10", "-10", "10", "--refine-box", "-10", "0.5", "-10", "0.5"],
    "adaptive_l1_3_m4_varcoef": ["--problem", "varcoef", "--solver", "fivepoint", "--min-level", "1", "--max-level", "3", "--nx", "4", "--domain", "-10", "10", "-10", "10", "--refine-box", "2", "10", "-3", "10"],
    "single_patch_m64_poisson": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "0", "--max-level", "0", "--nx", "64", "--domain", "0", PI, "0", PI],
    "adaptive_l1_3_m8_varcoef": ["--problem", "varcoef", "--solver", "fivepoint", "--min-level", "1", "--max-level", "3", "--nx", "8", "--domain", "-10", "10", "-10", "10", "--refine-box", "2", "10", "-3", "10"],
}


def main():
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        tmp = "/tmp/golden_%s.bin" % name
        out = subprocess.run([drv] + args + ["--dump", tmp], capture_output=True, text=True, check=True).stdout
        res = [l for l in out.splitlines() if l.startswith("REF_RESULT")][-1]
        D = read_dump(tmp)
        arrays = {}
        for k, v in D.items():
            arrays[k.replace("/", "|")] = np.array(v) if not isinstance(v, str) else np.array(v)
        arrays["args"] = np.array(" ".join(args))
        arrays["result"] = np.array(res[len("REF_RESULT "):])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, res, "%.1f KiB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
