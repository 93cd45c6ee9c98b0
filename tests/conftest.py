import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        out[k.replace("|", "/")] = str(v) if v.dtype.kind in "US" else v
    return out


def golden_case_args(gold):
    """Parse the ref_driver argument string stored in a fixture into oracle.run kwargs."""
    a = gold["args"].split()
    kw = dict(threshold=1.2, refine_box=None)
    i = 0
    while i < len(a):
        k = a[i]
        if k == "--problem":
            kw["problem_name"] = a[i + 1]; i += 2
        elif k == "--lambda":
            kw["problem_name"] = "helmholtz:%s" % a[i + 1]; i += 2
        elif k == "--solver":
            kw["solver_kind"] = a[i + 1]; i += 2
        elif k == "--min-level":
            kw["min_level"] = int(a[i + 1]); i += 2
        elif k == "--max-level":
            kw["max_level"] = int(a[i + 1]); i += 2
        elif k == "--nx":
            kw["nx"] = int(a[i + 1]); i += 2
        elif k == "--threshold":
            kw["threshold"] = float(a[i + 1]); i += 2
        elif k == "--domain":
            kw["box"] = tuple(float(v) for v in a[i + 1:i + 5]); i += 5
        elif k == "--refine-box":
            kw["refine_box"] = tuple(float(v) for v in a[i + 1:i + 5]); i += 5
        else:
            raise ValueError(k)
    return kw


GOLDEN_CASES = [
    "uniform_l2_m8_poisson",
    "uniform_l1_m16_helmholtz",
    "adaptive_l1_3_m8_poisson",
    "adaptive_tag2_m8_helmholtz_rect",
    "adaptive_l1_3_m8_varcoef",
    "uniform_l2_m8_helmholtz_indefinite",
    "adaptive_l1_4_m4_helmholtz",
    "adaptive_l1_3_m4_varcoef",
    "single_patch_m64_poisson",
]
