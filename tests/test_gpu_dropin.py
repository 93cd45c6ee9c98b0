"""Drop-in check on the GPU box: oracle/_ref/dropin_driver is the UNMODIFIED reference compiled
together with include/EllipticForestB200.hpp (the reference-side binding).  In one process it runs
the reference's CPU HPSAlgorithm and the B200 subclass on identical p4est meshes and reports the
relative max-norm difference of every node's operators and vectors.  Tolerance 1e-10 (north_star)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")
PI = "3.141592653589793"

CASES = {
    "uniform_poisson_m16": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "3", "--max-level", "3", "--nx", "16", "--domain", "0", PI, "0", PI],
    "adaptive_single_m8": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "0", "--max-level", "4", "--nx", "8", "--domain", "-10", "10", "-10", "10"],
    "adaptive_tag2_helmholtz": ["--problem", "helmholtz", "--solver", "fishpack", "--min-level", "1", "--max-level", "4", "--nx", "8", "--domain", "0", "2", "0", "1",
                                "--refine-box", "1.0", "2.0", "0.5", "1.0"],
    "homogeneous_cached": ["--problem", "poisson", "--solver", "fishpack", "--min-level", "2", "--max-level", "2", "--nx", "8", "--domain", "0", PI, "0", PI,
                           "--homogeneous", "1", "--cache", "1"],
    "varcoef_fivepoint_m8": ["--problem", "varcoef", "--solver", "fivepoint", "--min-level", "1", "--max-level", "3", "--nx", "8", "--domain", "-10", "10", "-10", "10",
                             "--refine-box", "2", "10", "-3", "10"],
    # the binding's opt-in threaded sampling of the std::function callbacks (sampling_threads = 4)
    "varcoef_fivepoint_m8_threads4": ["--problem", "varcoef", "--solver", "fivepoint", "--min-level", "1", "--max-level", "3", "--nx", "8", "--domain", "-10", "10", "-10", "10",
                                      "--refine-box", "2", "10", "-3", "10", "--threads", "4"],
    "robin_root_m8": ["--problem", "helmholtz", "--solver", "fishpack", "--min-level", "2", "--max-level", "2", "--nx", "8", "--domain", "0", PI, "0", PI, "--robin", "1"],
}


@pytest.mark.parametrize("case", list(CASES))
def test_reference_driver_with_b200_subclass(case):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/dropin_driver not built (needs /root/reference at build time)")
    p = subprocess.run([DRIVER] + CASES[case], capture_output=True, text=True, timeout=600)
    line = [l for l in p.stdout.splitlines() if l.startswith("DROPIN_RESULT")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(line[-1][len("DROPIN_RESULT "):])
    assert "error" not in res, res
    print(case, res)
    assert res["structure_ok"]
    for k in ("T", "S", "H", "X", "h", "w", "g", "u"):
        assert res[k] < 1e-10, (k, res[k])
