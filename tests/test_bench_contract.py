"""CPU tests of bench.py's contract: the reference arm prints one JSON line with the agreed keys (bounded sample of
the reference's own CPU path, oracle/_ref/ref_driver), and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_driver not built (needs /root/reference)")
@pytest.mark.parametrize("extra", [[], ["--adaptive", "0", "5", "--cpu-level", "3"]])
def test_reference_arm_line(extra):
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-level", "2"] + extra,
                         capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "DOFs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
