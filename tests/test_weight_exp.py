"""node_weighting()'s polynomial fast path (warp_common.cuh) against the host libm -- the exp() the oracle calls -- for EVERY float in its
interval [-0.5, -0]: 1,056,964,609 values, a few seconds of one core."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_weight_polynomial_equals_libm_exp_for_every_float(tmp_path):
    exe = tmp_path / "weight_exp_check"
    try:
        has_fma = " fma " in Path("/proc/cpuinfo").read_text()
    except OSError:
        has_fma = False
    subprocess.run(["gcc", "-O2", *(["-mfma"] if has_fma else []), "-o", str(exe), str(ROOT / "tests" / "c" / "weight_exp_check.c"), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert r.stdout.strip().endswith("n 1056964609 mismatches 0"), r.stdout


def test_cuda_source_uses_the_checked_constants():
    """the C check restates the polynomial; make sure the two cannot drift apart"""
    cu = (ROOT / "dynamicfusion_b200" / "csrc" / "warp_common.cuh").read_text()
    c = (ROOT / "tests" / "c" / "weight_exp_check.c").read_text()
    for token in ["1.0 / 6227020800.0", "1.0 / 479001600.0", "1.0 / 39916800.0", "1.0 / 3628800.0", "1.0 / 362880.0", "1.0 / 40320.0", "1.0 / 5040.0",
                  "1.0 / 720.0", "1.0 / 120.0", "1.0 / 24.0", "1.0 / 6.0", "0.77880078307140486825", "+ 0.25"]:
        assert token in cu and token in c, token
    assert cu.count("__fma_rn(p, r,") == 13 and c.count("fma(p, r,") == 13
