"""CPU model of the packed integrate kernel's division / square-root sequences (tests/c/packed_div_check.c): with a correctly
rounded seed they return the IEEE result on the kernel's whole checked domain; with a seed displaced by up to 2 ulp the division only
fails when the divisor's mantissa is all ones and the square root for a handful of operands -- which is why bit-exactness on the GPU is
pinned by the on-device self-test (tests/test_tsdf_gpu.py::test_packed_arithmetic_selftest_on_device), not by this model."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_sequences_are_exact_with_a_correctly_rounded_seed(tmp_path):
    exe = tmp_path / "packed_div_check"
    try:
        has_fma = " fma " in Path("/proc/cpuinfo").read_text()
    except OSError:
        has_fma = False
    subprocess.run(["gcc", "-O2", *(["-mfma"] if has_fma else []), "-o", str(exe), str(ROOT / "tests" / "c" / "packed_div_check.c"), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout
    rows = {ln.split()[0]: dict(zip(ln.split()[1::2], map(int, ln.split()[2::2]))) for ln in r.stdout.strip().splitlines()}
    assert set(rows) == {"division", "tiny_numerators", "square_root", "running_average", "every_mantissa"}
    for name, row in rows.items():
        assert row["n"] > 5_000_000 and row["k0"] == 0, (name, row)
    assert rows["division"]["outside_hard_case"] == 0              # perturbed seeds only fail for all-ones divisor mantissas
    assert rows["tiny_numerators"]["perturbed"] == 0 and rows["running_average"]["perturbed"] == 0
    assert rows["square_root"]["perturbed"] < 100                   # a handful in 3e7, at +-2 ulp mostly


def test_cuda_source_uses_the_modelled_sequences():
    """the C model restates the kernel's sequences; make sure the two cannot drift apart"""
    cu = (ROOT / "dynamicfusion_b200" / "csrc" / "tsdf.cu").read_text()
    for token in ["const f32x2 E = fma2(R0, NZ[h], ONE);", "const f32x2 R1 = fma2(R0, E, R0);", "const f32x2 QX0 = mul2(R1, X[h]), QY0 = mul2(R1, Y[h]);",
                  "const f32x2 RX = fma2(QX0, NZ[h], X[h]), RY = fma2(QY0, NZ[h], Y[h]);", "const f32x2 QX = fma2(R1, RX, QX0), QY = fma2(R1, RY, QY0);",
                  "const f32x2 S = mul2(N2, RS), H = mul2(RS, HALF);", "const f32x2 E = fma2(mul2(S, MINUS1), S, N2);", "const f32x2 S1 = fma2(E, H, S);",
                  "rcp.approx.ftz.f32", "rsqrt.approx.ftz.f32", "fma.rn.f32x2", "mul.rn.f32x2", "add.rn.f32x2"]:
        assert token in cu, token
    c = (ROOT / "tests" / "c" / "packed_div_check.c").read_text()
    for token in ["fmaf(r0, -z, 1.f)", "fmaf(r0, e, r0)", "r1 * x", "fmaf(q0, -z, x)", "fmaf(r1, rem, q0)", "fmaf(-s, s, v)", "fmaf(e, h, s)"]:
        assert token in c, token
