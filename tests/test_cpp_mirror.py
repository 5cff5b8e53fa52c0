"""The C++ mirror of the reference's public API (include/kfusion/*.hpp + libkfusion.so): a demo.cpp-like program
(tests/cpp/demo_like.cpp) must compile and link against it on the CPU box, and run green on the GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "tests" / "cpp" / "demo_like"


def _build():
    from dynamicfusion_b200 import build
    return build.build_cpp_program(ROOT / "tests" / "cpp" / "demo_like.cpp", EXE)


def test_demo_like_program_compiles_and_links():
    exe = _build()
    assert exe.exists()
    # the mirror library exports the reference's class symbols
    out = subprocess.run(["nm", "-DC", str(ROOT / "dynamicfusion_b200" / "libkfusion.so")], capture_output=True, text=True).stdout
    for sym in ("kfusion::KinFu::operator()", "kfusion::KinFuParams::default_params_dynamicfusion", "kfusion::WarpField::warp",
                "kfusion::cuda::TsdfVolume::integrate", "kfusion::cuda::ProjectiveICP::estimateTransform", "kfusion::cuda::computeDists",
                "kfusion::WarpFieldOptimiser::optimiseWarpData", "kfusion::cuda::DeviceMemory2D::upload"):
        assert sym in out, sym


@pytest.mark.gpu
def test_demo_like_program_runs():
    exe = _build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
