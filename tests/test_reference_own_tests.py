"""The reference's OWN test files, compiled UNCHANGED from /root/reference/tests with a minimal GoogleTest stand-in
(tests/cpp/refcompat/gtest/gtest.h; GoogleTest is absent from this image):

  * tests/utils/test_quaternion.cc, test_dual_quaternion.cc  -- header-only; built twice, against this repo's headers (include/)
    and against the reference's own headers: every test must get the SAME verdict in both builds.  Two of their assertions are
    wrong in the reference itself (QuaternionTest.rodrigues, DualQuaternionTest.DualQuaternionConstructor -- SURVEY section 4
    "self-inconsistent"); they fail identically on both sides, everything else passes.
  * tests/warp_test.cpp  -- drives WarpField / WarpFieldOptimiser exactly as the reference's CI would; runs on the GPU.
    EnergyDataSingleVertexTest (1e-5), MultipleNodesTest and NonRigidTest (1e-3) pass at the reference's own tolerances.
    EnergyDataRigidTest and WarpAndReverseTest (the same geometry: five collinear vertices on the cube diagonal against the eight
    symmetric corner nodes, i.e. four distinct weight classes for five equations per axis) miss the asserted 1e-3 by the exact
    least-squares residual, 6e-3: the asserted values are unattainable for ANY minimiser of the reference's energy (DESIGN.md
    section 5; tests/test_stages_gpu.py pins the solver to the float64 least-squares optimum of these scenarios instead).

The binaries are built where /root/reference is mounted (dynamicfusion_b200.build.build_reference_tests, also called by
__graft_entry__.build()) and travel to the GPU box with the snapshot."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BUILD = ROOT / "tests" / "cpp" / "_build"


def _verdicts(exe):
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    ok = set(re.findall(r"\[       OK \] (\S+)", r.stdout))
    bad = set(re.findall(r"\[  FAILED  \] (\S+\.\S+)", r.stdout))
    return ok, bad, r.stdout


def _ensure_built():
    from dynamicfusion_b200 import build
    if build.REF_TESTS_DIR.exists():
        build.build_reference_tests()


def test_reference_quaternion_tests_same_verdicts_on_both_header_sets():
    _ensure_built()
    if not (BUILD / "ref_test_quaternion").exists():
        pytest.skip("reference tests were not built (needs /root/reference at build time)")
    expected_fail = {"test_quaternion": {"QuaternionTest.rodrigues"}, "test_dual_quaternion": {"DualQuaternionTest.DualQuaternionConstructor"}}
    expected_pass = {"test_quaternion": {"QuaternionTest.encodeRotation", "QuaternionTest.quat_product", "QuaternionTest.dotProduct",
                                         "QuaternionTest.normalize", "QuaternionTest.rotate", "QuaternionTest.normal"},
                     "test_dual_quaternion": {"DualQuaternionTest.canSplitOperations", "DualQuaternionTest.isAssociative"}}
    for t in ("test_quaternion", "test_dual_quaternion"):
        ok_m, bad_m, _ = _verdicts(BUILD / f"mine_{t}")
        ok_r, bad_r, _ = _verdicts(BUILD / f"ref_{t}")
        assert ok_m == ok_r and bad_m == bad_r, (t, ok_m ^ ok_r, bad_m ^ bad_r)
        assert ok_m == expected_pass[t] and bad_m == expected_fail[t], (t, ok_m, bad_m)


@pytest.mark.gpu
def test_reference_warp_test_on_the_gpu():
    exe = BUILD / "mine_warp_test"
    if not exe.exists():
        pytest.skip("tests/cpp/_build/mine_warp_test was not built (needs /root/reference at build time)")
    ok, bad, out = _verdicts(exe)
    assert ok | bad == {f"WARP_FIELD_TEST.{n}" for n in ("EnergyDataSingleVertexTest", "EnergyDataRigidTest", "WarpAndReverseTest",
                                                         "MultipleNodesTest", "NonRigidTest")}, out[-2000:]
    assert ok == {"WARP_FIELD_TEST.EnergyDataSingleVertexTest", "WARP_FIELD_TEST.MultipleNodesTest", "WARP_FIELD_TEST.NonRigidTest"}, out[-3000:]
    assert bad == {"WARP_FIELD_TEST.EnergyDataRigidTest", "WARP_FIELD_TEST.WarpAndReverseTest"}, out[-3000:]
    # the over-determined scenarios: the first failing coordinate is off by the least-squares residual (6e-3), not by more
    diffs = [abs(float(m.group(1)) - float(m.group(2))) for m in re.finditer(r"first: (\S+)\n second: (\S+)", out)]
    assert len(diffs) == 2 and all(4e-3 < d < 8e-3 for d in diffs), diffs


@pytest.mark.gpu
def test_reference_ceres_warp_test_on_the_gpu():
    """tests/ceres_warp_test.cpp drives WarpField::energy_data (Ceres in the reference, the device LM/PCG here) -- unchanged source.
    WarpAndReverseTest asserts the over-determined five-vertex geometry to 1e-3 and misses by the least-squares residual, as above."""
    exe = BUILD / "mine_ceres_warp_test"
    if not exe.exists():
        pytest.skip("tests/cpp/_build/mine_ceres_warp_test was not built (needs /root/reference at build time)")
    ok, bad, out = _verdicts(exe)
    assert ok == {"CERES_WARP_TEST.EnergyDataSingleVertexTest", "CERES_WARP_TEST.EnergyDataRigidTest"}, out[-3000:]
    assert bad == {"CERES_WARP_TEST.WarpAndReverseTest"}, out[-3000:]
    diffs = [abs(float(m.group(1)) - float(m.group(2))) for m in re.finditer(r"first: (\S+)\n second: (\S+)", out)]
    assert len(diffs) == 1 and 4e-3 < diffs[0] < 8e-3, diffs
