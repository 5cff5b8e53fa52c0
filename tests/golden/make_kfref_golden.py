"""Regenerate tests/golden/kfref_golden.json: SHA-256 digests of what the REFERENCE's own CUDA kernels
(/root/reference/kfusion/src/cuda/{tsdf_volume,imgproc,proj_icp}.cu, compiled for the host by oracle/ref_shim into
oracle/_ref/libkfref.so) produce on the seeded scenes of tests/kfref_cases.py.  Needs /root/reference (build container only):

    make -C oracle ref && python tests/golden/make_kfref_golden.py

tests/test_oracle_vs_reference_kernels.py checks the oracle against these digests everywhere (also where the reference and
its host build are absent)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import orc  # noqa: E402
import kfref_cases  # noqa: E402

if __name__ == "__main__":
    orc.build()
    assert orc.reference_available(), "oracle/_ref/libkfref.so missing: run `make -C oracle ref` where /root/reference exists"
    out = {}
    for name, arr in kfref_cases.all_cases(orc, ref=True).items():
        out[name] = {"sha256": kfref_cases.digest(arr), "dtype": str(arr.dtype), "shape": list(arr.shape)}
    (Path(__file__).parent / "kfref_golden.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print(f"wrote {len(out)} digests")
