"""CPU-side checks of the drop-in boundary: libdfusion.so loads and exports every symbol include/dfusion.h declares
(no compute calls: there is no GPU here), and the product never reaches into oracle/."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "dfusion.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(df_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dynamicfusion_b200 import build, capi
    build.build()
    lib = capi.load()
    assert capi.MISSING == []
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dfusion.h but not exported"
        assert name in capi.PROTOTYPES, f"{name} has no ctypes prototype in capi.py"
    assert lib.df_version() >= 100
    assert lib.df_error_string(0) == b"success"


def test_params_struct_layouts_agree():
    import ctypes as C
    from dynamicfusion_b200 import capi
    from oracle import orc_pipe
    assert C.sizeof(capi.KinfuParams) == C.sizeof(orc_pipe.KinfuParams)
    p = capi.KinfuParams()
    capi.load().df_kinfu_default_params(C.byref(p), 0)
    q = orc_pipe.default_params(0)
    assert bytes(p) == bytes(orc_pipe.params_from(p))
    for f, _ in capi.KinfuParams._fields_:
        a, b = getattr(p, f), getattr(q, f)
        if hasattr(a, "_length_"):
            assert list(a) == list(b), f
        elif hasattr(a, "_fields_"):
            assert bytes(a) == bytes(b), f
        else:
            assert a == b, f


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under dynamicfusion_b200/ or include/ may import, link or execute it"""
    bad = []
    for path in list((ROOT / "dynamicfusion_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.suffix not in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"}:
            continue
        for line in path.read_text(errors="ignore").splitlines():
            if re.search(r"^\s*(from|import)\s+oracle|from\s+\.\.?\s*oracle|#include\s+[\"<].*orc_|liborc\.so|CDLL\(.*orc", line):
                bad.append((str(path), line.strip()))
    assert not bad, bad


def test_loader_fails_loudly_without_library(monkeypatch, tmp_path):
    import pytest
    from dynamicfusion_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load()
