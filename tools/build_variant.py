"""Build an instrumented / A-B variant of the library next to the product one: libdfusion_<tag>.so, loaded with DF_LIB_VARIANT=<tag>.
    python tools/build_variant.py lmprof -DDF_LM_PROFILE        (per-phase clock64 accounting of the LM/PCG kernel, printed per CTA)
Only the sources named with --src (default: solve.cu) are recompiled with the extra flags; the other objects come from the product build."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dynamicfusion_b200 import build as B

tag = sys.argv[1]
srcs = ["solve.cu"]
extra = []
it = iter(sys.argv[2:])
for a in it:
    if a == "--src": srcs = next(it).split(",")
    else: extra.append(a)
B.build()
objdir = B.HERE / "build"
objs = []
for src in B.sources():
    obj = objdir / (src.stem + ".o")
    if src.name in srcs:
        obj = objdir / f"{src.stem}_{tag}.o"
        subprocess.run(["/usr/local/cuda/bin/nvcc", *B.NVCC_FLAGS, *extra, "-x", "cu", "-c", str(src), "-o", str(obj)], check=True)
    objs.append(obj)
out = B.HERE / f"libdfusion_{tag}.so"
subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-o", str(out), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(out)
