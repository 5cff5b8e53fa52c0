"""Summarise an ncu report of ONE kernel by source line without the GUI: joins `ncu --page source --csv` (per-SASS-instruction counts)
with `nvdisasm -g` line info of the cubin inside libdfusion.so.
   python tools/ncu_by_line.py gpurun_out/fusion_v1.ncu-rep fusion integrate_warped [--top 25]
Prints the kernel's headline metrics (raw page) and the share of executed warp instructions / stall samples per source line."""
import argparse
import collections
import csv
import io
import re
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("cubin", help="translation unit name inside libdfusion.so, e.g. fusion")
    ap.add_argument("kernel", help="substring of the kernel's mangled name")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--by-samples", action="store_true", help="rank the lines by stall samples instead of executed instructions")
    a = ap.parse_args()
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__occupancy_limit_registers"]
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"  {h:70s} {v} {u}")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", a.cubin + ".sm_100a.cubin", str(ROOT / "dynamicfusion_b200" / "libdfusion.so")], cwd=td, capture_output=True)
        sass = subprocess.run(["nvdisasm", "-g", "-c", a.cubin + ".sm_100a.cubin"], cwd=td, capture_output=True, text=True).stdout.splitlines()
    in_k, cur, addr2loc = False, None, {}
    for ln in sass:
        if ln.startswith(".text."):
            in_k = a.kernel in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
        if m and cur is not None:
            addr2loc[int(m.group(1), 16)] = cur
    src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
    h = src[1]
    iA, iI, iT, iS = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    base, tot, samples = None, 0, 0
    inst, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in src[2:]:
        if len(r) <= iT or not r[iA]:
            continue
        if r[iA] == "Address":                      # a report with several kernels: the next kernel's table begins -- only the first is summarised
            break
        ad = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
        base = ad if base is None else base
        loc = addr2loc.get(ad - base, ("?", 0))
        n, t, s = int(r[iI] or 0), int(r[iT] or 0), int(r[iS] or 0)
        inst[loc] += n; thr[loc] += t; smp[loc] += s
        tot += n; samples += s
    print(f"  executed warp instructions {tot}, stall samples {samples}")
    print("  share of warp instructions | share of samples | avg active threads | source line")
    ranked = [(loc, inst[loc]) for loc, _ in smp.most_common(a.top)] if a.by_samples else inst.most_common(a.top)
    for loc, n in ranked:
        text = ""
        f = next((p for p in (ROOT / "dynamicfusion_b200" / "csrc").rglob(loc[0])), None) if loc[0] != "?" else None
        if f:
            lines = f.read_text().splitlines()
            text = lines[loc[1] - 1].strip()[:100] if 0 < loc[1] <= len(lines) else ""
        print(f"  {n / max(tot, 1):6.1%}  {smp[loc] / max(samples, 1):6.1%}  {thr[loc] / max(n, 1):5.1f}  {loc[0]}:{loc[1]}  {text}")


if __name__ == "__main__":
    main()
