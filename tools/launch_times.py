"""Sum the per-kernel GPU times of an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (development aid):
   python tools/launch_times.py gpurun_out/launches.csv [--last-frames N --per-frame K]"""
import csv
import collections
import sys

rows = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('"')]
acc, cnt = collections.Counter(), collections.Counter()
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    acc[name] += v
    cnt[name] += 1
tot = sum(acc.values())
for n, v in acc.most_common(24):
    print(f"{v:10.1f} us  {100 * v / tot:5.1f} %  x{cnt[n]:5d}  {v / cnt[n]:8.2f} us/launch  {n[:90]}")
print(f"{tot:10.1f} us total")
