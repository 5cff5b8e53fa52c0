"""Workload of bench.py's live DRAM-traffic probe: the bench sequence's frames 0..N-1 through the device-resident frame loop, nothing
else.  bench.py runs it ONCE under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:<integrate|ray-cast>` and reads
the per-launch byte counts of the last frames from ncu's CSV (roofline.traffic / roofline_raycast.traffic).  Numbers printed by a
run under ncu are never bench values: this script prints none."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=13)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--max-nodes", type=int, default=2048)
    a = ap.parse_args()
    import torch
    from dynamicfusion_b200 import kinfu as kf, synth
    p = kf.KinFuParams.default_params_dynamicfusion()
    kf.KinFuParams.set_volume(p, a.dim, 1.0)
    p.max_nodes = a.max_nodes
    p.cloud_capacity = 4_000_000
    k = kf.KinFu(p)
    for t in range(a.frames):
        d = torch.from_numpy(synth.umbrella_depth(t, seed=0).view(np.int16).copy()).cuda()
        k(d)
    torch.cuda.synchronize()
    k.close()


if __name__ == "__main__":
    main()
