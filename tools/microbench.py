"""Per-kernel CUDA-event timings at the bench configuration (512^3, 640x480) -- development aid, not the bench."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dynamicfusion_b200 import host, synth  # noqa: E402


def timeit(fn, iters=10, flush=None):
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--frames", type=int, default=3)
    a = ap.parse_args()
    K = synth.DEFAULT_K
    dim = a.dim
    vol = host.TsdfVolume((dim, dim, dim))
    vol.setTruncDist(0.04); vol.setMaxWeight(64); vol.setSize((1.0, 1.0, 1.0)); vol.setPose(synth.volume_pose(1.0))
    vol.setRaycastStepFactor(0.75); vol.setGradientDeltaFactor(0.5); vol.clear()
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")      # 256 MiB > 126 MB L2
    depth = synth.umbrella_depth(0)
    d = host.u16_to_device(depth)
    dists = host.computeDists(d, K)
    n_upd = torch.zeros(1, dtype=torch.int64, device="cuda")
    pose = host.identity_pose()
    vol.integrate(dists, pose, K, n_upd)
    torch.cuda.synchronize()
    nupd = int(n_upd.item())
    out = {"dim": dim, "n_upd": nupd}
    med, best = timeit(lambda: vol.integrate(dists, pose, K), flush=flush)
    bytes_int = 8 * nupd + 2 * 640 * 480
    out["integrate_ms"] = med; out["integrate_best_ms"] = best; out["integrate_GBs"] = bytes_int / med / 1e6
    out["integrate_dense_GBs"] = (8 * dim ** 3) / med / 1e6
    med, best = timeit(lambda: vol.raycast(pose, K, 640, 480), flush=flush)
    out["raycast_ms"] = med; out["raycast_best_ms"] = best
    med, best = timeit(lambda: vol.clear(), flush=flush)
    out["clear_ms"] = med; out["clear_GBs"] = 4 * dim ** 3 / med / 1e6
    med, best = timeit(lambda: host.computeDists(d, K))
    out["compute_dists_ms"] = med
    med, best = timeit(lambda: host.depthBilateralFilter(d, 7, 4.5, 0.04))
    out["bilateral_ms"] = med
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
