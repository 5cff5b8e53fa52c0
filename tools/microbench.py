"""Per-kernel CUDA-event timings at the bench configuration (512^3, 640x480) -- development aid, not the bench.
   python tools/microbench.py [--dim 512] [--integrate-impl 1|2] [--zchunk N] [--pipeline]"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--integrate-impl", type=int, default=3)
    ap.add_argument("--zchunk", type=int, default=0)
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    os.environ["DF_INTEGRATE_IMPL"] = str(a.integrate_impl)
    if a.zchunk:
        os.environ["DF_INTEGRATE_ZCHUNK"] = str(a.zchunk)
    import torch
    from dynamicfusion_b200 import host, kinfu as kf, synth

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

    K = synth.DEFAULT_K
    dim = a.dim
    out = {"dim": dim, "integrate_impl": a.integrate_impl, "zchunk": a.zchunk}
    if a.pipeline:
        p = kf.KinFuParams.default_params_dynamicfusion()
        kf.KinFuParams.set_volume(p, dim, 1.0)
        p.max_nodes = 2048; p.cloud_capacity = 4_000_000; p.flags = kf.STAGE_TIMING
        k = kf.KinFu(p)
        acc, n = {}, 0
        for t in range(a.frames):
            d = torch.from_numpy(synth.umbrella_depth(t).view(np.int16).copy()).cuda()
            k(d)
            if a.trace and (t < 6 or t % 5 == 0):
                i, sm = k.info(), k.stage_ms()
                print(f"frame {t:3d} lm {i['lm_iters']} pcg {i['pcg_iters']:4d} cloud {i['cloud_points']:7d} n_upd {i['n_updated']:9d} "
                      f"solve {sm['solve']:.3f} icp {sm['icp']:.3f} integ {sm['integrate']:.3f} extract {sm['extract']:.3f} total {sum(sm.values()):.3f}", flush=True)
            if t >= 3:
                for name, v in k.stage_ms().items():
                    acc[name] = acc.get(name, 0.0) + v
                n += 1
        out["stage_ms"] = {kk: round(v / n, 4) for kk, v in acc.items()}
        out["frame_ms_sum"] = round(sum(out["stage_ms"].values()), 4)
        out["info"] = k.info()
        out["solve_stats"] = [float(v) for v in k.buffer("solve_stats")]
        k.close()
        print(json.dumps(out, indent=1))
        return
    vol = host.TsdfVolume((dim, dim, dim))
    vol.setTruncDist(0.04); vol.setMaxWeight(64); vol.setSize((1.0, 1.0, 1.0)); vol.setPose(synth.volume_pose(1.0))
    vol.setRaycastStepFactor(0.75); vol.setGradientDeltaFactor(0.5); vol.clear()
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")      # 256 MiB > 126 MB L2
    d = host.u16_to_device(synth.umbrella_depth(0))
    dists = host.computeDists(d, K)
    n_upd = torch.zeros(1, dtype=torch.int64, device="cuda")
    pose = host.identity_pose()
    vol.integrate(dists, pose, K, n_upd)
    torch.cuda.synchronize()
    nupd = int(n_upd.item())
    out["n_upd"] = nupd
    med, best = timeit(lambda: vol.integrate(dists, pose, K), flush=flush)
    bytes_int = 8 * nupd + 2 * 640 * 480
    out["integrate_ms"] = med; out["integrate_best_ms"] = best; out["integrate_GBs"] = bytes_int / med / 1e6
    med, best = timeit(lambda: vol.raycast(pose, K, 640, 480), flush=flush)
    out["raycast_ms"] = med; out["raycast_best_ms"] = best
    cap = 4_000_000
    med, best = timeit(lambda: vol.fetchCloud(cap), flush=flush, iters=5)
    out["extract_cloud_ms"] = med
    pts, cnt = vol.fetchCloud(cap)
    n = int(cnt.item()); out["cloud_points"] = n
    med, best = timeit(lambda: vol.fetchNormals(pts, n), flush=flush, iters=5)
    out["extract_normals_ms"] = med
    med, best = timeit(lambda: vol.clear(), flush=flush)
    out["clear_ms"] = med; out["clear_GBs"] = 4 * dim ** 3 / med / 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
