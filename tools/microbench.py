"""Per-kernel CUDA-event timings at the bench configuration (512^3, 640x480) -- development aid, not the bench.
   python tools/microbench.py [--dim 512] [--integrate-impl 1|3|5] [--zchunk N] [--pipeline]"""
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
    ap.add_argument("--integrate-impl", type=int, default=5)
    ap.add_argument("--zchunk", type=int, default=0)
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--trace-all", action="store_true", help="with --trace: a line for every frame")
    ap.add_argument("--warped", action="store_true", help="with --pipeline: DF_KINFU_WARPED_INTEGRATE (per-voxel warped fusion, SURVEY 8f(1))")
    ap.add_argument("--weight-scale", type=float, default=100.0)
    ap.add_argument("--hd", action="store_true", help="config C4: 1280x720 depth, volume edge 1.5 m (use with --dim 768)")
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
        p.max_nodes = 2048; p.cloud_capacity = 4_000_000; p.flags = kf.STAGE_TIMING | (kf.WARPED_INTEGRATE if a.warped else 0)
        p.fusion_weight_scale = a.weight_scale
        k = kf.KinFu(p)
        acc, n = {}, 0
        for t in range(a.frames):
            d = torch.from_numpy(synth.umbrella_depth(t).view(np.int16).copy()).cuda()
            k(d)
            if a.trace and (t < 6 or t % 5 == 0 or a.trace_all):
                i, sm = k.info(), k.stage_ms()
                st = [float(v) for v in k.buffer("solve_stats")]
                print(f"frame {t:3d} lm {i['lm_iters']} pcg {i['pcg_iters']:4d} cloud {i['cloud_points']:7d} n_upd {i['n_updated']:9d} n_warped {i['n_warped']:9d} "
                      f"solve {sm['solve']:.3f} icp {sm['icp']:.3f} integ {sm['integrate']:.3f} extract {sm['extract']:.3f} total {sum(sm.values()):.3f} nnz {int(st[6])} lm-diag {st[7]}", flush=True)
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
    cols, rows, size = (1280, 720, 1.5) if a.hd else (640, 480, 1.0)
    if a.hd:
        K = (K[0] * 2, K[1] * 2, 640.0, 360.0)
    vol = host.TsdfVolume((dim, dim, dim), track_activity=True)
    vol.setTruncDist(0.04); vol.setMaxWeight(64); vol.setSize((size, size, size)); vol.setPose(synth.volume_pose(size))
    vol.setRaycastStepFactor(0.75); vol.setGradientDeltaFactor(0.5); vol.clear()
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")      # 256 MiB > 126 MB L2
    d = host.u16_to_device(synth.umbrella_depth(0, cols=cols, rows=rows, K=K))
    dists = host.computeDists(d, K)
    n_upd = torch.zeros(1, dtype=torch.int64, device="cuda")
    pose = host.identity_pose()
    vol.integrate(dists, pose, K, n_upd)
    torch.cuda.synchronize()
    nupd = int(n_upd.item())
    out["n_upd"] = nupd
    med, best = timeit(lambda: vol.integrate(dists, pose, K), flush=flush)
    bytes_int = 8 * nupd + 2 * cols * rows
    out["integrate_ms"] = med; out["integrate_best_ms"] = best; out["integrate_GBs"] = bytes_int / med / 1e6
    med, best = timeit(lambda: vol.raycast(pose, K, cols, rows), flush=flush)
    out["raycast_ms"] = med; out["raycast_best_ms"] = best
    cap = 4_000_000
    med, best = timeit(lambda: vol.fetchCloud(cap), flush=flush, iters=5)
    out["extract_cloud_ms"] = med
    pts, cnt = vol.fetchCloud(cap)
    n = int(cnt.item()); out["cloud_points"] = n
    med, best = timeit(lambda: vol.fetchNormals(pts, n), flush=flush, iters=5)
    out["extract_normals_ms"] = med
    # occupancy at 8^3-brick granularity (dense-equivalent vs occupied-brick bytes)
    v = vol.data_.view(dim // 8, 8, dim // 8, 8, dim // 8, 8)
    observed = ((v >> 16) & 0xffff) != 0
    bricks_obs = int(observed.any(dim=5).any(dim=3).any(dim=1).sum().item())
    surf = observed & ((v & 0xffff) != 0x3c00)
    bricks_surf = int(surf.any(dim=5).any(dim=3).any(dim=1).sum().item())
    out["dense_bytes"] = 4 * dim ** 3
    out["bricks_total"] = (dim // 8) ** 3
    out["bricks_observed"] = bricks_obs; out["observed_brick_bytes"] = bricks_obs * 2048
    out["bricks_with_surface"] = bricks_surf; out["surface_brick_bytes"] = bricks_surf * 2048
    out["activity_fraction"] = float((vol.activity_ != 0).float().mean().item())
    med, best = timeit(lambda: vol.clear(), flush=flush)
    out["clear_ms"] = med; out["clear_GBs"] = 4 * dim ** 3 / med / 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
