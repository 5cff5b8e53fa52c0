#!/usr/bin/env python
"""bench.py -- frames/s of the DynamicFusion per-frame hot path at BASELINE.json's quoted configuration
(configs[1]: synthetic "umbrella" sequence, 640x480 depth, 512^3 TSDF over 1 m^3, ~2k warp nodes, full
preprocess -> ICP -> raycast -> k-NN/DQB warp -> data-term solve -> warp -> project/remove -> integrate -> extract ->
raycast loop on one B200).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one frame.  N > 1 (under torchrun, one rank per GPU): every rank runs an independent sequence (seed = rank,
config 5): weak scaling, no data-path collective; timing is the max over ranks.

One JSON line is printed by rank 0.  `value` = frames/s with the depth frames already resident in HBM; `e2e` = the same
through the reference-facing call with HOST depth buffers (df_kinfu_process_host: H2D of the frame and D2H of the
ICP status + pose inside the timed region); `roofline` = the integrate kernel's algorithmic bytes / its CUDA-event
duration against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle's restated loop on the same workload.
--impl reference times that CPU restatement (the reference itself cannot be built here: CUDA 12.9 dropped texture
references, OpenCV/Opt/Terra/Ceres are absent -- see DESIGN.md) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DIM, SIZE, COLS, ROWS, MAX_NODES = 512, 1.0, 640, 480, 2048
WORKLOAD = "C2 synthetic umbrella sequence: 640x480 u16 depth, 512^3 TSDF / 1 m^3, ~2k warp nodes, full per-frame loop"
METRIC = "frames/sec @512^3 TSDF, 640x480 depth (full warp+integrate+raycast loop)"


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        # the sampler is started before the warm-up frames (nvidia-smi needs ~0.3 s to emit its first line); only the samples that
        # arrived between mark_begin() and mark_end() -- i.e. while the timed frames were executing -- are reported.  A timed
        # region shorter than the sampling period falls back to the samples within 0.25 s around it and says so.
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        note = None
        if not inside and self.t0 is not None:
            inside = [r for (t, r) in self.rows if self.t0 - 0.25 <= t <= (self.t1 or t) + 0.25]
            note = "timed region shorter than the sampling period: samples within 0.25 s of it"
        sm = [float(r[0]) for r in inside if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def make_frames(n: int, seed: int):
    from dynamicfusion_b200 import synth
    return np.stack([synth.umbrella_depth(t, seed=seed) for t in range(n)])


def cpu_params():
    from oracle import orc_pipe
    p = orc_pipe.default_params(0, dim=DIM, size=SIZE)
    p.max_nodes = MAX_NODES
    p.cloud_capacity = 4_000_000
    return p


def run_cpu(frames: np.ndarray, warm: int, steps: int):
    """the oracle's restated per-frame loop on the host cores; frame 0 initialises, then `warm` untimed, `steps` timed"""
    from oracle import orc, orc_pipe
    orc.build()
    k = orc_pipe.KinFu(cpu_params())
    k(frames[0])
    for t in range(1, 1 + warm):
        k(frames[t])
    t0 = time.perf_counter()
    for t in range(1 + warm, 1 + warm + steps):
        k(frames[t])
    dt = time.perf_counter() - t0
    info = k.info()
    k.close()
    return dt, info


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args, rank, world):
    if rank != 0:
        return
    budget_s = 150.0
    est = 1.3                                   # s/frame measured on 8 vCPUs; re-estimated from the first timed frame below
    steps = max(1, min(args.steps, int(budget_s / est)))
    warm = min(args.warmup, 2)
    frames = make_frames(1 + warm + steps, 0)
    dt, info = run_cpu(frames, warm, steps)
    fps = steps / dt
    cores = host_threads()
    sample = f"{steps} timed frames (+1 init, +{warm} warm-up) of the same 512^3 workload, OpenMP on {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f16/u16 volume)",
            "data": "synthetic", "config": {"workload": WORKLOAD, "nodes": info["nodes"], "note": "CPU restatement (oracle port): the reference's own "
                                            "CUDA/Opt/OpenCV build is not possible in this image"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--roofline-frames", type=int, default=12)
    ap.add_argument("--no-warped", action="store_true", help="skip the short pass through the per-voxel warped fusion variant (SURVEY 8f(1))")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from dynamicfusion_b200 import distrib, kinfu as kf

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    distrib.init("nccl", device)

    def barrier():
        distrib.barrier(device)

    K, W = args.steps, args.warmup
    nframes = 1 + W + K
    frames = make_frames(nframes, seed=distrib.sequence_seed(rank))   # independent sequence per rank (config 5)
    frames_i16 = torch.from_numpy(frames.view(np.int16))
    frames_dev = frames_i16.cuda()
    frames_pinned = frames_i16.pin_memory()

    def params(flags=0):
        p = kf.KinFuParams.default_params_dynamicfusion()
        kf.KinFuParams.set_volume(p, DIM, SIZE)
        p.max_nodes = MAX_NODES
        p.cloud_capacity = 4_000_000
        p.flags = flags
        return p

    def timed(run_frame):
        """frame 0 + W warm-up frames untimed, then exactly K frames between barrier+sync, CUDA events on the launching stream"""
        sampler = ClockSampler(local_rank)
        sampler.start()
        k = kf.KinFu(params())
        ok = 0
        for t in range(1 + W):
            run_frame(k, t)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.mark_begin()
        e0.record()
        for t in range(1 + W, 1 + W + K):
            ok += run_frame(k, t)
        e1.record()
        barrier()
        sampler.mark_end()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        info = k.info()
        k.close()
        ms, _total, ok = distrib.aggregate(ms, ok, device)       # max time over ranks; ok = fewest fused frames on any rank
        return ms, ok, info, clocks

    pitch = COLS * 2
    # value: inputs already resident in HBM
    ms_dev, ok_dev, info, clocks = timed(lambda k, t: k.lib.df_kinfu_process_device(k.h, frames_dev[t].data_ptr(), pitch))
    # e2e: the reference-facing call with HOST buffers (pinned), H2D + D2H inside the timed region
    ms_e2e, ok_e2e, _, clocks_e2e = timed(lambda k, t: k.lib.df_kinfu_process_host(k.h, frames_pinned[t].data_ptr(), pitch))
    assert ok_dev == K and ok_e2e == K, f"tracking was lost during the timed region ({ok_dev}/{ok_e2e} of {K} frames fused)"

    # roofline of the dominant kernel (integrate): per-stage CUDA events + voxels written, on a separate short pass
    k = kf.KinFu(params(kf.STAGE_TIMING))
    stage_acc, nupd_acc, nroof = {}, 0, 0
    for t in range(min(nframes, 3 + args.roofline_frames)):
        k.lib.df_kinfu_process_device(k.h, frames_dev[t].data_ptr(), pitch)
        if t >= 3:
            for name, v in k.stage_ms().items():
                stage_acc[name] = stage_acc.get(name, 0.0) + v
            nupd_acc += k.info()["n_updated"]
            nroof += 1
    # ray-cast roofline (the metric names integrate + ray-cast): re-run the frame's last ray-cast on the final volume with the counting
    # instantiation of the kernel (df_raycast_points_stats) to get U = unique voxels read; its time is the stage's CUDA-event time
    raycast_info = None
    try:
        import ctypes as C
        from dynamicfusion_b200 import capi, host
        ptr, pitch_, c_, r_ = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        capi.check(k.lib.df_kinfu_get_buffer(k.h, kf.BUF["volume"], C.byref(ptr), C.byref(pitch_), C.byref(c_), C.byref(r_)))
        view = host.TsdfVolume.__new__(host.TsdfVolume)
        view.device, view.activity_, view._ws, view._proj_ws = device, None, None, None
        view.dims_ = np.array([DIM] * 3, np.int32)
        view.size_ = np.array([SIZE] * 3, np.float32)
        view.trunc_dist_ = max(0.04, 2.1 * SIZE / DIM)
        view.max_weight_ = 64
        view.pose_ = (np.eye(3, dtype=np.float32), np.array([-SIZE / 2, -SIZE / 2, 0.5], np.float32))
        view.raycast_step_factor_, view.gradient_delta_factor_ = 0.75, 0.5
        view._vol = lambda: capi.make_volume(ptr.value, view.dims_, view.getVoxelSize(), view.trunc_dist_, view.max_weight_)
        st = view.raycast_stats(k.getCameraPose(), (570.342, 570.342, 320.0, 240.0), COLS, ROWS)
        raycast_info = {kk: st[kk] for kk in ("unique_voxels", "hit_rays", "march_samples", "algorithmic_bytes")}
    except Exception as e:                                                      # informational: never lose the headline line
        raycast_info = {"error": repr(e)}
    k.close()
    stage_ms = {n: v / max(nroof, 1) for n, v in stage_acc.items()}
    n_upd = nupd_acc / max(nroof, 1)
    peak, peak_src = measured_peaks()
    alg_bytes = 8.0 * n_upd + 2.0 * COLS * ROWS                # 4 B read + 4 B write per updated voxel + the fp16 dists image
    integ_ms = stage_ms.get("integrate", float("nan"))
    achieved = alg_bytes / (integ_ms * 1e-3) / 1e9 if integ_ms and integ_ms > 0 else float("nan")
    traffic = None
    tf = ROOT / "profiles" / "integrate_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    total_frames = K * world
    fps = total_frames / (ms_dev * 1e-3)
    fps_e2e = total_frames / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f16/u16 volume)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes": info["nodes"], "cloud_points": info["cloud_points"], "knn": 8,
                   "solver": "LM 5 x PCG 100 (early-out)", "sequences": world, "parallelism": f"{world} independent sequences" if world > 1 else "single sequence",
                   "l2": "working set (512 MiB volume, re-read every frame) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": COLS * ROWS * 2, "d2h_bytes_per_step": 52,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": int(info["launches"]) * K,
        "clocks": clocks,
        "roofline": {"kernel": {"1": "integrate_kernel<4>", "2": "integrate_kernel_v2<4>"}.get(os.environ.get("DF_INTEGRATE_IMPL", "3"), "integrate_kernel_v3"), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "voxels_written_per_launch": n_upd, "kernel_ms": integ_ms,
                     "dense_upper_bound_bytes": 8.0 * DIM ** 3 + 2.0 * COLS * ROWS},
        "stage_ms": stage_ms,
    }
    if raycast_info and "error" not in raycast_info:
        rc_ms = stage_ms.get("raycast_prev", float("nan"))
        rc_ach = raycast_info["algorithmic_bytes"] / (rc_ms * 1e-3) / 1e9 if rc_ms and rc_ms > 0 else float("nan")
        line["roofline_raycast"] = {"kernel": "raycast_points_kernel", "bound": "hbm", "achieved": rc_ach, "peak": peak, "unit": "GB/s",
                                    "frac": rc_ach / peak if peak else None, "traffic": None, "kernel_ms": rc_ms,
                                    "algorithmic_bytes_per_launch": raycast_info["algorithmic_bytes"], "unique_voxels_read": raycast_info["unique_voxels"],
                                    "hit_rays": raycast_info["hit_rays"], "march_samples": raycast_info["march_samples"],
                                    "uncached_upper_bound_bytes": 4 * (raycast_info["march_samples"] + 64 * raycast_info["hit_rays"]) + 32 * COLS * ROWS,
                                    "note": "U counted by the counting instantiation of the same kernel on the last frame's volume and pose; kernel_ms = "
                                            "the stage's CUDA-event time averaged over the roofline frames"}
    elif raycast_info:
        line["roofline_raycast"] = raycast_info
    if rank == 0 and world == 1 and not args.no_warped:
        # SURVEY 8f(1), reported beside the headline (never part of it): the same sequence with the fusion step of every frame done by
        # df_integrate_warped (DF_KINFU_WARPED_INTEGRATE) instead of the reference's rigid fallback.  Short separate pass.
        try:
            p = params(kf.STAGE_TIMING | kf.WARPED_INTEGRATE)
            p.fusion_weight_scale = 100.0
            kw = kf.KinFu(p)
            nw = min(nframes, 3 + 8)
            acc, n_upd_w, n_warped, cnt = {}, 0, 0, 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for t in range(nw):
                if t == 3:
                    torch.cuda.synchronize()
                    e0.record()
                kw.lib.df_kinfu_process_device(kw.h, frames_dev[t].data_ptr(), pitch)
                if t >= 3:
                    for name, v in kw.stage_ms().items():
                        acc[name] = acc.get(name, 0.0) + v
                    i = kw.info()
                    n_upd_w += i["n_updated"]; n_warped += i["n_warped"]; cnt += 1
            e1.record()
            torch.cuda.synchronize()
            kw.close()
            line["warped_fusion"] = {"what": "same workload, fusion step = per-voxel warped integration (df_integrate_warped, weight_scale 100)",
                                     "frames": cnt, "ms_per_frame_incl_stage_readback": e0.elapsed_time(e1) / max(cnt, 1),
                                     "stage_ms": {n: v / max(cnt, 1) for n, v in acc.items()},
                                     "voxels_warped_per_frame": n_warped / max(cnt, 1), "voxels_written_per_frame": n_upd_w / max(cnt, 1),
                                     "volume_voxels": DIM ** 3}
        except Exception as e:                                                  # informational only: never lose the headline line
            line["warped_fusion"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        steps_cpu = 12
        dt, cinfo = run_cpu(frames[: 1 + 1 + steps_cpu], 1, steps_cpu)
        line["cpu_baseline"] = {"value": steps_cpu / dt, "unit": "frames/s", "cores": host_threads(), "kind": "port",
                                "sample": f"{steps_cpu} timed frames (+1 init, +1 warm-up) of the same sequence through the CPU oracle (OpenMP)"}
        # SURVEY 8d also asks for the single-thread figure (the reference's own warp / k-NN loops are serial): 2 frames, 1 OpenMP thread
        try:
            import ctypes
            gomp = ctypes.CDLL("libgomp.so.1")
            gomp.omp_set_num_threads(1)
            dt1, _ = run_cpu(frames[:3], 0, 2)
            gomp.omp_set_num_threads(host_threads())
            line["cpu_baseline"]["single_thread_value"] = 2 / dt1
        except Exception as e:                                                  # informational only
            line["cpu_baseline"]["single_thread_value"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
