import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from dynamicfusion_b200 import kinfu as kf, synth
from oracle import orc, orc_pipe
import test_pipeline_gpu as tp
p = tp._params(64, kf.RIGID_ONLY)
cpu = orc_pipe.KinFu(orc_pipe.params_from(p))
frames = [synth.umbrella_depth(t) for t in range(4)]
for d in frames: cpu(d)
vc = cpu.buffer("volume"); fc, wc = tp._tsdf(vc)
for rep in range(3):
    gpu = kf.KinFu(p)
    for d in frames: gpu(d)
    vg = gpu.buffer("volume"); fg, wg = tp._tsdf(vg)
    same = wg == wc
    dp = max(np.abs(gpu.getCameraPose(t)[1] - cpu.getCameraPose(t)[1]).max() for t in range(4))
    dr = max(np.abs(gpu.getCameraPose(t)[0] - cpu.getCameraPose(t)[0]).max() for t in range(4))
    print("PDL", os.environ.get("DF_PDL"), "rep", rep, "w_mismatch", np.mean(wg != wc), "f_mismatch", np.mean(np.abs(fg[same] - fc[same]) > 2e-3), "v_mismatch", np.mean(vg != vc), "pose dt", dp, "dR", dr)
    gpu.close()
