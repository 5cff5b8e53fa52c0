import os, sys
sys.path.insert(0, '/root/repo')
os.environ["DF_RAYCAST_TMA"] = "1"
import numpy as np, torch
from dynamicfusion_b200 import host, synth
K = synth.DEFAULT_K
dim = 128
v = host.TsdfVolume((dim, dim, dim), track_activity=True)
v.setTruncDist(0.04); v.setMaxWeight(64); v.setSize((1.0, 1.0, 1.0)); v.setPose(synth.volume_pose(1.0)); v.setRaycastStepFactor(0.75); v.setGradientDeltaFactor(0.5); v.clear()
dists = host.computeDists(host.u16_to_device(synth.umbrella_depth(0)), K)
v.integrate(dists, host.identity_pose(), K)
torch.cuda.synchronize()
print("integrated")
p, n, _ = v.raycast(host.identity_pose(), K, 640, 480)
torch.cuda.synchronize()
print("tma raycast ok", int((~torch.isnan(p[..., 0])).sum()))
