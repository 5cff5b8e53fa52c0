"""Digests of the oracle's outputs for the SURVEY 8f additions (oracle/orc_fusion.c) on seeded inputs -> tests/golden/fusion_golden.json.
These steps have no reference output to pin to (the reference never wrote them); the digests pin the RESTATEMENT against drift, so that
the GPU parity tests keep comparing with the same arithmetic that the closed-form CPU tests (tests/test_fusion_oracle.py) vouched for.
   python tests/golden/make_fusion_golden.py"""
import ctypes as C
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def cases():
    from dynamicfusion_b200 import synth
    from oracle import orc
    K = synth.DEFAULT_K
    out = {}
    rng = np.random.default_rng(2024)
    M, dim = 240, 48
    d = rng.normal(size=(M, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v = (np.array([0.0, 0.0, 1.0]) + 0.25 * d).astype(np.float32)
    nodes = orc.make_nodes(v)
    for i in range(M):
        t = (rng.uniform(-1, 1, 3) * 0.003).astype(np.float32)
        orc.load().orc_node_encode_translation(C.c_void_p(nodes[i].ctypes.data), C.c_float(float(t[0])), C.c_float(float(t[1])), C.c_float(float(t[2])))
    depth = synth.sphere_wall_depth(seed=77)
    a, b = np.deg2rad(4.0), np.deg2rad(-2.0)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    w2c = ((Rx @ Ry).astype(np.float32), np.array([0.01, -0.02, 0.015], np.float32))
    for scale in (0.0, 100.0):
        vol = np.zeros(dim ** 3, np.uint32)
        n = 0
        for _ in range(2):
            n += orc.integrate_warped(vol, (dim,) * 3, (1.0 / dim,) * 3, 0.04, 64, depth, synth.volume_pose(1.0), w2c, K, nodes, scale)
        out[f"integrate_warped_{dim}_scale{int(scale)}"] = {"written": int(n), "sha256": hashlib.sha256(vol.tobytes()).hexdigest()}
    cloud = np.zeros((9000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-0.4, 0.4, (9000, 3)) + np.array([0, 0, 1.0], np.float32)
    cloud[::23, 0] = np.nan
    ext = orc.extend_field(nodes, cloud, 0.07, 50, 1024)
    out["extend_field"] = {"nodes": int(len(ext)), "sha256": hashlib.sha256(ext.tobytes()).hexdigest()}
    return out


if __name__ == "__main__":
    res = cases()
    (Path(__file__).parent / "fusion_golden.json").write_text(json.dumps(res, indent=1) + "\n")
    print(json.dumps(res, indent=1))
