"""CPU check of the RULE behind the visibility culling of csrc/fusion.cu (fusion_run_invisible), restated in numpy float32: whenever the
rule declares a run of a warp's sub-brick invisible, the oracle (oracle/orc_fusion.c) must not write a single voxel of it.  This pins the
geometry of the rule (displacement bound, grown box, half-space tests, three-level depth maxima) without a GPU; the kernel's own use of it
is checked on the GPU by tests/test_fusion_gpu.py::test_culling_is_invisible (culled and unculled runs store identical volumes)."""
import ctypes as C

import numpy as np
import pytest

from dynamicfusion_b200 import synth

K = synth.DEFAULT_K
F32 = np.float32
SUB, TILE, COARSE = 8, 16, 4


def _node_translations(orc, nodes):
    t = orc.node_translations(nodes)            # (w, x, y, z) per node, DualQuaternion::getTranslation
    return t[:, 1:4]


def _tile_maxima(depth):
    rows, cols = depth.shape
    ty, tx = -(-rows // TILE), -(-cols // TILE)
    m = (depth.astype(F32) * F32(0.001))
    pad = np.zeros((ty * TILE, tx * TILE), F32)
    pad[:rows, :cols] = m
    fine = pad.reshape(ty, TILE, tx, TILE).max(axis=(1, 3))
    cy, cx = -(-ty // COARSE), -(-tx // COARSE)
    padc = np.zeros((cy * COARSE, cx * COARSE), F32)
    padc[:ty, :tx] = fine
    coarse = padc.reshape(cy, COARSE, cx, COARSE).max(axis=(1, 3))
    return fine, coarse, F32(fine.max())


def _run_invisible(xa, xb, ya, yb, za, zb, vs, delta, v2c_R, v2c_t, cols, rows, trunc, fine, coarse, gmax):
    """numpy float32 restatement of fusion_run_invisible for one sub-brick"""
    fx, fy, cx, cy = (F32(v) for v in K)
    corners = []
    for k in range(8):
        c = np.array([(F32(xb + 1) * vs[0] + delta) if k & 1 else (F32(xa - 1) * vs[0] - delta),
                      (F32(yb + 1) * vs[1] + delta) if k & 2 else (F32(ya - 1) * vs[1] - delta),
                      (F32(zb + 1) * vs[2] + delta) if k & 4 else (F32(za - 1) * vs[2] - delta)], F32)
        corners.append((v2c_R @ c + v2c_t).astype(F32))
    pc = np.array(corners, F32)
    if np.all(pc[:, 2] < F32(-1e-3)):
        return True
    if np.any(~(pc[:, 2] > F32(1e-2))):
        return False
    u = fx * (pc[:, 0] / pc[:, 2]) + cx
    v = fy * (pc[:, 1] / pc[:, 2]) + cy
    if np.all(u < -1) or np.all(v < -1) or np.all(u > cols + 1) or np.all(v > rows + 1):
        return True
    px0, px1 = max(0, int(np.floor(u.min() - 1))), min(cols - 1, int(np.floor(u.max() + 1)))
    py0, py1 = max(0, int(np.floor(v.min() - 1))), min(rows - 1, int(np.floor(v.max() + 1)))
    if px1 < px0 or py1 < py0:
        return True
    tx0, tx1, ty0, ty1 = px0 // TILE, px1 // TILE, py0 // TILE, py1 // TILE
    if (tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= 32:
        m = fine[ty0:ty1 + 1, tx0:tx1 + 1].max()
    else:
        tx0, tx1, ty0, ty1 = tx0 // COARSE, tx1 // COARSE, ty0 // COARSE, ty1 // COARSE
        m = coarse[ty0:ty1 + 1, tx0:tx1 + 1].max() if (tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= 32 else gmax
    return bool(pc[:, 2].min() - F32(1e-3) > m + trunc)


def _tilted_pose(scale=1.0):
    a, b = np.deg2rad(5.0 * scale), np.deg2rad(-3.0 * scale)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    return (Rx @ Ry).astype(F32), np.array([0.02, -0.015, 0.03], F32)


@pytest.mark.parametrize("dim,M,t_scale,seed", [(48, 200, 0.002, 1), (48, 60, 0.012, 2), (64, 400, 0.0, 3), (40, 120, 0.03, 4)])
def test_runs_declared_invisible_are_never_written(orc, dim, M, t_scale, seed):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(M, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    verts = (np.array([0.0, 0.0, 1.0]) + 0.25 * d).astype(F32)
    verts[::3, 2] = 1.38
    verts[::3, :2] = rng.uniform(-0.45, 0.45, size=(len(verts[::3]), 2))
    nodes = orc.make_nodes(verts)
    for i in range(M):
        t = (rng.uniform(-1, 1, 3) * t_scale).astype(F32)
        orc.load().orc_node_encode_translation(C.c_void_p(nodes[i].ctypes.data), C.c_float(float(t[0])), C.c_float(float(t[1])), C.c_float(float(t[2])))
    depth = synth.sphere_wall_depth(seed=seed)
    depth[100:200, 300:420] = 0                               # a hole: tiles without any depth
    size, trunc = 1.0, F32(0.04)
    vs = np.array([size / dim] * 3, F32)
    pose_vol = synth.volume_pose(size)
    Rc, tc = _tilted_pose()
    Ri = np.linalg.inv(Rc.astype(np.float64)).astype(F32)
    world2cam = (Ri, (-(Ri @ tc)).astype(F32))
    vol = np.zeros(dim ** 3, np.uint32)
    n = orc.integrate_warped(vol, (dim,) * 3, vs, float(trunc), 64, depth, pose_vol, world2cam, K, nodes, 100.0)
    written = (vol >> 16).reshape(dim, dim, dim) != 0          # [z][y][x]
    assert n == int(written.sum()) and n > 1000

    tr = _node_translations(orc, nodes)
    delta = F32(8.0) * F32(np.sqrt((tr.astype(F32) ** 2).sum(1)).max()) * F32(1.0001) + F32(1e-6)
    v2c_R = (world2cam[0].astype(np.float64) @ pose_vol[0].astype(np.float64)).astype(F32)
    v2c_t = (world2cam[0].astype(np.float64) @ pose_vol[1].astype(np.float64) + world2cam[1].astype(np.float64)).astype(F32)
    fine, coarse, gmax = _tile_maxima(depth)
    culled_voxels = bad = 0
    for z0 in range(0, dim, SUB):
        for y0 in range(0, dim, 4):
            for x0 in range(0, dim, 8):
                xb, yb, zb = min(x0 + 7, dim - 1), min(y0 + 3, dim - 1), min(z0 + SUB, dim) - 1
                if _run_invisible(x0, xb, y0, yb, z0, zb, vs, delta, v2c_R, v2c_t, 640, 480, trunc, fine, coarse, gmax):
                    blk = written[z0:zb + 1, y0:yb + 1, x0:xb + 1]
                    culled_voxels += blk.size
                    bad += int(blk.sum())
    print(f"{dim}^3, delta {float(delta) * 1000:.1f} mm: rule culls {culled_voxels / dim ** 3:.1%} of the volume, oracle writes {n / dim ** 3:.1%}")
    assert bad == 0, f"{bad} written voxels lie in runs the rule would skip"
    if t_scale <= 0.002:
        assert culled_voxels > 0.08 * dim ** 3                 # the rule must actually cull when the displacement bound is small (half of this 1 m cube is free space in front of the wall)
