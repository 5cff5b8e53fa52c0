"""CPU checks of the oracle's per-voxel warped integration (oracle/orc_fusion.c).  The reference never finished this step
(tsdf_volume.cpp:240-252 is commented out), so there is no reference output to pin it to: "parity unpinned".  What CAN be checked
without a reference is checked here -- closed-form scenes and the properties the construction must have."""
import numpy as np

from dynamicfusion_b200 import synth

K = synth.DEFAULT_K
DIM, SIZE, TRUNC, MAXW = 48, 0.6, 0.04, 64
VS = (SIZE / DIM,) * 3
POSE_VOL = (np.eye(3, dtype=np.float32), np.array([-SIZE / 2, -SIZE / 2, 0.5], np.float32))
IDENT = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))


def _half(bits):
    return bits.astype(np.uint16).view(np.float16).astype(np.float32)


def _plane_depth(z0_mm):
    return np.full((480, 640), z0_mm, np.uint16)


def _node_sheet(orc, z=0.8, n=12, weight=3.0):
    g = np.linspace(-0.25, 0.25, n, dtype=np.float32)
    xs, ys = np.meshgrid(g, g)
    v = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, z, np.float32)], -1)
    return orc.make_nodes(v, weight)


def _voxel_cam_z():
    z = np.arange(DIM, dtype=np.float32) * np.float32(VS[2]) + np.float32(0.5)
    return z


def test_identity_field_plane_gives_the_closed_form_profile(orc):
    """identity field, camera at the origin, fronto-parallel plane at z0: rho = z0 - z_voxel along every pixel's ray"""
    z0 = 0.8
    vol = np.zeros(DIM ** 3, np.uint32)
    nodes = _node_sheet(orc)
    n = orc.integrate_warped(vol, (DIM,) * 3, VS, TRUNC, MAXW, _plane_depth(800), POSE_VOL, IDENT, K, nodes, 0.0)
    v = vol.reshape(DIM, DIM, DIM)                       # [z][y][x]
    w = v >> 16
    f = _half(v & 0xffff)
    zc = _voxel_cam_z()
    rho = np.float32(z0) - zc
    upd = rho > -TRUNC
    # which voxels project into the image (float64 here; a one-pixel rim is left out of the comparison)
    xs = np.arange(DIM) * VS[0] - SIZE / 2
    zz, yy, xx = np.meshgrid(zc.astype(np.float64), xs, xs, indexing="ij")
    u, vv = K[0] * xx / zz + K[2], K[1] * yy / zz + K[3]
    inside = (u >= 1) & (vv >= 1) & (u < 639) & (vv < 479)
    outside = (u < -1) | (vv < -1) | (u > 641) | (vv > 481)
    upd3 = np.broadcast_to(upd[:, None, None], inside.shape)
    assert np.all(w[inside & upd3] == 1) and np.all(w[~upd3] == 0) and np.all(w[outside] == 0)
    assert int((inside & upd3).sum()) <= n <= int((~outside & upd3).sum())
    want = np.minimum(1.0, rho / TRUNC).astype(np.float32)
    got = f[:, DIM // 2, DIM // 2]
    np.testing.assert_allclose(got[upd], want[upd], atol=2e-3)       # fp16 storage + the 0.001f millimetre product
    assert np.all(f[~upd] == 0)


def test_unit_weight_matches_the_rigid_rule_where_ray_length_equals_depth(orc):
    """on the optical axis |vc| == vc.z, so the rigid integrate (ray-length sdf) and the warped one (projective z sdf) with an
    identity field and weight 1 must store the same voxel, twice in a row (running average included)"""
    depth = _plane_depth(800)
    nodes = _node_sheet(orc)
    a = np.zeros(DIM ** 3, np.uint32)
    b = np.zeros(DIM ** 3, np.uint32)
    dists = orc.compute_dists(depth, K)
    for _ in range(2):
        orc.integrate_warped(a, (DIM,) * 3, VS, TRUNC, MAXW, depth, POSE_VOL, IDENT, K, nodes, 0.0)
        orc.integrate(b, (DIM,) * 3, VS, TRUNC, MAXW, dists, POSE_VOL, K)
    va, vb = a.reshape(DIM, DIM, DIM), b.reshape(DIM, DIM, DIM)
    # the axis passes through voxel (x, y) = (24, 24): x*vs - 0.3 == 0
    col_a, col_b = va[:, DIM // 2, DIM // 2], vb[:, DIM // 2, DIM // 2]
    assert np.array_equal(col_a >> 16, col_b >> 16)
    fa, fb = _half(col_a & 0xffff), _half(col_b & 0xffff)
    np.testing.assert_allclose(fa, fb, atol=1.3e-2)      # dists are fp16 ray lengths (0.5 mm steps at 0.8 m = 0.012 trunc units), depth is exact mm


def test_uniform_translation_field_equals_a_shifted_camera(orc):
    """node weight 1e4 makes every blend weight exactly 1.0f, so the eight un-normalised translations add up to 8*t exactly
    (warp_field.cpp:203-217 does not normalise): warping by the field == moving the camera by -8t"""
    t = np.float32(1.0 / 256.0)
    nodes = _node_sheet(orc, weight=1.0e4)
    for i in range(len(nodes)):
        orc.load().orc_node_encode_translation(nodes[i].ctypes.data_as(__import__("ctypes").c_void_p), __import__("ctypes").c_float(0.0),
                                               __import__("ctypes").c_float(0.0), __import__("ctypes").c_float(float(t)))
    depth = synth.sphere_wall_depth(seed=3, centre=(0.0, 0.0, 0.8), radius=0.15, wall_z=1.0)
    a = np.zeros(DIM ** 3, np.uint32)
    b = np.zeros(DIM ** 3, np.uint32)
    ident_nodes = _node_sheet(orc, weight=1.0e4)
    na = orc.integrate_warped(a, (DIM,) * 3, VS, TRUNC, MAXW, depth, POSE_VOL, IDENT, K, nodes, 0.0)
    shifted = (np.eye(3, dtype=np.float32), np.array([0, 0, 8 * t], np.float32))
    nb = orc.integrate_warped(b, (DIM,) * 3, VS, TRUNC, MAXW, depth, POSE_VOL, shifted, K, ident_nodes, 0.0)
    differ = np.count_nonzero(a != b)
    assert abs(na - nb) <= 1e-4 * nb and differ <= 1e-4 * DIM ** 3, (na, nb, differ)
    assert na > 1000


def test_sample_weight_is_the_quantised_mean_node_distance(orc):
    """TsdfVolume::weighting (tsdf_volume.cpp:300-306) on the voxel's own 8 neighbours, recomputed here by brute force"""
    nodes = _node_sheet(orc)
    scale = 100.0
    vol = np.zeros(DIM ** 3, np.uint32)
    orc.integrate_warped(vol, (DIM,) * 3, VS, TRUNC, MAXW, _plane_depth(800), POSE_VOL, IDENT, K, nodes, scale)
    v = vol.reshape(DIM, DIM, DIM)
    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(400):
        z, y, x = (int(c) for c in rng.integers(0, DIM, 3))
        w = int(v[z, y, x] >> 16)
        if w == 0:
            continue
        xc = np.array([x, y, z], np.float32) * np.float32(VS[0]) + POSE_VOL[1]
        d = np.sqrt(np.sort(((nodes[:, :3] - xc) ** 2).sum(1))[:8].astype(np.float32)).astype(np.float32)
        want = int(np.clip(np.rint(d.sum(dtype=np.float32) / 8 * np.float32(scale)), 1, MAXW))
        assert abs(w - want) <= 1, (w, want)              # numpy sums in a different order: allow the rounding boundary
        checked += 1
    assert checked > 100


def test_weighted_average_saturates_at_max_weight(orc):
    nodes = _node_sheet(orc)
    vol = np.zeros(DIM ** 3, np.uint32)
    for _ in range(3):
        orc.integrate_warped(vol, (DIM,) * 3, VS, TRUNC, 40, _plane_depth(800), POSE_VOL, IDENT, K, nodes, 100.0)
    w = vol >> 16
    assert w.max() == 40 and np.all(w <= 40)


# ------------------------------------------------------------------ extending the warp field (orc_extend_field) ----------------
def _brute_unsupported(nodes, cloud, radius):
    d2 = ((cloud[:, None, :3].astype(np.float32) - nodes[None, :, :3]) ** 2)
    d2 = (d2[..., 0] + d2[..., 1]) + d2[..., 2]                      # the reference's float evaluation order
    valid = ~np.isnan(cloud[:, :3]).any(1)
    return valid & (d2.min(1) > np.float32(radius) * np.float32(radius))


def test_extend_field_appends_every_step_th_unsupported_point(orc):
    rng = np.random.default_rng(5)
    nodes = orc.make_nodes(rng.uniform(-0.1, 0.1, (120, 3)))
    cloud = np.zeros((6000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-0.3, 0.3, (6000, 3))
    cloud[::13, 2] = np.nan
    out = orc.extend_field(nodes, cloud, 0.05, 50, 4096)
    uns = np.flatnonzero(_brute_unsupported(nodes, cloud, 0.05))
    picked = uns[::50]
    assert len(out) == len(nodes) + len(picked)
    assert np.array_equal(out[: len(nodes)], nodes)
    assert np.array_equal(out[len(nodes):, :3], cloud[picked, :3])
    new = out[len(nodes):]
    assert np.all(new[:, 3] == 1) and np.all(new[:, 4:7] == 0) and np.all(new[:, 7] == 1) and np.all(new[:, 8:11] == 0) and np.all(new[:, 11] == 3)


def test_extend_field_respects_capacity_and_support(orc):
    rng = np.random.default_rng(6)
    nodes = orc.make_nodes(rng.uniform(-0.1, 0.1, (100, 3)))
    cloud = np.zeros((5000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-0.3, 0.3, (5000, 3))
    assert len(orc.extend_field(nodes, cloud, 0.05, 50, 110)) == 110            # full table
    assert len(orc.extend_field(nodes, cloud, 10.0, 50, 4096)) == 100           # everything supported
    grown = orc.extend_field(nodes, cloud, 0.05, 1, 100000)                     # step 1: every unsupported point becomes a node ...
    assert len(orc.extend_field(grown, cloud, 0.05, 1, 100000)) == len(grown)   # ... after which the cloud is fully supported (idempotent)


def test_oracle_outputs_match_committed_digests(orc):
    """tests/golden/fusion_golden.json (made by tests/golden/make_fusion_golden.py): the restatement the GPU kernels are compared with
    must not drift"""
    import importlib.util
    import json
    from pathlib import Path
    here = Path(__file__).resolve().parent / "golden"
    spec = importlib.util.spec_from_file_location("make_fusion_golden", here / "make_fusion_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.cases() == json.loads((here / "fusion_golden.json").read_text())
