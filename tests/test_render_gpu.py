"""Display kernels (df_render_image, df_render_tangent_colors -- what KinFu::renderImage and the reference's demo show) against the
REFERENCE's own render kernels, compiled for the host into oracle/_ref/libkfref.so (kfusion/src/cuda/imgproc.cu:474-583).  There is
no oracle restatement for these: the comparison is directly with the reference's code.  Bar: tangent colours identical wherever the
normal is defined; the Phong image within 1 grey level (the reference's __powf / rsqrt are approximate GPU intrinsics)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dynamicfusion_b200 import capi, host, synth  # noqa: E402

K = synth.DEFAULT_K


def test_render_kernels_match_the_reference_code(orc):
    if not orc.reference_available():
        pytest.skip("oracle/_ref/libkfref.so not built")
    ref = orc.load_ref()
    dim = 128
    vol = host.TsdfVolume((dim, dim, dim))
    vol.setTruncDist(0.04); vol.setMaxWeight(64); vol.setSize((1.0, 1.0, 1.0)); vol.setPose(synth.volume_pose(1.0))
    vol.setRaycastStepFactor(0.75); vol.setGradientDeltaFactor(0.5); vol.clear()
    dists = host.computeDists(host.u16_to_device(synth.umbrella_depth(0)), K)
    pose = host.identity_pose()
    vol.integrate(dists, pose, K)
    pts, nrm, _ = vol.raycast(pose, K, 640, 480)
    assert int((~torch.isnan(pts[..., 0])).sum().item()) > 100_000
    light = (C.c_float * 3)(0.3, -0.2, -0.5)
    lib = capi.load()
    img = torch.zeros((480, 640, 4), dtype=torch.uint8, device="cuda")
    tan = torch.zeros_like(img)
    capi.check(lib.df_render_image(pts.data_ptr(), 640 * 16, nrm.data_ptr(), 640 * 16, 640, 480, light, img.data_ptr(), 640 * 4, None))
    capi.check(lib.df_render_tangent_colors(nrm.data_ptr(), 640 * 16, 640, 480, tan.data_ptr(), 640 * 4, None))
    torch.cuda.synchronize()
    hp, hn = np.ascontiguousarray(pts.cpu().numpy()), np.ascontiguousarray(nrm.cpu().numpy())
    rimg, rtan = np.zeros((480, 640, 4), np.uint8), np.zeros((480, 640, 4), np.uint8)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    ref.kfref_render_image(vp(hp), C.c_size_t(640 * 16), vp(hn), C.c_size_t(640 * 16), 640, 480, orc.intr(*K), light, vp(rimg), C.c_size_t(640 * 4))
    ref.kfref_render_tangent_colors(vp(hn), C.c_size_t(640 * 16), 640, 480, vp(rtan), C.c_size_t(640 * 4))
    gi, gt = img.cpu().numpy(), tan.cpu().numpy()
    assert np.abs(gi.astype(np.int16) - rimg.astype(np.int16)).max() <= 1
    assert np.mean(gi != rimg) < 0.02
    valid = ~np.isnan(hn[..., 0])
    assert np.array_equal(gt[valid], rtan[valid])
    assert len(np.unique(gi[valid][:, 0])) > 50 and len(np.unique(gi[~valid][:, 0])) > 50      # a shaded surface over the gradient background
