"""DF_WARP_REUSE_KNN: a warp that re-uses the neighbours + weights of an earlier pass over the same points (what the
pipeline does after the data-term solve) must give bit-identical results to a warp that searches again."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dynamicfusion_b200 import capi, host  # noqa: E402


def test_warp_reuse_equals_fresh_search(orc):
    rng = np.random.default_rng(3)
    M, N = 700, 30000
    node_pts = rng.uniform(-0.3, 0.3, (M, 3)).astype(np.float32)
    wf = host.WarpField()
    wf.init(node_pts)
    nodes = wf.nodes_.cpu().numpy()
    for m in range(M):
        t = rng.normal(scale=0.01, size=3).astype(np.float32)
        orc.load().orc_node_encode_translation(C.c_void_p(nodes[m].ctypes.data), C.c_float(t[0]), C.c_float(t[1]), C.c_float(t[2]))
    wf.setNodes(torch.from_numpy(nodes).cuda())
    pts = np.zeros((N, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.3, 0.3, (N, 3))
    nrm = rng.normal(size=(N, 4)).astype(np.float32)
    pts[::11, 0] = np.nan
    live = pts.copy()
    live[:, :3] += 0.002
    live[::7, 1] = np.nan                                   # rows skipped by the solve must still be warped afterwards

    # fresh search
    p1, n1 = torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda()
    wf.warp(p1, n1)
    # solve (builds idx / w for `pts`), then warp with re-use.  nonlinear_iters = 0 keeps the nodes unchanged.
    lib = capi.load()
    need = lib.df_solve_workspace_bytes(M, N)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    p2, n2, l2 = torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda(), torch.from_numpy(live).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    capi.check(lib.df_solve_data_term(wf.nodes_.data_ptr(), M, wf._grid(), p2.data_ptr(), l2.data_ptr(), N, 4, 0, 0, 0, stats.data_ptr(), ws.data_ptr(), stream))
    assert np.array_equal(wf.nodes_.cpu().numpy(), nodes), "zero LM iterations must leave the nodes untouched"
    idx, w = C.c_void_p(), C.c_void_p()
    capi.check(lib.df_solve_knn_buffers(ws.data_ptr(), M, N, C.byref(idx), C.byref(w)))
    ident = capi.make_aff(np.eye(3), np.zeros(3))
    capi.check(lib.df_warp(wf.nodes_.data_ptr(), M, wf._grid(), p2.data_ptr(), n2.data_ptr(), N, 4, ident, 4, idx, w, stream))
    assert np.array_equal(p1.cpu().numpy().view(np.uint32), p2.cpu().numpy().view(np.uint32))
    assert np.array_equal(n1.cpu().numpy().view(np.uint32), n2.cpu().numpy().view(np.uint32))
    assert stats.cpu().numpy()[3] == np.sum(~np.isnan(pts[:, 0]) & ~np.isnan(live[:, 1]))
