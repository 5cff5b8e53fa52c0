"""ctypes binding of the oracle's per-frame loop (oracle/orc_pipeline.c) and of the full-size helpers.
TEST INFRASTRUCTURE ONLY (see oracle/orc.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import orc


class KinfuParams(C.Structure):
    """orc_kinfu_params: identical layout to df_kinfu_params (include/dfusion.h)"""
    _fields_ = [("cols", C.c_int), ("rows", C.c_int), ("intr", orc.Intr),
                ("volume_dims", C.c_int * 3), ("volume_size", C.c_float * 3), ("volume_pose", orc.Aff3f),
                ("bilateral_sigma_depth", C.c_float), ("bilateral_sigma_spatial", C.c_float), ("bilateral_kernel_size", C.c_int),
                ("icp_truncate_depth_dist", C.c_float), ("icp_dist_thres", C.c_float), ("icp_angle_thres", C.c_float),
                ("icp_iter_num", C.c_int * 4),
                ("tsdf_min_camera_movement", C.c_float), ("tsdf_trunc_dist", C.c_float), ("tsdf_max_weight", C.c_int),
                ("raycast_step_factor", C.c_float), ("gradient_delta_factor", C.c_float),
                ("light_pose", C.c_float * 3),
                ("solver_nonlinear_iters", C.c_int), ("solver_linear_iters", C.c_int),
                ("max_nodes", C.c_int), ("node_step", C.c_int), ("cloud_capacity", C.c_int), ("flags", C.c_int),
                ("fusion_weight_scale", C.c_float), ("extend_radius", C.c_float)]


def params_from(product_params) -> KinfuParams:
    """byte-copy a dynamicfusion_b200.capi.KinfuParams (same layout)"""
    p = KinfuParams()
    assert C.sizeof(p) == C.sizeof(product_params)
    C.memmove(C.byref(p), C.byref(product_params), C.sizeof(p))
    return p


def default_params(which=0, dim=None, size=None) -> KinfuParams:
    """KinFuParams::default_params_dynamicfusion (kinfu.cpp:14-49) without needing libdfusion.so"""
    p = KinfuParams()
    p.cols, p.rows = 640, 480
    p.intr = orc.Intr(570.342, 570.342, 320.0, 240.0) if which == 0 else orc.Intr(525.0, 525.0, 319.5, 239.5)
    d, s = (256, 1.0) if which == 0 else (512, 3.0)
    d, s = dim or d, size or s
    for i in range(3):
        p.volume_dims[i] = d
        p.volume_size[i] = s
    for i in range(9):
        p.volume_pose.R[i] = 1.0 if i % 4 == 0 else 0.0
    p.volume_pose.t[0], p.volume_pose.t[1], p.volume_pose.t[2] = -s / 2, -s / 2, 0.5
    p.bilateral_sigma_depth, p.bilateral_sigma_spatial, p.bilateral_kernel_size = 0.04, 4.5, 7
    p.icp_truncate_depth_dist, p.icp_dist_thres, p.icp_angle_thres = 0.0, 0.1, np.float32(30.0) * np.float32(0.017453293)
    for i, v in enumerate((10, 5, 4, 0)):
        p.icp_iter_num[i] = v
    p.tsdf_min_camera_movement, p.tsdf_trunc_dist, p.tsdf_max_weight = 0.0, 0.04, 64
    p.raycast_step_factor, p.gradient_delta_factor = 0.75, 0.5
    p.solver_nonlinear_iters, p.solver_linear_iters = 5, 100
    p.max_nodes, p.node_step, p.cloud_capacity, p.flags = 4096, 50, 256 * 256 * 256 // 4, 0
    return p


class KinFu:
    def __init__(self, params: KinfuParams):
        self.lib = orc.load()
        self.lib.orc_kinfu_create.restype = C.c_void_p
        self.lib.orc_kinfu_buffer.restype = C.c_void_p
        self.params = params
        self.h = C.c_void_p(self.lib.orc_kinfu_create(C.byref(params)))

    def close(self):
        if self.h:
            self.lib.orc_kinfu_destroy(self.h)
            self.h = None

    def __call__(self, depth: np.ndarray) -> bool:
        assert depth.dtype == np.uint16 and depth.flags.c_contiguous
        return bool(self.lib.orc_kinfu_process(self.h, C.c_void_p(depth.ctypes.data), C.c_size_t(depth.strides[0])))

    def info(self) -> dict:
        v = (C.c_longlong * 8)()
        self.lib.orc_kinfu_info(self.h, v)
        keys = ["frame_counter", "nodes", "cloud_points", "poses", "icp_ok", "launches", "resets", "lm_iters"]
        return dict(zip(keys, [int(x) for x in v]))

    def _buf(self, which, dtype, shape):
        ptr = self.lib.orc_kinfu_buffer(self.h, which)
        n = int(np.prod(shape))
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))
        return arr.reshape(shape).copy()

    def buffer(self, name: str) -> np.ndarray:
        p = self.params
        img = (p.rows, p.cols, 4)
        nvox = p.volume_dims[0] * p.volume_dims[1] * p.volume_dims[2]
        i = self.info()
        table = {"volume": (0, np.uint32, (nvox,)), "dists": (1, np.uint16, (p.rows, p.cols)), "curr_depth": (2, np.uint16, (p.rows, p.cols)),
                 "curr_points": (3, np.float32, img), "curr_normals": (4, np.float32, img), "prev_points": (5, np.float32, img),
                 "prev_normals": (6, np.float32, img), "canonical": (7, np.float32, img), "canonical_normals": (8, np.float32, img),
                 "cloud": (9, np.float32, (max(i["cloud_points"], 0), 4)), "cloud_normals": (10, np.float32, (max(i["cloud_points"], 0), 4)),
                 "nodes": (11, np.float32, (max(i["nodes"], 0), 12)), "canonical_visible": (12, np.float32, img),
                 "solve_stats": (13, np.float64, (8,)), "poses": (14, np.float32, (i["poses"], 12)), "stage_s": (15, np.float64, (10,))}
        which, dt, shape = table[name]
        if int(np.prod(shape)) == 0:
            return np.zeros(shape, dt)
        return self._buf(which, dt, shape)

    def getCameraPose(self, time=-1):
        poses = self.buffer("poses")
        a = poses[time]
        return a[:9].reshape(3, 3), a[9:]


def knn8_fast(nodes, queries):
    q = np.ascontiguousarray(queries, np.float32)
    N, stride = q.shape
    idx = np.empty((N, 8), np.int32)
    d2 = np.empty((N, 8), np.float32)
    orc.load().orc_knn8_fast(orc._p(nodes), len(nodes), orc._p(q), C.c_longlong(N), stride, orc._p(idx), orc._p(d2))
    return idx, d2


def solve_data_term_big(nodes, canon, live, flags=0, lm_iters=5, lin_iters=100):
    c = np.ascontiguousarray(canon, np.float32)
    l = np.ascontiguousarray(live, np.float32)
    N, stride = c.shape
    stats = np.zeros(8, np.float64)
    orc.load().orc_solve_data_term_big(orc._p(nodes), len(nodes), orc._p(c), orc._p(l), C.c_longlong(N), stride, flags, lm_iters, lin_iters, orc._p(stats))
    return stats
