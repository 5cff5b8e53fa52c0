// warp.cu -- fused k-NN + node weights + dual-quaternion-blend warp of vertex/normal sets on sm_100a.
// Replaces WarpField::KNN / getWeightsAndUpdateKNN / DQB / warp of the reference (kfusion/src/warp_field.cpp:180-251),
// which run single-threaded on the CPU through a nanoflann kd-tree (~1 M queries per frame, the reference's dominant
// cost).  The node table (M x 48 B) is GPU-resident; a block stages node positions tile by tile
// in shared memory and every thread scans them for its own point.
#include "warp_common.cuh"

using namespace dfb;

namespace {

__global__ void __launch_bounds__(256) knn8_kernel(const float *__restrict__ nodes, int M, const float *__restrict__ queries, int N,
                                                   int qstride, int *__restrict__ idx, float *__restrict__ d2)
{
    __shared__ KnnSmem sm;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    bool valid = false;
    if (q < N) {
        const float *p = queries + (size_t)q * qstride;
        qx = p[0]; qy = p[1]; qz = p[2];
        valid = !(isnan(qx) || isnan(qy) || isnan(qz));
    }
    int bi[8]; float bd[8];
    knn8_scan(nodes, M, valid, qx, qy, qz, sm, bi, bd);
    if (q < N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { idx[(size_t)q * 8 + i] = bi[i]; d2[(size_t)q * 8 + i] = bd[i]; }
    }
}

struct WarpParams {
    const float *nodes; int M;
    float *points; float *normals; int N; int stride;
    Aff w2l;
    int flags;
    int *idx_out; float *w_out;
    const int *rank;          // REF_NORMAL_INDEX mode: rank of each point among coordinate-valid points
    const int *first_nan;     // REF_NORMAL_INDEX mode: first normal index whose x is NaN (or N)
};

// cv::Affine3f * Vec3f (opencv2/core/affine.hpp): m0*x + m1*y + m2*z + m3 evaluated left to right
__device__ __forceinline__ float3 aff_apply_cv(const Aff &a, const float3 v)
{
    return make_float3(a.r0.x * v.x + a.r0.y * v.y + a.r0.z * v.z + a.t.x,
                       a.r1.x * v.x + a.r1.y * v.y + a.r1.z * v.z + a.t.y,
                       a.r2.x * v.x + a.r2.y * v.y + a.r2.z * v.z + a.t.z);
}

// WarpField::warp, warp_field.cpp:180-195
__global__ void __launch_bounds__(256) warp_kernel(const WarpParams p)
{
    __shared__ KnnSmem sm;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    float3 pt = make_float3(0.f, 0.f, 0.f), nr = pt;
    bool valid = false;
    int ni = q;
    if (q < p.N) {
        const float *pp = p.points + (size_t)q * p.stride;
        pt = make_float3(pp[0], pp[1], pp[2]);
        valid = !isnan(pt.x);
        if (p.flags & DF_WARP_REF_NORMAL_INDEX) {
            // the reference's normal cursor advances only on warped points and stalls for ever at the first NaN normal
            ni = p.rank[q];
            valid = valid && ni < *p.first_nan;
        }
        if (valid) {
            const float *np = p.normals + (size_t)ni * p.stride;
            nr = make_float3(np[0], np[1], np[2]);
            valid = !isnan(nr.x);
        }
    }
    int bi[8]; float bd[8];
    knn8_scan(p.nodes, p.M, valid, pt.x, pt.y, pt.z, sm, bi, bd);
    if (q >= p.N) return;
    float w8[8];
    if (valid) {
        const Dqb d = dqb_blend(p.nodes, bi, bd, w8);
        float3 wp = aff_apply_cv(p.w2l, dq_transform(d, pt));
        float3 wn;
        if (p.flags & DF_WARP_NORMAL_ROTATE_ONLY) {
            const float3 r = qrotate(d.rot, nr);
            wn = make_float3(p.w2l.r0.x * r.x + p.w2l.r0.y * r.y + p.w2l.r0.z * r.z,
                             p.w2l.r1.x * r.x + p.w2l.r1.y * r.y + p.w2l.r1.z * r.z,
                             p.w2l.r2.x * r.x + p.w2l.r2.y * r.y + p.w2l.r2.z * r.z);
        } else {
            wn = aff_apply_cv(p.w2l, dq_transform(d, nr));       // the reference also translates normals
        }
        float *pp = p.points + (size_t)q * p.stride;
        pp[0] = wp.x; pp[1] = wp.y; pp[2] = wp.z;
        float *np = p.normals + (size_t)ni * p.stride;
        np[0] = wn.x; np[1] = wn.y; np[2] = wn.z;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = 0.f;
    }
    if (p.idx_out) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { p.idx_out[(size_t)q * 8 + i] = bi[i]; p.w_out[(size_t)q * 8 + i] = w8[i]; }
    }
}

}  // namespace

extern "C" int df_knn8(const float *nodes, int M, const float *queries, int N, int qstride, int32_t *idx, float *d2, void *stream)
{
    if (N <= 0) return 0;
    knn8_kernel<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(nodes, M, queries, N, qstride, idx, d2);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" int df_warp(const float *nodes, int M, float *points, float *normals, int N, int stride, df_aff3f warp_to_live, int flags,
                       int32_t *idx_out, float *w_out, void *stream)
{
    if (N <= 0) return 0;
    if (flags & DF_WARP_REF_NORMAL_INDEX) return (int)cudaErrorNotSupported;   // reference normal-cursor quirk: not built yet
    WarpParams p;
    p.nodes = nodes; p.M = M; p.points = points; p.normals = normals; p.N = N; p.stride = stride;
    p.w2l = make_aff(warp_to_live); p.flags = flags; p.idx_out = idx_out; p.w_out = w_out;
    p.rank = nullptr; p.first_nan = nullptr;
    warp_kernel<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(p);
    DF_LAUNCH_CHECK();
    return 0;
}
