// extract.cu -- zero-crossing cloud + normal extraction from the TSDF volume, sm_100a.
// Replaces FullScan6 / ExtractNormals of the reference (kfusion/src/cuda/tsdf_volume.cu:486-831).
//
// The reference appends points through one global atomicAdd cursor, so its output ORDER is nondeterministic (and the
// warp nodes are picked from that order, warp_field.cpp:49-51).  Here extraction is deterministic: a counting pass,
// an exclusive scan over block counts and an emit pass produce the points in ascending (z, y, x) voxel order, +x,+y,+z
// edge within a voxel -- the same order the CPU oracle uses, so the result is comparable element by element.
//
// Work decomposition: the volume is one contiguous array cut into blocks of EX_THREADS quads (1024 voxels when VX = 4), a
// quad = VX consecutive voxels = one 16-byte load.  A quad whose voxels are all unobserved (W == 0) or free space (F == 1)
// needs nothing else -- on real scenes that is >90 % of them; only surface quads fetch the +x / +y / +z neighbours (L1/L2
// hits).  Kernels: see the comment above extract_count_kernel.
#include "df_common.cuh"

using namespace dfb;

namespace {

constexpr int EX_THREADS = 256;

struct ExtractParams {
    const uint32_t *data;
    int Dx, Dy, Dz;
    float3 vs;
    Aff pose;
    size_t nvox;
    int nblocks;
    const unsigned char *activity;   // optional (dfusion.h, DF_ACTIVITY_VOXELS): blocks whose stretch of the volume is inactive are skipped
};

__device__ __forceinline__ float vox_f(uint32_t v) { return half_bits_to_float((unsigned short)(v & 0xffffu)); }
__device__ __forceinline__ bool sign_change(float F, float Fn) { return (F > 0 && Fn < 0) || (F < 0 && Fn > 0); }

// Calls emit(point) for every zero crossing owned by this thread, in (voxel, axis) order.  tsdf_volume.cu:548-633.
// the thread's own VX voxels (one 16-byte load when VX = 4); zeros past the end of the volume
template <int VX>
__device__ __forceinline__ void load_quad(const ExtractParams &p, size_t v0, uint32_t (&own)[VX])
{
#pragma unroll
    for (int j = 0; j < VX; ++j) own[j] = 0u;
    if (v0 >= p.nvox) return;
    if (VX == 4) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p.data + v0));
        own[0] = q.x; own[1 % VX] = q.y; own[2 % VX] = q.z; own[3 % VX] = q.w;
    } else {
        own[0] = __ldg(p.data + v0);
    }
}

template <int VX, typename Emit>
__device__ __forceinline__ void thread_crossings(const ExtractParams &p, size_t v0, const uint32_t (&own)[VX], Emit emit)
{
    bool any = false;
#pragma unroll
    for (int j = 0; j < VX; ++j) any |= vox_active(own[j]);
    if (!any) return;

    const size_t slice = (size_t)p.Dx * p.Dy;
    const int z = (int)(v0 / slice);
    if (z >= p.Dz - 1) return;                               // loop bound z < dims.z - 1, tsdf_volume.cu:553
    const int rem = (int)(v0 - (size_t)z * slice);
    const int y = rem / p.Dx;
    const int x0 = rem - y * p.Dx;

#pragma unroll
    for (int j = 0; j < VX; ++j) {
        if (!vox_active(own[j])) continue;
        const int x = x0 + j;
        const float F = vox_f(own[j]);
        const float3 V = make_float3(((float)x + 0.5f) * p.vs.x, ((float)y + 0.5f) * p.vs.y, ((float)z + 0.5f) * p.vs.z);
        if (x + 1 < p.Dx) {
            const uint32_t nv = (j + 1 < VX) ? own[(j + 1) % VX] : __ldg(p.data + v0 + VX);
            if (vox_active(nv)) {
                const float Fn = vox_f(nv);
                if (sign_change(F, Fn)) {
                    const float Vnx = V.x + p.vs.x;
                    const float d_inv = 1.f / (fabsf(F) + fabsf(Fn));
                    emit(aff_mul(p.pose, make_float3((V.x * fabsf(Fn) + Vnx * fabsf(F)) * d_inv, V.y, V.z)));
                }
            }
        }
        if (y + 1 < p.Dy) {
            const uint32_t nv = __ldg(p.data + v0 + j + p.Dx);
            if (vox_active(nv)) {
                const float Fn = vox_f(nv);
                if (sign_change(F, Fn)) {
                    const float Vny = V.y + p.vs.y;
                    const float d_inv = 1.f / (fabsf(F) + fabsf(Fn));
                    emit(aff_mul(p.pose, make_float3(V.x, (V.y * fabsf(Fn) + Vny * fabsf(F)) * d_inv, V.z)));
                }
            }
        }
        {
            const uint32_t nv = __ldg(p.data + v0 + j + slice);
            if (vox_active(nv)) {
                const float Fn = vox_f(nv);
                if (sign_change(F, Fn)) {
                    const float Vnz = V.z + p.vs.z;
                    const float d_inv = 1.f / (fabsf(F) + fabsf(Fn));
                    emit(aff_mul(p.pose, make_float3(V.x, V.y, (V.z * fabsf(Fn) + Vnz * fabsf(F)) * d_inv)));
                }
            }
        }
    }
}

// Count / emit.  A block = EX_THREADS * EX_QPT quads (4096 voxels when VX = 4), a thread owns quads tid, tid + 256, ...: all of
// its 16-byte loads are issued before any of them is used, and each load instruction of a warp is one contiguous 512-byte
// request.  History (profiles/): one quad per thread meant 131,072 CTAs for 512^3 -- the emit pass, whose CTAs mostly return at
// once, still took 0.17 ms: block scheduling, not memory, was the limit; one WARP per block (8 quads per lane, shuffle scans)
// removed the scheduling cost but serialised the dependent neighbour fetches of the surface quads and was 4x slower.
// With an activity map the quads of inactive stretches are not loaded at all.
constexpr int EX_QPT = 4;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[EX_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < EX_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < EX_THREADS / 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < EX_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const int base = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[EX_THREADS / 32 - 1];
    __syncthreads();
    return base + incl - v;
}

template <int VX>
__device__ __forceinline__ int quad_count(const ExtractParams &p, size_t v0, const uint32_t (&own)[VX])
{
    int n = 0;
    thread_crossings<VX>(p, v0, own, [&](const float3) { ++n; });
    return n;
}

template <int VX>
__device__ __forceinline__ void load_block_quads(const ExtractParams &p, size_t q0, uint32_t (&own)[EX_QPT][VX])
{
#pragma unroll
    for (int i = 0; i < EX_QPT; ++i) {
        const size_t v0 = (q0 + (size_t)(i * EX_THREADS) + threadIdx.x) * VX;
#pragma unroll
        for (int j = 0; j < VX; ++j) own[i][j] = 0u;
        if (p.activity && v0 < p.nvox && !p.activity[v0 / DF_ACTIVITY_VOXELS]) continue;    // cannot hold surface: not even loaded
        load_quad<VX>(p, v0, own[i]);
    }
}

template <int VX>
__global__ void __launch_bounds__(EX_THREADS) extract_count_kernel(const ExtractParams p, int *block_counts)
{
    DF_PDL_ENTRY();
    const size_t q0 = (size_t)blockIdx.x * EX_THREADS * EX_QPT;
    uint32_t own[EX_QPT][VX];
    load_block_quads<VX>(p, q0, own);
    int n = 0;
#pragma unroll
    for (int i = 0; i < EX_QPT; ++i) n += quad_count<VX>(p, (q0 + (size_t)(i * EX_THREADS) + threadIdx.x) * VX, own[i]);
    const int any = __syncthreads_or(n);
    if (!any) { if (threadIdx.x == 0) block_counts[blockIdx.x] = 0; return; }
    int total;
    block_exclusive_scan(n, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// two-level exclusive scan over the per-block counts: 1024 counts per "super" block (warp-shuffle scan), then one block
// over the <= 1024 super totals.  (A single-block serial-chunk scan of the 131,072 counts of a 512^3 volume took 0.26 ms.)
__global__ void __launch_bounds__(1024) extract_scan_local_kernel(const int *counts, int *offsets, int n, int *super_tot)
{
    DF_PDL_ENTRY();
    __shared__ int wsum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i = blockIdx.x * 1024 + t;
    const int c = i < n ? counts[i] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
        wsum[lane] = w;
    }
    __syncthreads();
    const int base = warp ? wsum[warp - 1] : 0;
    if (i < n) offsets[i] = base + incl - c;
    if (t == 1023) super_tot[blockIdx.x] = wsum[31];
}

__global__ void __launch_bounds__(1024) extract_scan_super_kernel(const int *super_tot, int nsuper, int *super_off, int capacity, int *count_out)
{
    DF_PDL_ENTRY();
    __shared__ int wsum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c = t < nsuper ? super_tot[t] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
        wsum[lane] = w;
    }
    __syncthreads();
    const int base = warp ? wsum[warp - 1] : 0;
    if (t < nsuper) super_off[t] = base + incl - c;
    if (t == 1023) *count_out = min(wsum[31], capacity);
}

template <int VX>
__global__ void __launch_bounds__(EX_THREADS) extract_emit_kernel(const ExtractParams p, const int *block_counts, const int *offsets,
                                                                  const int *super_off, float4 *out, int capacity)
{
    DF_PDL_ENTRY();
    if (block_counts[blockIdx.x] == 0) return;
    const size_t q0 = (size_t)blockIdx.x * EX_THREADS * EX_QPT;
    uint32_t own[EX_QPT][VX];
    load_block_quads<VX>(p, q0, own);
    int run = super_off[blockIdx.x >> 10] + offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < EX_QPT; ++i) {                                // points leave in quad order i * EX_THREADS + tid
        const size_t v0 = (q0 + (size_t)(i * EX_THREADS) + threadIdx.x) * VX;
        const int c = quad_count<VX>(p, v0, own[i]);
        int total;
        int k = run + block_exclusive_scan(c, &total);
        if (c)
            thread_crossings<VX>(p, v0, own[i], [&](const float3 q) {
                if (k < capacity) out[k] = make_float4(q.x, q.y, q.z, 0.f);
                ++k;
            });
        run += total;
    }
}

ExtractParams make_params(const df_volume &vol, const df_aff3f &pose, int vx)
{
    ExtractParams p;
    p.data = vol.data;
    p.Dx = vol.dims[0]; p.Dy = vol.dims[1]; p.Dz = vol.dims[2];
    p.vs = make_float3(vol.voxel_size[0], vol.voxel_size[1], vol.voxel_size[2]);
    p.pose = make_aff(pose);
    p.nvox = (size_t)vol.dims[0] * vol.dims[1] * vol.dims[2];
    p.nblocks = (int)((p.nvox + (size_t)EX_THREADS * EX_QPT * vx - 1) / ((size_t)EX_THREADS * EX_QPT * vx));
    p.activity = nullptr;
    return p;
}

int pick_vx(const df_volume &vol) { return (vol.dims[0] % 4 == 0 && ((uintptr_t)vol.data & 15u) == 0) ? 4 : 1; }

}  // namespace

extern "C" size_t df_extract_workspace_bytes(df_volume vol)
{
    const size_t nvox = (size_t)vol.dims[0] * vol.dims[1] * vol.dims[2];
    const size_t nblocks = (nvox + EX_THREADS - 1) / EX_THREADS;          // upper bound (VX = 1)
    return (2 * nblocks + 2048 + 64) * sizeof(int);
}

extern "C" int df_extract_cloud(df_volume vol, df_aff3f pose, float *out_points, int capacity, int *count, void *workspace, void *stream)
{
    return df_extract_cloud_tracked(vol, pose, out_points, capacity, count, workspace, nullptr, stream);
}

extern "C" int df_extract_cloud_tracked(df_volume vol, df_aff3f pose, float *out_points, int capacity, int *count, void *workspace,
                                        const unsigned char *activity, void *stream)
{
    const int vx = pick_vx(vol);
    ExtractParams p = make_params(vol, pose, vx);
    p.activity = activity;
    int *block_counts = (int *)workspace;
    int *offsets = block_counts + p.nblocks;
    cudaStream_t s = (cudaStream_t)stream;
    if (vx == 4) launch_pdl(extract_count_kernel<4>, dim3(p.nblocks), dim3(EX_THREADS), 0, s, p, block_counts);
    else launch_pdl(extract_count_kernel<1>, dim3(p.nblocks), dim3(EX_THREADS), 0, s, p, block_counts);
    DF_LAUNCH_CHECK();
    const int nsuper = (p.nblocks + 1023) / 1024;
    if (nsuper > 1024) return (int)cudaErrorInvalidValue;               // > 2^20 blocks (volume > 1024^3 voxels)
    int *super_tot = offsets + p.nblocks, *super_off = super_tot + 1024;
    launch_pdl(extract_scan_local_kernel, dim3(nsuper), dim3(1024), 0, s, block_counts, offsets, p.nblocks, super_tot);
    DF_LAUNCH_CHECK();
    launch_pdl(extract_scan_super_kernel, dim3(1), dim3(1024), 0, s, super_tot, nsuper, super_off, capacity, count);
    DF_LAUNCH_CHECK();
    if (vx == 4) launch_pdl(extract_emit_kernel<4>, dim3(p.nblocks), dim3(EX_THREADS), 0, s, p, block_counts, offsets, super_off, (float4 *)out_points, capacity);
    else launch_pdl(extract_emit_kernel<1>, dim3(p.nblocks), dim3(EX_THREADS), 0, s, p, block_counts, offsets, super_off, (float4 *)out_points, capacity);
    DF_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// extract normals: reference ExtractNormals::operator() tsdf_volume.cu:714-795 (launched with 8x redundant threads,
// :817-831); here one thread per point, count optionally read from device memory.
namespace {
struct NormalsParams {
    const uint32_t *data;
    int Dx, Dy, Dz;
    float3 vs_inv, gd;
    Aff pose;
    Mat3 Rinv;
    const float4 *points;
    int n;
    const int *count_dev;
    float4 *out;
};

__device__ __forceinline__ float en_tsdf(const NormalsParams &p, int x, int y, int z)
{ return half_bits_to_float((unsigned short)(__ldg(p.data + x + (size_t)p.Dx * y + (size_t)p.Dx * p.Dy * z) & 0xffffu)); }

__device__ __forceinline__ float en_interpolate(const NormalsParams &p, const float3 cf)
{
    const float fx = floorf(cf.x), fy = floorf(cf.y), fz = floorf(cf.z);
    if (!(fx >= 0) || !(fy >= 0) || !(fz >= 0) || !(fx < (float)(p.Dx - 1)) || !(fy < (float)(p.Dy - 1)) || !(fz < (float)(p.Dz - 1)))
        return qnan();
    const int gx = (int)fx, gy = (int)fy, gz = (int)fz;
    const float a = cf.x - (float)gx, b = cf.y - (float)gy, c = cf.z - (float)gz;
    const float v000 = en_tsdf(p, gx, gy, gz), v001 = en_tsdf(p, gx, gy, gz + 1);
    const float v010 = en_tsdf(p, gx, gy + 1, gz), v011 = en_tsdf(p, gx, gy + 1, gz + 1);
    const float v100 = en_tsdf(p, gx + 1, gy, gz), v101 = en_tsdf(p, gx + 1, gy, gz + 1);
    const float v110 = en_tsdf(p, gx + 1, gy + 1, gz), v111 = en_tsdf(p, gx + 1, gy + 1, gz + 1);
    float tsdf = 0.f;
    tsdf += v000 * (1 - a) * (1 - b) * (1 - c);
    tsdf += v001 * (1 - a) * (1 - b) * c;
    tsdf += v010 * (1 - a) * b * (1 - c);
    tsdf += v011 * (1 - a) * b * c;
    tsdf += v100 * a * (1 - b) * (1 - c);
    tsdf += v101 * a * (1 - b) * c;
    tsdf += v110 * a * b * (1 - c);
    tsdf += v111 * a * b * c;
    return tsdf;
}

__global__ void __launch_bounds__(256) extract_normals_kernel(const NormalsParams p)
{
    DF_PDL_ENTRY();
    const int n = p.count_dev ? min(*p.count_dev, p.n) : p.n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const float nanv = qnan();
        float3 nrm = make_float3(nanv, nanv, nanv);
        const float4 pt = p.points[idx];
        const float3 point = mat3_mul(p.Rinv, sub3(make_float3(pt.x, pt.y, pt.z), p.pose.t));
        const int gx = __float2int_rn(point.x * p.vs_inv.x), gy = __float2int_rn(point.y * p.vs_inv.y), gz = __float2int_rn(point.z * p.vs_inv.z);
        if (gx > 1 && gy > 1 && gz > 1 && gx < p.Dx - 2 && gy < p.Dy - 2 && gz < p.Dz - 2) {
            float3 t;
            t = point; t.x += p.gd.x; const float Fx1 = en_interpolate(p, mul3(t, p.vs_inv));
            t = point; t.x -= p.gd.x; const float Fx2 = en_interpolate(p, mul3(t, p.vs_inv));
            nrm.x = (Fx1 - Fx2) / p.gd.x;
            t = point; t.y += p.gd.y; const float Fy1 = en_interpolate(p, mul3(t, p.vs_inv));
            t = point; t.y -= p.gd.y; const float Fy2 = en_interpolate(p, mul3(t, p.vs_inv));
            nrm.y = (Fy1 - Fy2) / p.gd.y;
            t = point; t.z += p.gd.z; const float Fz1 = en_interpolate(p, mul3(t, p.vs_inv));
            t = point; t.z -= p.gd.z; const float Fz2 = en_interpolate(p, mul3(t, p.vs_inv));
            nrm.z = (Fz1 - Fz2) / p.gd.z;
            nrm = normalized3(mat_mul(p.pose.r0, p.pose.r1, p.pose.r2, nrm));
        }
        p.out[idx] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
    }
}
}  // namespace

extern "C" int df_extract_normals(df_volume vol, const float *points, int n_points, const int *count_dev, df_aff3f pose,
                                  const float *Rinv_host9, float delta_factor, float *out_normals, void *stream)
{
    if (n_points <= 0) return 0;
    NormalsParams p;
    p.data = vol.data;
    p.Dx = vol.dims[0]; p.Dy = vol.dims[1]; p.Dz = vol.dims[2];
    p.vs_inv = make_float3(1.f / vol.voxel_size[0], 1.f / vol.voxel_size[1], 1.f / vol.voxel_size[2]);
    p.gd = make_float3(vol.voxel_size[0] * delta_factor, vol.voxel_size[1] * delta_factor, vol.voxel_size[2] * delta_factor);
    p.pose = make_aff(pose);
    p.Rinv = make_mat3(Rinv_host9);
    p.points = (const float4 *)points; p.n = n_points; p.count_dev = count_dev; p.out = (float4 *)out_normals;
    const int blocks = count_dev ? 148 * 8 : div_up(n_points, 256);
    launch_pdl(extract_normals_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p);
    DF_LAUNCH_CHECK();
    return 0;
}
