// imgproc.cu -- depth pre-processing kernels for sm_100a (one thread per output pixel; all images are a few
// hundred KB and L2-resident, so these are launch/latency-bound, not bandwidth-bound).
// Replaces kfusion/src/cuda/imgproc.cu of the reference (cited per kernel).
#include "df_common.cuh"

using namespace dfb;

// compute_dists_kernel, imgproc.cu:259-272 (its guard `x < cols || y < rows` is always true inside the grid; the
// exact-fit guard is used here)
__global__ void __launch_bounds__(256) compute_dists_kernel(const unsigned short *depth, size_t dpitch, int cols, int rows,
                                                            float finvx, float finvy, float cx, float cy,
                                                            unsigned short *dists, size_t pitch)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float xl = ((float)x - cx) * finvx;
    const float yl = ((float)y - cy) * finvy;
    const float lambda = sqrtf(xl * xl + yl * yl + 1.f);
    row_ptr(dists, pitch, y)[x] = float_to_half_bits((float)row_ptr(depth, dpitch, y)[x] * lambda * 0.001f);
}

extern "C" int df_compute_dists(const uint16_t *depth, size_t depth_pitch, int cols, int rows, df_intr intr,
                                uint16_t *dists, size_t dists_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    launch_pdl(compute_dists_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, depth, depth_pitch, cols, rows, 1.f / intr.fx, 1.f / intr.fy,
                                                                   intr.cx, intr.cy, dists, dists_pitch);
    DF_LAUNCH_CHECK();
    return 0;
}

// bilateral_kernel, imgproc.cu:11-43.  __expf restated as expf (parity with the oracle: +-1 LSB on the u16 result).
__global__ void __launch_bounds__(256) bilateral_kernel(const unsigned short *src, size_t spitch, int cols, int rows,
                                                        unsigned short *dst, size_t dpitch, int ksz, float ss, float sd)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const int value = row_ptr(src, spitch, y)[x];
    const int tx = min(x - ksz / 2 + ksz, cols - 1);
    const int ty = min(y - ksz / 2 + ksz, rows - 1);
    float sum1 = 0, sum2 = 0;
    for (int cy = max(y - ksz / 2, 0); cy < ty; ++cy) {
        const unsigned short *srow = row_ptr(src, spitch, cy);
        for (int cx = max(x - ksz / 2, 0); cx < tx; ++cx) {
            const int depth = __ldg(srow + cx);
            const float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
            const float color2 = (float)((value - depth) * (value - depth));
            const float weight = expf(-(space2 * ss + color2 * sd));
            sum1 += (float)depth * weight;
            sum2 += weight;
        }
    }
    row_ptr(dst, dpitch, y)[x] = (unsigned short)__float2int_rn(sum1 / sum2);   // NaN -> 0, as on the reference GPU path
}

extern "C" int df_bilateral(const uint16_t *src, size_t src_pitch, int cols, int rows, uint16_t *dst, size_t dst_pitch,
                            int kernel_size, float sigma_spatial, float sigma_depth, void *stream)
{
    sigma_depth *= 1000;   // imgproc.cu:47
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    launch_pdl(bilateral_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, src, src_pitch, cols, rows, dst, dst_pitch, kernel_size,
                                                               0.5f / (sigma_spatial * sigma_spatial), 0.5f / (sigma_depth * sigma_depth));
    DF_LAUNCH_CHECK();
    return 0;
}

// truncate_depth_kernel, imgproc.cu:66-75
__global__ void __launch_bounds__(256) truncate_depth_kernel(unsigned short *depth, size_t pitch, int cols, int rows, unsigned short max_dist)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < cols && y < rows) {
        unsigned short *d = row_ptr(depth, pitch, y) + x;
        if (*d > max_dist) *d = 0;
    }
}

extern "C" int df_truncate_depth(uint16_t *depth, size_t pitch, int cols, int rows, float max_dist, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    launch_pdl(truncate_depth_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, depth, pitch, cols, rows, (unsigned short)(max_dist * 1000.f));
    DF_LAUNCH_CHECK();
    return 0;
}

// pyramid_kernel, imgproc.cu:94-123
__global__ void __launch_bounds__(256) pyramid_kernel(const unsigned short *src, size_t spitch, int scols, int srows,
                                                      unsigned short *dst, size_t dpitch, int dcols, int drows, float thr)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const int D = 5;
    const int center = row_ptr(src, spitch, 2 * y)[2 * x];
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int sum = 0, count = 0;
    for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy) {
        const unsigned short *srow = row_ptr(src, spitch, cy);
        for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
            const int val = __ldg(srow + cx);
            if ((float)abs(val - center) < thr) { sum += val; ++count; }
        }
    }
    row_ptr(dst, dpitch, y)[x] = (unsigned short)(count == 0 ? 0 : sum / count);
}

extern "C" int df_pyr_down(const uint16_t *src, size_t src_pitch, int src_cols, int src_rows, uint16_t *dst, size_t dst_pitch,
                           float sigma_depth, void *stream)
{
    sigma_depth *= 1000;   // imgproc.cu:127
    const int dcols = src_cols / 2, drows = src_rows / 2;
    dim3 block(32, 8), grid(div_up(dcols, 32), div_up(drows, 8));
    launch_pdl(pyramid_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, src, src_pitch, src_cols, src_rows, dst, dst_pitch, dcols, drows, sigma_depth * 3);
    DF_LAUNCH_CHECK();
    return 0;
}

// points_normals_kernel, imgproc.cu:210-239; Reprojector device.hpp:43-48
__device__ __forceinline__ float3 reproj(float finvx, float finvy, float cx, float cy, int u, int v, float z)
{ return make_float3(z * ((float)u - cx) * finvx, z * ((float)v - cy) * finvy, z); }

__global__ void __launch_bounds__(256) points_normals_kernel(float finvx, float finvy, float cx, float cy,
                                                             const unsigned short *depth, size_t dpitch, int cols, int rows,
                                                             float4 *points, size_t ppitch, float4 *normals, size_t npitch)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float nanv = qnan();
    float4 P = make_float4(nanv, nanv, nanv, nanv), N = P;
    if (x < cols - 1 && y < rows - 1) {
        const float z00 = (float)row_ptr(depth, dpitch, y)[x] * 0.001f;
        const float z01 = (float)row_ptr(depth, dpitch, y)[x + 1] * 0.001f;
        const float z10 = (float)row_ptr(depth, dpitch, y + 1)[x] * 0.001f;
        if (z00 * z01 * z10 != 0) {
            const float3 v00 = reproj(finvx, finvy, cx, cy, x, y, z00);
            const float3 v01 = reproj(finvx, finvy, cx, cy, x + 1, y, z01);
            const float3 v10 = reproj(finvx, finvy, cx, cy, x, y + 1, z10);
            const float3 n = normalized3(cross3(sub3(v01, v00), sub3(v10, v00)));
            N = make_float4(-n.x, -n.y, -n.z, 0.f);
            P = make_float4(v00.x, v00.y, v00.z, 0.f);
        }
    }
    row_ptr(points, ppitch, y)[x] = P;
    row_ptr(normals, npitch, y)[x] = N;
}

extern "C" int df_points_normals(df_intr intr, const uint16_t *depth, size_t depth_pitch, int cols, int rows,
                                 float *points, size_t points_pitch, float *normals, size_t normals_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    launch_pdl(points_normals_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, 1.f / intr.fx, 1.f / intr.fy, intr.cx, intr.cy, depth, depth_pitch,
                                                                    cols, rows, (float4 *)points, points_pitch, (float4 *)normals, normals_pitch);
    DF_LAUNCH_CHECK();
    return 0;
}

// resize_points_normals_kernel, imgproc.cu:368-400
__global__ void __launch_bounds__(256) resize_points_normals_kernel(const float4 *vsrc, size_t vspitch, const float4 *nsrc, size_t nspitch,
                                                                    float4 *vdst, size_t vdpitch, float4 *ndst, size_t ndpitch, int dcols, int drows)
{
    DF_PDL_ENTRY();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const float nanv = qnan();
    float4 V = make_float4(nanv, nanv, nanv, 0.f), N = V;
    const int xs = x * 2, ys = y * 2;
    const float4 d00 = row_ptr(vsrc, vspitch, ys)[xs], d01 = row_ptr(vsrc, vspitch, ys)[xs + 1];
    const float4 d10 = row_ptr(vsrc, vspitch, ys + 1)[xs], d11 = row_ptr(vsrc, vspitch, ys + 1)[xs + 1];
    if (!isnan(d00.x * d01.x * d10.x * d11.x)) {
        V = make_float4((d00.x + d01.x + d10.x + d11.x) * 0.25f, (d00.y + d01.y + d10.y + d11.y) * 0.25f,
                        (d00.z + d01.z + d10.z + d11.z) * 0.25f, 0.f);
        const float4 n00 = row_ptr(nsrc, nspitch, ys)[xs], n01 = row_ptr(nsrc, nspitch, ys)[xs + 1];
        const float4 n10 = row_ptr(nsrc, nspitch, ys + 1)[xs], n11 = row_ptr(nsrc, nspitch, ys + 1)[xs + 1];
        N = make_float4((n00.x + n01.x + n10.x + n11.x) * 0.25f, (n00.y + n01.y + n10.y + n11.y) * 0.25f,
                        (n00.z + n01.z + n10.z + n11.z) * 0.25f, 0.f);
    }
    row_ptr(vdst, vdpitch, y)[x] = V;
    row_ptr(ndst, ndpitch, y)[x] = N;
}

extern "C" int df_resize_points_normals(const float *vsrc, size_t vsrc_pitch, const float *nsrc, size_t nsrc_pitch,
                                        int src_cols, int src_rows, float *vdst, size_t vdst_pitch, float *ndst, size_t ndst_pitch,
                                        void *stream)
{
    const int dcols = src_cols / 2, drows = src_rows / 2;
    dim3 block(32, 8), grid(div_up(dcols, 32), div_up(drows, 8));
    launch_pdl(resize_points_normals_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const float4 *)vsrc, vsrc_pitch, (const float4 *)nsrc, nsrc_pitch,
                                                                           (float4 *)vdst, vdst_pitch, (float4 *)ndst, ndst_pitch, dcols, drows);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" const char *df_error_string(int status) { return status == 0 ? "success" : cudaGetErrorString((cudaError_t)status); }
extern "C" int df_version(void) { return 100; }
