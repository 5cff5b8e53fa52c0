// render.cu -- display-only shading kernels so that KinFu::renderImage (kinfu.cpp:312-336,407-436) has something to call.
// OUT OF SCOPE for performance: straightforward restatements of the reference's
// render_image_kernel (points variant) / tangent_colors_kernel (kfusion/src/cuda/imgproc.cu:484-572); __powf -> powf,
// rsqrt -> 1/sqrtf.
#include "df_common.cuh"

using namespace dfb;

namespace {

__device__ __forceinline__ uchar4 shade(bool background, float3 P, float3 N, float3 light, int y, int rows)
{
    float3 color;
    if (background) {
        const float3 bgr1 = make_float3(4.f / 255.f, 2.f / 255.f, 2.f / 255.f);
        const float3 bgr2 = make_float3(236.f / 255.f, 120.f / 255.f, 120.f / 255.f);
        const float w = (float)y / rows;
        color = add3(scale3(bgr1, 1 - w), scale3(bgr2, w));
    } else {
        const float Ka = 0.3f, Kd = 0.5f, Ks = 0.2f, n = 20.f;
        const float3 L = normalized3(sub3(light, P));
        const float3 V = normalized3(sub3(make_float3(0.f, 0.f, 0.f), P));
        const float3 R = normalized3(sub3(scale3(N, 2 * dot3(N, L)), L));
        const float Ix = Ka + Kd * fmaxf(0.f, dot3(N, L)) + Ks * powf(fmaxf(0.f, dot3(R, V)), n);
        color = make_float3(Ix, Ix, Ix);
    }
    uchar4 out;
    out.x = (unsigned char)(__saturatef(color.x) * 255.f);
    out.y = (unsigned char)(__saturatef(color.y) * 255.f);
    out.z = (unsigned char)(__saturatef(color.z) * 255.f);
    out.w = 0;
    return out;
}

// depth variant of render_image_kernel (imgproc.cu:420-472): vertex = Reprojector(x, y, d * 0.001f), background where d == 0
__global__ void __launch_bounds__(256) render_depth_kernel(const unsigned short *depth, size_t dpitch, const float4 *normals, size_t npitch,
                                                           float finvx, float finvy, float cx, float cy, float3 light, uchar4 *dst, size_t opitch,
                                                           int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const int d = row_ptr(depth, dpitch, y)[x];
    const float z = (float)d * 0.001f;
    const float3 P = make_float3(z * ((float)x - cx) * finvx, z * ((float)y - cy) * finvy, z);
    const float4 n4 = row_ptr(normals, npitch, y)[x];
    row_ptr(dst, opitch, y)[x] = shade(d == 0, P, make_float3(n4.x, n4.y, n4.z), light, y, rows);
}

// compute_normals_kernel + mask_depth_kernel (imgproc.cu:145-188): normals from the depth map itself, then depth := 0 where no normal
__global__ void __launch_bounds__(256) depth_normals_kernel(const unsigned short *depth, size_t dpitch, float finvx, float finvy, float cx, float cy,
                                                            float4 *normals, size_t npitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float nanv = qnan();
    float4 out = make_float4(nanv, nanv, nanv, 0.f);
    if (x < cols - 1 && y < rows - 1) {
        const float z00 = (float)row_ptr(depth, dpitch, y)[x] * 0.001f;
        const float z01 = (float)row_ptr(depth, dpitch, y)[x + 1] * 0.001f;
        const float z10 = (float)row_ptr(depth, dpitch, y + 1)[x] * 0.001f;
        if (z00 * z01 * z10 != 0) {
            const float3 v00 = make_float3(z00 * ((float)x - cx) * finvx, z00 * ((float)y - cy) * finvy, z00);
            const float3 v01 = make_float3(z01 * ((float)(x + 1) - cx) * finvx, z01 * ((float)y - cy) * finvy, z01);
            const float3 v10 = make_float3(z10 * ((float)x - cx) * finvx, z10 * ((float)(y + 1) - cy) * finvy, z10);
            const float3 n = normalized3(cross3(sub3(v01, v00), sub3(v10, v00)));
            out = make_float4(-n.x, -n.y, -n.z, 0.f);
        }
    }
    row_ptr(normals, npitch, y)[x] = out;
}

__global__ void __launch_bounds__(256) mask_depth_kernel(const float4 *normals, size_t npitch, unsigned short *depth, size_t dpitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    if (isnan(row_ptr(normals, npitch, y)[x].x)) row_ptr(depth, dpitch, y)[x] = 0;
}

// cloud_to_depth_kernel (imgproc.cu:277-287): z in metres -> u16 millimetres (C conversion: truncation; NaN -> 0 as cvt.rzi does)
__global__ void __launch_bounds__(256) cloud_to_depth_kernel(const float4 *cloud, size_t cpitch, unsigned short *depth, size_t dpitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    row_ptr(depth, dpitch, y)[x] = (unsigned short)(row_ptr(cloud, cpitch, y)[x].z * 1000);
}

// resize_depth_normals_kernel (imgproc.cu:307-344): 2x2 mean of depth (integer) and normals where all four depths are non-zero
__global__ void __launch_bounds__(256) resize_depth_normals_kernel(const unsigned short *dsrc, size_t dspitch, const float4 *nsrc, size_t nspitch,
                                                                   unsigned short *ddst, size_t ddpitch, float4 *ndst, size_t ndpitch, int dcols, int drows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const float nanv = qnan();
    unsigned short d = 0;
    float4 n = make_float4(nanv, nanv, nanv, nanv);
    const int xs = x * 2, ys = y * 2;
    const int d00 = row_ptr(dsrc, dspitch, ys)[xs], d01 = row_ptr(dsrc, dspitch, ys)[xs + 1];
    const int d10 = row_ptr(dsrc, dspitch, ys + 1)[xs], d11 = row_ptr(dsrc, dspitch, ys + 1)[xs + 1];
    if (d00 * d01 != 0 && d10 * d11 != 0) {
        d = (unsigned short)((d00 + d01 + d10 + d11) / 4);
        const float4 n00 = row_ptr(nsrc, nspitch, ys)[xs], n01 = row_ptr(nsrc, nspitch, ys)[xs + 1];
        const float4 n10 = row_ptr(nsrc, nspitch, ys + 1)[xs], n11 = row_ptr(nsrc, nspitch, ys + 1)[xs + 1];
        n.x = (float)((double)(n00.x + n01.x + n10.x + n11.x) * 0.25);
        n.y = (float)((double)(n00.y + n01.y + n10.y + n11.y) * 0.25);
        n.z = (float)((double)(n00.z + n01.z + n10.z + n11.z) * 0.25);
    }
    row_ptr(ddst, ddpitch, y)[x] = d;
    row_ptr(ndst, ndpitch, y)[x] = n;
}

__global__ void __launch_bounds__(256) render_points_kernel(const float4 *points, size_t ppitch, const float4 *normals, size_t npitch,
                                                            float3 light, uchar4 *dst, size_t dpitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float4 p4 = row_ptr(points, ppitch, y)[x];
    const float4 n4 = row_ptr(normals, npitch, y)[x];
    row_ptr(dst, dpitch, y)[x] = shade(isnan(p4.x), make_float3(p4.x, p4.y, p4.z), make_float3(n4.x, n4.y, n4.z), light, y, rows);
}

__global__ void __launch_bounds__(256) tangent_colors_kernel(const float4 *normals, size_t npitch, uchar4 *dst, size_t dpitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float4 n = row_ptr(normals, npitch, y)[x];
    const unsigned char r = (unsigned char)((5.f - n.x * 3.5f) * 25.5f);
    const unsigned char g = (unsigned char)((5.f - n.y * 2.5f) * 25.5f);
    const unsigned char b = (unsigned char)((5.f - n.z * 3.5f) * 25.5f);
    row_ptr(dst, dpitch, y)[x] = make_uchar4(b, g, r, 0);
}

}  // namespace

extern "C" int df_render_image(const float *points, size_t points_pitch, const float *normals, size_t normals_pitch, int cols, int rows,
                               const float *light_pose_host3, void *image_bgra, size_t image_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    render_points_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float4 *)points, points_pitch, (const float4 *)normals, normals_pitch,
                                                                   make_float3(light_pose_host3[0], light_pose_host3[1], light_pose_host3[2]),
                                                                   (uchar4 *)image_bgra, image_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" int df_render_tangent_colors(const float *normals, size_t normals_pitch, int cols, int rows, void *image_bgra, size_t image_pitch,
                                        void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    tangent_colors_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float4 *)normals, normals_pitch, (uchar4 *)image_bgra, image_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    return 0;
}

// ---- the reference's USE_DEPTH-path image operations (cuda/imgproc.hpp:15,21,23,31): not called by the default frame loop, provided so
// ---- that every function of the public header works.  Compared with the reference's own kernels in tests/test_render_gpu.py.
extern "C" int df_render_image_depth(const uint16_t *depth, size_t depth_pitch, const float *normals, size_t normals_pitch, int cols, int rows,
                                     df_intr intr, const float *light_pose_host3, void *image_bgra, size_t image_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    render_depth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(depth, depth_pitch, (const float4 *)normals, normals_pitch, 1.f / intr.fx, 1.f / intr.fy,
                                                                  intr.cx, intr.cy, make_float3(light_pose_host3[0], light_pose_host3[1], light_pose_host3[2]),
                                                                  (uchar4 *)image_bgra, image_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" int df_normals_mask_depth(df_intr intr, uint16_t *depth, size_t depth_pitch, int cols, int rows, float *normals, size_t normals_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    depth_normals_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(depth, depth_pitch, 1.f / intr.fx, 1.f / intr.fy, intr.cx, intr.cy, (float4 *)normals,
                                                                   normals_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    mask_depth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float4 *)normals, normals_pitch, depth, depth_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" int df_cloud_to_depth(const float *cloud, size_t cloud_pitch, int cols, int rows, uint16_t *depth, size_t depth_pitch, void *stream)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    cloud_to_depth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float4 *)cloud, cloud_pitch, depth, depth_pitch, cols, rows);
    DF_LAUNCH_CHECK();
    return 0;
}

extern "C" int df_resize_depth_normals(const uint16_t *dsrc, size_t dsrc_pitch, const float *nsrc, size_t nsrc_pitch, int src_cols, int src_rows,
                                       uint16_t *ddst, size_t ddst_pitch, float *ndst, size_t ndst_pitch, void *stream)
{
    const int dcols = src_cols / 2, drows = src_rows / 2;
    dim3 block(32, 8), grid(div_up(dcols, 32), div_up(drows, 8));
    resize_depth_normals_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dsrc, dsrc_pitch, (const float4 *)nsrc, nsrc_pitch, ddst, ddst_pitch,
                                                                          (float4 *)ndst, ndst_pitch, dcols, drows);
    DF_LAUNCH_CHECK();
    return 0;
}
