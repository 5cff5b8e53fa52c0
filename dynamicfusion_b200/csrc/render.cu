// render.cu -- display-only shading kernels so that KinFu::renderImage (kinfu.cpp:312-336,407-436) has something to call.
// OUT OF SCOPE for performance: straightforward restatements of the reference's
// render_image_kernel (points variant) / tangent_colors_kernel (kfusion/src/cuda/imgproc.cu:484-572); __powf -> powf,
// rsqrt -> 1/sqrtf.
#include "df_common.cuh"

using namespace dfb;

namespace {

__global__ void __launch_bounds__(256) render_points_kernel(const float4 *points, size_t ppitch, const float4 *normals, size_t npitch,
                                                            float3 light, uchar4 *dst, size_t dpitch, int cols, int rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    float3 color;
    const float4 p4 = row_ptr(points, ppitch, y)[x];
    if (isnan(p4.x)) {
        const float3 bgr1 = make_float3(4.f / 255.f, 2.f / 255.f, 2.f / 255.f);
        const float3 bgr2 = make_float3(236.f / 255.f, 120.f / 255.f, 120.f / 255.f);
        const float w = (float)y / rows;
        color = add3(scale3(bgr1, 1 - w), scale3(bgr2, w));
    } else {
        const float3 P = make_float3(p4.x, p4.y, p4.z);
        const float4 n4 = row_ptr(normals, npitch, y)[x];
        const float3 N = make_float3(n4.x, n4.y, n4.z);
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
    row_ptr(dst, dpitch, y)[x] = out;
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
