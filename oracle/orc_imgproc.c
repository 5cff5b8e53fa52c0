/* CPU ORACLE (test infrastructure only) -- depth pre-processing stages.
 * Restates kfusion/src/cuda/imgproc.cu of the reference; see orc_common.h for the numerics policy. */
#include "orc_common.h"
#include <stdlib.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* bilateral_kernel, imgproc.cu:11-43; launcher :45-57 (sigma_depth *= 1000; 0.5f/(s*s) computed on host).
 * __expf restated as expf; __float2int_rn as rintf. */
void orc_bilateral(const uint16_t *src, size_t spitch, int cols, int rows, uint16_t *dst, size_t dpitch, int ksz,
                   float sigma_spatial, float sigma_depth)
{
    sigma_depth *= 1000;
    const float ss = 0.5f / (sigma_spatial * sigma_spatial);
    const float sd = 0.5f / (sigma_depth * sigma_depth);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int value = orc_row_u16(src, spitch, y)[x];
            int tx = imin(x - ksz / 2 + ksz, cols - 1);
            int ty = imin(y - ksz / 2 + ksz, rows - 1);
            float sum1 = 0, sum2 = 0;
            for (int cy = imax(y - ksz / 2, 0); cy < ty; ++cy)
                for (int cx = imax(x - ksz / 2, 0); cx < tx; ++cx) {
                    int depth = orc_row_u16(src, spitch, cy)[cx];
                    float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
                    float color2 = (float)((value - depth) * (value - depth));
                    float weight = expf(-(space2 * ss + color2 * sd));
                    sum1 += (float)depth * weight;
                    sum2 += weight;
                }
            /* dst is ushort: int -> ushort conversion wraps; sum2 == 0 (empty window at the last row/col) gives
             * NaN -> __float2int_rn(NaN) = 0 on the GPU */
            float r = sum1 / sum2;
            int ri = (r == r) ? (int)rintf(r) : 0;
            orc_row_u16w(dst, dpitch, y)[x] = (uint16_t)ri;
        }
}

/* truncate_depth_kernel, imgproc.cu:66-85 */
void orc_truncate_depth(uint16_t *depth, size_t pitch, int cols, int rows, float max_dist)
{
    uint16_t md = (uint16_t)(max_dist * 1000.f);
    for (int y = 0; y < rows; ++y) {
        uint16_t *d = orc_row_u16w(depth, pitch, y);
        for (int x = 0; x < cols; ++x) if (d[x] > md) d[x] = 0;
    }
}

/* pyramid_kernel, imgproc.cu:94-123; launcher :125-136 (sigma_depth*1000*3); dst = src/2 */
void orc_pyr_down(const uint16_t *src, size_t spitch, int scols, int srows, uint16_t *dst, size_t dpitch, float sigma_depth)
{
    sigma_depth *= 1000;
    const float thr = sigma_depth * 3;
    const int dcols = scols / 2, drows = srows / 2;
    const int D = 5;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            int center = orc_row_u16(src, spitch, 2 * y)[2 * x];
            int tx = imin(2 * x - D / 2 + D, scols - 1);
            int ty = imin(2 * y - D / 2 + D, srows - 1);
            int sum = 0, count = 0;
            for (int cy = imax(0, 2 * y - D / 2); cy < ty; ++cy)
                for (int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx) {
                    int val = orc_row_u16(src, spitch, cy)[cx];
                    if ((float)abs(val - center) < thr) { sum += val; ++count; }
                }
            orc_row_u16w(dst, dpitch, y)[x] = (uint16_t)(count == 0 ? 0 : sum / count);
        }
}

/* Reprojector::operator(), device.hpp:43-48: x = z*(u-cx)*finv.x, left to right */
static inline orc_f3 orc_reproj(orc_intr k, float finvx, float finvy, int u, int v, float z)
{ return f3(z * ((float)u - k.cx) * finvx, z * ((float)v - k.cy) * finvy, z); }

/* points_normals_kernel, imgproc.cu:210-239 */
void orc_points_normals(orc_intr intr, const uint16_t *depth, size_t dpitch, int cols, int rows, float *points, size_t ppitch,
                        float *normals, size_t npitch)
{
    const float finvx = 1.f / intr.fx, finvy = 1.f / intr.fy;
    const float qnan = orc_qnan();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float *P = orc_row_f4w(points, ppitch, y) + 4 * x;
            float *Nn = orc_row_f4w(normals, npitch, y) + 4 * x;
            P[0] = P[1] = P[2] = P[3] = qnan;
            Nn[0] = Nn[1] = Nn[2] = Nn[3] = qnan;
            if (x >= cols - 1 || y >= rows - 1) continue;
            float z00 = (float)orc_row_u16(depth, dpitch, y)[x] * 0.001f;
            float z01 = (float)orc_row_u16(depth, dpitch, y)[x + 1] * 0.001f;
            float z10 = (float)orc_row_u16(depth, dpitch, y + 1)[x] * 0.001f;
            if (z00 * z01 * z10 != 0) {
                orc_f3 v00 = orc_reproj(intr, finvx, finvy, x, y, z00);
                orc_f3 v01 = orc_reproj(intr, finvx, finvy, x + 1, y, z01);
                orc_f3 v10 = orc_reproj(intr, finvx, finvy, x, y + 1, z10);
                orc_f3 n = orc_normalized(orc_cross(orc_sub(v01, v00), orc_sub(v10, v00)));
                Nn[0] = -n.x; Nn[1] = -n.y; Nn[2] = -n.z; Nn[3] = 0.f;
                P[0] = v00.x; P[1] = v00.y; P[2] = v00.z; P[3] = 0.f;
            }
        }
}

/* resize_points_normals_kernel, imgproc.cu:368-400; dst = src/2 */
void orc_resize_points_normals(const float *vsrc, size_t vspitch, const float *nsrc, size_t nspitch, int scols, int srows,
                               float *vdst, size_t vdpitch, float *ndst, size_t ndpitch)
{
    const int dcols = scols / 2, drows = srows / 2;
    const float qnan = orc_qnan();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            float *V = orc_row_f4w(vdst, vdpitch, y) + 4 * x;
            float *Nn = orc_row_f4w(ndst, ndpitch, y) + 4 * x;
            V[0] = V[1] = V[2] = qnan; V[3] = 0.f;
            Nn[0] = Nn[1] = Nn[2] = qnan; Nn[3] = 0.f;
            int xs = x * 2, ys = y * 2;
            const float *d00 = orc_row_f4(vsrc, vspitch, ys) + 4 * xs, *d01 = d00 + 4;
            const float *d10 = orc_row_f4(vsrc, vspitch, ys + 1) + 4 * xs, *d11 = d10 + 4;
            if (!isnan(d00[0] * d01[0] * d10[0] * d11[0])) {
                for (int c = 0; c < 3; ++c) V[c] = (d00[c] + d01[c] + d10[c] + d11[c]) * 0.25f;
                V[3] = 0.f;
                const float *n00 = orc_row_f4(nsrc, nspitch, ys) + 4 * xs, *n01 = n00 + 4;
                const float *n10 = orc_row_f4(nsrc, nspitch, ys + 1) + 4 * xs, *n11 = n10 + 4;
                for (int c = 0; c < 3; ++c) Nn[c] = (n00[c] + n01[c] + n10[c] + n11[c]) * 0.25f;
                Nn[3] = 0.f;
            }
        }
}
