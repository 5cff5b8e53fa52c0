/* CPU ORACLE (test infrastructure only) -- TSDF volume stages.
 * Restates kfusion/src/cuda/tsdf_volume.cu + device.hpp of the reference; see orc_common.h for the
 * numerics policy.  Every function cites the reference lines it follows. */
#include "orc_common.h"
#include <stdlib.h>

/* ---------------------------------------------------------------------------------------------
 * half <-> float, round-to-nearest-even  (__float2half_rn / __half2float, device.hpp:53-61) */
uint16_t orc_float2half_rn(float f)
{
    union { float f; uint32_t u; } c; c.f = f;
    uint32_t x = c.u;
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7fffu);          /* NaN -> canonical NaN (CUDA: 0x7fff) */
    if (absx >= 0x477ff000u) {                                          /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000001u) return (uint16_t)sign;                      /* <= 2^-25 rounds to zero (ties to even) */
    int e = (int)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;                        /* 24-bit significand */
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                  /* subnormal half */
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));                               /* carry propagates into exponent */
}

float orc_half2float(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    union { float f; uint32_t u; } c;
    if (e == 0) {
        if (m == 0) { c.u = sign; return c.f; }
        /* subnormal: m * 2^-24 */
        float v = (float)m * 5.9604644775390625e-08f;
        c.f = v; c.u |= sign; return c.f;
    }
    if (e == 31) { c.u = sign | 0x7f800000u | (m << 13); return c.f; }
    c.u = sign | ((e + 112u) << 23) | (m << 13);
    return c.f;
}

/* clear_volume_kernel, tsdf_volume.cu:15-28: every voxel = pack_tsdf(0.f, 0) */
void orc_clear_volume(orc_volume vol)
{
    size_t n = (size_t)vol.dims[0] * vol.dims[1] * vol.dims[2];
    uint32_t v = orc_pack_tsdf(0.f, 0);
    for (size_t i = 0; i < n; ++i) vol.data[i] = v;
}

/* compute_dists_kernel, imgproc.cu:259-272 (finv = 1/f on the host, imgproc.cu:291) */
void orc_compute_dists(const uint16_t *depth, size_t dpitch, int cols, int rows, orc_intr intr, uint16_t *dists, size_t pitch)
{
    float finvx = 1.f / intr.fx, finvy = 1.f / intr.fy;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y) {
        const uint16_t *d = orc_row_u16(depth, dpitch, y);
        uint16_t *o = orc_row_u16w(dists, pitch, y);
        for (int x = 0; x < cols; ++x) {
            float xl = ((float)x - intr.cx) * finvx;
            float yl = ((float)y - intr.cy) * finvy;
            float lambda = sqrtf(xl * xl + yl * yl + 1.f);
            o[x] = orc_float2half_rn((float)d[x] * lambda * 0.001f);
        }
    }
}

/* Projector::operator(), device.hpp:32-38 (division first, then fma) */
static inline void orc_project(orc_intr k, orc_f3 p, float *u, float *v)
{
    *u = fmaf(k.fx, p.x / p.z, k.cx);
    *v = fmaf(k.fy, p.y / p.z, k.cy);
}

/* TsdfIntegrator::operator(), tsdf_volume.cu:51-112; launcher :141-161 (tranc_dist_inv = 1.f/trunc).
 * Returns N_upd = number of voxels written (the algorithmic-byte count of SURVEY 8d). */
long long orc_integrate(orc_volume vol, const uint16_t *dists, size_t pitch, int cols, int rows, orc_aff3f vol2cam, orc_intr intr)
{
    const float trunc_inv = 1.f / vol.trunc_dist;
    const int Dx = vol.dims[0], Dy = vol.dims[1], Dz = vol.dims[2];
    const orc_f3 zstep = orc_scale(f3(vol2cam.R[2], vol2cam.R[5], vol2cam.R[8]), vol.voxel_size[2]);
    long long n_upd = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_upd)
    for (int y = 0; y < Dy; ++y)
        for (int x = 0; x < Dx; ++x) {
            orc_f3 vx = f3((float)x * vol.voxel_size[0], (float)y * vol.voxel_size[1], 0.f);
            orc_f3 vc = orc_aff_mul(&vol2cam, vx);
            uint32_t *vptr = vol.data + x + (size_t)Dx * y;
            for (int i = 0; i < Dz; ++i, vc = orc_add(vc, zstep), vptr += (size_t)Dx * Dy) {
                float u, v;
                orc_project(intr, vc, &u, &v);
                if (u < 0 || v < 0 || u >= (float)cols || v >= (float)rows) continue;
                /* The reference fetches Dp first and then tests (Dp == 0 || vc.z <= 0); testing vc.z first is
                 * equivalent (both 'continue') and keeps NaN coordinates (vc.z == 0) away from the lookup. */
                if (vc.z <= 0) continue;
                if (!(u == u) || !(v == v)) continue;
                float Dp = orc_half2float(orc_row_u16(dists, pitch, (int)v)[(int)u]);   /* point sampling = floor */
                if (Dp == 0) continue;
                float sdf = Dp - sqrtf(orc_dot(vc, vc));
                if (sdf >= -vol.trunc_dist) {
                    float tsdf = fminf(1.f, sdf * trunc_inv);
                    int weight_prev;
                    float tsdf_prev = orc_unpack_tsdf(*vptr, &weight_prev);
                    float tsdf_new = fmaf(tsdf_prev, (float)weight_prev, tsdf) / (float)(weight_prev + 1);
                    int weight_new = weight_prev + 1 < vol.max_weight ? weight_prev + 1 : vol.max_weight;
                    *vptr = orc_pack_tsdf(tsdf_new, weight_new);
                    ++n_upd;
                }
            }
        }
    return n_upd;
}

/* project_kernel, tsdf_volume.cu:114-137 (note: its guard is `x < cols || y < rows`; launch grid covers the
 * vertex map exactly so the guard is always true).  dists is read as half (tex) and written as u16 0. */
void orc_project_and_remove(uint16_t *dists, size_t pitch, int cols, int rows, orc_intr intr, float *points, size_t ppitch,
                            int pcols, int prows)
{
    /* Pass 1 reads, pass 2 scatters zeros: the reference races reads of the texture with the scatter; the
     * deterministic restatement reads the ORIGINAL depth everywhere (documented in DESIGN.md). */
    uint8_t *kill = (uint8_t *)calloc((size_t)cols * rows, 1);
    for (int y = 0; y < prows; ++y) {
        float *prow = orc_row_f4w(points, ppitch, y);
        for (int x = 0; x < pcols; ++x) {
            float *pt = prow + 4 * x;
            if (isnan(pt[0]) || isnan(pt[1]) || isnan(pt[2])) continue;
            float u, v;
            orc_project(intr, f3(pt[0], pt[1], pt[2]), &u, &v);
            if (!(u >= 0 && v >= 0 && v < (float)rows && u < (float)cols)) {   /* NaN coords count as off-image */
                pt[0] = pt[1] = pt[2] = orc_qnan(); pt[3] = 0.f;
                continue;
            }
            float Dp = orc_half2float(orc_row_u16(dists, pitch, (int)v)[(int)u]);
            kill[(size_t)(int)v * cols + (int)u] = 1;
            pt[0] = u * Dp; pt[1] = v * Dp; pt[2] = Dp; pt[3] = 0.f;
        }
    }
    for (int y = 0; y < rows; ++y) {
        uint16_t *d = orc_row_u16w(dists, pitch, y);
        for (int x = 0; x < cols; ++x) if (kill[(size_t)y * cols + x]) d[x] = 0;
    }
    free(kill);
}

/* interpolate(), tsdf_volume.cu:220-245 */
float orc_interpolate(const orc_volume *vol, orc_f3 cf)
{
    const int Dx = vol->dims[0], Dy = vol->dims[1], Dz = vol->dims[2];
    float fx = floorf(cf.x), fy = floorf(cf.y), fz = floorf(cf.z);
    if (!(fx >= 0) || !(fy >= 0) || !(fz >= 0) || !(fx < (float)(Dx - 1)) || !(fy < (float)(Dy - 1)) || !(fz < (float)(Dz - 1)))
        return orc_qnan();
    int gx = (int)fx, gy = (int)fy, gz = (int)fz;
    float a = cf.x - (float)gx, b = cf.y - (float)gy, c = cf.z - (float)gz;
    const uint32_t *d = vol->data;
    const size_t sy = (size_t)Dx, sz = (size_t)Dx * Dy;
    const size_t o = gx + sy * gy + sz * gz;
#define ORC_V(dx, dy, dz) orc_half2float((uint16_t)(d[o + (dx) + sy * (dy) + sz * (dz)] & 0xffffu))
    float tsdf = 0.f;
    tsdf += ORC_V(0, 0, 0) * (1 - a) * (1 - b) * (1 - c);
    tsdf += ORC_V(0, 0, 1) * (1 - a) * (1 - b) * c;
    tsdf += ORC_V(0, 1, 0) * (1 - a) * b * (1 - c);
    tsdf += ORC_V(0, 1, 1) * (1 - a) * b * c;
    tsdf += ORC_V(1, 0, 0) * a * (1 - b) * (1 - c);
    tsdf += ORC_V(1, 0, 1) * a * (1 - b) * c;
    tsdf += ORC_V(1, 1, 0) * a * b * (1 - c);
    tsdf += ORC_V(1, 1, 1) * a * b * c;
#undef ORC_V
    return tsdf;
}

typedef struct {
    orc_volume vol;
    orc_aff3f aff;
    float Rinv[9];
    orc_f3 volume_size, gradient_delta, voxel_size_inv;
    float time_step;
    orc_intr intr;
    float finvx, finvy;
} orc_raycaster;

/* fetch_tsdf(), tsdf_volume.cu:263-270: __float2int_rn == rintf; the reference does not bounds-check, the
 * restatement clamps for memory safety (a no-op whenever the reference's access is in bounds). */
static inline float orc_fetch_tsdf(const orc_raycaster *rc, orc_f3 p)
{
    int x = (int)rintf(p.x * rc->voxel_size_inv.x);
    int y = (int)rintf(p.y * rc->voxel_size_inv.y);
    int z = (int)rintf(p.z * rc->voxel_size_inv.z);
    const int *D = rc->vol.dims;
    x = x < 0 ? 0 : (x > D[0] - 1 ? D[0] - 1 : x);
    y = y < 0 ? 0 : (y > D[1] - 1 ? D[1] - 1 : y);
    z = z < 0 ? 0 : (z > D[2] - 1 ? D[2] - 1 : z);
    return orc_half2float((uint16_t)(rc->vol.data[x + (size_t)D[0] * y + (size_t)D[0] * D[1] * z] & 0xffffu));
}

/* compute_normal(), tsdf_volume.cu:409-426 (divides by delta, not 2*delta) */
static orc_f3 orc_compute_normal(const orc_volume *vol, orc_f3 p, orc_f3 gd, orc_f3 vinv)
{
    orc_f3 n;
    float Fx1 = orc_interpolate(vol, orc_mul(f3(p.x + gd.x, p.y, p.z), vinv));
    float Fx2 = orc_interpolate(vol, orc_mul(f3(p.x - gd.x, p.y, p.z), vinv));
    n.x = (Fx1 - Fx2) / gd.x;
    float Fy1 = orc_interpolate(vol, orc_mul(f3(p.x, p.y + gd.y, p.z), vinv));
    float Fy2 = orc_interpolate(vol, orc_mul(f3(p.x, p.y - gd.y, p.z), vinv));
    n.y = (Fy1 - Fy2) / gd.y;
    float Fz1 = orc_interpolate(vol, orc_mul(f3(p.x, p.y, p.z + gd.z), vinv));
    float Fz2 = orc_interpolate(vol, orc_mul(f3(p.x, p.y, p.z - gd.z), vinv));
    n.z = (Fz1 - Fz2) / gd.z;
    return orc_normalized(n);
}

/* TsdfRaycaster::operator()(points, normals), tsdf_volume.cu:341-405; intersect() :202-218;
 * launcher :459-474 (volume_size = vs*dims, time_step = trunc*factor, gradient_delta = vs*factor, vs_inv = 1/vs);
 * Reprojector device.hpp:43-48. */
void orc_raycast_points(orc_volume vol, orc_aff3f cam2vol, const float *Rinv, orc_intr intr, int cols, int rows,
                        float step_factor, float delta_factor, float *points, size_t ppitch, float *normals, size_t npitch,
                        long long *stats)
{
    orc_raycaster rc;
    rc.vol = vol; rc.aff = cam2vol; memcpy(rc.Rinv, Rinv, sizeof rc.Rinv);
    rc.volume_size = f3(vol.voxel_size[0] * (float)vol.dims[0], vol.voxel_size[1] * (float)vol.dims[1], vol.voxel_size[2] * (float)vol.dims[2]);
    rc.time_step = vol.trunc_dist * step_factor;
    rc.gradient_delta = f3(vol.voxel_size[0] * delta_factor, vol.voxel_size[1] * delta_factor, vol.voxel_size[2] * delta_factor);
    rc.voxel_size_inv = f3(1.f / vol.voxel_size[0], 1.f / vol.voxel_size[1], 1.f / vol.voxel_size[2]);
    rc.finvx = 1.f / intr.fx; rc.finvy = 1.f / intr.fy;
    const float qnan = orc_qnan();
    long long hits = 0, steps = 0, entered = 0;
    const orc_f3 vsz = f3(vol.voxel_size[0], vol.voxel_size[1], vol.voxel_size[2]);

#pragma omp parallel for schedule(dynamic, 4) reduction(+ : hits, steps, entered)
    for (int y = 0; y < rows; ++y) {
        float *prow = orc_row_f4w(points, ppitch, y);
        float *nrow = orc_row_f4w(normals, npitch, y);
        for (int x = 0; x < cols; ++x) {
            float *P = prow + 4 * x, *Nn = nrow + 4 * x;
            P[0] = P[1] = P[2] = P[3] = qnan;
            Nn[0] = Nn[1] = Nn[2] = Nn[3] = qnan;

            orc_f3 ray_org = f3(rc.aff.t[0], rc.aff.t[1], rc.aff.t[2]);
            /* reproj(x, y, 1.f): z * (u - c.x) * finv.x, left to right */
            orc_f3 rp = f3(1.f * ((float)x - intr.cx) * rc.finvx, 1.f * ((float)y - intr.cy) * rc.finvy, 1.f);
            orc_f3 ray_dir = orc_normalized(orc_mat_mul(rc.aff.R, rp));
            orc_f3 box_max = orc_sub(rc.volume_size, vsz);

            /* intersect(), :202-218 */
            orc_f3 invR = f3(1.f / ray_dir.x, 1.f / ray_dir.y, 1.f / ray_dir.z);
            orc_f3 tbot = orc_mul(invR, orc_sub(f3(0.f, 0.f, 0.f), ray_org));
            orc_f3 ttop = orc_mul(invR, orc_sub(box_max, ray_org));
            orc_f3 tmn = f3(fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z));
            orc_f3 tmx = f3(fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z));
            float tmin = fmaxf(fmaxf(tmn.x, tmn.y), fmaxf(tmn.x, tmn.z));
            float tmax = fminf(fminf(tmx.x, tmx.y), fminf(tmx.x, tmx.z));

            tmin = fmaxf(0.f, tmin);
            if (tmin >= tmax) continue;
            ++entered;

            tmax -= rc.time_step;
            orc_f3 vstep = orc_scale(ray_dir, rc.time_step);
            orc_f3 next = orc_add(ray_org, orc_scale(ray_dir, tmin));

            float tsdf_next = orc_fetch_tsdf(&rc, next);
            for (float tcurr = tmin; tcurr < tmax; tcurr += rc.time_step) {
                float tsdf_curr = tsdf_next;
                orc_f3 curr = next;
                next = orc_add(next, vstep);
                tsdf_next = orc_fetch_tsdf(&rc, next);
                ++steps;
                if (tsdf_curr < 0.f && tsdf_next > 0.f) break;
                if (tsdf_curr > 0.f && tsdf_next < 0.f) {
                    float Ft = orc_interpolate(&vol, orc_mul(curr, rc.voxel_size_inv));
                    float Ftdt = orc_interpolate(&vol, orc_mul(next, rc.voxel_size_inv));
                    float Ts = tcurr - (rc.time_step * Ft) / (Ftdt - Ft);
                    orc_f3 vertex = orc_add(ray_org, orc_scale(ray_dir, Ts));
                    orc_f3 normal = orc_compute_normal(&vol, vertex, rc.gradient_delta, rc.voxel_size_inv);
                    if (!isnan(normal.x * normal.y * normal.z)) {
                        normal = orc_mat_mul(rc.Rinv, normal);
                        vertex = orc_mat_mul(rc.Rinv, orc_sub(vertex, ray_org));
                        Nn[0] = normal.x; Nn[1] = normal.y; Nn[2] = normal.z; Nn[3] = 0.f;
                        P[0] = vertex.x; P[1] = vertex.y; P[2] = vertex.z; P[3] = 0.f;
                        ++hits;
                    }
                    break;
                }
            }
        }
    }
    if (stats) { stats[0] = hits; stats[1] = steps; stats[2] = entered; }
}

/* FullScan6::operator(), tsdf_volume.cu:511-710.  The reference emits points in a nondeterministic order
 * (global atomicAdd cursor); the restatement emits them in ascending (z, y, x) voxel order, +x, +y, +z edge
 * within a voxel -- parity with the reference is on the SET of points.  Returns min(count, capacity)
 * (output_count, :703) and stops writing when the buffer is full. */
long long orc_extract_cloud(orc_volume vol, orc_aff3f pose, float *out, long long capacity)
{
    const int Dx = vol.dims[0], Dy = vol.dims[1], Dz = vol.dims[2];
    const size_t sy = (size_t)Dx, sz = (size_t)Dx * Dy;
    long long count = 0;
    for (int z = 0; z < Dz - 1; ++z)
        for (int y = 0; y < Dy; ++y)
            for (int x = 0; x < Dx; ++x) {
                int W;
                float F = orc_unpack_tsdf(vol.data[x + sy * y + sz * z], &W);
                if (W == 0 || F == 1.f) continue;
                orc_f3 V = f3(((float)x + 0.5f) * vol.voxel_size[0], ((float)y + 0.5f) * vol.voxel_size[1], ((float)z + 0.5f) * vol.voxel_size[2]);
                for (int axis = 0; axis < 3; ++axis) {
                    int nx = x + (axis == 0), ny = y + (axis == 1), nz = z + (axis == 2);
                    if (nx >= Dx || ny >= Dy) continue;      /* z+1 < Dz guaranteed by the loop */
                    int Wn;
                    float Fn = orc_unpack_tsdf(vol.data[nx + sy * ny + sz * nz], &Wn);
                    if (Wn == 0 || Fn == 1.f) continue;
                    if (!((F > 0 && Fn < 0) || (F < 0 && Fn > 0))) continue;
                    orc_f3 p = V;
                    float d_inv = 1.f / (fabsf(F) + fabsf(Fn));
                    if (axis == 0) { float Vn = V.x + vol.voxel_size[0]; p.x = (V.x * fabsf(Fn) + Vn * fabsf(F)) * d_inv; }
                    if (axis == 1) { float Vn = V.y + vol.voxel_size[1]; p.y = (V.y * fabsf(Fn) + Vn * fabsf(F)) * d_inv; }
                    if (axis == 2) { float Vn = V.z + vol.voxel_size[2]; p.z = (V.z * fabsf(Fn) + Vn * fabsf(F)) * d_inv; }
                    orc_f3 q = orc_aff_mul(&pose, p);
                    if (count < capacity) { float *o = out + 4 * count; o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = 0.f; }
                    ++count;
                }
            }
    return count < capacity ? count : capacity;
}

/* ExtractNormals::operator(), tsdf_volume.cu:714-795 */
void orc_extract_normals(orc_volume vol, const float *pts, long long n, orc_aff3f pose, const float *Rinv, float delta_factor, float *out)
{
    const orc_f3 vinv = f3(1.f / vol.voxel_size[0], 1.f / vol.voxel_size[1], 1.f / vol.voxel_size[2]);
    const orc_f3 gd = f3(vol.voxel_size[0] * delta_factor, vol.voxel_size[1] * delta_factor, vol.voxel_size[2] * delta_factor);
    const float qnan = orc_qnan();
    const orc_f3 t = f3(pose.t[0], pose.t[1], pose.t[2]);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        orc_f3 nrm = f3(qnan, qnan, qnan);
        orc_f3 point = orc_mat_mul(Rinv, orc_sub(f3(pts[4 * i], pts[4 * i + 1], pts[4 * i + 2]), t));
        int gx = (int)rintf(point.x * vinv.x), gy = (int)rintf(point.y * vinv.y), gz = (int)rintf(point.z * vinv.z);
        if (gx > 1 && gy > 1 && gz > 1 && gx < vol.dims[0] - 2 && gy < vol.dims[1] - 2 && gz < vol.dims[2] - 2) {
            orc_f3 tt;
            tt = point; tt.x += gd.x; float Fx1 = orc_interpolate(&vol, orc_mul(tt, vinv));
            tt = point; tt.x -= gd.x; float Fx2 = orc_interpolate(&vol, orc_mul(tt, vinv));
            nrm.x = (Fx1 - Fx2) / gd.x;
            tt = point; tt.y += gd.y; float Fy1 = orc_interpolate(&vol, orc_mul(tt, vinv));
            tt = point; tt.y -= gd.y; float Fy2 = orc_interpolate(&vol, orc_mul(tt, vinv));
            nrm.y = (Fy1 - Fy2) / gd.y;
            tt = point; tt.z += gd.z; float Fz1 = orc_interpolate(&vol, orc_mul(tt, vinv));
            tt = point; tt.z -= gd.z; float Fz2 = orc_interpolate(&vol, orc_mul(tt, vinv));
            nrm.z = (Fz1 - Fz2) / gd.z;
            nrm = orc_normalized(orc_mat_mul(pose.R, nrm));
        }
        out[4 * i] = nrm.x; out[4 * i + 1] = nrm.y; out[4 * i + 2] = nrm.z; out[4 * i + 3] = 0.f;
    }
}
