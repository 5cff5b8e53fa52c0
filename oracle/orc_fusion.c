/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Per-voxel warped integration, SURVEY.md 8f(1).  PARITY UNPINNED BY THE REFERENCE: the reference never finished this
 * step -- TsdfVolume::surface_fusion (kfusion/src/tsdf_volume.cpp:228-254) computes psdf() for the warped ray-cast
 * points, runs the rigid integrate and leaves the per-entry update commented out (:248-251); TsdfVolume::Entry /
 * tsdf_entries_ (tsdf_volume.hpp:93-99) is an unused placeholder.  What is restated here is the update that code was
 * written towards (DynamicFusion, Newcombe et al. 2015, eq. 4-5), built ONLY from operations the reference does
 * define, each in the reference's operation order:
 *
 *   x_c   = pose_vol * (x*vs, y*vs, z*vs)                 voxel position as in TsdfIntegrator (tsdf_volume.cu:62-64: no
 *                                                         half-voxel offset), device Aff3f*float3 (device.hpp:71-74)
 *   N(x_c)= 8 nearest nodes, squared distances            WarpField::KNN (warp_field.cpp:247-251)
 *   x_w   = DQB(N).transform(x_c)                         WarpField::DQB + DualQuaternion::transform (:203-217; dual_quaternion.hpp:204-210)
 *   x_t   = world2cam * x_w                               cv::Affine3f * Vec3f, as the last step of WarpField::warp (:180-195)
 *   (u,v) = Projector(x_t)                                device.hpp:32-38, gates of project_kernel (tsdf_volume.cu:114-137)
 *   rho   = depth(floor v, floor u) * 0.001 - x_t.z       TsdfVolume::psdf (tsdf_volume.cpp:266-292): (K^-1 (u*Dp, v*Dp, Dp)).z - warped.z.
 *                                                         The reference hands psdf the u16 MILLIMETRE image where its kernel reads
 *                                                         half floats (kinfu.cpp:390) -- a units slip whose result it discards;
 *                                                         here the depth is taken in metres.
 *   update iff rho > -trunc                               surface_fusion :242 (`ro[i] > -trunc_dist_`)
 *   tsdf  = min(1, rho / trunc)                           `coeff = min(ro, trunc)` (:246), stored normalised like the rigid rule (tsdf_volume.cu:93)
 *   w(x)  = (sum_i sqrt(d_i^2)) / 8                       TsdfVolume::weighting (tsdf_volume.cpp:300-306)
 *   F'    = (F*W + tsdf*w) / (W + w),  W' = min(W + w, max_weight)      the commented lines :248-251 (the second one read as the
 *                                                         division it stands for) in the arithmetic of tsdf_volume.cu:97-103
 *
 * The volume keeps the reference's ushort2 voxel (f16 tsdf, u16 weight), so the sample weight is quantised:
 *   w_q = clamp(rint(w(x) * weight_scale), 1, max_weight);   weight_scale <= 0  =>  w_q = 1 (the rigid rule's weight).
 * Returns the number of voxels written.
 */
#include "orc_common.h"
#include <stdlib.h>

void orc_knn8_fast(const float *nodes, int M, const float *queries, long long N, int qstride, int32_t *idx, float *d2);
void orc_dq_transform(const float *rot4, const float *trans4, float *v);

long long orc_integrate_warped(orc_volume vol, const uint16_t *depth, size_t pitch, int cols, int rows, orc_aff3f vol2world,
                               orc_aff3f world2cam, orc_intr intr, const float *nodes, int M, float weight_scale)
{
    const float trunc_inv = 1.f / vol.trunc_dist;
    const int Dx = vol.dims[0], Dy = vol.dims[1], Dz = vol.dims[2];
    const size_t slice = (size_t)Dx * Dy;
    long long n_upd = 0;
    float *q = (float *)malloc(slice * 3 * sizeof(float));
    int32_t *idx = (int32_t *)malloc(slice * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc(slice * 8 * sizeof(float));
    for (int z = 0; z < Dz; ++z) {
        for (int y = 0; y < Dy; ++y)
            for (int x = 0; x < Dx; ++x) {
                const orc_f3 xc = orc_aff_mul(&vol2world, f3((float)x * vol.voxel_size[0], (float)y * vol.voxel_size[1], (float)z * vol.voxel_size[2]));
                float *p = q + ((size_t)y * Dx + x) * 3;
                p[0] = xc.x; p[1] = xc.y; p[2] = xc.z;
            }
        orc_knn8_fast(nodes, M, q, (long long)slice, 3, idx, d2);      /* same 8 (distance, index)-smallest as orc_knn8 */
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_upd)
        for (long long i = 0; i < (long long)slice; ++i) {
            float rot4[4], trans4[4];
            orc_dqb(nodes, idx + i * 8, d2 + i * 8, rot4, trans4, NULL);
            float p[3] = {q[i * 3], q[i * 3 + 1], q[i * 3 + 2]};
            orc_dq_transform(rot4, trans4, p);
            /* cv::Affine3f * Vec3f, left to right (warp_field.cpp:191) */
            const float tx = world2cam.R[0] * p[0] + world2cam.R[1] * p[1] + world2cam.R[2] * p[2] + world2cam.t[0];
            const float ty = world2cam.R[3] * p[0] + world2cam.R[4] * p[1] + world2cam.R[5] * p[2] + world2cam.t[1];
            const float tz = world2cam.R[6] * p[0] + world2cam.R[7] * p[1] + world2cam.R[8] * p[2] + world2cam.t[2];
            if (!(tz > 0)) continue;
            const float u = fmaf(intr.fx, tx / tz, intr.cx), v = fmaf(intr.fy, ty / tz, intr.cy);
            if (!(u >= 0 && v >= 0 && u < (float)cols && v < (float)rows)) continue;
            const uint16_t mm = orc_row_u16(depth, pitch, (int)v)[(int)u];
            if (mm == 0) continue;
            const float rho = (float)mm * 0.001f - tz;
            if (!(rho > -vol.trunc_dist)) continue;
            const float tsdf = fminf(1.f, rho * trunc_inv);
            int wq = 1;
            if (weight_scale > 0) {
                float sum = 0.f;
                for (int k = 0; k < 8; ++k) if (idx[i * 8 + k] >= 0) sum += sqrtf(d2[i * 8 + k]);
                const float w = sum / 8;
                const float s = rintf(w * weight_scale);
                wq = s < 1.f ? 1 : (s > (float)vol.max_weight ? vol.max_weight : (int)s);
            }
            uint32_t *vptr = vol.data + (size_t)z * slice + (size_t)i;
            int weight_prev;
            const float tsdf_prev = orc_unpack_tsdf(*vptr, &weight_prev);
            const float tsdf_new = fmaf(tsdf_prev, (float)weight_prev, tsdf * (float)wq) / (float)(weight_prev + wq);
            const int weight_new = weight_prev + wq < vol.max_weight ? weight_prev + wq : vol.max_weight;
            *vptr = orc_pack_tsdf(tsdf_new, weight_new);
            ++n_upd;
        }
    }
    free(q); free(idx); free(d2);
    return n_upd;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Extending the warp field, SURVEY.md 8f(3).  PARITY UNPINNED BY THE REFERENCE: Report.md ("4. Extending the warp field - stubbed out
 * functionality") describes the step, the code base has none.  Restated from the pieces the reference does define:
 *   - a point of the extracted canonical cloud is UNSUPPORTED when its nearest node (WarpField::KNN, warp_field.cpp:247-251: squared
 *     distance d0*d0 + d1*d1 + d2*d2 in float) is farther than `radius` (d^2 > radius*radius); NaN points are skipped;
 *   - the unsupported points are subsampled as WarpField::init subsamples the first cloud (every step-th, starting with the first:
 *     warp_field.cpp:49-60), in cloud order;
 *   - each becomes a node as init makes them (:68-80): vertex = the point, identity DualQuaternion() (rotation (1,0,0,0), dual (1,0,0,0)),
 *     weight 3; appended after the M existing nodes until max_nodes.
 * Support is judged against the M nodes present on entry only.  Returns the new node count. */
int orc_extend_field(float *nodes, int M, int max_nodes, const float *cloud, long long n_points, int stride, float radius, int step)
{
    if (M <= 0 || n_points <= 0 || step <= 0) return M;
    int32_t *idx = (int32_t *)malloc((size_t)n_points * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)n_points * 8 * sizeof(float));
    orc_knn8_fast(nodes, M, cloud, n_points, stride, idx, d2);
    const float r2 = radius * radius;
    long long rank = 0;
    int Mn = M;
    for (long long i = 0; i < n_points; ++i) {
        const float *p = cloud + (size_t)i * stride;
        if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) continue;
        if (!(d2[i * 8] > r2)) continue;
        if (rank % step == 0 && M + rank / step < max_nodes) {
            float *n = nodes + (size_t)(M + rank / step) * ORC_NODE_STRIDE;
            memset(n, 0, ORC_NODE_STRIDE * sizeof(float));
            n[0] = p[0]; n[1] = p[1]; n[2] = p[2];
            n[3] = 1.f; n[7] = 1.f; n[11] = 3.f;
            Mn = (int)(M + rank / step) + 1;
        }
        ++rank;
    }
    free(idx); free(d2);
    return Mn;
}
