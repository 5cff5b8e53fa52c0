/* CPU ORACLE (test infrastructure only) -- the per-frame loop.
 * Restates KinFu::operator() / KinFu::dynamicfusion (kfusion/src/kinfu.cpp:221-305,344-400) and
 * TsdfVolume::surface_fusion (kfusion/src/tsdf_volume.cpp:228-255) stage by stage on the CPU, with the same
 * deliberate choices as the product (DESIGN.md "Divergences"): filled nodes only, NaN rows skipped in the solve,
 * deterministic extraction order, project_and_remove sampling the original image, GUI call dropped.
 * Used by tests (short sequences, small volumes) and by bench.py as the CPU baseline. */
#include "orc_common.h"
#include "../include/df_hostmath.h"
#include <stdlib.h>
#include <stdio.h>

void orc_knn8_fast(const float *nodes, int M, const float *queries, long long N, int qstride, int32_t *idx, float *d2);

#define ORC_LEVELS_MAX 4

typedef struct {
    int cols, rows;
    orc_intr intr;
    int volume_dims[3];
    float volume_size[3];
    orc_aff3f volume_pose;
    float bilateral_sigma_depth, bilateral_sigma_spatial;
    int bilateral_kernel_size;
    float icp_truncate_depth_dist, icp_dist_thres, icp_angle_thres;
    int icp_iter_num[4];
    float tsdf_min_camera_movement, tsdf_trunc_dist;
    int tsdf_max_weight;
    float raycast_step_factor, gradient_delta_factor;
    float light_pose[3];
    int solver_nonlinear_iters, solver_linear_iters;
    int max_nodes, node_step, cloud_capacity, flags;
    float fusion_weight_scale;
    float extend_radius;
} orc_kinfu_params;     /* identical layout to df_kinfu_params (include/dfusion.h) */

typedef struct {
    orc_kinfu_params p;
    int levels;
    float trunc_dist, voxel_size[3];
    uint32_t *volume;
    uint16_t *dists, *cur_depth[ORC_LEVELS_MAX];
    float *cur_pts[ORC_LEVELS_MAX], *cur_nrm[ORC_LEVELS_MAX], *prev_pts[ORC_LEVELS_MAX], *prev_nrm[ORC_LEVELS_MAX];
    int lcols[ORC_LEVELS_MAX], lrows[ORC_LEVELS_MAX];
    float *canon, *canon_nrm, *canon_visible;
    float *cloud, *cloud_nrm; long long cloud_count;
    float *nodes; int M;
    float *poses; int nposes, poses_cap;
    int frame_counter, resets, last_ok;
    double solve_stats[8];
    double stage_s[10];
} orc_kinfu;

static double now_s(void);
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static orc_volume vol_of(orc_kinfu *k)
{
    orc_volume v;
    v.data = k->volume;
    for (int i = 0; i < 3; ++i) { v.dims[i] = k->p.volume_dims[i]; v.voxel_size[i] = k->voxel_size[i]; }
    v.trunc_dist = k->trunc_dist; v.max_weight = k->p.tsdf_max_weight;
    return v;
}
static orc_aff3f to_aff(const float *a) { orc_aff3f r; memcpy(r.R, a, 36); memcpy(r.t, a + 9, 12); return r; }

static void do_reset(orc_kinfu *k)
{
    if (k->frame_counter) ++k->resets;
    k->frame_counter = 0;
    k->nposes = 1;
    dfh_aff_identity(k->poses);
    orc_clear_volume(vol_of(k));
}

orc_kinfu *orc_kinfu_create(const orc_kinfu_params *pp)
{
    orc_kinfu *k = (orc_kinfu *)calloc(1, sizeof *k);
    k->p = *pp;
    int i = ORC_LEVELS_MAX - 1;
    for (; i >= 0 && !pp->icp_iter_num[i]; --i) {}
    k->levels = i + 1;
    float vmax = 0.f;
    for (int d = 0; d < 3; ++d) { k->voxel_size[d] = pp->volume_size[d] / pp->volume_dims[d]; if (k->voxel_size[d] > vmax) vmax = k->voxel_size[d]; }
    k->trunc_dist = pp->tsdf_trunc_dist > 2.1f * vmax ? pp->tsdf_trunc_dist : 2.1f * vmax;
    size_t nvox = (size_t)pp->volume_dims[0] * pp->volume_dims[1] * pp->volume_dims[2];
    k->volume = (uint32_t *)malloc(nvox * 4);
    size_t npix = (size_t)pp->cols * pp->rows;
    k->dists = (uint16_t *)malloc(npix * 2);
    int cols = pp->cols, rows = pp->rows;
    for (int l = 0; l < ORC_LEVELS_MAX; ++l) {
        size_t n = (size_t)cols * rows;
        k->lcols[l] = cols; k->lrows[l] = rows;
        k->cur_depth[l] = (uint16_t *)calloc(n ? n : 1, 2);
        k->cur_pts[l] = (float *)calloc(n ? n : 1, 16); k->cur_nrm[l] = (float *)calloc(n ? n : 1, 16);
        k->prev_pts[l] = (float *)calloc(n ? n : 1, 16); k->prev_nrm[l] = (float *)calloc(n ? n : 1, 16);
        cols /= 2; rows /= 2;
    }
    k->canon = (float *)malloc(npix * 16); k->canon_nrm = (float *)malloc(npix * 16); k->canon_visible = (float *)malloc(npix * 16);
    k->cloud = (float *)malloc((size_t)pp->cloud_capacity * 16); k->cloud_nrm = (float *)malloc((size_t)pp->cloud_capacity * 16);
    int maxM = pp->max_nodes > 0 ? pp->max_nodes : (pp->cloud_capacity + 49) / 50;
    k->nodes = (float *)calloc((size_t)maxM * ORC_NODE_STRIDE, 4);
    k->poses_cap = 4096; k->poses = (float *)malloc((size_t)k->poses_cap * 48);
    do_reset(k);
    return k;
}

void orc_kinfu_destroy(orc_kinfu *k)
{
    if (!k) return;
    free(k->volume); free(k->dists);
    for (int l = 0; l < ORC_LEVELS_MAX; ++l) { free(k->cur_depth[l]); free(k->cur_pts[l]); free(k->cur_nrm[l]); free(k->prev_pts[l]); free(k->prev_nrm[l]); }
    free(k->canon); free(k->canon_nrm); free(k->canon_visible); free(k->cloud); free(k->cloud_nrm); free(k->nodes); free(k->poses);
    free(k);
}

/* batch WarpField::warp (warp_field.cpp:180-195) with the grid k-NN; same arithmetic as orc_warp(flags = 0) */
static void warp_all(orc_kinfu *k, float *points, float *normals, long long N)
{
    int32_t *idx = (int32_t *)malloc((size_t)N * 8 * 4);
    float *d2 = (float *)malloc((size_t)N * 8 * 4);
    orc_knn8_fast(k->nodes, k->M, points, N, 4, idx, d2);
    orc_aff3f ident; float id12[12]; dfh_aff_identity(id12); ident = to_aff(id12);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < N; ++q) {
        float *pt = points + 4 * q, *nr = normals + 4 * q;
        if (isnan(pt[0]) || isnan(nr[0])) continue;
        float rot4[4], trans4[4];
        orc_dqb(k->nodes, idx + q * 8, d2 + q * 8, rot4, trans4, NULL);
        /* dq.transform(point); point = warp_to_live * point; same for the normal (translation included, as the reference) */
        float node[ORC_NODE_STRIDE] = {0}, t4[4];
        for (int c = 0; c < 4; ++c) { node[3 + c] = rot4[c]; node[7 + c] = trans4[c]; }
        orc_node_translation(node, t4);
        orc_quat_rotate_vec(rot4, pt); pt[0] += t4[1]; pt[1] += t4[2]; pt[2] += t4[3];
        orc_quat_rotate_vec(rot4, nr); nr[0] += t4[1]; nr[1] += t4[2]; nr[2] += t4[3];
        for (int pass = 0; pass < 2; ++pass) {
            float *v = pass ? nr : pt;
            float x = v[0], y = v[1], z = v[2];
            v[0] = ident.R[0] * x + ident.R[1] * y + ident.R[2] * z + ident.t[0];
            v[1] = ident.R[3] * x + ident.R[4] * y + ident.R[5] * z + ident.t[1];
            v[2] = ident.R[6] * x + ident.R[7] * y + ident.R[8] * z + ident.t[2];
        }
    }
    free(idx); free(d2);
}

int orc_solve_data_term_big(float *nodes, int M, const float *canon, const float *live, long long N, int stride, int flags,
                            int max_lm, int lin_iters, double *stats);

static void integrate_with(orc_kinfu *k, const uint16_t *dists, const float *cam_pose)
{
    float vol_pose[12], inv[12], vol2cam[12];
    memcpy(vol_pose, k->p.volume_pose.R, 36); memcpy(vol_pose + 9, k->p.volume_pose.t, 12);
    dfh_aff_inv(cam_pose, inv);
    dfh_aff_mul(inv, vol_pose, vol2cam);
    orc_integrate(vol_of(k), dists, (size_t)k->p.cols * 2, k->p.cols, k->p.rows, to_aff(vol2cam), k->p.intr);
}
static void raycast_to(orc_kinfu *k, const float *cam_pose, float *pts, float *nrm)
{
    float vol_pose[12], inv[12], cam2vol[12], Rinv[9];
    memcpy(vol_pose, k->p.volume_pose.R, 36); memcpy(vol_pose + 9, k->p.volume_pose.t, 12);
    dfh_aff_inv(vol_pose, inv);
    dfh_aff_mul(inv, cam_pose, cam2vol);
    dfh_mat3_inv(cam2vol, Rinv);
    orc_raycast_points(vol_of(k), to_aff(cam2vol), Rinv, k->p.intr, k->p.cols, k->p.rows, k->p.raycast_step_factor,
                       k->p.gradient_delta_factor, pts, (size_t)k->p.cols * 16, nrm, (size_t)k->p.cols * 16, NULL);
}
static void extract(orc_kinfu *k)
{
    float vol_pose[12], Rinv[9];
    memcpy(vol_pose, k->p.volume_pose.R, 36); memcpy(vol_pose + 9, k->p.volume_pose.t, 12);
    dfh_mat3_inv(vol_pose, Rinv);
    k->cloud_count = orc_extract_cloud(vol_of(k), k->p.volume_pose, k->cloud, k->p.cloud_capacity);
    orc_extract_normals(vol_of(k), k->cloud, k->cloud_count, k->p.volume_pose, Rinv, k->p.gradient_delta_factor, k->cloud_nrm);
}

/* returns 1 = fused + image, 0 = first frame / reset */
int orc_kinfu_process(orc_kinfu *k, const uint16_t *depth, size_t pitch)
{
    const orc_kinfu_params *p = &k->p;
    const int L = k->levels;
    const size_t p2 = (size_t)p->cols * 2;
    double t0 = now_s();
    memset(k->stage_s, 0, sizeof k->stage_s);
    orc_compute_dists(depth, pitch, p->cols, p->rows, p->intr, k->dists, p2);
    orc_bilateral(depth, pitch, p->cols, p->rows, k->cur_depth[0], p2, p->bilateral_kernel_size, p->bilateral_sigma_spatial, p->bilateral_sigma_depth);
    if (p->icp_truncate_depth_dist > 0) orc_truncate_depth(k->cur_depth[0], p2, p->cols, p->rows, p->icp_truncate_depth_dist);
    for (int i = 1; i < L; ++i)
        orc_pyr_down(k->cur_depth[i - 1], (size_t)k->lcols[i - 1] * 2, k->lcols[i - 1], k->lrows[i - 1], k->cur_depth[i], (size_t)k->lcols[i] * 2, p->bilateral_sigma_depth);
    for (int i = 0; i < L; ++i) {
        int div = 1 << i;
        orc_intr li = {p->intr.fx / div, p->intr.fy / div, p->intr.cx / div, p->intr.cy / div};
        orc_points_normals(li, k->cur_depth[i], (size_t)k->lcols[i] * 2, k->lcols[i], k->lrows[i], k->cur_pts[i], (size_t)k->lcols[i] * 16,
                           k->cur_nrm[i], (size_t)k->lcols[i] * 16);
    }
    double t1 = now_s(); k->stage_s[0] = t1 - t0;

    if (k->frame_counter == 0) {
        integrate_with(k, k->dists, k->poses + (size_t)(k->nposes - 1) * 12);
        extract(k);
        if (!(p->flags & 1)) {
            long long count = k->cloud_count;
            int step = p->node_step > 0 ? p->node_step : 50;
            int M = (int)((count + step - 1) / step);
            if (p->max_nodes > 0 && M > p->max_nodes) { step = (int)((count + p->max_nodes - 1) / p->max_nodes); M = (int)((count + step - 1) / step); }
            k->M = M;
            for (int m = 0; m < M; ++m) {
                float *n = k->nodes + (size_t)m * ORC_NODE_STRIDE;
                const float *c = k->cloud + (size_t)m * step * 4;
                n[0] = c[0]; n[1] = c[1]; n[2] = c[2];
                n[3] = 1.f; n[4] = n[5] = n[6] = 0.f; n[7] = 1.f; n[8] = n[9] = n[10] = 0.f; n[11] = 3.f;
            }
        }
        for (int i = 0; i < ORC_LEVELS_MAX; ++i) {
            float *t = k->cur_pts[i]; k->cur_pts[i] = k->prev_pts[i]; k->prev_pts[i] = t;
            t = k->cur_nrm[i]; k->cur_nrm[i] = k->prev_nrm[i]; k->prev_nrm[i] = t;
        }
        ++k->frame_counter;
        return 0;
    }

    {
        const float *vc[ORC_LEVELS_MAX], *nc[ORC_LEVELS_MAX], *vp[ORC_LEVELS_MAX], *np[ORC_LEVELS_MAX];
        int cols[ORC_LEVELS_MAX], rows[ORC_LEVELS_MAX]; size_t pit[ORC_LEVELS_MAX];
        for (int i = 0; i < L; ++i) { vc[i] = k->cur_pts[i]; nc[i] = k->cur_nrm[i]; vp[i] = k->prev_pts[i]; np[i] = k->prev_nrm[i];
                                      cols[i] = k->lcols[i]; rows[i] = k->lrows[i]; pit[i] = (size_t)k->lcols[i] * 16; }
        orc_aff3f T;
        int ok = orc_icp_estimate(vc, nc, vp, np, cols, rows, pit, L, p->icp_iter_num, p->intr, p->icp_dist_thres, p->icp_angle_thres, &T);
        k->last_ok = ok;
        if (!ok) { do_reset(k); return 0; }
        float T12[12]; memcpy(T12, T.R, 36); memcpy(T12 + 9, T.t, 12);
        if (k->nposes == k->poses_cap) { k->poses_cap *= 2; k->poses = (float *)realloc(k->poses, (size_t)k->poses_cap * 48); }
        dfh_aff_mul(k->poses + (size_t)(k->nposes - 1) * 12, T12, k->poses + (size_t)k->nposes * 12);
        ++k->nposes;
    }
    double t2 = now_s(); k->stage_s[1] = t2 - t1;
    const float *cam_pose = k->poses + (size_t)(k->nposes - 1) * 12;
    const long long npix = (long long)p->cols * p->rows;

    if (!(p->flags & 1) && k->M >= 8) {
        raycast_to(k, cam_pose, k->canon_visible, k->canon_nrm);
        float inv_pose[12];
        dfh_aff_inv(cam_pose, inv_pose);
        for (long long i = 0; i < npix; ++i) {
            const float *v = k->canon_visible + 4 * i;
            float x = v[0], y = v[1], z = v[2], w = v[3];
            float o[4];
            o[0] = inv_pose[0] * x + inv_pose[1] * y + inv_pose[2] * z + inv_pose[9];
            o[1] = inv_pose[3] * x + inv_pose[4] * y + inv_pose[5] * z + inv_pose[10];
            o[2] = inv_pose[6] * x + inv_pose[7] * y + inv_pose[8] * z + inv_pose[11];
            o[3] = w;
            memcpy(k->canon + 4 * i, o, 16); memcpy(k->canon_visible + 4 * i, o, 16);
        }
        double t3 = now_s(); k->stage_s[2] = t3 - t2;
        warp_all(k, k->canon, k->canon_nrm, npix);
        double t4 = now_s(); k->stage_s[3] = t4 - t3;
        orc_solve_data_term_big(k->nodes, k->M, k->canon, k->cur_pts[0], npix, 4, (p->flags & 4) ? 1 : 0, p->solver_nonlinear_iters,
                                p->solver_linear_iters, k->solve_stats);
        double t5 = now_s(); k->stage_s[4] = t5 - t4;
        warp_all(k, k->canon, k->canon_nrm, npix);
        double t6 = now_s(); k->stage_s[5] = t6 - t5;
        double t7, t8;
        if (p->flags & 8) {          /* DF_KINFU_WARPED_INTEGRATE: per-voxel warped fusion (orc_fusion.c), no pixel removal, no rigid integrate */
            t7 = now_s(); k->stage_s[6] = t7 - t6;
            orc_integrate_warped(vol_of(k), k->cur_depth[0], p2, p->cols, p->rows, p->volume_pose, to_aff(inv_pose), p->intr, k->nodes, k->M,
                                 p->fusion_weight_scale);
            t8 = now_s(); k->stage_s[7] = t8 - t7;
        } else {
        orc_project_and_remove(k->cur_depth[0], p2, p->cols, p->rows, p->intr, k->canon, (size_t)p->cols * 16, p->cols, p->rows);
        t7 = now_s(); k->stage_s[6] = t7 - t6;
        orc_compute_dists(k->cur_depth[0], p2, p->cols, p->rows, p->intr, k->dists, p2);
        integrate_with(k, k->dists, cam_pose);
        t8 = now_s(); k->stage_s[7] = t8 - t7;
        }
        extract(k);
        if ((p->flags & 16) && k->M > 0) {     /* DF_KINFU_EXTEND_FIELD: grow the field over unsupported canonical surface (orc_fusion.c) */
            const int maxM = p->max_nodes > 0 ? p->max_nodes : (p->cloud_capacity + 49) / 50;
            k->M = orc_extend_field(k->nodes, k->M, maxM, k->cloud, k->cloud_count, 4, p->extend_radius > 0 ? p->extend_radius : 0.03f,
                                    p->node_step > 0 ? p->node_step : 50);
        }
        double t9 = now_s(); k->stage_s[8] = t9 - t8;
        t2 = t9;
    } else {
        integrate_with(k, k->dists, cam_pose);
        double t8 = now_s(); k->stage_s[7] = t8 - t2; t2 = t8;
    }
    raycast_to(k, cam_pose, k->prev_pts[0], k->prev_nrm[0]);
    for (int i = 1; i < L; ++i)
        orc_resize_points_normals(k->prev_pts[i - 1], (size_t)k->lcols[i - 1] * 16, k->prev_nrm[i - 1], (size_t)k->lcols[i - 1] * 16, k->lcols[i - 1],
                                  k->lrows[i - 1], k->prev_pts[i], (size_t)k->lcols[i] * 16, k->prev_nrm[i], (size_t)k->lcols[i] * 16);
    k->stage_s[9] = now_s() - t2;
    ++k->frame_counter;
    return 1;
}

/* accessors for the ctypes binding */
void *orc_kinfu_buffer(orc_kinfu *k, int which)
{
    switch (which) {
        case 0: return k->volume; case 1: return k->dists; case 2: return k->cur_depth[0]; case 3: return k->cur_pts[0];
        case 4: return k->cur_nrm[0]; case 5: return k->prev_pts[0]; case 6: return k->prev_nrm[0]; case 7: return k->canon;
        case 8: return k->canon_nrm; case 9: return k->cloud; case 10: return k->cloud_nrm; case 11: return k->nodes;
        case 12: return k->canon_visible; case 13: return k->solve_stats; case 14: return k->poses; case 15: return k->stage_s;
    }
    return NULL;
}
void orc_kinfu_info(orc_kinfu *k, long long *info)
{
    info[0] = k->frame_counter; info[1] = k->M; info[2] = k->cloud_count; info[3] = k->nposes; info[4] = k->last_ok; info[5] = 0;
    info[6] = k->resets; info[7] = (long long)k->solve_stats[2];
}
