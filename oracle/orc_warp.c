/* CPU ORACLE (test infrastructure only) -- warp field: k-NN, node weights, dual-quaternion blend, warp.
 * Restates kfusion/src/warp_field.cpp, kfusion/src/utils/{quaternion,dual_quaternion,knn_point_cloud}.hpp
 * and the result-set semantics of the vendored nanoflann (kfusion/include/nanoflann/nanoflann.hpp:78-137).
 * Pinned by the reference's golden vectors: tests/utils/test_quaternion.cc, test_dual_quaternion.cc,
 * tests/nanoflann_test.cpp (orders regenerated from the vendored header by oracle/_ref/knn_ref).
 *
 * Quaternions are float[4] = (w, x, y, z).  A node is ORC_NODE_STRIDE floats:
 *   [0..2] vertex, [3..6] rotation_, [7..10] translation_ (the dual part, 0.5*t*r), [11] weight
 * (deformation_node, kfusion/include/kfusion/warp_field.hpp:35-40). */
#include "orc_common.h"

/* Quaternion::operator*, quaternion.hpp:191-199 */
void orc_quat_mul(const float *a, const float *b, float *out)
{
    float w = ((a[0] * b[0]) - (a[1] * b[1]) - (a[2] * b[2]) - (a[3] * b[3]));
    float x = ((a[0] * b[1]) + (a[1] * b[0]) + (a[2] * b[3]) - (a[3] * b[2]));
    float y = ((a[0] * b[2]) - (a[1] * b[3]) + (a[2] * b[0]) + (a[3] * b[1]));
    float z = ((a[0] * b[3]) + (a[1] * b[2]) - (a[2] * b[1]) + (a[3] * b[0]));
    out[0] = w; out[1] = x; out[2] = y; out[3] = z;
}

/* Quaternion::norm / normalize, quaternion.hpp:215-228: (1.0/theNorm) is a DOUBLE scalar, each component is
 * multiplied in double and narrowed to float by the constructor. */
static inline float quat_norm(const float *q) { return sqrtf((q[0] * q[0]) + (q[1] * q[1]) + (q[2] * q[2]) + (q[3] * q[3])); }
static inline void quat_normalize(float *q)
{
    float n = quat_norm(q);
    double s = 1.0 / (double)n;
    for (int i = 0; i < 4; ++i) q[i] = (float)(s * (double)q[i]);
}

/* Quaternion::encodeRotation, quaternion.hpp:74-82 (sin/cos of a float argument -> float overloads) */
void orc_quat_encode_rotation(float theta, float x, float y, float z, float *q)
{
    float sin_half = sinf(theta / 2);
    q[0] = cosf(theta / 2); q[1] = x * sin_half; q[2] = y * sin_half; q[3] = z * sin_half;
    quat_normalize(q);
}

/* Quaternion::rotate(T&,T&,T&), quaternion.hpp:106-115: q * (0,v) * q^*, NOT normalised */
void orc_quat_rotate_sandwich(const float *q, float *v)
{
    float qs[4] = {q[0], -q[1], -q[2], -q[3]};
    float p[4] = {0.f, v[0], v[1], v[2]}, t[4], r[4];
    orc_quat_mul(q, p, t);
    orc_quat_mul(t, qs, r);
    v[0] = r[1]; v[1] = r[2]; v[2] = r[3];
}

/* Quaternion::rotate(Vec3f&), quaternion.hpp:124-130: normalise a copy; v += (2 q_vec) x (q_vec x v + w v) */
void orc_quat_rotate_vec(const float *q, float *v)
{
    float r[4] = {q[0], q[1], q[2], q[3]};
    quat_normalize(r);
    orc_f3 qv = f3(r[1], r[2], r[3]);
    orc_f3 vv = f3(v[0], v[1], v[2]);
    orc_f3 inner = orc_add(orc_cross(qv, vv), orc_scale(vv, r[0]));
    orc_f3 c = orc_cross(orc_scale(qv, 2.f), inner);
    v[0] += c.x; v[1] += c.y; v[2] += c.z;
}

/* DualQuaternion(x,y,z,roll,pitch,yaw), dual_quaternion.hpp:36-51 (float trig; 0.5 is a double scalar) */
void orc_dq_from_euler(float x, float y, float z, float roll, float pitch, float yaw, float *rot, float *dual)
{
    rot[0] = cosf(roll / 2) * cosf(pitch / 2) * cosf(yaw / 2) + sinf(roll / 2) * sinf(pitch / 2) * sinf(yaw / 2);
    rot[1] = sinf(roll / 2) * cosf(pitch / 2) * cosf(yaw / 2) - cosf(roll / 2) * sinf(pitch / 2) * sinf(yaw / 2);
    rot[2] = cosf(roll / 2) * sinf(pitch / 2) * cosf(yaw / 2) + sinf(roll / 2) * cosf(pitch / 2) * sinf(yaw / 2);
    rot[3] = cosf(roll / 2) * cosf(pitch / 2) * sinf(yaw / 2) - sinf(roll / 2) * sinf(pitch / 2) * cosf(yaw / 2);
    float h[4] = {(float)(0.5 * 0.0), (float)(0.5 * (double)x), (float)(0.5 * (double)y), (float)(0.5 * (double)z)};
    orc_quat_mul(h, rot, dual);
}

/* DualQuaternion::getTranslation(), dual_quaternion.hpp:120-125: 2 * translation_ * conj(normalised rotation_) */
void orc_node_translation(const float *node, float *t4)
{
    float rot[4] = {node[3], node[4], node[5], node[6]};
    quat_normalize(rot);
    float conj[4] = {rot[0], -rot[1], -rot[2], -rot[3]};
    float two[4] = {2 * node[7], 2 * node[8], 2 * node[9], 2 * node[10]};
    orc_quat_mul(two, conj, t4);
}

/* DualQuaternion::encodeTranslation, dual_quaternion.hpp:82-85: translation_ = 0.5 * (0,x,y,z) * rotation_ */
void orc_node_encode_translation(float *node, float x, float y, float z)
{
    float h[4] = {(float)(0.5 * 0.0), (float)(0.5 * (double)x), (float)(0.5 * (double)y), (float)(0.5 * (double)z)};
    orc_quat_mul(h, node + 3, node + 7);
}

/* WarpField::KNN, warp_field.cpp:247-251, with nanoflann KNNResultSet::addPoint semantics
 * (nanoflann.hpp:110-131: strict '<' against the current worst, equal distances keep visiting order) and the
 * adaptor's distance d0*d0 + d1*d1 + d2*d2 (knn_point_cloud.hpp:26-32).  Exhaustive scan in ascending node
 * index: exact k-NN; ties resolve to the lower index.  M < 8 leaves idx = -1, d2 = FLT_MAX in the tail. */
void orc_knn8(const float *nodes, int M, const float *queries, long long N, int qstride, int32_t *idx, float *d2)
{
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < N; ++q) {
        const float *p = queries + (size_t)q * qstride;
        int32_t bi[8]; float bd[8];
        for (int i = 0; i < 8; ++i) { bi[i] = -1; bd[i] = 3.402823466e+38f; }
        int count = 0;
        if (!(p[0] != p[0] || p[1] != p[1] || p[2] != p[2]))
        for (int m = 0; m < M; ++m) {
            const float *v = nodes + (size_t)m * ORC_NODE_STRIDE;
            float d0 = p[0] - v[0], d1 = p[1] - v[1], dd2 = p[2] - v[2];
            float dist = d0 * d0 + d1 * d1 + dd2 * dd2;
            if (!(dist < bd[7])) continue;
            int i;
            for (i = count; i > 0; --i) {
                if (bd[i - 1] > dist) { if (i < 8) { bd[i] = bd[i - 1]; bi[i] = bi[i - 1]; } }
                else break;
            }
            if (i < 8) { bd[i] = dist; bi[i] = m; }
            if (count < 8) ++count;
        }
        for (int i = 0; i < 8; ++i) { idx[q * 8 + i] = bi[i]; d2[q * 8 + i] = bd[i]; }
    }
}

/* WarpField::weighting (warp_field.cpp:238-241, double exp of a float argument) +
 * WarpField::DQB (:203-217) + DualQuaternion(translation, rotation) ctor (dual_quaternion.hpp:59-63).
 * Outputs the blended dual quaternion (rot4 = rotation_, trans4 = translation_) and the 8 weights. */
void orc_dqb_weighted(const float *nodes, const int32_t *idx8, const float *w8, float *rot4, float *trans4);

void orc_dqb(const float *nodes, const int32_t *idx8, const float *d2_8, float *rot4, float *trans4, float *weights8)
{
    float w8[8];
    for (int i = 0; i < 8; ++i) {
        w8[i] = 0.f;
        if (idx8[i] >= 0) {
            float nw = nodes[(size_t)idx8[i] * ORC_NODE_STRIDE + 11];
            w8[i] = (float)exp((double)(-d2_8[i] / (2 * nw * nw)));
        }
        if (weights8) weights8[i] = w8[i];
    }
    orc_dqb_weighted(nodes, idx8, w8, rot4, trans4);
}

/* the blend itself (warp_field.cpp:207-216) for given weights; pinned bit for bit against the reference's own
 * Quaternion / DualQuaternion classes by tests/golden/dq_ref.json */
void orc_dqb_weighted(const float *nodes, const int32_t *idx8, const float *w8, float *rot4, float *trans4)
{
    float tsum[4] = {0, 0, 0, 0}, rsum[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
        if (idx8[i] < 0) continue;
        const float *node = nodes + (size_t)idx8[i] * ORC_NODE_STRIDE;
        float w = w8[i];
        float t4[4];
        orc_node_translation(node, t4);
        for (int c = 0; c < 4; ++c) tsum[c] = tsum[c] + w * t4[c];
        for (int c = 0; c < 4; ++c) rsum[c] = rsum[c] + w * node[3 + c];
    }
    quat_normalize(rsum);
    for (int c = 0; c < 4; ++c) rot4[c] = rsum[c];
    float h[4];
    for (int c = 0; c < 4; ++c) h[c] = (float)(0.5 * (double)tsum[c]);
    orc_quat_mul(h, rsum, trans4);
}

/* DualQuaternion::transform, dual_quaternion.hpp:204-210 */
void orc_dq_transform(const float *rot4, const float *trans4, float *v)
{
    float node[ORC_NODE_STRIDE] = {0};
    for (int c = 0; c < 4; ++c) { node[3 + c] = rot4[c]; node[7 + c] = trans4[c]; }
    float t4[4];
    orc_node_translation(node, t4);
    orc_quat_rotate_vec(rot4, v);
    v[0] += t4[1]; v[1] += t4[2]; v[2] += t4[3];
}

/* cv::Affine3f * Vec3f (opencv2/core/affine.hpp): m0*x + m1*y + m2*z + m3, left to right */
static void aff_apply_cv(const orc_aff3f *a, float *v)
{
    float x = v[0], y = v[1], z = v[2];
    v[0] = a->R[0] * x + a->R[1] * y + a->R[2] * z + a->t[0];
    v[1] = a->R[3] * x + a->R[4] * y + a->R[5] * z + a->t[1];
    v[2] = a->R[6] * x + a->R[7] * y + a->R[8] * z + a->t[2];
}

/* WarpField::warp, warp_field.cpp:180-195.
 * flags bit0 (ORC_WARP_REF_NORMAL_INDEX): reproduce the reference's normal cursor, which advances only on
 *   valid points (so normal j is paired with the j-th VALID point).  Default: normal i pairs with point i.
 * flags bit1 (ORC_WARP_NORMAL_ROTATE_ONLY): extension -- normals are rotated only (the reference also adds the
 *   blended translation and the warp_to_live translation to normals). */
void orc_warp(const float *nodes, int M, float *points, float *normals, long long N, int stride, orc_aff3f warp_to_live, int flags)
{
    long long cursor = 0;
    for (long long p = 0; p < N; ++p) {
        float *pt = points + (size_t)p * stride;
        long long ni = (flags & 1) ? cursor : p;
        float *nr = normals + (size_t)ni * stride;
        if (isnan(pt[0]) || isnan(nr[0])) continue;
        int32_t idx[8]; float d2[8];
        orc_knn8(nodes, M, pt, 1, stride, idx, d2);
        float rot4[4], trans4[4];
        orc_dqb(nodes, idx, d2, rot4, trans4, NULL);
        orc_dq_transform(rot4, trans4, pt);
        aff_apply_cv(&warp_to_live, pt);
        if (flags & 2) {
            orc_quat_rotate_vec(rot4, nr);
            float x = nr[0], y = nr[1], z = nr[2];
            nr[0] = warp_to_live.R[0] * x + warp_to_live.R[1] * y + warp_to_live.R[2] * z;
            nr[1] = warp_to_live.R[3] * x + warp_to_live.R[4] * y + warp_to_live.R[5] * z;
            nr[2] = warp_to_live.R[6] * x + warp_to_live.R[7] * y + warp_to_live.R[8] * z;
        } else {
            orc_dq_transform(rot4, trans4, nr);
            aff_apply_cv(&warp_to_live, nr);
        }
        ++cursor;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Grid-accelerated exact 8-NN with the SAME result as orc_knn8 (the 8 smallest by (distance, node index)); used by the
 * pipeline restatement and the CPU baseline so the baseline is not handicapped by an O(N*M) scan (the reference uses a
 * kd-tree, warp_field.cpp:247-251). */
#include <stdlib.h>
typedef struct { float mn[3]; float cell; int dim[3]; int *start; int *items; } orc_grid;

static void grid_build(orc_grid *g, const float *nodes, int M)
{
    float mx[3];
    for (int c = 0; c < 3; ++c) { g->mn[c] = 3.4e38f; mx[c] = -3.4e38f; }
    for (int m = 0; m < M; ++m)
        for (int c = 0; c < 3; ++c) {
            float v = nodes[(size_t)m * ORC_NODE_STRIDE + c];
            if (v < g->mn[c]) g->mn[c] = v;
            if (v > mx[c]) mx[c] = v;
        }
    float ext = 0.f;
    for (int c = 0; c < 3; ++c) if (mx[c] - g->mn[c] > ext) ext = mx[c] - g->mn[c];
    int res = (int)ceil(cbrt((double)M / 2.0));
    if (res < 1) res = 1;
    if (res > 128) res = 128;
    g->cell = ext > 0 ? ext / (float)res * 1.0001f : 1.f;
    for (int c = 0; c < 3; ++c) { g->dim[c] = (int)((mx[c] - g->mn[c]) / g->cell) + 1; }
    size_t ncell = (size_t)g->dim[0] * g->dim[1] * g->dim[2];
    g->start = (int *)calloc(ncell + 1, sizeof(int));
    g->items = (int *)malloc((size_t)(M > 0 ? M : 1) * sizeof(int));
    int *cellof = (int *)malloc((size_t)(M > 0 ? M : 1) * sizeof(int));
    for (int m = 0; m < M; ++m) {
        int ci[3];
        for (int c = 0; c < 3; ++c) {
            ci[c] = (int)((nodes[(size_t)m * ORC_NODE_STRIDE + c] - g->mn[c]) / g->cell);
            if (ci[c] >= g->dim[c]) ci[c] = g->dim[c] - 1;
            if (ci[c] < 0) ci[c] = 0;
        }
        cellof[m] = ci[0] + g->dim[0] * (ci[1] + g->dim[1] * ci[2]);
        g->start[cellof[m] + 1]++;
    }
    for (size_t i = 0; i < ncell; ++i) g->start[i + 1] += g->start[i];
    int *cur = (int *)malloc((ncell + 1) * sizeof(int));
    memcpy(cur, g->start, (ncell + 1) * sizeof(int));
    for (int m = 0; m < M; ++m) g->items[cur[cellof[m]]++] = m;
    free(cur); free(cellof);
}

static inline void knn_insert_lex(int32_t *bi, float *bd, float dist, int m)
{
    if (!(dist < bd[7] || (dist == bd[7] && m < bi[7]))) return;
    int i = 7;
    while (i > 0 && (bd[i - 1] > dist || (bd[i - 1] == dist && bi[i - 1] > m))) { bd[i] = bd[i - 1]; bi[i] = bi[i - 1]; --i; }
    bd[i] = dist; bi[i] = m;
}

void orc_knn8_fast(const float *nodes, int M, const float *queries, long long N, int qstride, int32_t *idx, float *d2)
{
    if (M < 64) { orc_knn8(nodes, M, queries, N, qstride, idx, d2); return; }
    orc_grid g;
    grid_build(&g, nodes, M);
#pragma omp parallel for schedule(dynamic, 256)
    for (long long q = 0; q < N; ++q) {
        const float *p = queries + (size_t)q * qstride;
        int32_t bi[8]; float bd[8];
        for (int i = 0; i < 8; ++i) { bi[i] = 0x7fffffff; bd[i] = 3.402823466e+38f; }
        if (!(p[0] != p[0] || p[1] != p[1] || p[2] != p[2])) {
            int c0[3];
            float dout = 0.f;                      /* distance from p to the grid's bounding box (0 inside) */
            for (int c = 0; c < 3; ++c) {
                float rel = (p[c] - g.mn[c]) / g.cell;
                c0[c] = (int)floorf(rel);
                float lo = g.mn[c], hi = g.mn[c] + g.cell * (float)g.dim[c];
                float dd = p[c] < lo ? lo - p[c] : (p[c] > hi ? p[c] - hi : 0.f);
                dout += dd * dd;
                if (c0[c] < 0) c0[c] = 0;
                if (c0[c] >= g.dim[c]) c0[c] = g.dim[c] - 1;
            }
            (void)dout;
            int maxr = g.dim[0] > g.dim[1] ? g.dim[0] : g.dim[1];
            if (g.dim[2] > maxr) maxr = g.dim[2];
            for (int r = 0; r <= maxr; ++r) {
                /* every unvisited cell is at least (r-1)*cell away from p along some axis once p's own cell is clamped
                 * into the grid; stop when that bound (conservatively shrunk) exceeds the current worst */
                if (r >= 2 && bi[7] != 0x7fffffff) {
                    float bound = (float)(r - 1) * g.cell * 0.999f;
                    if (bound * bound > bd[7]) break;
                }
                for (int z = c0[2] - r; z <= c0[2] + r; ++z) {
                    if (z < 0 || z >= g.dim[2]) continue;
                    for (int y = c0[1] - r; y <= c0[1] + r; ++y) {
                        if (y < 0 || y >= g.dim[1]) continue;
                        int shell = (z == c0[2] - r || z == c0[2] + r || y == c0[1] - r || y == c0[1] + r);
                        for (int x = c0[0] - r; x <= c0[0] + r; x += (shell ? 1 : (2 * r > 0 ? 2 * r : 1))) {
                            if (x < 0 || x >= g.dim[0]) continue;
                            size_t cid = (size_t)x + (size_t)g.dim[0] * ((size_t)y + (size_t)g.dim[1] * z);
                            for (int it = g.start[cid]; it < g.start[cid + 1]; ++it) {
                                int m = g.items[it];
                                const float *v = nodes + (size_t)m * ORC_NODE_STRIDE;
                                float d0 = p[0] - v[0], d1 = p[1] - v[1], dd2 = p[2] - v[2];
                                knn_insert_lex(bi, bd, d0 * d0 + d1 * d1 + dd2 * dd2, m);
                            }
                        }
                    }
                }
            }
        }
        for (int i = 0; i < 8; ++i) { idx[q * 8 + i] = bi[i] == 0x7fffffff ? -1 : bi[i]; d2[q * 8 + i] = bd[i]; }
    }
    free(g.start); free(g.items);
}
