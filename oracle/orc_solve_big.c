/* CPU ORACLE (test infrastructure only) -- data-term solve for full-size frames (N = 307,200 rows, M = thousands of
 * nodes), MATRIX-FREE like Opt's solver (deps/Opt/API/src/solverGPUGaussNewton.t:361-560: J^T J p is applied edge by
 * edge), in double: Levenberg-Marquardt around Jacobi-preconditioned CG with Opt's/Ceres' q-tolerance stopping rule
 * (:1093-1101) and trust-region update (:1122-1155).  Independent of the product's assembled-sparse-matrix formulation,
 * so agreement between the two is a real check.  Same conventions as orc_solve_data_term (orc_solve.c). */
#include "orc_common.h"
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_knn8_fast(const float *nodes, int M, const float *queries, long long N, int qstride, int32_t *idx, float *d2);

typedef struct {
    int M; long long N; const int32_t *idx; const float *w; const uint8_t *valid;
    double quirk_w; int quirk; double quirk_n;
} rows_t;

/* out = W^T (W in)   (3 right-hand sides, layout [d*M + n]) */
static void apply_JtJ(const rows_t *R, const double *in, double *out)
{
    const int M = R->M;
    for (int i = 0; i < 3 * M; ++i) out[i] = 0.0;
#pragma omp parallel
    {
        double *loc = (double *)calloc((size_t)3 * M, sizeof(double));
#pragma omp for schedule(static)
        for (long long v = 0; v < R->N; ++v) {
            if (!R->valid[v]) continue;
            double q[3] = {0, 0, 0};
            for (int k = 0; k < 8; ++k) { int n = R->idx[v * 8 + k]; if (n < 0) continue; double w = R->w[v * 8 + k];
                                          q[0] += w * in[n]; q[1] += w * in[M + n]; q[2] += w * in[2 * M + n]; }
            for (int k = 0; k < 8; ++k) { int n = R->idx[v * 8 + k]; if (n < 0) continue; double w = R->w[v * 8 + k];
                                          loc[n] += w * q[0]; loc[M + n] += w * q[1]; loc[2 * M + n] += w * q[2]; }
        }
#pragma omp critical
        for (int i = 0; i < 3 * M; ++i) out[i] += loc[i];
        free(loc);
    }
    if (R->quirk) for (int d = 0; d < 3; ++d) out[d * M] += R->quirk_n * R->quirk_w * R->quirk_w * in[d * M];
}

int orc_solve_data_term_big(float *nodes, int M, const float *canon, const float *live, long long N, int stride, int flags,
                            int max_lm, int lin_iters, double *stats)
{
    int32_t *idx = (int32_t *)malloc((size_t)N * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)N * 8 * sizeof(float));
    float *w = (float *)malloc((size_t)N * 8 * sizeof(float));
    uint8_t *valid = (uint8_t *)malloc((size_t)N);
    orc_knn8_fast(nodes, M, canon, N, stride, idx, d2);
    long long nvalid = 0;
    double c0 = 0.0;
    for (long long v = 0; v < N; ++v) {
        const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
        valid[v] = !(isnan(c[0]) || isnan(c[1]) || isnan(c[2]) || isnan(l[0]) || isnan(l[1]) || isnan(l[2]));
        for (int k = 0; k < 8; ++k) {
            int32_t n = idx[v * 8 + k];
            float nw = n >= 0 ? nodes[(size_t)n * ORC_NODE_STRIDE + 11] : 1.f;
            w[v * 8 + k] = (valid[v] && n >= 0) ? (float)exp((double)(-d2[v * 8 + k] / (2 * nw * nw))) : 0.f;
            if (!valid[v]) idx[v * 8 + k] = -1;
        }
        if (valid[v]) { float b0 = l[0] - c[0], b1 = l[1] - c[1], b2 = l[2] - c[2]; c0 += 0.5 * ((double)b0 * b0 + (double)b1 * b1 + (double)b2 * b2); }
        nvalid += valid[v];
    }
    rows_t R = {M, N, idx, w, valid, 0.0, 0, (double)N};
    const int M3 = 3 * M;
    double *x = (double *)calloc(M3, 8), *gb = (double *)calloc(M3, 8), *diag = (double *)calloc(M, 8);
    double *g = (double *)malloc(M3 * 8), *dl = (double *)malloc(M3 * 8), *r = (double *)malloc(M3 * 8), *z = (double *)malloc(M3 * 8);
    double *p = (double *)malloc(M3 * 8), *Ap = (double *)malloc(M3 * 8);
    for (int m = 0; m < M; ++m) { float t4[4]; orc_node_translation(nodes + (size_t)m * ORC_NODE_STRIDE, t4); x[m] = t4[1]; x[M + m] = t4[2]; x[2 * M + m] = t4[3]; }
    for (long long v = 0; v < N; ++v) {
        if (!valid[v]) continue;
        const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
        float b[3] = {l[0] - c[0], l[1] - c[1], l[2] - c[2]};
        for (int k = 0; k < 8; ++k) { int n = idx[v * 8 + k]; if (n < 0) continue; double wk = w[v * 8 + k];
                                      gb[n] += wk * b[0]; gb[M + n] += wk * b[1]; gb[2 * M + n] += wk * b[2]; diag[n] += wk * wk; }
    }
    if ((flags & 1) && N > 0 && valid[0]) {
        for (int k = 0; k < 8; ++k) R.quirk_w += (double)w[k];
        R.quirk = 1;
        float b[3] = {live[0] - canon[0], live[1] - canon[1], live[2] - canon[2]};
        for (int d = 0; d < 3; ++d) gb[d * M] += (double)N * R.quirk_w * b[d];
        diag[0] += (double)N * R.quirk_w * R.quirk_w;
        c0 += (double)N * 0.5 * ((double)b[0] * b[0] + (double)b[1] * b[1] + (double)b[2] * b[2]);
    }
    apply_JtJ(&R, x, Ap);
    double cost = c0;
    for (int i = 0; i < M3; ++i) cost += x[i] * (0.5 * Ap[i] - gb[i]);
    const double cost0 = cost;
    double radius = 1e4, decrease = 2.0;
    int it = 0, pcg_total = 0;
    for (; it < max_lm; ++it) {
        apply_JtJ(&R, x, Ap);
        double rz = 0.0;
        for (int i = 0; i < M3; ++i) {
            g[i] = gb[i] - Ap[i];
            double d = diag[i % M], cd = fmin(fmax(d, 1e-6), 1e32) / radius;
            dl[i] = 0.0; r[i] = g[i]; z[i] = g[i] / (d + cd); p[i] = z[i]; rz += g[i] * z[i];
        }
        double Q0 = 0.0;
        for (int l = 0; l < lin_iters && rz > 0.0; ++l) {
            apply_JtJ(&R, p, Ap);
            double pAp = 0.0;
            for (int i = 0; i < M3; ++i) { double d = diag[i % M]; Ap[i] += fmin(fmax(d, 1e-6), 1e32) / radius * p[i]; pAp += p[i] * Ap[i]; }
            if (!(pAp > 0.0)) break;
            double alpha = rz / pAp, rz_new = 0.0, qs = 0.0;
            for (int i = 0; i < M3; ++i) {
                double d = diag[i % M];
                dl[i] += alpha * p[i]; r[i] -= alpha * Ap[i];
                z[i] = r[i] / (d + fmin(fmax(d, 1e-6), 1e32) / radius);
                rz_new += r[i] * z[i]; qs += dl[i] * (r[i] + g[i]);
            }
            double Q1 = -0.5 * qs, beta = rz_new / rz;
            for (int i = 0; i < M3; ++i) p[i] = z[i] + beta * p[i];
            rz = rz_new; ++pcg_total;
            double zeta = (double)(l + 1) * (Q1 - Q0) / Q1;
            Q0 = Q1;
            if (zeta < 1e-4) break;
        }
        double mm = 0.0, aa = 0.0, dg = 0.0;
        for (int i = 0; i < M3; ++i) {
            double d = diag[i % M], cd = fmin(fmax(d, 1e-6), 1e32) / radius * dl[i];
            mm += dl[i] * (g[i] + r[i] + cd); aa += dl[i] * (g[i] - r[i] - cd); dg += dl[i] * g[i];
        }
        double model = 0.5 * mm, new_cost = cost - dg + 0.5 * aa, change = cost - new_cost;
        double rho = model > 0.0 ? change / model : 0.0;
        if (change >= 0.0 && rho > 1e-3) {
            for (int i = 0; i < M3; ++i) x[i] += dl[i];
            int stop = change <= cost * 1e-6;
            cost = new_cost;
            double f = 1.0 - (2.0 * rho - 1.0) * (2.0 * rho - 1.0) * (2.0 * rho - 1.0);
            radius /= (f > 1.0 / 3.0 ? f : 1.0 / 3.0);
            if (radius > 1e16) radius = 1e16;
            decrease = 2.0;
            if (stop) { ++it; break; }
        } else {
            radius /= decrease; decrease *= 2.0;
            if (radius <= 1e-32) break;
        }
    }
    for (int m = 0; m < M; ++m) orc_node_encode_translation(nodes + (size_t)m * ORC_NODE_STRIDE, (float)x[m], (float)x[M + m], (float)x[2 * M + m]);
    if (stats) { stats[0] = cost0; stats[1] = cost; stats[2] = it; stats[3] = (double)nvalid; stats[4] = pcg_total; stats[5] = 0; }
    free(idx); free(d2); free(w); free(valid); free(x); free(gb); free(diag); free(g); free(dl); free(r); free(z); free(p); free(Ap);
    return 1;
}
