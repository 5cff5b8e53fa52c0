/* CPU ORACLE (test infrastructure only) -- warp-field data-term solve.
 *
 * The reference solves  min_T  sum_v || live_v - canon_v - sum_k w_vk T_{n_vk} ||^2   over the node
 * translations T (kfusion/solvers/dynamicfusion.t:26-52) with Opt's matrix-free LM/PCG
 * (deps/Opt/API/src/solverGPUGaussNewton.t:1016-1177), starting from the nodes' current translations
 * (CombinedSolver.h:161-181) and writing the result back with encodeTranslation (CombinedSolver.h:189-197).
 * Opt needs Terra (pinned binary release-2016-03-25, not vendored, absent here), so this restates the
 * PUBLISHED algorithm: Levenberg-Marquardt with Ceres-style trust region (radius 1e4, diagonal scaling
 * clamped to [1e-6, 1e32]), each linear system solved EXACTLY here by dense Cholesky in double.
 * Anchors: the reference's tests/warp_test.cpp scenarios (tests/test_oracle_golden.py).
 *
 * Connectivity follows CombinedSolver::initializeConnectivity (CombinedSolver.h:66-84): per vertex the 8
 * nearest nodes of the (already warped) canonical vertex and their weights.  Rows with a NaN in
 * canon or live are skipped (the reference zero-fills them and keeps stale k-NN scratch: documented
 * divergence).  flags bit0 = reproduce the 2N-edge quirk (N extra copies of edge (v=0; n_k=0)).
 *
 * stats: [0] initial cost, [1] final cost, [2] LM iterations, [3] valid rows */
#include "orc_common.h"
#include <stdlib.h>

static int chol_solve(double *A, double *B, int n, int nrhs)
{
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0)) return 0;
        d = sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            const double *ai = A + (size_t)i * n, *aj = A + (size_t)j * n;
            for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int r = 0; r < nrhs; ++r) {
        double *b = B + (size_t)r * n;
        for (int i = 0; i < n; ++i) {
            double s = b[i];
            for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k];
            b[i] = s / A[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = b[i];
            for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k];
            b[i] = s / A[(size_t)i * n + i];
        }
    }
    return 1;
}

int orc_solve_data_term(float *nodes, int M, const float *canon, const float *live, long long N, int stride, int flags, int max_lm, double *stats)
{
    /* max_lm: total LM iterations = numIter x nonLinearIter without earlyOut (kinfu.cpp:114-120: 5; tests: 20 x 15) */
    int32_t *idx = (int32_t *)malloc((size_t)N * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)N * 8 * sizeof(float));
    float *w = (float *)malloc((size_t)N * 8 * sizeof(float));
    uint8_t *valid = (uint8_t *)malloc((size_t)N);
    orc_knn8(nodes, M, canon, N, stride, idx, d2);
    long long nvalid = 0;
    for (long long v = 0; v < N; ++v) {
        const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
        valid[v] = !(isnan(c[0]) || isnan(c[1]) || isnan(c[2]) || isnan(l[0]) || isnan(l[1]) || isnan(l[2]));
        for (int k = 0; k < 8; ++k) {
            int32_t n = idx[v * 8 + k];
            float nw = n >= 0 ? nodes[(size_t)n * ORC_NODE_STRIDE + 11] : 1.f;
            w[v * 8 + k] = (valid[v] && n >= 0) ? (float)exp((double)(-d2[v * 8 + k] / (2 * nw * nw))) : 0.f;
        }
        nvalid += valid[v];
    }
    /* unknowns: current node translations (CombinedSolver.h:165-172) */
    double *T = (double *)calloc((size_t)M * 3, sizeof(double));
    for (int m = 0; m < M; ++m) {
        float t4[4];
        orc_node_translation(nodes + (size_t)m * ORC_NODE_STRIDE, t4);
        T[m * 3 + 0] = t4[1]; T[m * 3 + 1] = t4[2]; T[m * 3 + 2] = t4[3];
    }
    /* constant normal matrix JtJ = W^T W (scalar blocks) */
    double *JtJ = (double *)calloc((size_t)M * M, sizeof(double));
    for (long long v = 0; v < N; ++v) {
        if (!valid[v]) continue;
        for (int a = 0; a < 8; ++a) {
            if (idx[v * 8 + a] < 0) continue;
            for (int b = 0; b < 8; ++b) {
                if (idx[v * 8 + b] < 0) continue;
                JtJ[(size_t)idx[v * 8 + a] * M + idx[v * 8 + b]] += (double)w[v * 8 + a] * (double)w[v * 8 + b];
            }
        }
    }
    double quirk_w = 0.0; int quirk = 0;
    if ((flags & 1) && N > 0 && valid[0]) {     /* N copies of the edge (v = 0, n_k = node 0 for all k) */
        for (int k = 0; k < 8; ++k) quirk_w += (double)w[k];
        JtJ[0] += (double)N * quirk_w * quirk_w;
        quirk = 1;
    }
    double *g = (double *)malloc((size_t)M * 3 * sizeof(double));
    double *A = (double *)malloc((size_t)M * M * sizeof(double));
    double *delta = (double *)malloc((size_t)M * 3 * sizeof(double));
    double *Tn = (double *)malloc((size_t)M * 3 * sizeof(double));

#define COST_AND_GRAD(Tv, costp, gradp)                                                                     \
    do {                                                                                                    \
        double cst = 0.0;                                                                                   \
        if (gradp) for (int i_ = 0; i_ < M * 3; ++i_) (gradp)[i_] = 0.0;                                    \
        for (long long v = 0; v < N; ++v) {                                                                 \
            if (!valid[v]) continue;                                                                        \
            const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;                    \
            double r[3] = {(double)(l[0] - c[0]), (double)(l[1] - c[1]), (double)(l[2] - c[2])};            \
            for (int k = 0; k < 8; ++k) { int n = idx[v * 8 + k]; if (n < 0) continue;                      \
                for (int d = 0; d < 3; ++d) r[d] -= (double)w[v * 8 + k] * (Tv)[n * 3 + d]; }               \
            cst += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];                                                 \
            if (gradp) for (int k = 0; k < 8; ++k) { int n = idx[v * 8 + k]; if (n < 0) continue;           \
                for (int d = 0; d < 3; ++d) (gradp)[d * M + n] += (double)w[v * 8 + k] * r[d]; }            \
        }                                                                                                   \
        if (quirk) {                                                                                        \
            double r[3] = {(double)(live[0] - canon[0]), (double)(live[1] - canon[1]), (double)(live[2] - canon[2])}; \
            for (int d = 0; d < 3; ++d) r[d] -= quirk_w * (Tv)[d];                                          \
            cst += (double)N * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);                                   \
            if (gradp) for (int d = 0; d < 3; ++d) (gradp)[d * M + 0] += (double)N * quirk_w * r[d];        \
        }                                                                                                   \
        *(costp) = 0.5 * cst;                                                                               \
    } while (0)

    double cost, cost0;
    COST_AND_GRAD(T, &cost, g);
    cost0 = cost;
    double radius = 1e4, decrease = 2.0;        /* solverGPUGaussNewton.t:26-39 */
    int it = 0;
    for (; it < max_lm; ++it) {
        for (size_t i = 0; i < (size_t)M * M; ++i) A[i] = JtJ[i];
        for (int m = 0; m < M; ++m) {
            double dgn = JtJ[(size_t)m * M + m];
            double c = dgn < 1e-6 ? 1e-6 : (dgn > 1e32 ? 1e32 : dgn);
            A[(size_t)m * M + m] += c / radius;
        }
        for (int i = 0; i < M * 3; ++i) delta[i] = g[i];
        if (!chol_solve(A, delta, M, 3)) break;
        /* model cost change for a linear least-squares problem: 0.5*(2 g.d - d^T JtJ d) */
        double model = 0.0;
        for (int d = 0; d < 3; ++d)
            for (int m = 0; m < M; ++m) {
                double jd = 0.0;
                for (int n = 0; n < M; ++n) jd += JtJ[(size_t)m * M + n] * delta[d * M + n];
                model += delta[d * M + m] * (2.0 * g[d * M + m] - jd);
            }
        model *= 0.5;
        for (int m = 0; m < M; ++m) for (int d = 0; d < 3; ++d) Tn[m * 3 + d] = T[m * 3 + d] + delta[d * M + m];
        double new_cost;
        COST_AND_GRAD(Tn, &new_cost, (double *)0);
        double change = cost - new_cost;
        double rho = model > 0 ? change / model : 0.0;
        if (change >= 0 && rho > 1e-3) {
            for (int i = 0; i < M * 3; ++i) T[i] = Tn[i];
            int stop = change <= cost * 1e-6;   /* function_tolerance, CombinedSolver.h:88 */
            cost = new_cost;
            double f = 1.0 - (2.0 * rho - 1.0) * (2.0 * rho - 1.0) * (2.0 * rho - 1.0);
            radius /= (f > 1.0 / 3.0 ? f : 1.0 / 3.0);
            if (radius > 1e16) radius = 1e16;
            decrease = 2.0;
            if (stop && !(flags & 2)) { ++it; break; }
            COST_AND_GRAD(T, &cost, g);
        } else {
            radius /= decrease; decrease *= 2.0;
            if (radius <= 1e-32) break;
        }
    }
    for (int m = 0; m < M; ++m)
        orc_node_encode_translation(nodes + (size_t)m * ORC_NODE_STRIDE, (float)T[m * 3], (float)T[m * 3 + 1], (float)T[m * 3 + 2]);
    if (stats) { stats[0] = cost0; stats[1] = cost; stats[2] = it; stats[3] = (double)nvalid; }
    free(idx); free(d2); free(w); free(valid); free(T); free(JtJ); free(g); free(A); free(delta); free(Tn);
    return 1;
#undef COST_AND_GRAD
}
