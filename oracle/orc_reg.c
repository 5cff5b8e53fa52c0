/* CPU ORACLE (test infrastructure only) -- SURVEY.md 8f(2): the robust data term over 6-DoF node increments plus the
 * regularisation term, i.e. the energy the reference defines piecewise and never assembles:
 *
 *   - 6-wide parameter blocks per node, (rotation increment, translation increment): DynamicFusionDataEnergy / DynamicFusionRegEnergy
 *     register KNN_NEIGHBOURS blocks of 6 (kfusion/include/kfusion/optimisation.hpp:108-110,141-143); the data functor reads only
 *     epsilon[3..5] (:47) -- the rotation half exists and has a zero Jacobian;
 *   - tukeyPenalty(x, c = 0.01) = x (1 - x^2/c^2)^2 for |x| <= c, else 0 (optimisation.hpp:84-88; kfusion/solvers/dynamicfusion.t:43-49,
 *     applied to the data residual in the commented-out Energy line :51): Tukey's biweight INFLUENCE function psi = rho';
 *   - huberPenalty(a, delta = 1e-4) = a^2/2 for |a| <= delta, else delta |a| - delta^2/2 (optimisation.hpp:134-138; dynamicfusion.t:34-40):
 *     Huber's LOSS rho;
 *   - the regularisation functor is an empty stub (optimisation.hpp:125-132), WarpField::energy_reg is empty (warp_field.cpp:168-172) and
 *     KinFu::edges_ (pairs of dual quaternions, kinfu.hpp:95) is never filled.
 *
 * PARITY UNPINNED: the reference has no code that evaluates this energy, so nothing here can be checked against it.  The restatement
 * uses only operations the reference defines, read as the M-estimator problem of the DynamicFusion paper (eq. 6-8) that the names point to:
 *
 *   E(eps) = sum_v sum_c rho_T(r_vc)  +  lambda sum_(i,j) alpha_ij sum_c rho_H(d_ijc)
 *   r_v    = live_v - [ rotate(qhat_v, canon_v) + sum_k w_vk t_k ],  qhat_v = normalize(sum_k w_vk q_k)   (WarpField::DQB + transform,
 *            warp_field.cpp:203-217, dual_quaternion.hpp:204-210: the warp the rest of the pipeline applies)
 *   d_ij   = T_i(g_j) - T_j(g_j),  T_k(p) = rotate(q_k, p) + t_k  (DualQuaternion::transform),  g_j = node j's position,
 *            j in the reg_k nearest other nodes of i,  alpha_ij = max(weight_i, weight_j)
 *   rho_T' = tukeyPenalty  (rho_T(x) = c^2/6 (1 - (1 - x^2/c^2)^3), constant c^2/6 beyond c),  rho_H = huberPenalty
 *   node increment eps_k = (omega_k, tau_k):  q_k <- exp(omega_k) q_k,  t_k <- t_k + tau_k
 *
 * minimised by Gauss-Newton / IRLS (weights rho'(x)/x: (1 - x^2/c^2)^2 and min(1, delta/|a|)) with Levenberg damping mu * diag(H); every
 * linear system is solved exactly (dense Cholesky, double), so this is only meant for M up to a few hundred nodes.  The CUDA solver
 * (csrc/regsolve.cu: matrix-free block-Jacobi PCG) is compared with it on energies and node parameters.
 *
 * flags: bit0 optimise the rotation increments (otherwise translation-only), bit1 Tukey on the data term (otherwise squared loss),
 *        bit2 Huber on the regularisation term (otherwise squared loss).
 * stats (16 doubles): [0] energy before, [1] energy after, [2] GN iterations, [3] valid vertices, [4] data energy after, [5] reg energy after,
 *        [6] edges, [8 + it] energy before GN iteration it (it < 8). */
#include "orc_common.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    double lambda, tukey_c, huber_delta, lm_mu;
    int gn_iters, reg_k, flags, lin_iters;     /* lin_iters: PCG cap of the CUDA solver (unused here: exact solves) */
} orc_f2_params;

static void qmul(const double *a, const double *b, double *o)
{
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
static void qrot(const double *q, const double *p, double *o)      /* unit q: o = q (0,p) q* */
{
    const double t[3] = {2 * (q[2] * p[2] - q[3] * p[1]), 2 * (q[3] * p[0] - q[1] * p[2]), 2 * (q[1] * p[1] - q[2] * p[0])};
    o[0] = p[0] + q[0] * t[0] + (q[2] * t[2] - q[3] * t[1]);
    o[1] = p[1] + q[0] * t[1] + (q[3] * t[0] - q[1] * t[2]);
    o[2] = p[2] + q[0] * t[2] + (q[1] * t[1] - q[2] * t[0]);
}
static void node_state(const float *n, double *q, double *t)        /* unit rotation quaternion and translation of a node, in double */
{
    double nn = 0;
    for (int i = 0; i < 4; ++i) { q[i] = n[3 + i]; nn += q[i] * q[i]; }
    nn = sqrt(nn);
    for (int i = 0; i < 4; ++i) q[i] /= nn;
    const double d[4] = {2.0 * n[7], 2.0 * n[8], 2.0 * n[9], 2.0 * n[10]}, c[4] = {q[0], -q[1], -q[2], -q[3]};
    double r[4];
    qmul(d, c, r);                                                   /* getTranslation: 2 * dual * conj(rot) */
    t[0] = r[1]; t[1] = r[2]; t[2] = r[3];
}
static double rho_tukey(double x, double c) { if (fabs(x) > c) return c * c / 6.0; const double u = 1.0 - x * x / (c * c); return c * c / 6.0 * (1.0 - u * u * u); }
static double w_tukey(double x, double c) { if (fabs(x) > c) return 0.0; const double u = 1.0 - x * x / (c * c); return u * u; }
static double rho_huber(double a, double d) { return fabs(a) <= d ? a * a / 2 : d * fabs(a) - d * d / 2; }
static double w_huber(double a, double d) { return fabs(a) <= d ? 1.0 : d / fabs(a); }

static int chol_solve1(double *A, double *b, int n)
{
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0)) return 0;
        d = sqrt(d);
        A[(size_t)j * n + j] = d;
#pragma omp parallel for schedule(static)
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            const double *ai = A + (size_t)i * n, *aj = A + (size_t)j * n;
            for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    return 1;
}

/* the 3x3 matrix B with  phi = B omega_k : the rotation-vector increment of the blended rotation qhat for an increment omega_k of node k,
 * qhat = Q/|Q|, Q = sum_k w_k q_k, q_k <- exp(omega_k) q_k ~ q_k + 1/2 (0, omega_k) q_k:
 *   dQ = w_k/2 (0, omega) q_k;  dqhat = (I - qhat qhat^T) dQ / |Q|;  (0, phi) = 2 dqhat qhat*  */
static void blend_B(const double *qhat, double nQ, double wk, const double *qk, double *B)
{
    const double qc[4] = {qhat[0], -qhat[1], -qhat[2], -qhat[3]};
    for (int a = 0; a < 3; ++a) {
        double e[4] = {0, 0, 0, 0}, dq[4], out[4];
        e[1 + a] = 1.0;
        qmul(e, qk, dq);
        double dot = 0;
        for (int i = 0; i < 4; ++i) dot += qhat[i] * dq[i];
        for (int i = 0; i < 4; ++i) dq[i] = (dq[i] - qhat[i] * dot) * (wk / nQ);      /* the 1/2 and the 2 cancel */
        qmul(dq, qc, out);
        B[0 * 3 + a] = out[1]; B[1 * 3 + a] = out[2]; B[2 * 3 + a] = out[3];
    }
}

/* edges: for node i the reg_k nearest OTHER nodes (k-NN over the node positions, ties to the lower index); returns the edge count */
int orc_f2_edges(const float *nodes, int M, int reg_k, int32_t *edge_j)
{
    int32_t *idx = (int32_t *)malloc((size_t)M * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)M * 8 * sizeof(float));
    orc_knn8(nodes, M, nodes, M, ORC_NODE_STRIDE, idx, d2);
    int total = 0;
    for (int i = 0; i < M; ++i) {
        int got = 0;
        for (int k = 0; k < 8 && got < reg_k; ++k) {
            const int j = idx[i * 8 + k];
            if (j < 0 || j == i) continue;
            edge_j[(size_t)i * reg_k + got++] = j;
        }
        for (; got < reg_k; ++got) edge_j[(size_t)i * reg_k + got] = -1;
    }
    for (size_t e = 0; e < (size_t)M * reg_k; ++e) total += edge_j[e] >= 0;
    free(idx); free(d2);
    return total;
}

int orc_solve_f2(float *nodes, int M, const float *canon, const float *live, long long N, int stride, const orc_f2_params *prm, double *stats)
{
    const int twist = prm->flags & 1, rob_d = prm->flags & 2, rob_r = prm->flags & 4;
    const int reg_k = prm->reg_k > 7 ? 7 : (prm->reg_k < 0 ? 0 : prm->reg_k);
    const int n6 = 6 * M;
    int32_t *idx = (int32_t *)malloc((size_t)N * 8 * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)N * 8 * sizeof(float));
    float *w = (float *)malloc((size_t)N * 8 * sizeof(float));
    uint8_t *valid = (uint8_t *)malloc((size_t)N);
    orc_knn8(nodes, M, canon, N, stride, idx, d2);
    long long nvalid = 0;
    for (long long v = 0; v < N; ++v) {
        const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
        valid[v] = !(isnan(c[0]) || isnan(c[1]) || isnan(c[2]) || isnan(l[0]) || isnan(l[1]) || isnan(l[2])) && idx[v * 8 + 7] >= 0;
        for (int k = 0; k < 8; ++k) {
            const int32_t n = idx[v * 8 + k];
            const float nw = n >= 0 ? nodes[(size_t)n * ORC_NODE_STRIDE + 11] : 1.f;
            w[v * 8 + k] = (valid[v] && n >= 0) ? (float)exp((double)(-d2[v * 8 + k] / (2 * nw * nw))) : 0.f;   /* warp_field.cpp:238-241 */
        }
        nvalid += valid[v];
    }
    int32_t *edge_j = (int32_t *)malloc((size_t)M * (reg_k ? reg_k : 1) * sizeof(int32_t));
    const int nedges = reg_k ? orc_f2_edges(nodes, M, reg_k, edge_j) : 0;
    double *H = (double *)malloc((size_t)n6 * n6 * sizeof(double));
    double *g = (double *)malloc((size_t)n6 * sizeof(double));
    double *Q = (double *)malloc((size_t)M * 4 * sizeof(double)), *T = (double *)malloc((size_t)M * 3 * sizeof(double));
    memset(stats, 0, 16 * sizeof(double));
    double e_data = 0, e_reg = 0;
    int it = 0;
    for (;; ++it) {
        for (int m = 0; m < M; ++m) node_state(nodes + (size_t)m * ORC_NODE_STRIDE, Q + 4 * m, T + 3 * m);
        const int assemble = it < prm->gn_iters;
        if (assemble) { memset(H, 0, (size_t)n6 * n6 * sizeof(double)); memset(g, 0, (size_t)n6 * sizeof(double)); }
        e_data = 0; e_reg = 0;
        for (long long v = 0; v < N; ++v) {
            if (!valid[v]) continue;
            const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
            const double p[3] = {c[0], c[1], c[2]};
            double Qs[4] = {0, 0, 0, 0}, ts[3] = {0, 0, 0};
            for (int k = 0; k < 8; ++k) {
                const int n = idx[v * 8 + k];
                const double wk = w[v * 8 + k];
                for (int i = 0; i < 4; ++i) Qs[i] += wk * Q[4 * n + i];
                for (int i = 0; i < 3; ++i) ts[i] += wk * T[3 * n + i];
            }
            const double nQ = sqrt(Qs[0] * Qs[0] + Qs[1] * Qs[1] + Qs[2] * Qs[2] + Qs[3] * Qs[3]);
            double qh[4] = {Qs[0] / nQ, Qs[1] / nQ, Qs[2] / nQ, Qs[3] / nQ}, y[3];
            qrot(qh, p, y);
            double r[3], W[3];
            for (int i = 0; i < 3; ++i) {
                r[i] = (double)l[i] - (y[i] + ts[i]);
                W[i] = rob_d ? w_tukey(r[i], prm->tukey_c) : 1.0;
                e_data += rob_d ? rho_tukey(r[i], prm->tukey_c) : 0.5 * r[i] * r[i];
            }
            if (!assemble) continue;
            /* J (3 x 48): d warped / d (omega_k, tau_k) = [ -[y]x B_k , w_k I ] */
            double J[8][3][6];
            for (int k = 0; k < 8; ++k) {
                const int n = idx[v * 8 + k];
                double B[9];
                memset(J[k], 0, sizeof J[k]);
                if (twist) {
                    blend_B(qh, nQ, w[v * 8 + k], Q + 4 * n, B);
                    /* -[y]x B : row i = -(y x B_col) -> (B_col x y) */
                    for (int a = 0; a < 3; ++a) {
                        const double b0 = B[0 * 3 + a], b1 = B[1 * 3 + a], b2 = B[2 * 3 + a];
                        J[k][0][a] = b1 * y[2] - b2 * y[1];
                        J[k][1][a] = b2 * y[0] - b0 * y[2];
                        J[k][2][a] = b0 * y[1] - b1 * y[0];
                    }
                }
                for (int i = 0; i < 3; ++i) J[k][i][3 + i] = w[v * 8 + k];
            }
            for (int ka = 0; ka < 8; ++ka) {
                const int na = idx[v * 8 + ka];
                for (int a = 0; a < 6; ++a) {
                    double ga = 0;
                    for (int i = 0; i < 3; ++i) ga += J[ka][i][a] * W[i] * r[i];
                    g[6 * na + a] += ga;
                    for (int kb = 0; kb < 8; ++kb) {
                        const int nb = idx[v * 8 + kb];
                        for (int b = 0; b < 6; ++b) {
                            double h = 0;
                            for (int i = 0; i < 3; ++i) h += J[ka][i][a] * W[i] * J[kb][i][b];
                            H[(size_t)(6 * na + a) * n6 + 6 * nb + b] += h;
                        }
                    }
                }
            }
        }
        for (int i = 0; i < M && reg_k; ++i)
            for (int e = 0; e < reg_k; ++e) {
                const int j = edge_j[(size_t)i * reg_k + e];
                if (j < 0) continue;
                const double gj[3] = {nodes[(size_t)j * ORC_NODE_STRIDE], nodes[(size_t)j * ORC_NODE_STRIDE + 1], nodes[(size_t)j * ORC_NODE_STRIDE + 2]};
                double yi[3], yj[3], d[3], Wd[3];
                qrot(Q + 4 * i, gj, yi); qrot(Q + 4 * j, gj, yj);
                const double alpha = fmax((double)nodes[(size_t)i * ORC_NODE_STRIDE + 11], (double)nodes[(size_t)j * ORC_NODE_STRIDE + 11]) * prm->lambda;
                for (int c = 0; c < 3; ++c) {
                    d[c] = (yi[c] + T[3 * i + c]) - (yj[c] + T[3 * j + c]);
                    Wd[c] = alpha * (rob_r ? w_huber(d[c], prm->huber_delta) : 1.0);
                    e_reg += alpha * (rob_r ? rho_huber(d[c], prm->huber_delta) : 0.5 * d[c] * d[c]);
                }
                if (!assemble) continue;
                /* d(delta) = d + Ji eps_i + Jj eps_j;  Ji = [ -[yi]x , I ],  Jj = [ +[yj]x , -I ] */
                double Je[2][3][6];
                memset(Je, 0, sizeof Je);
                if (twist) {
                    const double sk_i[9] = {0, -yi[2], yi[1], yi[2], 0, -yi[0], -yi[1], yi[0], 0}, sk_j[9] = {0, -yj[2], yj[1], yj[2], 0, -yj[0], -yj[1], yj[0], 0};
                    for (int r_ = 0; r_ < 3; ++r_) for (int a = 0; a < 3; ++a) { Je[0][r_][a] = -sk_i[r_ * 3 + a]; Je[1][r_][a] = sk_j[r_ * 3 + a]; }
                }
                for (int c = 0; c < 3; ++c) { Je[0][c][3 + c] = 1.0; Je[1][c][3 + c] = -1.0; }
                const int nn[2] = {i, j};
                for (int sa = 0; sa < 2; ++sa)
                    for (int a = 0; a < 6; ++a) {
                        double ga = 0;
                        for (int c = 0; c < 3; ++c) ga += Je[sa][c][a] * Wd[c] * d[c];
                        g[6 * nn[sa] + a] -= ga;
                        for (int sb = 0; sb < 2; ++sb)
                            for (int b = 0; b < 6; ++b) {
                                double h = 0;
                                for (int c = 0; c < 3; ++c) h += Je[sa][c][a] * Wd[c] * Je[sb][c][b];
                                H[(size_t)(6 * nn[sa] + a) * n6 + 6 * nn[sb] + b] += h;
                            }
                    }
            }
        const double energy = e_data + e_reg;
        if (it == 0) stats[0] = energy;
        if (it < 8) stats[8 + it] = energy;
        if (!assemble) { stats[1] = energy; break; }
        /* Levenberg damping; the rotation increments of a translation-only solve are pinned to zero by a unit diagonal */
        for (int i = 0; i < n6; ++i) {
            double *hd = H + (size_t)i * n6 + i;
            if (!twist && (i % 6) < 3) { *hd = 1.0; g[i] = 0.0; }
            else *hd += prm->lm_mu * (*hd) + 1e-12;
        }
        if (!chol_solve1(H, g, n6)) break;
        for (int m = 0; m < M; ++m) {
            float *nd = nodes + (size_t)m * ORC_NODE_STRIDE;
            const double *dl = g + 6 * m;
            double q[4] = {Q[4 * m], Q[4 * m + 1], Q[4 * m + 2], Q[4 * m + 3]};
            const double th = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
            if (twist && th > 0) {
                const double s = sin(th / 2) / th, e[4] = {cos(th / 2), s * dl[0], s * dl[1], s * dl[2]};
                double qn[4];
                qmul(e, q, qn);
                const double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
                for (int i = 0; i < 4; ++i) q[i] = qn[i] / nn;
            }
            const double t[3] = {T[3 * m] + dl[3], T[3 * m + 1] + dl[4], T[3 * m + 2] + dl[5]};
            const double h[4] = {0, 0.5 * t[0], 0.5 * t[1], 0.5 * t[2]};
            double dual[4];
            qmul(h, q, dual);                                       /* DualQuaternion(t, r): dual part = 1/2 (0, t) r */
            for (int i = 0; i < 4; ++i) { nd[3 + i] = (float)q[i]; nd[7 + i] = (float)dual[i]; }
        }
    }
    stats[2] = it; stats[3] = (double)nvalid; stats[4] = e_data; stats[5] = e_reg; stats[6] = nedges;
    free(idx); free(d2); free(w); free(valid); free(edge_j); free(H); free(g); free(Q); free(T);
    return 0;
}
