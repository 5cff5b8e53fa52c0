// regsolve.cu -- SURVEY.md 8f(2): robust data term over 6-DoF node increments + regularisation term, on sm_100a.  OPT-IN, beside the
// reference-behaviour solve of solve.cu (translation-only, squared loss, no regulariser: what the reference actually runs).
//
// What the reference defines and never assembles (PARITY UNPINNED, see oracle/orc_reg.c for the full citation list): 6-wide parameter
// blocks per node (optimisation.hpp:108-110,141-143), tukeyPenalty = Tukey's influence function (optimisation.hpp:84-88,
// dynamicfusion.t:43-51), huberPenalty = Huber's loss (optimisation.hpp:134-138, dynamicfusion.t:34-40), an empty regularisation
// functor / WarpField::energy_reg (optimisation.hpp:125-132, warp_field.cpp:168-172), the never-filled KinFu::edges_ (kinfu.hpp:95).
// Read as the M-estimator problem those names point to (DynamicFusion eq. 6-8):
//   E = sum_v sum_c rho_T(live_v - warp(canon_v))_c + lambda sum_(i,j) alpha_ij sum_c rho_H(T_i(g_j) - T_j(g_j))_c
// with warp = the DQB warp the rest of the pipeline applies, node increment (omega_k, tau_k): q_k <- exp(omega_k) q_k, t_k <- t_k + tau_k.
//
// Solver: Gauss-Newton / IRLS; each step's 6M x 6M system (H + mu diag H) delta = g is solved MATRIX-FREE by block-Jacobi PCG in double:
//   f2_linearize   per vertex: blend, residual, Tukey weights, the eight 3 x 3 rotation Jacobians (kept in float), gradient and the
//                  6 x 6 diagonal blocks (double atomics);  f2_reg: the same per regularisation edge;
//   f2_blocks      per node: damping + inverse of its 6 x 6 block (the preconditioner);
//   per PCG step   f2_apply_data (per vertex: t = J p, u = W t, scatter J^T u), f2_apply_reg (per edge), f2_pcg_update (ONE block:
//                  both dot products, alpha, beta, the vector updates and the preconditioner for all 6M entries) -- no host round trip,
//                  a device-side flag turns the remaining launches of a converged solve into no-ops.
// Design note on tensor cores (north_star: "only for the 6x6/node J^T J block contractions where they actually dominate"): the block
// contractions here are 8 x (3 x 6)^T (3 x 6) products per vertex inside f2_linearize -- 2,300 FMAs against ~350 bytes of traffic and
// 288 atomics per vertex; measured (profiles/) the kernel is bound by the double atomics, not by the multiply-adds, so they
// stay on the FP64 pipe (a tcgen05 tile is 64 x 8 at minimum; these are 6 x 6 and double).
// Sums use double atomics: results are reproducible to ~1e-12 relative, not bit-reproducible run to run (documented; the
// reference-behaviour solve of solve.cu is).
#include "warp_common.cuh"
#include <cstdlib>

using namespace dfb;

namespace {

struct F2Ws {
    int *idx; float *d2; float *w; unsigned char *valid;        // per vertex
    float *Jw;                                                  // per vertex: 8 x 9 rotation Jacobians
    double *Wr;                                                 // per vertex: 3 robust weights
    int *edge_j; double *edge_y; double *edge_W;                // per edge slot (M * reg_k): neighbour, (yi, yj), 3 weights
    double *Q, *T;                                              // per node: unit rotation quaternion (4), translation (3)
    double *g, *D, *Dinv, *damp;                                // 6M, 36M, 36M, 6M
    double *x, *r, *z, *p, *Ap;                                 // 6M each
    double *scal;                                               // [0] rz, [1] rz0, [2] done flag, [3] pcg iterations, [4] e_data, [5] e_reg, [6] valid count
    int *knn_nodes; float *knn_nodes_d2;                        // M x 8: neighbours of the node positions (edge construction)
};

size_t f2_align(size_t v) { return (v + 255) & ~(size_t)255; }

size_t f2_layout(F2Ws &ws, char *base, int M, int N, int reg_k)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += f2_align(bytes); return p; };
    const size_t E = (size_t)M * (reg_k > 0 ? reg_k : 1);
    ws.idx = (int *)take((size_t)N * 8 * 4); ws.d2 = (float *)take((size_t)N * 8 * 4); ws.w = (float *)take((size_t)N * 8 * 4);
    ws.valid = (unsigned char *)take((size_t)N);
    ws.Jw = (float *)take((size_t)N * 72 * 4); ws.Wr = (double *)take((size_t)N * 3 * 8);
    ws.edge_j = (int *)take(E * 4); ws.edge_y = (double *)take(E * 6 * 8); ws.edge_W = (double *)take(E * 3 * 8);
    ws.Q = (double *)take((size_t)M * 4 * 8); ws.T = (double *)take((size_t)M * 3 * 8);
    ws.g = (double *)take((size_t)M * 6 * 8); ws.D = (double *)take((size_t)M * 36 * 8); ws.Dinv = (double *)take((size_t)M * 36 * 8);
    ws.damp = (double *)take((size_t)M * 6 * 8);
    ws.x = (double *)take((size_t)M * 6 * 8); ws.r = (double *)take((size_t)M * 6 * 8); ws.z = (double *)take((size_t)M * 6 * 8);
    ws.p = (double *)take((size_t)M * 6 * 8); ws.Ap = (double *)take((size_t)M * 6 * 8);
    ws.scal = (double *)take(64 * 8);
    ws.knn_nodes = (int *)take((size_t)M * 8 * 4); ws.knn_nodes_d2 = (float *)take((size_t)M * 8 * 4);
    return o;
}

struct Qd { double w, x, y, z; };
__device__ __forceinline__ Qd qdmul(const Qd a, const Qd b)
{
    return Qd{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ void qdrot(const Qd q, const double (&p)[3], double (&o)[3])
{
    const double t0 = 2 * (q.y * p[2] - q.z * p[1]), t1 = 2 * (q.z * p[0] - q.x * p[2]), t2 = 2 * (q.x * p[1] - q.y * p[0]);
    o[0] = p[0] + q.w * t0 + (q.y * t2 - q.z * t1);
    o[1] = p[1] + q.w * t1 + (q.z * t0 - q.x * t2);
    o[2] = p[2] + q.w * t2 + (q.x * t1 - q.y * t0);
}
__device__ __forceinline__ double rho_tukey(double x, double c) { if (fabs(x) > c) return c * c / 6.0; const double u = 1.0 - x * x / (c * c); return c * c / 6.0 * (1.0 - u * u * u); }
__device__ __forceinline__ double w_tukey(double x, double c) { if (fabs(x) > c) return 0.0; const double u = 1.0 - x * x / (c * c); return u * u; }
__device__ __forceinline__ double rho_huber(double a, double d) { return fabs(a) <= d ? a * a / 2 : d * fabs(a) - d * d / 2; }
__device__ __forceinline__ double w_huber(double a, double d) { return fabs(a) <= d ? 1.0 : d / fabs(a); }

// validity + node weights of every vertex (warp_field.cpp:238-241), once per call: the node positions never move
__global__ void __launch_bounds__(256) f2_prepare_kernel(const float *__restrict__ nodes, const float *__restrict__ canon, const float *__restrict__ live,
                                                        int N, int stride, F2Ws ws)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
    const bool ok = !(isnan(c[0]) || isnan(c[1]) || isnan(c[2]) || isnan(l[0]) || isnan(l[1]) || isnan(l[2])) && ws.idx[(size_t)v * 8 + 7] >= 0;
    ws.valid[v] = ok;
    for (int k = 0; k < 8; ++k) {
        const int n = ws.idx[(size_t)v * 8 + k];
        ws.w[(size_t)v * 8 + k] = (ok && n >= 0) ? node_weighting(ws.d2[(size_t)v * 8 + k], __ldg(nodes + (size_t)n * DF_NODE_STRIDE + 11)) : 0.f;
    }
    if (ok) atomicAdd(ws.scal + 6, 1.0);
}

// edges: the reg_k nearest OTHER nodes of every node, from the 8-NN of the node positions (ties to the lower index)
__global__ void __launch_bounds__(256) f2_edges_kernel(int M, int reg_k, F2Ws ws)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int got = 0;
    for (int k = 0; k < 8 && got < reg_k; ++k) {
        const int j = ws.knn_nodes[i * 8 + k];
        if (j < 0 || j == i) continue;
        ws.edge_j[(size_t)i * reg_k + got++] = j;
    }
    for (; got < reg_k; ++got) ws.edge_j[(size_t)i * reg_k + got] = -1;
}

__global__ void __launch_bounds__(256) f2_state_kernel(const float *__restrict__ nodes, int M, F2Ws ws)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float *n = nodes + (size_t)m * DF_NODE_STRIDE;
    Qd q = {n[3], n[4], n[5], n[6]};
    const double nn = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    q.w /= nn; q.x /= nn; q.y /= nn; q.z /= nn;
    const Qd t = qdmul(Qd{2.0 * n[7], 2.0 * n[8], 2.0 * n[9], 2.0 * n[10]}, Qd{q.w, -q.x, -q.y, -q.z});   // getTranslation: 2 * dual * conj(rot)
    ws.Q[4 * m] = q.w; ws.Q[4 * m + 1] = q.x; ws.Q[4 * m + 2] = q.y; ws.Q[4 * m + 3] = q.z;
    ws.T[3 * m] = t.x; ws.T[3 * m + 1] = t.y; ws.T[3 * m + 2] = t.z;
    for (int a = 0; a < 6; ++a) { ws.g[6 * m + a] = 0.0; }
    for (int a = 0; a < 36; ++a) ws.D[36 * (size_t)m + a] = 0.0;
}

struct F2Run { double tukey_c, huber_delta, lambda, mu; int twist, rob_d, rob_r, reg_k, assemble; };

__global__ void __launch_bounds__(128) f2_linearize_kernel(const float *__restrict__ canon, const float *__restrict__ live, int N, int stride, F2Ws ws, F2Run run)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double e_local = 0.0;
    if (v < N && ws.valid[v]) {
        const float *c = canon + (size_t)v * stride, *l = live + (size_t)v * stride;
        const double p[3] = {c[0], c[1], c[2]};
        int nk[8]; double wk[8];
        Qd Qs = {0, 0, 0, 0};
        double ts[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            nk[k] = ws.idx[(size_t)v * 8 + k]; wk[k] = ws.w[(size_t)v * 8 + k];
            const double *q = ws.Q + 4 * nk[k], *t = ws.T + 3 * nk[k];
            Qs.w += wk[k] * q[0]; Qs.x += wk[k] * q[1]; Qs.y += wk[k] * q[2]; Qs.z += wk[k] * q[3];
            ts[0] += wk[k] * t[0]; ts[1] += wk[k] * t[1]; ts[2] += wk[k] * t[2];
        }
        const double nQ = sqrt(Qs.w * Qs.w + Qs.x * Qs.x + Qs.y * Qs.y + Qs.z * Qs.z);
        const Qd qh = {Qs.w / nQ, Qs.x / nQ, Qs.y / nQ, Qs.z / nQ};
        double y[3];
        qdrot(qh, p, y);
        double r[3], W[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            r[i] = (double)l[i] - (y[i] + ts[i]);
            W[i] = run.rob_d ? w_tukey(r[i], run.tukey_c) : 1.0;
            e_local += run.rob_d ? rho_tukey(r[i], run.tukey_c) : 0.5 * r[i] * r[i];
            ws.Wr[(size_t)v * 3 + i] = W[i];
        }
        if (run.assemble) {
            const Qd qc = {qh.w, -qh.x, -qh.y, -qh.z};
            for (int k = 0; k < 8; ++k) {
                // J = [ Jw (3x3) , w_k I ]: Jw column a = B_a x y, (0, B_a) = (w_k/|Q|) P[(0, e_a) q_k] qhat*   (oracle/orc_reg.c blend_B)
                double J[3][6];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int a = 0; a < 6; ++a) J[i][a] = 0.0;
                if (run.twist) {
                    const double *qk = ws.Q + 4 * nk[k];
                    const Qd q = {qk[0], qk[1], qk[2], qk[3]};
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const Qd e = {0.0, a == 0 ? 1.0 : 0.0, a == 1 ? 1.0 : 0.0, a == 2 ? 1.0 : 0.0};
                        Qd dq = qdmul(e, q);
                        const double dot = qh.w * dq.w + qh.x * dq.x + qh.y * dq.y + qh.z * dq.z, s = wk[k] / nQ;
                        dq.w = (dq.w - qh.w * dot) * s; dq.x = (dq.x - qh.x * dot) * s; dq.y = (dq.y - qh.y * dot) * s; dq.z = (dq.z - qh.z * dot) * s;
                        const Qd o = qdmul(dq, qc);
                        J[0][a] = o.y * y[2] - o.z * y[1];
                        J[1][a] = o.z * y[0] - o.x * y[2];
                        J[2][a] = o.x * y[1] - o.y * y[0];
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) J[i][3 + i] = wk[k];
                float *jw = ws.Jw + (size_t)v * 72 + k * 9;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int a = 0; a < 3; ++a) jw[i * 3 + a] = (float)J[i][a];
                // the stored (float) Jacobian is the one the PCG applies: use it for the gradient and the diagonal block too
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int a = 0; a < 3; ++a) J[i][a] = (double)jw[i * 3 + a];
                double *gn = ws.g + 6 * nk[k], *Dn = ws.D + 36 * (size_t)nk[k];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    double ga = 0.0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) ga += J[i][a] * W[i] * r[i];
                    if (ga != 0.0) atomicAdd(gn + a, ga);
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        double h = 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) h += J[i][a] * W[i] * J[i][b];
                        if (h != 0.0) atomicAdd(Dn + a * 6 + b, h);
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) e_local += __shfl_xor_sync(0xffffffffu, e_local, o);
    if ((threadIdx.x & 31) == 0 && e_local != 0.0) atomicAdd(ws.scal + 4, e_local);
}

__global__ void __launch_bounds__(128) f2_reg_kernel(const float *__restrict__ nodes, int M, F2Ws ws, F2Run run)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double e_local = 0.0;
    if (s < M * run.reg_k) {
        const int i = s / run.reg_k, j = ws.edge_j[s];
        if (j >= 0) {
            const double gj[3] = {nodes[(size_t)j * DF_NODE_STRIDE], nodes[(size_t)j * DF_NODE_STRIDE + 1], nodes[(size_t)j * DF_NODE_STRIDE + 2]};
            const Qd qi = {ws.Q[4 * i], ws.Q[4 * i + 1], ws.Q[4 * i + 2], ws.Q[4 * i + 3]}, qj = {ws.Q[4 * j], ws.Q[4 * j + 1], ws.Q[4 * j + 2], ws.Q[4 * j + 3]};
            double yi[3], yj[3], d[3], Wd[3];
            qdrot(qi, gj, yi); qdrot(qj, gj, yj);
            const double alpha = fmax((double)nodes[(size_t)i * DF_NODE_STRIDE + 11], (double)nodes[(size_t)j * DF_NODE_STRIDE + 11]) * run.lambda;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c] = (yi[c] + ws.T[3 * i + c]) - (yj[c] + ws.T[3 * j + c]);
                Wd[c] = alpha * (run.rob_r ? w_huber(d[c], run.huber_delta) : 1.0);
                e_local += alpha * (run.rob_r ? rho_huber(d[c], run.huber_delta) : 0.5 * d[c] * d[c]);
                ws.edge_W[(size_t)s * 3 + c] = Wd[c];
                ws.edge_y[(size_t)s * 6 + c] = yi[c]; ws.edge_y[(size_t)s * 6 + 3 + c] = yj[c];
            }
            if (run.assemble) {
                // d(eps) = d + Ji eps_i + Jj eps_j, Ji = [ -[yi]x , I ], Jj = [ +[yj]x , -I ]
                double Je[2][3][6];
#pragma unroll
                for (int sd = 0; sd < 2; ++sd)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int a = 0; a < 6; ++a) Je[sd][c][a] = 0.0;
                if (run.twist) {
                    Je[0][0][1] = yi[2]; Je[0][0][2] = -yi[1]; Je[0][1][0] = -yi[2]; Je[0][1][2] = yi[0]; Je[0][2][0] = yi[1]; Je[0][2][1] = -yi[0];
                    Je[1][0][1] = -yj[2]; Je[1][0][2] = yj[1]; Je[1][1][0] = yj[2]; Je[1][1][2] = -yj[0]; Je[1][2][0] = -yj[1]; Je[1][2][1] = yj[0];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { Je[0][c][3 + c] = 1.0; Je[1][c][3 + c] = -1.0; }
                const int nn[2] = {i, j};
#pragma unroll
                for (int sd = 0; sd < 2; ++sd)
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double ga = 0.0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) ga += Je[sd][c][a] * Wd[c] * d[c];
                        if (ga != 0.0) atomicAdd(ws.g + 6 * nn[sd] + a, -ga);
#pragma unroll
                        for (int b = 0; b < 6; ++b) {
                            double h = 0.0;
#pragma unroll
                            for (int c = 0; c < 3; ++c) h += Je[sd][c][a] * Wd[c] * Je[sd][c][b];
                            if (h != 0.0) atomicAdd(ws.D + 36 * (size_t)nn[sd] + a * 6 + b, h);
                        }
                    }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) e_local += __shfl_xor_sync(0xffffffffu, e_local, o);
    if ((threadIdx.x & 31) == 0 && e_local != 0.0) atomicAdd(ws.scal + 5, e_local);
}

// per node: Levenberg damping of the diagonal, pin of the rotation increments in a translation-only solve, inverse of the 6 x 6 block
__global__ void __launch_bounds__(128) f2_blocks_kernel(int M, F2Ws ws, F2Run run)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double A[36];
    for (int a = 0; a < 36; ++a) A[a] = ws.D[36 * (size_t)m + a];
    for (int a = 0; a < 6; ++a) {
        double dmp;
        if (!run.twist && a < 3) {
            for (int b = 0; b < 6; ++b) A[a * 6 + b] = A[b * 6 + a] = 0.0;
            dmp = 1.0;                                            // unit diagonal, zero gradient: the increment stays zero
            ws.g[6 * m + a] = 0.0;
        } else dmp = run.mu * A[a * 6 + a] + 1e-12;
        ws.damp[6 * m + a] = dmp;
        A[a * 6 + a] += dmp;
    }
    // inverse by Cholesky: A = L L^T, solve for the six unit vectors
    double L[36];
    for (int a = 0; a < 36; ++a) L[a] = 0.0;
    bool ok = true;
    for (int j = 0; j < 6; ++j) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        d = sqrt(d);
        L[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / d;
        }
    }
    for (int c = 0; c < 6; ++c) {
        double yv[6], xv[6];
        for (int i = 0; i < 6; ++i) { double s = i == c ? 1.0 : 0.0; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * yv[k]; yv[i] = s / L[i * 6 + i]; }
        for (int i = 5; i >= 0; --i) { double s = yv[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * xv[k]; xv[i] = s / L[i * 6 + i]; }
        for (int i = 0; i < 6; ++i) ws.Dinv[36 * (size_t)m + i * 6 + c] = ok ? xv[i] : (i == c ? 1.0 / A[c * 6 + c] : 0.0);
    }
}

__device__ __forceinline__ double f2_block_sum(double v, double *sm)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}

__device__ __forceinline__ void f2_precond(const double *Dinv, const double *r, double *z, int m)
{
    const double *B = Dinv + 36 * (size_t)m;
    for (int a = 0; a < 6; ++a) {
        double s = 0.0;
        for (int b = 0; b < 6; ++b) s += B[a * 6 + b] * r[6 * m + b];
        z[6 * m + a] = s;
    }
}

// PCG start (ONE block): x = 0, r = g, z = M^-1 r, p = z, Ap = 0
__global__ void __launch_bounds__(1024) f2_pcg_init_kernel(int M, F2Ws ws)
{
    __shared__ double sm[40];
    double part = 0.0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        for (int a = 0; a < 6; ++a) { ws.x[6 * m + a] = 0.0; ws.r[6 * m + a] = ws.g[6 * m + a]; ws.Ap[6 * m + a] = 0.0; }
        f2_precond(ws.Dinv, ws.r, ws.z, m);
        for (int a = 0; a < 6; ++a) { ws.p[6 * m + a] = ws.z[6 * m + a]; part += ws.r[6 * m + a] * ws.z[6 * m + a]; }
    }
    const double rz = f2_block_sum(part, sm);
    if (threadIdx.x == 0) { ws.scal[0] = rz; ws.scal[1] = rz; ws.scal[2] = rz > 0.0 ? 0.0 : 1.0; }
}

__global__ void __launch_bounds__(128) f2_apply_data_kernel(int N, F2Ws ws)
{
    if (ws.scal[2] != 0.0) return;                                 // converged: the remaining launches are no-ops
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N || !ws.valid[v]) return;
    int nk[8]; double wk[8];
    double t[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        nk[k] = ws.idx[(size_t)v * 8 + k]; wk[k] = ws.w[(size_t)v * 8 + k];
        const double *pk = ws.p + 6 * nk[k];
        const float *jw = ws.Jw + (size_t)v * 72 + k * 9;
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] += (double)jw[i * 3] * pk[0] + (double)jw[i * 3 + 1] * pk[1] + (double)jw[i * 3 + 2] * pk[2] + wk[k] * pk[3 + i];
    }
    const double u[3] = {ws.Wr[(size_t)v * 3] * t[0], ws.Wr[(size_t)v * 3 + 1] * t[1], ws.Wr[(size_t)v * 3 + 2] * t[2]};
    if (u[0] == 0.0 && u[1] == 0.0 && u[2] == 0.0) return;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float *jw = ws.Jw + (size_t)v * 72 + k * 9;
        double *out = ws.Ap + 6 * nk[k];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double s = (double)jw[a] * u[0] + (double)jw[3 + a] * u[1] + (double)jw[6 + a] * u[2];
            if (s != 0.0) atomicAdd(out + a, s);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) atomicAdd(out + 3 + i, wk[k] * u[i]);
    }
}

__global__ void __launch_bounds__(128) f2_apply_reg_kernel(int M, F2Ws ws, F2Run run)
{
    if (ws.scal[2] != 0.0) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M * run.reg_k) return;
    const int i = s / run.reg_k, j = ws.edge_j[s];
    if (j < 0) return;
    const double *yi = ws.edge_y + (size_t)s * 6, *yj = yi + 3, *W = ws.edge_W + (size_t)s * 3;
    const double *pi = ws.p + 6 * i, *pj = ws.p + 6 * j;
    // d_lin = omega_i x yi + tau_i - omega_j x yj - tau_j
    double d[3] = {pi[3] - pj[3], pi[4] - pj[4], pi[5] - pj[5]};
    if (run.twist) {
        d[0] += (pi[1] * yi[2] - pi[2] * yi[1]) - (pj[1] * yj[2] - pj[2] * yj[1]);
        d[1] += (pi[2] * yi[0] - pi[0] * yi[2]) - (pj[2] * yj[0] - pj[0] * yj[2]);
        d[2] += (pi[0] * yi[1] - pi[1] * yi[0]) - (pj[0] * yj[1] - pj[1] * yj[0]);
    }
    const double u[3] = {W[0] * d[0], W[1] * d[1], W[2] * d[2]};
    double *oi = ws.Ap + 6 * i, *oj = ws.Ap + 6 * j;
    if (run.twist) {                                               // J_i^T u = yi x u (rotation part), -(yj x u) for node j
        atomicAdd(oi + 0, yi[1] * u[2] - yi[2] * u[1]); atomicAdd(oi + 1, yi[2] * u[0] - yi[0] * u[2]); atomicAdd(oi + 2, yi[0] * u[1] - yi[1] * u[0]);
        atomicAdd(oj + 0, -(yj[1] * u[2] - yj[2] * u[1])); atomicAdd(oj + 1, -(yj[2] * u[0] - yj[0] * u[2])); atomicAdd(oj + 2, -(yj[0] * u[1] - yj[1] * u[0]));
    }
    for (int c = 0; c < 3; ++c) { atomicAdd(oi + 3 + c, u[c]); atomicAdd(oj + 3 + c, -u[c]); }
}

// one PCG step's scalar + vector work for all 6M entries (ONE block): Ap += damp p; alpha; x, r; z = M^-1 r; beta; p; Ap <- 0
__global__ void __launch_bounds__(1024) f2_pcg_update_kernel(int M, F2Ws ws, double tol2)
{
    __shared__ double sm[40];
    if (ws.scal[2] != 0.0) return;
    const double rz = ws.scal[0];
    double part = 0.0;
    for (int i = threadIdx.x; i < 6 * M; i += blockDim.x) {
        const double ap = ws.Ap[i] + ws.damp[i] * ws.p[i];
        ws.Ap[i] = ap;
        part += ws.p[i] * ap;
    }
    const double pAp = f2_block_sum(part, sm);
    if (!(pAp > 0.0)) { if (threadIdx.x == 0) ws.scal[2] = 1.0; return; }
    const double alpha = rz / pAp;
    for (int i = threadIdx.x; i < 6 * M; i += blockDim.x) { ws.x[i] += alpha * ws.p[i]; ws.r[i] -= alpha * ws.Ap[i]; }
    __syncthreads();
    part = 0.0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        f2_precond(ws.Dinv, ws.r, ws.z, m);
        for (int a = 0; a < 6; ++a) part += ws.r[6 * m + a] * ws.z[6 * m + a];
    }
    const double rz_new = f2_block_sum(part, sm);
    const double beta = rz_new / rz;
    for (int i = threadIdx.x; i < 6 * M; i += blockDim.x) { ws.p[i] = ws.z[i] + beta * ws.p[i]; ws.Ap[i] = 0.0; }
    if (threadIdx.x == 0) {
        ws.scal[0] = rz_new;
        ws.scal[3] += 1.0;
        if (!(rz_new > tol2 * ws.scal[1])) ws.scal[2] = 1.0;
    }
}

// q_k <- exp(omega_k) q_k, t_k <- t_k + tau_k, node re-encoded as DualQuaternion(t, r) (dual part = 1/2 (0, t) r)
__global__ void __launch_bounds__(256) f2_update_nodes_kernel(float *nodes, int M, F2Ws ws, F2Run run)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double *dl = ws.x + 6 * m;
    Qd q = {ws.Q[4 * m], ws.Q[4 * m + 1], ws.Q[4 * m + 2], ws.Q[4 * m + 3]};
    const double th = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    if (run.twist && th > 0.0) {
        const double s = sin(th / 2) / th;
        const Qd qn = qdmul(Qd{cos(th / 2), s * dl[0], s * dl[1], s * dl[2]}, q);
        const double nn = sqrt(qn.w * qn.w + qn.x * qn.x + qn.y * qn.y + qn.z * qn.z);
        q = Qd{qn.w / nn, qn.x / nn, qn.y / nn, qn.z / nn};
    }
    const double t[3] = {ws.T[3 * m] + dl[3], ws.T[3 * m + 1] + dl[4], ws.T[3 * m + 2] + dl[5]};
    const Qd dual = qdmul(Qd{0.0, 0.5 * t[0], 0.5 * t[1], 0.5 * t[2]}, q);
    float *nd = nodes + (size_t)m * DF_NODE_STRIDE;
    nd[3] = (float)q.w; nd[4] = (float)q.x; nd[5] = (float)q.y; nd[6] = (float)q.z;
    nd[7] = (float)dual.w; nd[8] = (float)dual.x; nd[9] = (float)dual.y; nd[10] = (float)dual.z;
}

__global__ void f2_record_kernel(F2Ws ws, double *stats, int slot, int final_pass, int edges)
{
    const double e = ws.scal[4] + ws.scal[5];
    if (slot == 0) stats[0] = e;
    if (slot < 8) stats[8 + slot] = e;
    if (final_pass) { stats[1] = e; stats[2] = (double)slot; stats[3] = ws.scal[6]; stats[4] = ws.scal[4]; stats[5] = ws.scal[5]; stats[6] = (double)edges; stats[7] = ws.scal[3]; }
    ws.scal[4] = 0.0; ws.scal[5] = 0.0;
}

}  // namespace

extern "C" size_t df_solve_f2_workspace_bytes(int M, int N, int reg_k)
{
    F2Ws ws;
    return f2_layout(ws, nullptr, M, N, reg_k < 0 ? 0 : (reg_k > 7 ? 7 : reg_k)) + 256;
}

extern "C" int df_solve_f2(float *nodes, int M, const void *node_grid, const float *canon, const float *live, int N, int stride,
                           const df_f2_params *prm, double *stats_dev, void *workspace, void *stream)
{
    if (M < 8 || N <= 0 || !prm || !workspace) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    const int reg_k = prm->reg_k < 0 ? 0 : (prm->reg_k > 7 ? 7 : prm->reg_k);
    F2Ws ws;
    char *base = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    f2_layout(ws, base, M, N, reg_k);
    F2Run run;
    run.tukey_c = prm->tukey_c; run.huber_delta = prm->huber_delta; run.lambda = prm->reg_lambda; run.mu = prm->lm_mu;
    run.twist = prm->flags & DF_F2_TWIST ? 1 : 0; run.rob_d = prm->flags & DF_F2_TUKEY ? 1 : 0; run.rob_r = prm->flags & DF_F2_HUBER ? 1 : 0;
    run.reg_k = (prm->reg_lambda > 0.0) ? reg_k : 0;
    run.assemble = 1;
    cudaError_t e = cudaMemsetAsync(ws.scal, 0, 64 * 8, s);
    if (e != cudaSuccess) return (int)e;
    int st = df_knn8(nodes, M, node_grid, canon, N, stride, ws.idx, ws.d2, s);
    if (st) return st;
    f2_prepare_kernel<<<div_up(N, 256), 256, 0, s>>>(nodes, canon, live, N, stride, ws);
    int edges = 0;
    if (run.reg_k) {
        st = df_knn8(nodes, M, node_grid, nodes, M, DF_NODE_STRIDE, ws.knn_nodes, ws.knn_nodes_d2, s);
        if (st) return st;
        f2_edges_kernel<<<div_up(M, 256), 256, 0, s>>>(M, run.reg_k, ws);
        edges = M * run.reg_k;                                     // upper bound reported in stats[6] (slots; -1 entries are skipped)
    }
    const int gn = prm->gn_iters < 0 ? 0 : prm->gn_iters;
    const int lin = prm->lin_iters > 0 ? prm->lin_iters : 100;
    for (int it = 0; it <= gn; ++it) {
        run.assemble = it < gn;
        f2_state_kernel<<<div_up(M, 256), 256, 0, s>>>(nodes, M, ws);
        f2_linearize_kernel<<<div_up(N, 128), 128, 0, s>>>(canon, live, N, stride, ws, run);
        if (run.reg_k) f2_reg_kernel<<<div_up(M * run.reg_k, 128), 128, 0, s>>>(nodes, M, ws, run);
        f2_record_kernel<<<1, 1, 0, s>>>(ws, stats_dev, it, it == gn, edges);
        if (it == gn) break;
        f2_blocks_kernel<<<div_up(M, 128), 128, 0, s>>>(M, ws, run);
        f2_pcg_init_kernel<<<1, 1024, 0, s>>>(M, ws);
        for (int l = 0; l < lin; ++l) {
            f2_apply_data_kernel<<<div_up(N, 128), 128, 0, s>>>(N, ws);
            if (run.reg_k) f2_apply_reg_kernel<<<div_up(M * run.reg_k, 128), 128, 0, s>>>(M, ws, run);
            f2_pcg_update_kernel<<<1, 1024, 0, s>>>(M, ws, 1e-24);
        }
        f2_update_nodes_kernel<<<div_up(M, 256), 256, 0, s>>>(nodes, M, ws, run);
    }
    DF_LAUNCH_CHECK();
    return 0;
}
