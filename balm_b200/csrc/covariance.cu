// covariance.cu -- pose-covariance propagation of the consistency experiment:
//
//   Rcov_raw = sum over observations (a, j) of  Ls c_cov_j Ls^T ,  Ls = d(gradient)/d(9 parameters of cluster (a, j))
//   Rcov     = H^-1 Rcov_raw H^-T
// = VOX_HESS::left_jacobian_point + BALM2::multi_second + the last lines of BALM2::damping_iter
//   (/root/reference/src/simulation/BAs_left.hpp:342-473, 995-1023, 1089-1096), with c_cov as PointCluster::push
//   accumulates it under isotropic point noise (/root/reference/src/simulation/toolss.hpp:311-343) or given per observation.
//
// The reference forms a dense 6N x 9 matrix Ls per observation and adds the 6N x 6N product Ls c Ls^T. Written out, the
// block of Ls for pose p is (c2 = coe 2/NN)
//     Ls_j[p] = c2 ( Q_p Z_j + [j == p] E_p ),   Q_p = [A_p | a_p] (6x4),  Z_j = [G_j ; -h_j^T/NN] (4x9),  E_p (6x9)
// (A_p = g2_combos[p], a_p = (U_l TC_p)[:,3] -- the same vector as in the Hessian, G_j as in the reference), hence per voxel
//     sum_j Ls_j c_j Ls_j^T [p,q] = c2^2 ( Q_p Wz Q_q^T + Q_p Y_q + Y_p^T Q_q^T + [p == q] D_p ),
//     Wz = sum_j Z_j c_j Z_j^T (4x4),  Y_q = Z_q c_q E_q^T (4x6),  D_p = E_p c_p E_p^T (6x6)
//   = c2^2 ( Q V^T + V Q^T )[p,q] + [p == q] c2^2 D_p ,   V_p = Q_p Wz/2 + Y_p^T .
// So the whole sum is a symmetric rank-8M product of the same shape as the Hessian accumulation: eight rows per voxel
// (four of c2 Q^T, four of c2 V^T, 6N columns each), accumulated by the fp64 DMMA SYRK kernel with the A operand taken
// from the partner row (row ^ 4), plus 6x6 diagonal blocks -- O(M N^2) tensor-core work instead of the reference's
// O(K N^2) dense rank-9 updates. H^-1 (.) H^-T reuses the LDL^T factors of the solve with n right-hand sides.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "internal.cuh"

#define CTRY(x) do { int _s = (x); if (_s != BALM_OK) return _s; } while (0)

int launch_syrk_f64_on(balm_ctx *c, const double *G, int64_t rows, bool first_batch, int row_xor);

namespace {

// ---- small dense helpers on local arrays ----
__device__ __forceinline__ void hat3(const double *v, double *M /*3x3 row-major*/) {
  M[0] = 0; M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2]; M[4] = 0; M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}
// g1(w): 4 x 9 (BAs_left.hpp:321-331), row-major
__device__ __forceinline__ void g1_mat(const double *w, double *X /*[4][9]*/) {
#pragma unroll
  for (int e = 0; e < 36; e++) X[e] = 0.0;
  X[0] = w[0]; X[1] = w[1]; X[2] = w[2]; X[6] = w[3];
  X[9 + 1] = w[0]; X[9 + 3] = w[1]; X[9 + 4] = w[2]; X[9 + 7] = w[3];
  X[18 + 2] = w[0]; X[18 + 4] = w[1]; X[18 + 5] = w[2]; X[18 + 8] = w[3];
  X[27 + 6] = w[0]; X[27 + 7] = w[1]; X[27 + 8] = w[2];
}

// c_cov (9x9, row-major) of a cluster under isotropic point noise, from its moments (toolss.hpp:311-343):
// Bi = x Ex + y Ey + z Ez  ->  sum Bf Bf^T = [[sum_kl P_kl E_k E_l^T, sum_k v_k E_k], [sym, N I]]
__device__ __forceinline__ void cluster_cov_iso(const double *o, double pn2, double *C /*[81]*/) {
  const double P[3][3] = {{o[0], o[1], o[2]}, {o[1], o[3], o[4]}, {o[2], o[4], o[5]}};
  // E_k as (row, col, value) triplets: Ex: (0,0,2) (1,1,1) (2,2,1); Ey: (1,0,1) (3,1,2) (4,2,1); Ez: (2,0,1) (4,1,1) (5,2,2)
  const int Er[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  const double Ev[3][3] = {{2, 1, 1}, {1, 2, 1}, {1, 1, 2}};
#pragma unroll
  for (int e = 0; e < 81; e++) C[e] = 0.0;
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++)
      for (int cc = 0; cc < 3; cc++)  // (E_k E_l^T)[r1][r2] = sum_c E_k[r1][c] E_l[r2][c]; column c holds one entry per E
        C[Er[k][cc] * 9 + Er[l][cc]] += P[k][l] * Ev[k][cc] * Ev[l][cc];
  for (int k = 0; k < 3; k++)
    for (int cc = 0; cc < 3; cc++) {
      C[Er[k][cc] * 9 + 6 + cc] += o[6 + k] * Ev[k][cc];
      C[(6 + cc) * 9 + Er[k][cc]] += o[6 + k] * Ev[k][cc];
    }
  for (int d = 0; d < 3; d++) C[(6 + d) * 9 + 6 + d] = o[9];
#pragma unroll
  for (int e = 0; e < 81; e++) C[e] *= pn2;
}

struct CovArgs {
  const double *obs;
  int64_t Kp;
  const int *pose_idx;
  const long long *row_ptr;
  const double *poses;
  const double *stats;   // [nv][20] of the batch (v0-based)
  const double *ccov;    // [K][81] row-major or null
  double pn2;
  int64_t v0, v1;
  int N, ldg;
  double *Wz;            // [nv][10]
  double *G;             // [8*nv][ldg]
  double *Dobs;          // [K][21] (global observation index)
};

// Everything one observation contributes; shared by both passes.
struct ObsTerms {
  double Z[4][9];   // [G_j ; -h_j^T / NN]
  double E[6][9];
  double Q[6][4];   // [A_p | a_p]
};

__device__ __forceinline__ void obs_terms(const double *o, const double *pose12, const double *st, ObsTerms &t) {
  double r[9], p[3];
  load_pose(pose12, r, p);  // row-major R
  const double *vb = st, *u0 = st + 3, *u1 = st + 6, *u2 = st + 9;
  const double inv = st[12];
  const double lam0 = st[17], lam1 = st[18], lam2 = st[19];
  const WC w = world_cluster(o, r, p);
  // m = P' - v' vb^T, mb = v' - n vb  (= TC [R | p - vb]^T, as in the observation sweep)
  const double m[3][3] = {{w.p00 - w.v0 * vb[0], w.p01 - w.v0 * vb[1], w.p02 - w.v0 * vb[2]},
                          {w.p01 - w.v1 * vb[0], w.p11 - w.v1 * vb[1], w.p12 - w.v1 * vb[2]},
                          {w.p02 - w.v2 * vb[0], w.p12 - w.v2 * vb[1], w.p22 - w.v2 * vb[2]}};
  const double mb[3] = {w.v0 - w.n * vb[0], w.v1 - w.n * vb[1], w.v2 - w.n * vb[2]};
  double t0[3], s0 = 0.0;
  for (int a = 0; a < 3; a++) { t0[a] = m[a][0] * u0[0] + m[a][1] * u0[1] + m[a][2] * u0[2]; s0 += mb[a] * u0[a]; }
  // A_p = [hat(t0) - hat(u0) m ; s0 I + u0 mb^T]
  double Ht[9], Hu[9];
  hat3(t0, Ht);
  hat3(u0, Hu);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      t.Q[a][b] = Ht[a * 3 + b] - (Hu[a * 3] * m[0][b] + Hu[a * 3 + 1] * m[1][b] + Hu[a * 3 + 2] * m[2][b]);
      t.Q[3 + a][b] = (a == b ? s0 : 0.0) + u0[a] * mb[b];
    }
  // a_p = [-u0 x v' ; n u0]
  const double vw[3] = {w.v0, w.v1, w.v2};
  double cuv[3];
  cross3(u0, vw, cuv);
  for (int a = 0; a < 3; a++) { t.Q[a][3] = -cuv[a]; t.Q[3 + a][3] = w.n * u0[a]; }
  // ub = R^T u0, w4 = [ub ; p.u0], w5 = [ub ; (p - vb).u0]
  double ub[3];
  for (int a = 0; a < 3; a++) ub[a] = r[a] * u0[0] + r[3 + a] * u0[1] + r[6 + a] * u0[2];
  const double pu = p[0] * u0[0] + p[1] * u0[1] + p[2] * u0[2];
  const double vu = vb[0] * u0[0] + vb[1] * u0[1] + vb[2] * u0[2];
  const double w4[4] = {ub[0], ub[1], ub[2], pu}, w5[4] = {ub[0], ub[1], ub[2], pu - vu};
  double X[36];
  // B_j = [R | p - vb] g1(w4) - (vb.u0) [0 | R]   (3 x 9)
  g1_mat(w4, X);
  double B[3][9];
  for (int a = 0; a < 3; a++)
    for (int e = 0; e < 9; e++)
      B[a][e] = r[a * 3] * X[e] + r[a * 3 + 1] * X[9 + e] + r[a * 3 + 2] * X[18 + e] + (p[a] - vb[a]) * X[27 + e];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) B[a][6 + b] -= vu * r[a * 3 + b];
  // G_j = Mk B_j / NN,  Mk = sum_{k=1,2} u_k u_k^T / (lam0 - lam_k)
  const double f1 = 1.0 / (lam0 - lam1), f2 = 1.0 / (lam0 - lam2);
  for (int e = 0; e < 9; e++) {
    const double d1 = u1[0] * B[0][e] + u1[1] * B[1][e] + u1[2] * B[2][e];
    const double d2 = u2[0] * B[0][e] + u2[1] * B[1][e] + u2[2] * B[2][e];
    for (int a = 0; a < 3; a++) t.Z[a][e] = (f1 * u1[a] * d1 + f2 * u2[a] * d2) * inv;
    t.Z[3][e] = 0.0;
  }
  for (int b = 0; b < 3; b++) t.Z[3][6 + b] = -ub[b] * inv;  // -h_j^T / NN, h_j = [0_6, R^T u0]
  // E_p = U_l T g1(w5): top = -hat(u0) (R X03 + p X3), bottom = u0 X3
  g1_mat(w5, X);
  for (int e = 0; e < 9; e++) {
    double y[3];
    for (int a = 0; a < 3; a++) y[a] = r[a * 3] * X[e] + r[a * 3 + 1] * X[9 + e] + r[a * 3 + 2] * X[18 + e] + p[a] * X[27 + e];
    double cy[3];
    cross3(u0, y, cy);
    for (int a = 0; a < 3; a++) { t.E[a][e] = -cy[a]; t.E[3 + a][e] = u0[a] * X[27 + e]; }
  }
}

__device__ __forceinline__ void load_obs(const CovArgs &a, long long s, double *o) {
#pragma unroll
  for (int c = 0; c < 10; c++) o[c] = a.obs[c * a.Kp + s];
}
__device__ __forceinline__ void load_ccov(const CovArgs &a, long long s, const double *o, double *C) {
  if (a.ccov) {
    for (int e = 0; e < 81; e++) C[e] = a.ccov[(size_t)s * 81 + e];
  } else {
    cluster_cov_iso(o, a.pn2, C);
  }
}

// pass 1: Wz[v] = sum_j Z_j c_j Z_j^T   (one warp per voxel, fixed-order shuffle reduction)
__global__ void __launch_bounds__(128) cov_wz_kernel(CovArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t v = a.v0 + (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (v >= a.v1) return;
  const long long s0 = a.row_ptr[v], s1 = a.row_ptr[v + 1];
  const double *st = a.stats + (v - a.v0) * BALM_STATS_STRIDE;
  double acc[10];
  for (int e = 0; e < 10; e++) acc[e] = 0.0;
  for (long long s = s0 + lane; s < s1; s += 32) {
    double o[10], C[81];
    load_obs(a, s, o);
    load_ccov(a, s, o, C);
    ObsTerms t;
    obs_terms(o, a.poses + 12 * a.pose_idx[s], st, t);
    double ZC[4][9];
    for (int i = 0; i < 4; i++)
      for (int e = 0; e < 9; e++) {
        double x = 0.0;
        for (int f = 0; f < 9; f++) x += t.Z[i][f] * C[f * 9 + e];
        ZC[i][e] = x;
      }
    int q = 0;
    for (int i = 0; i < 4; i++)
      for (int j = i; j < 4; j++) {
        double x = 0.0;
        for (int e = 0; e < 9; e++) x += ZC[i][e] * t.Z[j][e];
        acc[q++] += x;
      }
  }
  for (int e = 0; e < 10; e++) acc[e] = warp_sum(acc[e]);
  if (lane == 0)
    for (int e = 0; e < 10; e++) a.Wz[(v - a.v0) * 10 + e] = acc[e];
}

// pass 2: the eight rows of a voxel (c2 Q^T | c2 V^T) and the diagonal block c2^2 D_p of every observation
__global__ void __launch_bounds__(128) cov_rows_kernel(CovArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t v = a.v0 + (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (v >= a.v1) return;
  const long long s0 = a.row_ptr[v], s1 = a.row_ptr[v + 1];
  const double *st = a.stats + (v - a.v0) * BALM_STATS_STRIDE;
  const double c2 = st[16] * 2.0 * st[12];  // coe * 2 / NN
  double Wz[4][4];
  {
    const double *wz = a.Wz + (v - a.v0) * 10;
    int q = 0;
    for (int i = 0; i < 4; i++)
      for (int j = i; j < 4; j++) { Wz[i][j] = wz[q]; Wz[j][i] = wz[q]; q++; }
  }
  for (long long s = s0 + lane; s < s1; s += 32) {
    double o[10], C[81];
    load_obs(a, s, o);
    load_ccov(a, s, o, C);
    ObsTerms t;
    const int p = a.pose_idx[s];
    obs_terms(o, a.poses + 12 * p, st, t);
    // EC = E c (6x9);  Y^T = E c Z^T (6x4);  D = E c E^T (6x6)
    double EC[6][9];
    for (int i = 0; i < 6; i++)
      for (int e = 0; e < 9; e++) {
        double x = 0.0;
        for (int f = 0; f < 9; f++) x += t.E[i][f] * C[f * 9 + e];
        EC[i][e] = x;
      }
    double *g = a.G + (size_t)(8 * (v - a.v0)) * a.ldg + 6 * p;
    for (int i = 0; i < 6; i++) {
      for (int mcol = 0; mcol < 4; mcol++) {
        double yt = 0.0;
        for (int e = 0; e < 9; e++) yt += EC[i][e] * t.Z[mcol][e];
        double qw = 0.0;
        for (int f = 0; f < 4; f++) qw += t.Q[i][f] * Wz[f][mcol];
        g[(size_t)mcol * a.ldg + i] = c2 * t.Q[i][mcol];
        g[(size_t)(4 + mcol) * a.ldg + i] = c2 * (0.5 * qw + yt);
      }
    }
    double *d = a.Dobs + (size_t)s * 21;
    int q = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        double x = 0.0;
        for (int e = 0; e < 9; e++) x += EC[i][e] * t.E[j][e];
        d[q++] = c2 * c2 * x;
      }
  }
}

// per pose: sum of its observations' diagonal blocks in ascending voxel order (dense: s = v N + p; sparse: the lists)
__global__ void cov_diag_reduce_kernel(const double *Dobs, int N, int64_t M, int dense, const int *csc_ptr,
                                       const int *csc_obs, double *Dacc /*[21][N]*/, int add, int64_t v0, int64_t v1,
                                       const int *csc_lo, const int *csc_hi) {
  const int p = blockIdx.x, q = threadIdx.x;
  if (q >= 21) return;
  double s = 0.0;
  if (dense) {
    for (int64_t v = v0; v < v1; v++) s += Dobs[((size_t)v * N + p) * 21 + q];
  } else {
    for (int t = csc_lo[p]; t < csc_hi[p]; t++) s += Dobs[(size_t)csc_obs[t] * 21 + q];
  }
  Dacc[(size_t)q * N + p] = add ? Dacc[(size_t)q * N + p] + s : s;
}

// Rcov_raw = + sum_splits S + blockdiag(D), mirrored (the Hessian assembly with the opposite sign and no gradient)
__global__ void __launch_bounds__(1024) cov_assemble_kernel(const double *part, int splits, int tiles, int nb, const double *Dacc,
                                                            int N, int n, double *R) {
  constexpr int TILE = BALM_SYRK_TILE;
  const int bx = blockIdx.x, by = blockIdx.y;
  if (by < bx) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r = bx * 32 + ty, c = by * 32 + tx;
  if (r >= n || c >= n) return;
  const int rr = r < c ? r : c, cc = r < c ? c : r;
  const int bi = rr / TILE, bj = cc / TILE;
  const int t = bi * nb - bi * (bi - 1) / 2 + (bj - bi);
  const size_t off = (size_t)t * (TILE * TILE) + (size_t)(rr % TILE) * TILE + (cc % TILE);
  double s = 0.0;
  for (int sp = 0; sp < splits; sp++) s += part[(size_t)sp * tiles * (TILE * TILE) + off];
  if (rr / 6 == cc / 6) {
    const int r6 = rr % 6, c6 = cc % 6;
    const int q = r6 * 6 - r6 * (r6 - 1) / 2 + (c6 - r6);
    s += Dacc[(size_t)q * N + rr / 6];
  }
  R[(size_t)r * n + c] = s;
  R[(size_t)c * n + r] = s;
}

// ---- n right-hand sides through the LDL^T factors of the solve (A: L below the diagonal, Xinv: L_jj^-1, dinv) ----
constexpr int NB = BALM_NB;
constexpr int MRHS_SMEM = 2 * NB * (NB + 1) * (int)sizeof(double);
// B[j0 .. j0+nbw) rows, column tile blockIdx.x*64:  B <- X B  (TRANS = 0) or  B <- X^T B  (TRANS = 1)
template <int TRANS>
__global__ void __launch_bounds__(256) mrhs_diag_kernel(const double *Xcm, double *B, int n, int j0, int nbw, const double *rowscale) {
  extern __shared__ double cov_smem[];
  double (*sX)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem);
  double (*sB)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem + NB * (NB + 1));
  const int tid = threadIdx.x, c0 = blockIdx.x * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int rI = e & (NB - 1), cI = e >> 6;
    sX[rI][cI] = Xcm[cI * NB + rI];  // X[r][c]
    const int col = c0 + cI;
    double b = (rI < nbw && col < n) ? B[(size_t)col * n + j0 + rI] : 0.0;
    if (rowscale && rI < nbw) b *= rowscale[j0 + rI];
    sB[rI][cI] = b;
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int rI = e & (NB - 1), cI = e >> 6;
    double s = 0.0;
    if (TRANS) { for (int k = rI; k < NB; k++) s += sX[k][rI] * sB[k][cI]; }   // X^T is upper triangular
    else { for (int k = 0; k <= rI; k++) s += sX[rI][k] * sB[k][cI]; }
    const int col = c0 + cI;
    if (rI < nbw && col < n) B[(size_t)col * n + j0 + rI] = s;
  }
}
// forward: B[i0.., c0..] -= L[i0.., j0..j0+nbw) B[j0.., c0..]   (grid: row tiles below x column tiles)
__global__ void __launch_bounds__(256) mrhs_fwd_update_kernel(const double *A, double *B, int n, int j0, int nbw) {
  extern __shared__ double cov_smem[];
  double (*sL)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem);
  double (*sB)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem + NB * (NB + 1));
  const int tid = threadIdx.x, i0 = j0 + nbw + blockIdx.x * NB, c0 = blockIdx.y * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int rI = e & (NB - 1), kI = e >> 6;
    sL[rI][kI] = (i0 + rI < n && kI < nbw) ? A[(size_t)(j0 + kI) * n + i0 + rI] : 0.0;
    sB[rI][kI] = (rI < nbw && c0 + kI < n) ? B[(size_t)(c0 + kI) * n + j0 + rI] : 0.0;  // sB[k][col]
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int rI = e & (NB - 1), cI = e >> 6;
    if (i0 + rI >= n || c0 + cI >= n) continue;
    double s = 0.0;
    for (int k = 0; k < NB; k++) s += sL[rI][k] * sB[k][cI];
    B[(size_t)(c0 + cI) * n + i0 + rI] -= s;
  }
}
// backward: B[j0.., c0..] -= sum over row blocks r > j of L[r.., j0..]^T B[r.., c0..]   (one CTA per column tile)
__global__ void __launch_bounds__(256) mrhs_bwd_update_kernel(const double *A, double *B, int n, int j0, int nbw) {
  extern __shared__ double cov_smem[];
  double (*sL)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem);
  double (*sB)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(cov_smem + NB * (NB + 1));
  const int tid = threadIdx.x, c0 = blockIdx.x * NB;
  double acc[16];
  for (int q = 0; q < 16; q++) acc[q] = 0.0;
  for (int i0 = j0 + nbw; i0 < n; i0 += NB) {
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 256) {
      const int rI = e & (NB - 1), kI = e >> 6;
      sL[rI][kI] = (i0 + rI < n && kI < nbw) ? A[(size_t)(j0 + kI) * n + i0 + rI] : 0.0;   // L[i0+r][j0+k]
      sB[rI][kI] = (i0 + rI < n && c0 + kI < n) ? B[(size_t)(c0 + kI) * n + i0 + rI] : 0.0;  // B[i0+r][c0+col]
    }
    __syncthreads();
    for (int q = 0; q < 16; q++) {
      const int e = tid + q * 256, kI = e & (NB - 1), cI = e >> 6;  // output (row j0+kI, col c0+cI)
      double s = 0.0;
      for (int rI = 0; rI < NB; rI++) s += sL[rI][kI] * sB[rI][cI];
      acc[q] += s;
    }
  }
  for (int q = 0; q < 16; q++) {
    const int e = tid + q * 256, kI = e & (NB - 1), cI = e >> 6;
    if (kI < nbw && c0 + cI < n) B[(size_t)(c0 + cI) * n + j0 + kI] -= acc[q];
  }
}
__global__ void transpose_inplace_kernel(double *B, int n) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 32 + threadIdx.y;
  if (i < n && j < i) {
    const double a = B[(size_t)j * n + i], b = B[(size_t)i * n + j];
    B[(size_t)j * n + i] = b;
    B[(size_t)i * n + j] = a;
  }
}

__global__ void scale_rows_kernel(double *B, const double *dinv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n) B[(size_t)j * n + i] *= dinv[i];
}

// B <- (L D L^T)^-1 B for all n columns: forward sweep (B_j <- X_j B_j, rows below -= L21 B_j), rows scaled by 1/d,
// backward sweep (B_j <- X_j^T (B_j - sum_{r>j} L_rj^T B_r))
int solve_many(balm_ctx *c, double *B) {
  const int n = c->n, ct = (n + NB - 1) / NB;
  cudaStream_t st = c->stream;
  int launches = 0;
  for (int j0 = 0, pi = 0; j0 < n; j0 += NB, pi++) {
    const int nbw = std::min(NB, n - j0), m = n - j0 - nbw;
    mrhs_diag_kernel<0><<<ct, 256, MRHS_SMEM, st>>>(c->Xinv + (size_t)pi * NB * NB, B, n, j0, nbw, nullptr);
    if (m > 0) mrhs_fwd_update_kernel<<<dim3((m + NB - 1) / NB, ct), 256, MRHS_SMEM, st>>>(c->A, B, n, j0, nbw);
    launches += 2;
  }
  scale_rows_kernel<<<dim3((n + 255) / 256, n), 256, 0, st>>>(B, c->dinv, n);
  launches++;
  const int npan = (n + NB - 1) / NB;
  for (int pi = npan - 1; pi >= 0; pi--) {
    const int j0 = pi * NB, nbw = std::min(NB, n - j0);
    if (j0 + nbw < n) mrhs_bwd_update_kernel<<<ct, 256, MRHS_SMEM, st>>>(c->A, B, n, j0, nbw);
    mrhs_diag_kernel<1><<<ct, 256, MRHS_SMEM, st>>>(c->Xinv + (size_t)pi * NB * NB, B, n, j0, nbw, nullptr);
    launches += 2;
  }
  c->launches += launches;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}

}  // namespace

int covariance_setup() {
  CUDA_TRY(cudaFuncSetAttribute(mrhs_diag_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRHS_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(mrhs_diag_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MRHS_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(mrhs_fwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MRHS_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(mrhs_bwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MRHS_SMEM));
  return BALM_OK;
}

struct DevPool {  // device scratch of one call, freed on every exit path
  std::vector<void *> p;
  ~DevPool() { for (void *q : p) cudaFree(q); }
  template <typename T> int get(T **out, size_t count) {
    CUDA_TRY(cudaMalloc((void **)out, sizeof(T) * (count ? count : 1)));
    p.push_back((void *)*out);
    return BALM_OK;
  }
};

// Defined in balm_api.cu
int balm_cov_evaluate_hessian(balm_ctx *c, const double *poses_dev, bool include_fix);

int pose_covariance_dev(balm_ctx *c, const double *poses_dev, const double *ccov_host, double point_noise, bool include_fix,
                        double *Rraw_host, double *Rcov_host) {
  const int n = c->n, N = c->N;
  const int64_t M = c->M, K = c->K;
  cudaStream_t st = c->stream;
  if (M < 1) { balm_set_error("balm_pose_covariance: no voxels registered"); return BALM_ERR_INVALID; }
  DevPool pool;
  // voxels per batch: the stats table holds VB voxels; the row buffer gets its own budget
  size_t budget = (size_t)4 << 30;
  if (const char *e = getenv("BALM_COV_BUDGET_MB")) budget = (size_t)atoll(e) << 20;
  int64_t vb = std::min<int64_t>(c->VB, (int64_t)(budget / ((size_t)8 * c->ldg * sizeof(double))));
  vb = std::max<int64_t>(1, std::min<int64_t>(vb, M));
  double *G = nullptr, *Wz = nullptr, *Dobs = nullptr, *Dacc = nullptr, *R = nullptr, *ccov = nullptr;
  CTRY(pool.get(&G, (size_t)8 * vb * c->ldg));
  CTRY(pool.get(&Wz, (size_t)vb * 10));
  CTRY(pool.get(&Dobs, (size_t)K * 21));
  CTRY(pool.get(&Dacc, (size_t)21 * N));
  CTRY(pool.get(&R, (size_t)n * n));
  if (ccov_host) {
    CTRY(pool.get(&ccov, (size_t)K * 81));
    CUDA_TRY(cudaMemcpyAsync(ccov, ccov_host, sizeof(double) * (size_t)K * 81, cudaMemcpyHostToDevice, st));
  }
  bool first = true;
  for (int64_t v0 = 0; v0 < M; v0 += vb) {
    const int64_t v1 = std::min(M, v0 + vb), nv = v1 - v0;
    CTRY(launch_csc_batch(c, v0, v1));
    double *rdummy = c->scal + 21;
    CUDA_TRY(cudaMemsetAsync(rdummy, 0, sizeof(double), st));
    CTRY(launch_voxel_stats(c, poses_dev, v0, v1, c->stats, include_fix, rdummy));
    CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)8 * nv * c->ldg, st));
    CovArgs a{c->obs, c->Kp, c->pose_idx, c->row_ptr, poses_dev, c->stats, ccov, point_noise * point_noise, v0, v1, N,
              c->ldg, Wz, G, Dobs};
    const unsigned grid = (unsigned)((nv + 3) / 4);
    cov_wz_kernel<<<grid, 128, 0, st>>>(a);
    cov_rows_kernel<<<grid, 128, 0, st>>>(a);
    cov_diag_reduce_kernel<<<N, 32, 0, st>>>(Dobs, N, M, c->dense ? 1 : 0, c->csc_ptr, c->csc_obs, Dacc, first ? 0 : 1, v0,
                                             v1, c->csc_lo, c->csc_hi);
    c->launches += 3;
    CTRY(launch_syrk_f64_on(c, G, 8 * nv, first, /*row_xor=*/4));
    first = false;
  }
  CTRY(launch_csc_batch(c, 0, M));
  const int nbb = (n + 31) / 32;
  cov_assemble_kernel<<<dim3(nbb, nbb), dim3(32, 32), 0, st>>>(c->syrk_part, c->syrk_splits, c->syrk_tiles, c->syrk_nb, Dacc, N,
                                                             n, R);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
  if (Rraw_host) CUDA_TRY(cudaMemcpyAsync(Rraw_host, R, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, st));
  if (Rcov_host) {
    // H at these poses (divide_thread, BAs_left.hpp:1091), factored without damping, then H^-1 R H^-T
    CTRY(balm_cov_evaluate_hessian(c, poses_dev, include_fix));
    CTRY(launch_ldlt_solve(c, 0.0));
    CUDA_TRY(cudaMemcpyAsync(c->h_flags, c->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // a singular H (free gauge: no fix clusters in it) does not produce an exact zero pivot in floating point but
    // pivots at rounding level relative to the others: refuse instead of returning 1e16-sized garbage
    std::vector<double> dinv_h((size_t)n);
    CUDA_TRY(cudaMemcpy(dinv_h.data(), c->dinv, sizeof(double) * n, cudaMemcpyDeviceToHost));
    double dmin = 1e300, dmax = 0.0;
    for (int i = 0; i < n; i++) { const double d = fabs(1.0 / dinv_h[i]); dmin = std::min(dmin, d); dmax = std::max(dmax, d); }
    if (c->h_flags[0] || !(dmin > 1e-11 * dmax)) {
      balm_set_error("balm_pose_covariance: H is singular (gauge freedom not fixed: no fix clusters in the Hessian)");
      return BALM_ERR_NOT_PD;
    }
    CTRY(solve_many(c, R));                                   // H^-1 R
    transpose_inplace_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 32), 0, st>>>(R, n);
    CTRY(solve_many(c, R));                                   // H^-1 (H^-1 R)^T = H^-1 R H^-T
    c->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(Rcov_host, R, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return BALM_OK;
}
