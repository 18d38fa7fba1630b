/*
 * balm_oracle.c -- CPU restatement of the BALM 2.0 BA hot path.  TEST INFRASTRUCTURE ONLY
 * (see balm_oracle.h for the scope, the reference file:line map and the parity-pin note).
 *
 * Written for fidelity to the reference loop nest, not for speed: the O(k^2) 6x6 pair loop with
 * read-modify-write into a dense column-major n x n matrix (bavoxel.hpp:404-418), the upper->lower
 * mirror (:422-424), the 4-thread split with in-order reduction (:1044-1056) and the unblocked pivoted
 * LDL^T are kept on purpose because this file doubles as the CPU baseline (bench.py cpu_baseline leg).
 * The two defects of bavoxel.hpp:312-320 (the Cs table built for ALL voxels by every thread and never
 * freed) are neutralised: per-voxel 4x4 forms are built on the fly for [head,end) only.
 */
#include "balm_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------- small dense helpers (row-major scratch matrices) ---------- */

static void mat_mul(const double *A, const double *B, double *C, int m, int k, int n) {
  /* C(m x n) = A(m x k) * B(k x n), all row-major */
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int t = 0; t < k; t++) s += A[i * k + t] * B[t * n + j];
      C[i * n + j] = s;
    }
}

static void mat_mul_bt(const double *A, const double *B, double *C, int m, int k, int n) {
  /* C(m x n) = A(m x k) * B(n x k)^T */
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int t = 0; t < k; t++) s += A[i * k + t] * B[j * k + t];
      C[i * n + j] = s;
    }
}

/* tools.hpp:99-106 */
static void hat3(const double v[3], double O[9]) {
  O[0] = 0;     O[1] = -v[2]; O[2] = v[1];
  O[3] = v[2];  O[4] = 0;     O[5] = -v[0];
  O[6] = -v[1]; O[7] = v[0];  O[8] = 0;
}

/* pose12 (R col-major + p) -> row-major 3x3 R and p */
static void pose_unpack(const double *pose12, double R[9], double p[3]) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[r * 3 + c] = pose12[c * 3 + r];
  p[0] = pose12[9]; p[1] = pose12[10]; p[2] = pose12[11];
}

static void pose_pack(const double R[9], const double p[3], double *pose12) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) pose12[c * 3 + r] = R[r * 3 + c];
  pose12[9] = p[0]; pose12[10] = p[1]; pose12[11] = p[2];
}

/* T = [R p; 0 1] (bavoxel.hpp:308-310) */
static void pose_T4(const double *pose12, double T[16]) {
  double R[9], p[3];
  pose_unpack(pose12, R, p);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = p[r];
  }
  T[12] = T[13] = T[14] = 0; T[15] = 1;
}

/* Co = [P v; v^T N] (bavoxel.hpp:318) */
static void cluster_C4(const double *c10, double C[16]) {
  C[0] = c10[0]; C[1] = c10[1]; C[2] = c10[2];  C[3] = c10[6];
  C[4] = c10[1]; C[5] = c10[3]; C[6] = c10[4];  C[7] = c10[7];
  C[8] = c10[2]; C[9] = c10[4]; C[10] = c10[5]; C[11] = c10[8];
  C[12] = c10[6]; C[13] = c10[7]; C[14] = c10[8]; C[15] = c10[9];
}

/* ---------- SO(3) ---------- */

void orc_exp_so3(const double phi[3], double Rcm[9]) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double nrm = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  if (nrm >= 1e-11) { /* tools.hpp:60 */
    double ax[3] = {phi[0] / nrm, phi[1] / nrm, phi[2] / nrm};
    double K[9], KK[9];
    hat3(ax, K);
    mat_mul(K, K, KK, 3, 3, 3);
    double s = sin(nrm), c1 = 1.0 - cos(nrm);
    for (int i = 0; i < 9; i++) R[i] += s * K[i] + c1 * KK[i];
  }
  double p0[3] = {0, 0, 0}, tmp[12];
  pose_pack(R, p0, tmp);
  memcpy(Rcm, tmp, 9 * sizeof(double));
}

void orc_log_so3(const double Rcm[9], double phi[3]) {
  /* R(r,c) = Rcm[c*3+r]; tools.hpp:92-97 */
  double tr = Rcm[0] + Rcm[4] + Rcm[8];
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : acos(0.5 * (tr - 1));
  double K[3] = {Rcm[1 * 3 + 2] - Rcm[2 * 3 + 1],  /* R(2,1)-R(1,2) */
                 Rcm[2 * 3 + 0] - Rcm[0 * 3 + 2],  /* R(0,2)-R(2,0) */
                 Rcm[0 * 3 + 1] - Rcm[1 * 3 + 0]}; /* R(1,0)-R(0,1) */
  double f = (fabs(theta) < 0.001) ? 0.5 : 0.5 * theta / sin(theta);
  for (int i = 0; i < 3; i++) phi[i] = f * K[i];
}

/* ---------- 3x3 symmetric eigen-solver (cyclic Jacobi, eigenvalues ascending) ---------- */

void orc_eig3(const double Acm[9], double lam[3], double Ucm[9]) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) a[r][c] = 0.5 * (Acm[c * 3 + r] + Acm[r * 3 + c]);
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
    static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int e = 0; e < 3; e++) {
      int p = PQ[e][0], q = PQ[e][1];
      double apq = a[p][q];
      if (apq == 0.0) continue;
      double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      /* A <- J^T A J with J the (p,q) rotation */
      for (int k = 0; k < 3; k++) {
        double akp = a[k][p], akq = a[k][q];
        a[k][p] = c * akp - s * akq;
        a[k][q] = s * akp + c * akq;
      }
      for (int k = 0; k < 3; k++) {
        double apk = a[p][k], aqk = a[q][k];
        a[p][k] = c * apk - s * aqk;
        a[q][k] = s * apk + c * aqk;
      }
      a[p][q] = a[q][p] = 0.0;
      for (int k = 0; k < 3; k++) {
        double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - s * vkq;
        v[k][q] = s * vkp + c * vkq;
      }
    }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (d[idx[j]] > d[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  for (int k = 0; k < 3; k++) {
    lam[k] = d[idx[k]];
    for (int r = 0; r < 3; r++) Ucm[k * 3 + r] = v[r][idx[k]];
  }
}

/* ---------- PointCluster::transform (tools.hpp:333-339) ---------- */

void orc_cluster_transform(const double in10[10], const double pose12[12], double out10[10]) {
  double R[9], p[3];
  pose_unpack(pose12, R, p);
  double P[9] = {in10[0], in10[1], in10[2], in10[1], in10[3], in10[4], in10[2], in10[4], in10[5]};
  double v[3] = {in10[6], in10[7], in10[8]};
  double N = in10[9];
  double Rv[3], RP[9], RPRt[9];
  mat_mul(R, v, Rv, 3, 3, 1);
  mat_mul(R, P, RP, 3, 3, 3);
  mat_mul_bt(RP, R, RPRt, 3, 3, 3);
  double Pn[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double rp = Rv[r] * p[c], rpT = Rv[c] * p[r]; /* rp = R v p^T (tools.hpp:337) */
      Pn[r * 3 + c] = RPRt[r * 3 + c] + rp + rpT + N * p[r] * p[c];
    }
  out10[0] = Pn[0]; out10[1] = Pn[1]; out10[2] = Pn[2];
  out10[3] = Pn[4]; out10[4] = Pn[5]; out10[5] = Pn[8];
  for (int r = 0; r < 3; r++) out10[6 + r] = Rv[r] + N * p[r];
  out10[9] = N;
}

/* ---------- left_evaluate_acc2 (bavoxel.hpp:304-426) ---------- */

void orc_left_evaluate_acc2(const orc_problem *pb, const double *poses12, int64_t head, int64_t end,
                            int include_fix, double *H, double *g, double *residual) {
  const int N = pb->n_poses;
  const int n = 6 * N;
  memset(H, 0, sizeof(double) * (size_t)n * n);
  memset(g, 0, sizeof(double) * n);
  *residual = 0;
  const int l = 0;

  double *T = (double *)malloc(sizeof(double) * 16 * N);
  for (int i = 0; i < N; i++) pose_T4(poses12 + 12 * i, T + 16 * i);

  /* per-voxel scratch sized for the largest voxel in range */
  int64_t kmax = 0;
  for (int64_t a = head; a < end; a++) {
    int64_t k = pb->row_ptr[a + 1] - pb->row_ptr[a];
    if (k > kmax) kmax = k;
  }
  double *TC = (double *)malloc(sizeof(double) * 16 * (kmax + 1));
  double *TCT = (double *)malloc(sizeof(double) * 16 * (kmax + 1));
  double *gkl = (double *)malloc(sizeof(double) * 18 * (kmax + 1)); /* [k][slot][6] */
  double *UlTCF = (double *)malloc(sizeof(double) * 6 * (kmax + 1));

  for (int64_t a = head; a < end; a++) {
    const double coe = pb->coe[a];
    const int64_t s0 = pb->row_ptr[a];
    const int k_obs = (int)(pb->row_ptr[a + 1] - s0);
    const int32_t *pidx = pb->pose_idx + s0;
    double C[16];
    memset(C, 0, sizeof(C)); /* bavoxel.hpp:325 */
    if (include_fix && pb->fix10) cluster_C4(pb->fix10 + 10 * a, C); /* benchmark_virtual.cpp:242-243 */

    for (int s = 0; s < k_obs; s++) { /* :331-339, CSR holds exactly the slots with N>0 */
      double Co[16];
      cluster_C4(pb->obs10 + 10 * (s0 + s), Co);
      const double *Tj = T + 16 * pidx[s];
      mat_mul(Tj, Co, TC + 16 * s, 4, 4, 4);
      mat_mul_bt(TC + 16 * s, Tj, TCT + 16 * s, 4, 4, 4);
      for (int e = 0; e < 16; e++) C[e] += TCT[16 * s + e];
    }

    const double NN = C[15]; /* :341 */
    for (int e = 0; e < 16; e++) C[e] /= NN;
    const double vbar[3] = {C[3], C[7], C[11]};
    double A[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[c * 3 + r] = C[r * 4 + c] - vbar[r] * vbar[c]; /* :345 */
    double lmbd[3], Ucm[9];
    orc_eig3(A, lmbd, Ucm);
    *residual += coe * lmbd[l]; /* :349 */

    double u[3][3], U[3][24]; /* U[k] is 6x4 row-major (:351-360) */
    for (int k = 0; k < 3; k++) {
      for (int r = 0; r < 3; r++) u[k][r] = Ucm[k * 3 + r];
      double nu[3] = {-u[k][0], -u[k][1], -u[k][2]}, Hm[9];
      hat3(nu, Hm);
      memset(U[k], 0, sizeof(U[k]));
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[k][r * 4 + c] = Hm[r * 3 + c];
      for (int r = 0; r < 3; r++) U[k][(3 + r) * 4 + 3] = u[k][r];
    }

    for (int s = 0; s < k_obs; s++) { /* :365-395 */
      const int i = pidx[s];
      const double *Ti = T + 16 * i;
      double temp[12]; /* 3x4 = T[i].block<3,4>(0,0), last column minus vbar (:368-369) */
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 4; c++) temp[r * 4 + c] = Ti[r * 4 + c];
        temp[r * 4 + 3] -= vbar[r];
      }
      double M[12]; /* 4x3 TC_TCFSp = TC[i] * temp^T (:370) */
      mat_mul_bt(TC + 16 * s, temp, M, 4, 4, 3);
      for (int k = 0; k < 3; k++) {
        double Mul[4], Muk[4], g1[6], g2[6];
        mat_mul(M, u[l], Mul, 4, 3, 1);
        mat_mul(M, u[k], Muk, 4, 3, 1);
        mat_mul(U[k], Mul, g1, 6, 4, 1); /* :374 */
        mat_mul(U[l], Muk, g2, 6, 4, 1); /* :375 */
        for (int r = 0; r < 6; r++) gkl[(k * (kmax + 1) + s) * 6 + r] = (g1[r] + g2[r]) / NN;
      }
      double UTC[24]; /* U[l]*TC[i] : 6x4; column 3 kept (:380) */
      mat_mul(U[l], TC + 16 * s, UTC, 6, 4, 4);
      double *ai = UlTCF + 6 * s;
      for (int r = 0; r < 6; r++) ai[r] = UTC[r * 4 + 3];
      const double *gl = gkl + (l * (kmax + 1) + s) * 6;
      for (int r = 0; r < 6; r++) g[6 * i + r] += coe * gl[r]; /* :381 */

      double Ha[36];
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) Ha[r * 6 + c] = -2.0 / NN / NN * ai[r] * ai[c]; /* :385 */
      double Mu3[3] = {M[0] * u[l][0] + M[1] * u[l][1] + M[2] * u[l][2],
                       M[3] * u[l][0] + M[4] * u[l][1] + M[5] * u[l][2],
                       M[6] * u[l][0] + M[7] * u[l][1] + M[8] * u[l][2]};
      double h1[9], h2[9], Ell[9];
      hat3(Mu3, h1);
      hat3(u[l], h2);
      mat_mul(h1, h2, Ell, 3, 3, 3);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ha[r * 6 + c] += (Ell[r * 3 + c] + Ell[c * 3 + r]) / NN; /* :387-388 */
      for (int k = 0; k < 3; k++)
        if (k != l) {
          const double w = 2.0 / (lmbd[l] - lmbd[k]);
          const double *gk = gkl + (k * (kmax + 1) + s) * 6;
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) Ha[r * 6 + c] += w * gk[r] * gk[c]; /* :392 */
        }
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) H[(size_t)(6 * i + c) * n + 6 * i + r] += coe * Ha[r * 6 + c]; /* :394 */
    }

    for (int s = 0; s < k_obs; s++) { /* :397-402 */
      const int i = pidx[s];
      double UT[24], Hb[36];
      mat_mul(U[l], TCT + 16 * s, UT, 6, 4, 4);
      mat_mul_bt(UT, U[l], Hb, 6, 4, 6);
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) H[(size_t)(6 * i + c) * n + 6 * i + r] += 2.0 / NN * coe * Hb[r * 6 + c];
    }

    for (int s = 0; s < k_obs - 1; s++) { /* :404-418 */
      const int i = pidx[s];
      const double *ai = UlTCF + 6 * s;
      for (int t = s + 1; t < k_obs; t++) {
        const int j = pidx[t];
        const double *aj = UlTCF + 6 * t;
        double Ha[36];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) Ha[r * 6 + c] = -2.0 / NN / NN * ai[r] * aj[c];
        for (int k = 0; k < 3; k++)
          if (k != l) {
            const double w = 2.0 / (lmbd[l] - lmbd[k]);
            const double *gi = gkl + (k * (kmax + 1) + s) * 6;
            const double *gj = gkl + (k * (kmax + 1) + t) * 6;
            for (int r = 0; r < 6; r++)
              for (int c = 0; c < 6; c++) Ha[r * 6 + c] += w * gi[r] * gj[c];
          }
        for (int c = 0; c < 6; c++) {
          double *col = H + (size_t)(6 * j + c) * n + 6 * i;
          for (int r = 0; r < 6; r++) col[r] += coe * Ha[r * 6 + c];
        }
      }
    }
  }

  for (int i = 1; i < N; i++) /* :422-424 mirror upper blocks to lower */
    for (int j = 0; j < i; j++)
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++)
          H[(size_t)(6 * j + c) * n + 6 * i + r] = H[(size_t)(6 * i + r) * n + 6 * j + c];

  free(T); free(TC); free(TCT); free(gkl); free(UlTCF);
}

/* ---------- evaluate_only_residual (bavoxel.hpp:428-470) ---------- */

void orc_evaluate_only_residual(const orc_problem *pb, const double *poses12, double *residual) {
  double r = 0;
  for (int64_t a = 0; a < pb->n_voxels; a++) {
    double sig[10];
    if (pb->fix10) memcpy(sig, pb->fix10 + 10 * a, sizeof(sig));
    else memset(sig, 0, sizeof(sig));
    for (int64_t s = pb->row_ptr[a]; s < pb->row_ptr[a + 1]; s++) {
      double tr[10];
      orc_cluster_transform(pb->obs10 + 10 * s, poses12 + 12 * pb->pose_idx[s], tr);
      for (int e = 0; e < 10; e++) sig[e] += tr[e];
    }
    const double Nn = sig[9];
    double vb[3] = {sig[6] / Nn, sig[7] / Nn, sig[8] / Nn};
    double P[9] = {sig[0], sig[1], sig[2], sig[1], sig[3], sig[4], sig[2], sig[4], sig[5]};
    double A[9];
    for (int rr = 0; rr < 3; rr++)
      for (int c = 0; c < 3; c++) A[c * 3 + rr] = P[rr * 3 + c] / Nn - vb[rr] * vb[c];
    double lam[3], U[9];
    orc_eig3(A, lam, U);
    r += pb->coe[a] * lam[0];
  }
  *residual = r;
}

/* ---------- divide_thread_left (bavoxel.hpp:1025-1059) ---------- */

typedef struct {
  const orc_problem *pb;
  const double *poses;
  int64_t head, end;
  int include_fix;
  double *H, *g;
  double r;
} thr_arg;

static void *thr_main(void *p) {
  thr_arg *a = (thr_arg *)p;
  orc_left_evaluate_acc2(a->pb, a->poses, a->head, a->end, a->include_fix, a->H, a->g, &a->r);
  return NULL;
}

double orc_divide_thread_left(const orc_problem *pb, const double *poses12, int threads, int include_fix,
                              double *H, double *g) {
  const int n = 6 * pb->n_poses;
  int tn = threads;
  if (pb->n_voxels < tn) tn = 1; /* :1042 */
  if (tn <= 1) {
    double r;
    orc_left_evaluate_acc2(pb, poses12, 0, pb->n_voxels, include_fix, H, g, &r);
    return r;
  }
  memset(H, 0, sizeof(double) * (size_t)n * n);
  memset(g, 0, sizeof(double) * n);
  thr_arg *args = (thr_arg *)calloc(tn, sizeof(thr_arg));
  pthread_t *th = (pthread_t *)calloc(tn, sizeof(pthread_t));
  double part = 1.0 * (double)pb->n_voxels / tn; /* :1045 */
  for (int i = 0; i < tn; i++) {
    args[i].pb = pb; args[i].poses = poses12; args[i].include_fix = include_fix;
    args[i].head = (int64_t)(part * i);       /* implicit double->int conversion at :1047 */
    args[i].end = (int64_t)(part * (i + 1));
    args[i].H = (double *)malloc(sizeof(double) * (size_t)n * n);
    args[i].g = (double *)malloc(sizeof(double) * n);
    pthread_create(&th[i], NULL, thr_main, &args[i]);
  }
  double residual = 0;
  for (int i = 0; i < tn; i++) { /* join in order, add in order (:1049-1056) */
    pthread_join(th[i], NULL);
    for (size_t e = 0; e < (size_t)n * n; e++) H[e] += args[i].H[e];
    for (int e = 0; e < n; e++) g[e] += args[i].g[e];
    residual += args[i].r;
    free(args[i].H); free(args[i].g);
  }
  free(args); free(th);
  return residual;
}

/* ---------- dense LDL^T with diagonal pivoting (Eigen LDLT semantics, bavoxel.hpp:1114) ----------
 * Eigen/src/Cholesky/LDLT.h (unblocked, Lower): at step k pick the largest |diagonal| of the remaining
 * block, apply the symmetric transposition, then the left-looking update of column k.  P A P^T = L D L^T. */

int orc_ldlt_solve(int n, const double *A, const double *b, double *x) {
  double *L = (double *)malloc(sizeof(double) * (size_t)n * n); /* col-major, lower used */
  int *tr = (int *)malloc(sizeof(int) * n);
  double *tmp = (double *)malloc(sizeof(double) * n);
  memcpy(L, A, sizeof(double) * (size_t)n * n);
  int zero_pivot = 0;
#define LM(r, c) L[(size_t)(c) * n + (r)]
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = fabs(LM(k, k));
    for (int i = k + 1; i < n; i++)
      if (fabs(LM(i, i)) > best) { best = fabs(LM(i, i)); piv = i; }
    tr[k] = piv;
    if (piv != k) {
      /* symmetric swap of rows/cols k and piv on the lower triangle */
      for (int c = 0; c < k; c++) { double t = LM(k, c); LM(k, c) = LM(piv, c); LM(piv, c) = t; }
      for (int r = piv + 1; r < n; r++) { double t = LM(r, k); LM(r, k) = LM(r, piv); LM(r, piv) = t; }
      for (int i = k + 1; i < piv; i++) { double t = LM(i, k); LM(i, k) = LM(piv, i); LM(piv, i) = t; }
      { double t = LM(k, k); LM(k, k) = LM(piv, piv); LM(piv, piv) = t; }
    }
    if (k > 0) {
      for (int c = 0; c < k; c++) tmp[c] = LM(c, c) * LM(k, c); /* D * A10^T */
      double s = 0;
      for (int c = 0; c < k; c++) s += LM(k, c) * tmp[c];
      LM(k, k) -= s;
      for (int c = 0; c < k; c++) { /* A21 -= A20 * tmp, column sweep (col-major friendly) */
        const double tc = tmp[c];
        const double *col = &LM(0, c);
        double *dst = &LM(0, k);
        for (int r = k + 1; r < n; r++) dst[r] -= col[r] * tc;
      }
    }
    const double akk = LM(k, k);
    if (fabs(akk) > 0) {
      for (int r = k + 1; r < n; r++) LM(r, k) /= akk;
    } else {
      zero_pivot = 1;
    }
  }
  /* solve: x = P^T L^-T D^-1 L^-1 P b */
  for (int i = 0; i < n; i++) x[i] = b[i];
  for (int k = 0; k < n; k++)
    if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
  for (int c = 0; c < n; c++) {
    const double xc = x[c];
    for (int r = c + 1; r < n; r++) x[r] -= LM(r, c) * xc;
  }
  for (int i = 0; i < n; i++) {
    const double d = LM(i, i);
    x[i] = (fabs(d) > 2.2250738585072014e-308) ? x[i] / d : 0.0; /* Eigen: tolerance = 1/highest() */
  }
  for (int c = n - 1; c >= 0; c--) {
    double s = x[c];
    for (int r = c + 1; r < n; r++) s -= LM(r, c) * x[r];
    x[c] = s;
  }
  for (int k = n - 1; k >= 0; k--)
    if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
#undef LM
  free(L); free(tr); free(tmp);
  return zero_pivot;
}

/* ---------- left update (bavoxel.hpp:1116-1126) ---------- */

void orc_left_update(int N, const double *poses12, const double *dx, double *out) {
  for (int j = 0; j < N; j++) {
    double dRcm[9], dR[9], R[9], p[3], Rn[9], pn[3], p0[3] = {0, 0, 0}, tmp[12];
    orc_exp_so3(dx + 6 * j, dRcm);
    memcpy(tmp, dRcm, sizeof(dRcm)); tmp[9] = tmp[10] = tmp[11] = 0;
    pose_unpack(tmp, dR, p0);
    pose_unpack(poses12 + 12 * j, R, p);
    mat_mul(dR, R, Rn, 3, 3, 3);
    mat_mul(dR, p, pn, 3, 3, 1);
    for (int r = 0; r < 3; r++) pn[r] += dx[6 * j + 3 + r];
    pose_pack(Rn, pn, out + 12 * j);
  }
}

int orc_lm_step(int N, const double *H, const double *g, double u, const double *poses12, double *dx,
                double *trial, double *q1) {
  const int n = 6 * N;
  double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *rhs = (double *)malloc(sizeof(double) * n);
  memcpy(A, H, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; i++) { A[(size_t)i * n + i] += u * H[(size_t)i * n + i]; rhs[i] = -g[i]; } /* :1113-1114 */
  int zp = orc_ldlt_solve(n, A, rhs, dx);
  orc_left_update(N, poses12, dx, trial);
  double s = 0; /* q1 = 0.5 * dx . (u*D*dx - g), D = diag(H) (:1127) */
  for (int i = 0; i < n; i++) s += dx[i] * (u * H[(size_t)i * n + i] * dx[i] - g[i]);
  *q1 = 0.5 * s;
  free(A); free(rhs);
  return zp;
}

/* ---------- damping_iter (bavoxel.hpp:1069-1166) ---------- */

static void gauge_to_pose0(int N, double *poses12, int force_identity) {
  double R0[9], p0[3];
  pose_unpack(poses12, R0, p0);
  int first = force_identity ? 1 : 0;
  for (int i = first; i < N; i++) {
    double R[9], p[3], d[3], pn[3], Rn[9];
    pose_unpack(poses12 + 12 * i, R, p);
    for (int r = 0; r < 3; r++) d[r] = p[r] - p0[r];
    for (int r = 0; r < 3; r++) pn[r] = R0[0 * 3 + r] * d[0] + R0[1 * 3 + r] * d[1] + R0[2 * 3 + r] * d[2];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        Rn[r * 3 + c] = R0[0 * 3 + r] * R[0 * 3 + c] + R0[1 * 3 + r] * R[1 * 3 + c] + R0[2 * 3 + r] * R[2 * 3 + c];
    pose_pack(Rn, pn, poses12 + 12 * i);
  }
  if (force_identity) { /* benchmark_virtual.cpp:478-479 */
    double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
    pose_pack(I, z, poses12);
  }
}

int orc_damping_iter(const orc_problem *pb, double *poses12, const orc_lm_opts *o, orc_trace *trace,
                     int *n_iters, double *poses_per_iter) {
  const int N = pb->n_poses, n = 6 * N;
  *n_iters = 0;
  if (o->min_planes_per_pose > 0) { /* :1071-1085 */
    int *planes = (int *)calloc(N, sizeof(int));
    for (int64_t s = 0; s < pb->row_ptr[pb->n_voxels]; s++) planes[pb->pose_idx[s]]++;
    int mn = planes[0];
    for (int i = 1; i < N; i++) if (planes[i] < mn) mn = planes[i];
    free(planes);
    if (mn < o->min_planes_per_pose) return 4;
  }
  double u = o->u0, v = o->v0;
  double *H = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *g = (double *)malloc(sizeof(double) * n);
  double *dx = (double *)malloc(sizeof(double) * n);
  double *trial = (double *)malloc(sizeof(double) * 12 * N);
  double r1 = 0, r2 = 0;
  int calc_hess = 1;
  for (int it = 0; it < o->max_iter; it++) {
    if (calc_hess) r1 = orc_divide_thread_left(pb, poses12, o->threads, o->hess_includes_fix, H, g);
    double q1;
    orc_lm_step(N, H, g, u, poses12, dx, trial, &q1);
    orc_evaluate_only_residual(pb, trial, &r2);
    double q = r1 - r2;
    orc_trace *t = &trace[it];
    t->r1 = r1; t->r2 = r2; t->u = u; t->v = v; t->q = q; t->q1 = q1; t->recomputed_hess = calc_hess;
    if (q > 0) { /* :1134-1143 */
      memcpy(poses12, trial, sizeof(double) * 12 * N);
      double rho = q / q1;
      v = 2;
      double f = 1 - pow(2 * rho - 1, 3);
      u *= (f < (1.0 / 3.0) ? (1.0 / 3.0) : f);
      calc_hess = 1;
      t->accepted = 1;
    } else { /* :1144-1149 */
      u = u * v;
      v = 2 * v;
      calc_hess = 0;
      t->accepted = 0;
    }
    if (poses_per_iter) memcpy(poses_per_iter + (size_t)it * 12 * N, poses12, sizeof(double) * 12 * N);
    *n_iters = it + 1;
    if (o->rel_tol >= 0 && fabs(r1 - r2) / r1 < o->rel_tol) break; /* :1155 */
  }
  if (o->gauge_mode == 0) gauge_to_pose0(N, poses12, 0);
  else if (o->gauge_mode == 1) gauge_to_pose0(N, poses12, 1);
  free(H); free(g); free(dx); free(trial);
  return 0;
}

/* ---------- rsme (benchmark_virtual.cpp:48-61) ---------- */

void orc_rmse(int N, const double *est, const double *gt, double *rot, double *tran) {
  double sr = 0, st = 0;
  for (int i = 0; i < N; i++) {
    double Rg[9], pg[3], Re[9], pe[3], D[9], tmp[12], phi[3], z[3] = {0, 0, 0};
    pose_unpack(gt + 12 * i, Rg, pg);
    pose_unpack(est + 12 * i, Re, pe);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        D[r * 3 + c] = Rg[0 * 3 + r] * Re[0 * 3 + c] + Rg[1 * 3 + r] * Re[1 * 3 + c] + Rg[2 * 3 + r] * Re[2 * 3 + c];
    pose_pack(D, z, tmp);
    orc_log_so3(tmp, phi);
    sr += phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    for (int r = 0; r < 3; r++) st += (pe[r] - pg[r]) * (pe[r] - pg[r]);
  }
  *rot = sqrt(sr / N);
  *tran = sqrt(st / N);
}
