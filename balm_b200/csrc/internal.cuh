// internal.cuh -- shared declarations of libbalm_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/balm_b200.h"

#define BALM_SCAL_RCUR 16
#define BALM_SCAL_RTRIAL 17
#define BALM_STATS_STRIDE 40  // doubles per voxel in the stats table: [0..19] eigen data (vb, u0..u2, 1/NN, the three
                              // sqrt(coe|w_k|), coe, lambda), [20..39] the observation sweep's record (see obs_pass_kernel)
#define BALM_SWEEP_OFF 20
#define BALM_SWEEP_WORDS 20
#define BALM_ACC 33           // per-pose accumulators of the observation pass: g(6) + sym 6x6 diag block (21)
                              // + exact fp64 sums of squares of the 6 G' columns (diagonal of G'^T G')
#define BALM_NB 64            // LDL^T panel width
#define BALM_SYRK_TILE 128

void balm_set_error(const std::string &s);

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      balm_set_error(_b);                                                                     \
      return BALM_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)


struct balm_ctx {
  int N = 0, n = 0, ldg = 0, Np = 0, device = 0, prec = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream of the solve's look-ahead (lower priority)
  cudaEvent_t sev[3] = {nullptr, nullptr, nullptr};  // panel done / rest done (two parities)
  bool solve_lookahead = true;
  double *ysol = nullptr;          // [n] forward-substitution result y (kept apart from the running rhs `sol`)
  bool lm_active = false;     // inside balm_damping_iter (scale speculation and stats hand-over are loop-local)
  bool reuse_ws = false;      // the registration in progress has the shape of the previous one: buffers are kept
  bool defer = false;         // inside balm_damping_iter: phases are enqueued back to back, one host sync per iteration
  bool pending_eval = false;  // an evaluation's events / flags still have to be read (defer mode)
  bool pending_stats_cached = false;
  int sm_count = 148;

  // ---- registered problem (device) ----
  int64_t M = 0, K = 0, Kp = 0;
  bool dense = false;   // every voxel observed by every pose, slot j == pose j
  int max_k = 0;
  double *obs = nullptr;          // SoA [10][Kp], voxel-major (CSR order)
  int *pose_idx = nullptr;        // [K]
  long long *row_ptr = nullptr;   // [M+1]
  double *coe = nullptr;          // [M]
  double *fix = nullptr;          // SoA [10][M] or null
  int *csc_ptr = nullptr, *csc_obs = nullptr, *csc_vox = nullptr;  // pose-major lists (sparse problems)
  int csc_max_len = 0;
  int *csc_lo = nullptr, *csc_hi = nullptr;  // [N] segment of every pose's list inside the current voxel batch (null: whole list)
  int min_planes = 0;             // min over poses of #voxels observing it (precheck, bavoxel.hpp:1071-1085)
  int *planes = nullptr;          // [2][N] device: this rank's per-pose voxel counts | their all-reduced sum
  bool registered = false;        // a voxel set (possibly empty: a rank whose shard has no voxels) is registered
  unsigned long long *vox_key = nullptr;  // [M] 63-bit octree key of every voxel (ascending) when the set came from
                                          // balm_cut_voxels: lets balm_append_scan find the leaf a new point falls into
  int *vox_layer = nullptr;       // [M] octree layer of every voxel (a key digit 7 is BOTH octant 7 and "not split":
                                  // the layer says which)
  bool has_keys = false;
  double assoc_voxel_size = 1.0;  // association parameters of that call
  int assoc_layer_limit = 2, assoc_min_ps = 15;
  double assoc_eig[3] = {1.0 / 16, 1.0 / 16, 1.0 / 16};

  // ---- optimiser state (device) ----
  double *poses = nullptr, *poses_trial = nullptr;  // [12N]
  double *H = nullptr, *g = nullptr;                // [n*n + n + 8] contiguous: H | g | r  (one all-reduce)
  double *A = nullptr;                              // [n*n] factor workspace
  double *W = nullptr;                              // [2][n*NB] panel * D (double-buffered by step parity)
  double *dx = nullptr;                             // [n]
  double *dvec = nullptr;                           // [n] diag(H)
  double *Xinv = nullptr;                           // [panels][64*64] inverses of the unit-lower diagonal blocks
  double *dinv = nullptr;                           // [n] 1/d
  double *sol = nullptr;                            // [n] right-hand side / forward-substitution vector
  double *dval = nullptr;                           // [n] d (the tile-DAG factorisation scales its operands with it)
  double *rres = nullptr;                           // [n] residual b - A x of the last solve (backward-error check)
  double *rdelta = nullptr;                         // [n] refinement correction
  int *dag_flags = nullptr;                         // tile-DAG dependency counters: upd | pdone [nt*nt], xdone | rhs_cnt [nt]
  bool solve_dag = true;                            // persistent tile-DAG factorisation (BALM_SOLVE_MULTIKERNEL=1: old path)
  int dag_grid = 0, dag_near = 0;
  unsigned long long *dag_trace = nullptr;          // BALM_DAG_TRACE: device timestamps of the last factorisation
  void *solve_graph = nullptr;                      // cudaGraphExec_t of the solve sequence
  int solve_launches = 0;
  double *scal = nullptr;                           // device scalars [32]: 0 r(eval) 1 q1 2 r(residual) 3 u 4..15 gauge
                                                    // snapshot, 16/17 this rank's residual at the current / trial poses
  double *h_scal = nullptr;                         // pinned host mirror [16]
  int *flags = nullptr;                             // device flags [4] (0: ldlt bad pivot)
  int *h_flags = nullptr;

  // ---- evaluation workspaces ----
  int64_t VB = 0;                 // voxels per batch
  double *stats = nullptr;        // [VB][20]
  double *vsums = nullptr;        // [VB][10] per-voxel cluster sums (stats pass: streaming kernel -> eigen kernel)
  double *stats_trial = nullptr;  // [M][20] stats of the LM trial poses (single-batch problems): an accepted step
                                  // hands them to the next evaluation instead of recomputing them
  double *G = nullptr;            // [3*VB][ldg] fp64 scaled factor matrix G' (MN-major: pose index contiguous)
  int obs_chunks = 0;
  double *obs_part = nullptr;     // [obs_chunks][27][Np]
  double *accum = nullptr;        // [BALM_ACC][Np] reduced g / diag-block / sum-of-squares accumulators
  double *accum_batch = nullptr;  // same, current batch only
  int res_blocks = 0;
  double *res_part = nullptr;     // [res_blocks]
  int syrk_splits = 0, syrk_tiles = 0, syrk_nb = 0;
  double *syrk_part = nullptr;    // [splits][tiles][128*128]

  // ---- tensor path (int8 split-integer) ----
  int8_t *Gq = nullptr;           // [4][rows_alloc][ldg] int8 digit planes of rint(G' * sc) (3 or 4 used per batch)
  unsigned long long *colmax = nullptr;  // [ldg] bit patterns of max |G'[:,j]| of the current batch
  void *tmap = nullptr;           // CUtensorMap storage

  // ---- multi-GPU ----
  void *comm = nullptr;
  double *Hpack = nullptr;        // [n(n+1)/2 + n + 1] lower triangle of H | g | r: what the all-reduce carries
  int rank = 0, world = 1;

  // ---- instrumentation ----
  balm_timings tm{};
  cudaEvent_t ev[16] = {};
  long long launches = 0;
};

// ---------------- device math ----------------
#ifdef __CUDACC__

struct WC {  // world-frame cluster
  double p00, p01, p02, p11, p12, p22, v0, v1, v2, n;
};

// PointCluster::transform (tools.hpp:333-339) on packed symmetric storage. R row-major r[9].
__device__ __forceinline__ WC world_cluster(const double *o, const double *r, const double *p) {
  // Rv
  const double rv0 = r[0] * o[6] + r[1] * o[7] + r[2] * o[8];
  const double rv1 = r[3] * o[6] + r[4] * o[7] + r[5] * o[8];
  const double rv2 = r[6] * o[6] + r[7] * o[7] + r[8] * o[8];
  // X = R * P  (P symmetric: o0 o1 o2 / o1 o3 o4 / o2 o4 o5)
  const double x00 = r[0] * o[0] + r[1] * o[1] + r[2] * o[2];
  const double x01 = r[0] * o[1] + r[1] * o[3] + r[2] * o[4];
  const double x02 = r[0] * o[2] + r[1] * o[4] + r[2] * o[5];
  const double x10 = r[3] * o[0] + r[4] * o[1] + r[5] * o[2];
  const double x11 = r[3] * o[1] + r[4] * o[3] + r[5] * o[4];
  const double x12 = r[3] * o[2] + r[4] * o[4] + r[5] * o[5];
  const double x20 = r[6] * o[0] + r[7] * o[1] + r[8] * o[2];
  const double x21 = r[6] * o[1] + r[7] * o[3] + r[8] * o[4];
  const double x22 = r[6] * o[2] + r[7] * o[4] + r[8] * o[5];
  const double N = o[9];
  WC w;
  // R P R^T + (Rv) p^T + p (Rv)^T + N p p^T
  w.p00 = x00 * r[0] + x01 * r[1] + x02 * r[2] + 2.0 * rv0 * p[0] + N * p[0] * p[0];
  w.p01 = x00 * r[3] + x01 * r[4] + x02 * r[5] + rv0 * p[1] + rv1 * p[0] + N * p[0] * p[1];
  w.p02 = x00 * r[6] + x01 * r[7] + x02 * r[8] + rv0 * p[2] + rv2 * p[0] + N * p[0] * p[2];
  w.p11 = x10 * r[3] + x11 * r[4] + x12 * r[5] + 2.0 * rv1 * p[1] + N * p[1] * p[1];
  w.p12 = x10 * r[6] + x11 * r[7] + x12 * r[8] + rv1 * p[2] + rv2 * p[1] + N * p[1] * p[2];
  w.p22 = x20 * r[6] + x21 * r[7] + x22 * r[8] + 2.0 * rv2 * p[2] + N * p[2] * p[2];
  w.v0 = rv0 + N * p[0];
  w.v1 = rv1 + N * p[1];
  w.v2 = rv2 + N * p[2];
  w.n = N;
  return w;
}

// pose12 (R column-major, p) -> row-major r[9], p[3]
__device__ __forceinline__ void load_pose(const double *__restrict__ pose12, double *r, double *p) {
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int rr = 0; rr < 3; rr++) r[rr * 3 + c] = __ldg(pose12 + c * 3 + rr);
  p[0] = __ldg(pose12 + 9);
  p[1] = __ldg(pose12 + 10);
  p[2] = __ldg(pose12 + 11);
}

// streaming load: read-only path, do not allocate in L1 (the observations are read exactly once per kernel)
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
  return v;
}

template <bool PLAIN>
__device__ __forceinline__ void load_pose_any(const double *pose12, double *r, double *p) {
  if (PLAIN) {  // shared-memory table
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int rr = 0; rr < 3; rr++) r[rr * 3 + c] = pose12[c * 3 + rr];
    p[0] = pose12[9]; p[1] = pose12[10]; p[2] = pose12[11];
  } else {
    load_pose(pose12, r, p);
  }
}

#define BALM_JROT(app, aqq, apq, arp, arq, vp0, vp1, vp2, vq0, vq1, vq2)        \
  if (apq != 0.0) {                                                             \
    const double th = (aqq - app) / (2.0 * apq);                                \
    const double t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1.0));      \
    const double c = rsqrt(t * t + 1.0);                                        \
    const double s = t * c;                                                     \
    app -= t * apq;                                                             \
    aqq += t * apq;                                                             \
    apq = 0.0;                                                                  \
    const double trp = arp, trq = arq;                                          \
    arp = c * trp - s * trq;                                                    \
    arq = s * trp + c * trq;                                                    \
    double a_, b_;                                                              \
    a_ = vp0; b_ = vq0; vp0 = c * a_ - s * b_; vq0 = s * a_ + c * b_;           \
    a_ = vp1; b_ = vq1; vp1 = c * a_ - s * b_; vq1 = s * a_ + c * b_;           \
    a_ = vp2; b_ = vq2; vp2 = c * a_ - s * b_; vq2 = s * a_ + c * b_;           \
  }

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues ascending, eigenvectors u0,u1,u2.
// Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> (bavoxel.hpp:345,452).
__device__ __forceinline__ void eig3_jacobi(double a00, double a01, double a02, double a11, double a12, double a22,
                                            double *lam, double *u0, double *u1, double *u2) {
  double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;  // v[r][c]
  const double scale = fabs(a00) + fabs(a11) + fabs(a22);
#pragma unroll 1
  for (int sweep = 0; sweep < 16; sweep++) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    if (off <= 1e-40 * scale) break;
    BALM_JROT(a00, a11, a01, a02, a12, v00, v10, v20, v01, v11, v21)  // (p,q)=(0,1), r=2
    BALM_JROT(a00, a22, a02, a01, a12, v00, v10, v20, v02, v12, v22)  // (0,2), r=1
    BALM_JROT(a11, a22, a12, a01, a02, v01, v11, v21, v02, v12, v22)  // (1,2), r=0
  }
  double d0 = a00, d1 = a11, d2 = a22;
  double e0[3] = {v00, v10, v20}, e1[3] = {v01, v11, v21}, e2[3] = {v02, v12, v22};
#define BALM_CSWAP(da, db, ea, eb)                                        \
  if (da > db) {                                                          \
    double t_ = da; da = db; db = t_;                                     \
    _Pragma("unroll") for (int q_ = 0; q_ < 3; q_++) { t_ = ea[q_]; ea[q_] = eb[q_]; eb[q_] = t_; } \
  }
  BALM_CSWAP(d0, d1, e0, e1)
  BALM_CSWAP(d1, d2, e1, e2)
  BALM_CSWAP(d0, d1, e0, e1)
#undef BALM_CSWAP
  lam[0] = d0; lam[1] = d1; lam[2] = d2;
#pragma unroll
  for (int q = 0; q < 3; q++) { u0[q] = e0[q]; u1[q] = e1[q]; u2[q] = e2[q]; }
}

__device__ __forceinline__ void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// Exp (tools.hpp:56-71), row-major output
__device__ __forceinline__ void exp_so3(const double *phi, double *R) {
  const double nrm = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  if (nrm >= 1e-11) {
    const double x = phi[0] / nrm, y = phi[1] / nrm, z = phi[2] / nrm;
    const double s = sin(nrm), c1 = 1.0 - cos(nrm);
    const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    const double KK[9] = {-(y * y + z * z), x * y, x * z, x * y, -(x * x + z * z), y * z, x * z, y * z, -(x * x + y * y)};
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] += s * K[i] + c1 * KK[i];
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

// ---------------- kernel launchers (defined in the .cu files) ----------------
int launch_voxel_stats(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, double *stats_out, bool use_fix,
                       double *residual_out_dev);
int launch_obs_pass(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, bool first_batch);
int launch_obs_colmax(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, bool first_batch);
int launch_obs_int8(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, bool first_batch, const double *sc,
                    int8_t *Gq, int64_t plane_stride, const int *S_dev, int S_alloc, int64_t rows_padded, bool fused,
                    const int *skip, int wpc = 4);
int tensor_obs_and_syrk(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, bool first_batch);
int launch_syrk_f64(balm_ctx *c, int64_t rows, bool first_batch);
int launch_assemble(balm_ctx *c);
int build_pose_major_lists(balm_ctx *c, const int *planes_host);
int launch_csc_batch(balm_ctx *c, int64_t v0, int64_t v1);
int launch_ldlt_solve(balm_ctx *c, double u);
int refine_solution(balm_ctx *c, int *still_bad);
int launch_pose_update(balm_ctx *c, const double *poses_in, const double *dx, double *poses_out);
int launch_gauge(balm_ctx *c, double *poses, int mode);
int launch_synth(balm_ctx *c, int64_t n_voxels, int64_t first_voxel, int pts, double noise, double range,
                 uint64_t seed, const double *poses_gt_dev);
int factor_kernels_setup();
int syrk_f64_setup();
int ldlt_setup();
int covariance_setup();
int tensor_syrk_init(balm_ctx *c);
void tensor_syrk_free(balm_ctx *c);
int tensor_syrk_check(balm_ctx *c);
int tensor_syrk_check_enqueue(balm_ctx *c);
int tensor_syrk_check_finish(balm_ctx *c);
void tensor_syrk_new_problem(balm_ctx *c);
