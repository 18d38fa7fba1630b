/*
 * balm_b200.h -- C ABI of libbalm_b200.so: the B200-native replacement for the inner optimisation loop of
 * BALM 2.0 (VOX_HESS factor evaluation + BALM2::damping_iter()).
 *
 * The reference has no FFI/plugin interface: its seam is the header-level C++ API of
 * /root/reference/src/benchmark/bavoxel.hpp.  Every entry point below names the reference interface it
 * replaces (file:line relative to /root/reference).  include/bavoxel_b200.hpp is the C++ shim that keeps the
 * reference's class/method names on top of this ABI; balm_b200/ (Python) mirrors the same names via ctypes.
 * INTEGRATION.md shows the binding a BALM maintainer adds.
 *
 * Conventions (same as the reference):
 *   poses12 : N x 12 doubles; per pose R column-major (Eigen default) then p    -- IMUST::R, IMUST::p
 *                                                                    (include/tools.hpp:141-201)
 *   obs10   : K x 10 doubles; per (voxel,pose) observation P00,P01,P02,P11,P12,P22,v0,v1,v2,N
 *                                                      -- PointCluster{P,v,N} (include/tools.hpp:290-349)
 *   row_ptr : M+1 int64 CSR offsets voxel -> observations; pose_idx : K int32, ascending inside a voxel.
 *             A slot is present iff PointCluster::N != 0 (bavoxel.hpp:34,332).
 *   coe     : M doubles, push_voxel weight (bavoxel.hpp:42-48)
 *   fix10   : M x 10 doubles or NULL, the fixed / marginalised cluster sig_vecs[a] (bavoxel.hpp:47,441)
 *   H       : n x n doubles, column-major, full symmetric, n = 6N; block (i,j) at rows 6i.., cols 6j..
 *             (bavoxel.hpp:394,416); g : n; unknown order per pose [phi(3), dt(3)] (DVEL = 6, tools.hpp:20)
 * All pointers are HOST pointers unless the name ends in _dev.  No torch/Eigen types cross this boundary.
 * A context owns one CUDA stream and is not re-entrant; distinct contexts are independent.
 */
#ifndef BALM_B200_H
#define BALM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct balm_ctx balm_ctx;

typedef enum {
  BALM_OK = 0,
  BALM_ERR_INVALID = 1,        /* bad argument / call order */
  BALM_ERR_CUDA = 2,           /* CUDA runtime error, see balm_last_error() */
  BALM_ERR_NOT_PD = 3,         /* (H + u D) had a zero/tiny pivot; LM treats it as a rejected step */
  BALM_ERR_TOO_FEW_PLANES = 4, /* reference prints + exit(0) (bavoxel.hpp:1079-1085) */
  BALM_ERR_NCCL = 5,
  BALM_ERR_UNSUPPORTED = 6
} balm_status;

/* Arithmetic of the rank-3M Hessian accumulation  H -= G' G'^T  (bavoxel.hpp:385,392,410-416). */
typedef enum {
  BALM_PREC_FP64 = 0,   /* fp64 DMMA accumulation (BASELINE config C2, "fp64 accumulation") */
  BALM_PREC_TENSOR = 1  /* tcgen05 int8 split-integer (Ozaki) accumulation, exact int32 sums, fp64 result */
} balm_precision;

typedef struct {
  int max_iter;             /* 10  (bavoxel.hpp:1104); benchmark_virtual twin: 20 (benchmark_virtual.cpp:408) */
  double u0;                /* 0.01 (bavoxel.hpp:1087); twin: 0.1 (benchmark_virtual.cpp:380) */
  double v0;                /* 2 */
  double rel_tol;           /* 1e-6 (bavoxel.hpp:1155); < 0 disables the convergence exit (timing runs) */
  int hess_includes_fix;    /* 0: bavoxel.hpp:325 (fix ignored in H,g); 1: benchmark_virtual.cpp:241-243 */
  int gauge_mode;           /* 0: bavoxel.hpp:1159-1164; 1: benchmark_virtual.cpp:472-479; 2: none */
  int min_planes_per_pose;  /* 20 (bavoxel.hpp:1079); 0 disables the precheck */
  int verbose;              /* 1: print the reference's per-iteration line (bavoxel.hpp:1132) */
  int force_hess;           /* 1: re-evaluate H,g every iteration even after a rejected step (benchmarks: makes
                               every LM iteration the same amount of work; the reference reuses H, :1148) */
} balm_lm_opts;

typedef struct {
  double r1, r2, u, v, q, q1; /* exactly the values printed at bavoxel.hpp:1132 */
  int accepted;
  int recomputed_hess;
  int not_pd;
} balm_trace;

/* Per-phase device time (CUDA events on the ctx stream), SUMMED since balm_reset_counters(), with the
 * number of evaluations / solves / residual passes that contributed and the kernels launched. */
typedef struct {
  float ms_stats, ms_obs, ms_slice, ms_syrk, ms_assemble, ms_allreduce, ms_solve, ms_residual, ms_update;
  int launches;   /* kernels launched by the library since balm_reset_counters() */
  int n_eval, n_solve, n_residual;
  int digit_planes; /* tensor path: int8 digit planes used by the last evaluation (3 or 4, chosen on the device) */
  int single_sweeps; /* tensor path: evaluations whose one fused observation sweep (column scales speculated from the
                        previous evaluation) was accepted ... */
  int redone_sweeps; /* ... and those whose scales failed the check, so the digit-plane sweep ran again */
  int refinements;   /* solves whose backward-error check asked for fp64 iterative refinement (indefinite H + uD) */
  int n_stats;       /* evaluations that ran the per-voxel eigen pass themselves (ms_stats sums only these): inside
                        balm_damping_iter the residual pass of the previous step usually hands its results over */
} balm_timings;

const char *balm_last_error(void);
int balm_version(void);

/* ctx lifetime. `device` is the CUDA ordinal. Replaces constructing VOX_HESS + BALM2 (bavoxel.hpp:21,984). */
int balm_create(balm_ctx **out, int n_poses, int device, int precision);
int balm_destroy(balm_ctx *ctx);

/* Registers all plane voxels at once = the sequence of VOX_HESS::push_voxel calls (bavoxel.hpp:30-51) made by
 * OCTO_TREE_NODE::tras_opt (bavoxel.hpp:908-929). Host arrays are copied (H2D) and may be freed on return. */
int balm_set_voxels(balm_ctx *ctx, int64_t n_voxels, const int64_t *row_ptr, const int32_t *pose_idx,
                    const double *obs10, const double *fix10, const double *coe);
/* n_voxels == 0 registers an EMPTY shard (row_ptr = {0}): evaluations contribute zeros and the rank still takes part
 * in every collective of a multi-GPU job. */
/* Same, but the arrays already live in device memory of ctx's GPU (used for HBM-resident timing). */
int balm_set_voxels_dev(balm_ctx *ctx, int64_t n_voxels, const int64_t *row_ptr_dev, const int32_t *pose_idx_dev,
                        const double *obs10_dev, const double *fix10_dev, const double *coe_dev, int64_t n_obs);

/* VOX_HESS::left_evaluate_acc2(xs, head, end, Hess, JacT, residual) (bavoxel.hpp:304-426).
 * H (n*n) and g (n) may be NULL to keep the result on the device only. With a communicator attached
 * (balm_comm_init) the result is the all-reduced sum over ranks = divide_thread_left's reduction
 * (bavoxel.hpp:1049-1056). head/end select a voxel range of THIS rank's voxels ([0,M) = all). */
int balm_evaluate(balm_ctx *ctx, const double *poses12, int64_t head, int64_t end, int include_fix, double *H,
                  double *g, double *residual);
/* VOX_HESS::evaluate_only_residual(xs, residual) (bavoxel.hpp:428-470) */
int balm_residual(balm_ctx *ctx, const double *poses12, double *residual);
/* D = diag(H); dx = (H + u D)^-1 (-g); q1 = 0.5 dx.(u D dx - g) on the H,g of the last balm_evaluate
 * (bavoxel.hpp:1113-1114,1127). not_pd set when the LDL^T met a zero/tiny pivot. dx (n) may be NULL. */
int balm_solve(balm_ctx *ctx, double u, double *dx, double *q1, int *not_pd);
/* BALM2::damping_iter(x_stats, voxhess) (bavoxel.hpp:1069-1166). poses12 updated in place.
 * trace: room for opts->max_iter entries (may be NULL). poses_per_iter: max_iter*N*12 doubles or NULL,
 * receives the accepted pose set after every iteration (before the gauge step) for parity checks. */
int balm_damping_iter(balm_ctx *ctx, double *poses12, const balm_lm_opts *opts, balm_trace *trace, int *n_iters,
                      double *poses_per_iter);
void balm_default_lm_opts(balm_lm_opts *opts);

/* Association on the GPU: raw scans + initial poses -> plane voxels, registered in the
 * ctx exactly as balm_set_voxels would. Replaces the loop  cut_voxel(surf_map, *pl_fulls[i], x_buf[i], i)  over all scans
 * followed by  recut(win_size); tras_opt(voxhess, win_size)  over all root voxels
 * (benchmark_realworld.cpp:187-200 -> bavoxel.hpp:1170-1223, 737-776, 908-929, 30-51).
 * xyz: n_points x 3 float32 in the BODY frame (the PCD floats), frame: n_points pose indices, poses12: N x 12.
 * opts may be NULL (reference constants: voxel_size 1, layer_limit 2, min_ps 15, eigen ratios 1/16). */
typedef struct {
  double voxel_size;          /* bavoxel.hpp:15; benchmark_realworld.launch sets 2 */
  int layer_limit;            /* bavoxel.hpp:8  (0..2) */
  int min_ps;                 /* bavoxel.hpp:12 */
  double eigen_value_array[3];/* bavoxel.hpp:11; benchmark_realworld.cpp:183-185 sets {1/16, 1/16, 1/9} */
} balm_assoc_opts;
void balm_default_assoc_opts(balm_assoc_opts *opts);
int balm_cut_voxels(balm_ctx *ctx, int64_t n_points, const float *xyz, const int32_t *frame, const double *poses12,
                    const balm_assoc_opts *opts, int64_t *n_voxels_out, int64_t *n_obs_out);

/* Sliding-window marginalisation of the registered voxel set =
 * OCTO_TREE_ROOT::marginalize -> OCTO_TREE_NODE::to_margi on every plane leaf (bavoxel.hpp:948-963, 778-816) followed by
 * the re-registration tras_opt -> push_voxel would make (bavoxel.hpp:908-929, 30-51): the clusters of the oldest mg_size
 * scans, transformed by poses12 (the optimised poses, N x 12), are added to the voxel's fix cluster when that holds fewer
 * than 50 points (:790), the remaining scans shift down by mg_size (pose slot i := old slot i + mg_size; the caller shifts
 * its pose array the same way, consistency.cpp:137-140), voxels left with fewer than min_ps points or fewer than two
 * observing scans are not pushed, coe = points left. Everything stays in HBM; the context is re-registered as if
 * balm_set_voxels had been called with the new set (n_poses unchanged: the last mg_size slots are empty until new scans
 * are associated). balm_download_fix copies the fix clusters (M x 10) back. */
int balm_marginalize(balm_ctx *ctx, int mg_size, const double *poses12, int min_ps, int64_t *n_voxels_out,
                     int64_t *n_obs_out);
int balm_download_fix(balm_ctx *ctx, double *fix10);
/* The other half of the sliding window: a NEW scan is associated with the voxel set already in HBM (which must come from
 * balm_cut_voxels: it carries the 63-bit octree key of every voxel). Every point (xyz: n_points x 3 float32, body frame)
 * is transformed by poses12[slot], finds the plane leaf that contains it with the arithmetic of cut_voxel / cut_func
 * (bavoxel.hpp:1170-1223, 700-735), and the body-frame moments of a leaf's new points become its observation in pose
 * slot `slot` (which must lie above every scan already in the window); then EVERY voxel is re-judged as recut does
 * (bavoxel.hpp:737-776 -> judge_eigen :654-699) from its fix cluster and all its clusters transformed by poses12 (N x 12):
 * still planar (eigen ratio below the threshold of its layer), more than min_ps window points, at least two scans ->
 * kept with coe = window points; otherwise dropped. Not done here (it happens at the next full balm_cut_voxels): new
 * roots / leaves from points outside the registered planes, re-splitting a leaf that stopped being planar.
 * n_matched_out: points that fell into a registered voxel. balm_download_keys copies the M voxel keys
 * ([root x|y|z biased by 2^18, 19 bits each | octant at layer 1 | octant at layer 2], 7 = not split) and their octree
 * layers (0..2; a digit 7 is also octant 7, the layer tells the two apart); either pointer may be NULL. */
int balm_append_scan(balm_ctx *ctx, int64_t n_points, const float *xyz, const double *poses12, int slot,
                     int64_t *n_voxels_out, int64_t *n_obs_out, int64_t *n_matched_out);
int balm_download_keys(balm_ctx *ctx, uint64_t *keys, int32_t *layers);

/* Pose-covariance propagation of the consistency experiment:
 *   Rcov_raw = sum over observations of  Ls c_cov Ls^T      -- VOX_HESS::left_jacobian_point + BALM2::multi_second
 *                                                              (src/simulation/BAs_left.hpp:342-473, 995-1023)
 *   Rcov     = H^-1 Rcov_raw H^-T                            -- the tail of BALM2::damping_iter (BAs_left.hpp:1089-1096)
 * at the poses given (normally the output of balm_damping_iter). c_cov81: K x 81 doubles, the 9 x 9 covariance of every
 * cluster's parameters (P00,P01,P02,P11,P12,P22,v0,v1,v2) in the order of obs10, i.e. PointCluster::c_cov
 * (src/simulation/toolss.hpp:288,341); NULL derives it from the cluster moments for isotropic point noise of standard
 * deviation point_noise, which is exactly what PointCluster::push accumulates (toolss.hpp:311-343). include_fix: the fix
 * clusters take part in the voxel statistics and in H (BAs_left.hpp:183-185); without them H is singular (gauge) and the
 * call fails with BALM_ERR_NOT_PD when Rcov is requested. Either output (n x n, column-major, symmetric) may be NULL. */
int balm_pose_covariance(balm_ctx *ctx, const double *poses12, const double *c_cov81, double point_noise, int include_fix,
                         double *Rcov_raw, double *Rcov);

/* Multi-GPU: voxels are sharded across ranks by the caller (each rank registers its own shard); the library
 * all-reduces [H | g | r] with NCCL after every evaluation and the scalar after every residual pass.
 * unique_id: 128 bytes from balm_comm_unique_id() on rank 0, broadcast by the caller (torch.distributed, MPI). */
int balm_comm_unique_id(void *out128);
int balm_comm_init(balm_ctx *ctx, int rank, int world, const void *unique_id128);

/* Instrumentation */
int balm_get_timings(balm_ctx *ctx, balm_timings *out);
int balm_reset_counters(balm_ctx *ctx);
int balm_sync(balm_ctx *ctx);
/* CUDA-event bracket on the ctx stream: begin records, end records + synchronises and returns elapsed ms. */
int balm_timer_begin(balm_ctx *ctx);
int balm_timer_end(balm_ctx *ctx, float *ms);
/* Debug aid: with BALM_DAG_TRACE=1 in the environment, device timestamps (ns) of the last factorisation: 4 per 64-column
 * step of the chain, then 3 per CTA (start, end, tasks executed). */
int balm_debug_dag_trace(balm_ctx *ctx, unsigned long long *out, int max_entries);
/* Device pointers of the last evaluation's H (n*n col-major), g (n) -- for zero-copy consumers. */
int balm_device_views(balm_ctx *ctx, double **H_dev, double **g_dev);

/* Synthetic scene of the benchmark_virtual shape generated directly in HBM
 * (benchmark_virtual.cpp:547-606 scene, :491-503 pose noise, :391 coe): every pose sees every plane,
 * pts_size points per observation. Outputs are device buffers owned by the ctx and registered as its voxels;
 * poses_gt / poses_init (N*12 host doubles) receive ground truth and perturbed start. first_voxel offsets the
 * plane ids so ranks generate disjoint shards of one global scene. */
int balm_synth_virtual(balm_ctx *ctx, int64_t n_voxels, int64_t first_voxel, int pts_size, double point_noise,
                       double surf_range, uint64_t seed, double *poses_gt, double *poses_init);
/* Copies the registered voxels back to host arrays (K*10, K, M+1, M) -- lets tests/bench feed the oracle and
 * the host-buffer (e2e) path with the very scene generated above. Any pointer may be NULL. */
int balm_download_voxels(balm_ctx *ctx, int64_t *row_ptr, int32_t *pose_idx, double *obs10, double *coe);
/* Same for the voxels [head, end) only (row_ptr re-based to 0: end-head+1 entries); n_obs_out receives the number of
 * observations in the range, so a first call with NULL arrays sizes the buffers. Lets a parity test compare
 * balm_evaluate(head, end) of a full-size problem (BASELINE C3: 4 GB of observations) with the CPU oracle on exactly
 * the same voxels. */
int balm_download_voxel_range(balm_ctx *ctx, int64_t head, int64_t end, int64_t *row_ptr, int32_t *pose_idx,
                              double *obs10, double *coe, int64_t *n_obs_out);
int64_t balm_num_obs(balm_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
