/*
 * balm_oracle.h -- CPU restatement of the BALM 2.0 BA hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for balm_b200: a plain-C restatement of
 *   /root/reference/include/tools.hpp:56-71,92-97,99-106,290-349   (Exp, Log, hat, PointCluster)
 *   /root/reference/src/benchmark/bavoxel.hpp:30-51                (push_voxel weight)
 *   /root/reference/src/benchmark/bavoxel.hpp:304-426              (left_evaluate_acc2)
 *   /root/reference/src/benchmark/bavoxel.hpp:428-470              (evaluate_only_residual)
 *   /root/reference/src/benchmark/bavoxel.hpp:1025-1059            (divide_thread_left, 4 threads)
 *   /root/reference/src/benchmark/bavoxel.hpp:1069-1166            (damping_iter)
 *   /root/reference/src/benchmark/benchmark_virtual.cpp:218-482    (single-thread twin, fix cluster included)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libbalm_b200.so) never links or calls it.
 *
 * PARITY PIN: the reference has no tests, golden vectors or fixed seeds, and Eigen / PCL / ROS are
 * absent from this image, so the reference cannot be built as shipped.  Its OWN headers do compile, verbatim and where they
 * lie, against stand-ins for the few Eigen / PCL / ROS names they use (oracle/ref_stubs/, oracle/ref_harness*.cpp ->
 * oracle/_ref/libbalm_ref*.so, recipe in oracle/Makefile): tests/test_reference_pin.py checks this oracle and the numpy
 * restatements against that code (Exp/Log, left_evaluate_acc2, evaluate_only_residual, divide_thread_left, the solve line,
 * damping_iter end to end, cut_voxel/recut/tras_opt, marginalize, the simulation's left_jacobian_point / covariance push).
 * STILL UNPINNED: Eigen's own arithmetic -- SelfAdjointEigenSolver<Matrix3d> (bavoxel.hpp:345,452) and MatrixXd::ldlt()
 * (bavoxel.hpp:1114); version unpinned, "3.3.7 recommended", README.md:29 -- is a stand-in there too; H, g, r are invariant
 * to the eigenvector sign and the solve is checked by its residual, and that boundary is covered by analytic identities,
 * finite differences and an independent numpy restatement (tests/test_oracle.py).
 *
 * Data layout (shared with include/balm_b200.h):
 *   poses12 : N x 12 doubles, per pose R column-major (R00,R10,R20,R01,...,R22) then p (3)   [IMUST R,p]
 *   obs10   : K x 10 doubles, per observation P00,P01,P02,P11,P12,P22,v0,v1,v2,N              [PointCluster]
 *   row_ptr : M+1 int64, CSR voxel -> observation range; pose_idx : K int32 (ascending inside a voxel)
 *   coe     : M doubles (push_voxel weight = sum of N over the voxel, bavoxel.hpp:42-44)
 *   fix10   : M x 10 doubles or NULL (the "fix" / marginalised cluster, sig_vecs)
 *   H       : n x n doubles column-major full symmetric, n = 6N; g : n; dx order per pose [phi(3), dt(3)]
 */
#ifndef BALM_ORACLE_H
#define BALM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int n_poses;
  int64_t n_voxels;
  const int64_t *row_ptr;
  const int32_t *pose_idx;
  const double *obs10;
  const double *fix10; /* may be NULL */
  const double *coe;
} orc_problem;

typedef struct {
  int max_iter;        /* 10 in bavoxel.hpp:1104, 20 in benchmark_virtual.cpp:408 */
  double u0;           /* 0.01 (bavoxel.hpp:1087) / 0.1 (benchmark_virtual.cpp:380) */
  double v0;           /* 2 */
  double rel_tol;      /* 1e-6 (bavoxel.hpp:1155); <0 disables the convergence exit (timing runs) */
  int hess_includes_fix; /* 0 = bavoxel.hpp:325 (C starts at zero); 1 = benchmark_virtual.cpp:241-243 */
  int threads;         /* 4 = divide_thread_left (bavoxel.hpp:1027); 1 = benchmark_virtual twin */
  int gauge_mode;      /* 0 = bavoxel.hpp:1159-1164 (all poses relative to pose 0);
                          1 = benchmark_virtual.cpp:472-479 (same, pose 0 forced to exact identity);
                          2 = leave poses untouched (per-iteration parity runs) */
  int min_planes_per_pose; /* 20 (bavoxel.hpp:1079); 0 disables the precheck */
} orc_lm_opts;

typedef struct {
  double r1, r2, u, v, q, q1; /* the values printed at bavoxel.hpp:1132 (u,v before the update) */
  int accepted;
  int recomputed_hess;
} orc_trace;

/* tools.hpp:56-71 */
void orc_exp_so3(const double phi[3], double R_colmajor[9]);
/* tools.hpp:92-97 */
void orc_log_so3(const double R_colmajor[9], double phi[3]);
/* symmetric 3x3 eigen-decomposition, eigenvalues ascending, eigenvectors in columns (col-major U).
 * Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> (bavoxel.hpp:345). Cyclic Jacobi. */
void orc_eig3(const double A_colmajor[9], double lam[3], double U_colmajor[9]);
/* tools.hpp:333-339 on packed 10-tuples */
void orc_cluster_transform(const double in10[10], const double pose12[12], double out10[10]);

/* bavoxel.hpp:304-426 over voxels [head,end). H (n x n col-major), g (n) are overwritten. */
void orc_left_evaluate_acc2(const orc_problem *pb, const double *poses12, int64_t head, int64_t end,
                            int include_fix, double *H, double *g, double *residual);
/* bavoxel.hpp:428-470 */
void orc_evaluate_only_residual(const orc_problem *pb, const double *poses12, double *residual);
/* bavoxel.hpp:1025-1059: split [0,M) into `threads` contiguous ranges, evaluate, sum in thread order */
double orc_divide_thread_left(const orc_problem *pb, const double *poses12, int threads, int include_fix,
                              double *H, double *g);
/* Dense LDL^T with diagonal pivoting + solve, stands in for Eigen's MatrixXd::ldlt().solve()
 * (bavoxel.hpp:1114). A (n x n col-major, full symmetric) is not modified. Returns 0, or 1 if a zero pivot
 * was met (Eigen would silently continue; so do we). */
int orc_ldlt_solve(int n, const double *A, const double *b, double *x);
/* bavoxel.hpp:1116-1126 */
void orc_left_update(int n_poses, const double *poses12, const double *dx, double *poses12_out);
/* bavoxel.hpp:1069-1166. poses12 updated in place. trace has room for opts->max_iter entries.
 * poses_per_iter (may be NULL): (max_iter) x N x 12, the accepted pose set after each iteration
 * (before the final gauge step). Returns 0 ok, 4 too few planes (reference: exit(0)). */
int orc_damping_iter(const orc_problem *pb, double *poses12, const orc_lm_opts *opts, orc_trace *trace,
                     int *n_iters, double *poses_per_iter);
/* One LM trial step from given H,g (bavoxel.hpp:1113-1127): dx, trial poses and q1. */
int orc_lm_step(int n_poses, const double *H, const double *g, double u, const double *poses12, double *dx,
                double *poses12_trial, double *q1);
/* benchmark_virtual.cpp:48-61 */
void orc_rmse(int n_poses, const double *poses_est, const double *poses_gt, double *rot, double *tran);

#ifdef __cplusplus
}
#endif
#endif
