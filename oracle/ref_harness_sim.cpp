// oracle/ref_harness_sim.cpp -- C entry points around the REFERENCE'S OWN consistency-experiment code, compiled from where it
// lies:  #include "BAs_left.hpp"  ->  /root/reference/src/simulation/BAs_left.hpp (+ toolss.hpp), against oracle/ref_stubs.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libbalm_ref_sim.so). Pins tests/numpy_cov.py (and through it balm_pose_covariance)
// against  VOX_HESS::left_jacobian_point  (:342-473),  PointCluster::push's c_cov accumulation (toolss.hpp:311-343)  and the
// sim's left_evaluate_acc2 (:162-290, the variant that includes the fix cluster). Nothing of the reference is restated here.
#include <stdint.h>
#include "BAs_left.hpp"

namespace {
void unpack_pose(const double *p12, IMUST &x) {
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) x.R(r, c) = p12[c * 3 + r];
  for (int r = 0; r < 3; r++) x.p[r] = p12[9 + r];
}
struct Problem {
  int N;
  std::vector<std::vector<PointCluster>> slots;
  std::vector<PointCluster> fix;
  VOX_HESS vox;
  std::vector<IMUST> xs;
};
}  // namespace

extern "C" {

void *ref_sim_problem_create(int N, int64_t M, const int64_t *row_ptr, const int32_t *pose_idx, const double *obs10,
                             const double *fix10, const double *ccov81) {
  win_size = N;
  Problem *p = new Problem();
  p->N = N;
  p->slots.assign((size_t)M, std::vector<PointCluster>((size_t)N));
  p->fix.assign((size_t)M, PointCluster());
  auto unpack = [](const double *o, PointCluster &c) {
    c.P(0, 0) = o[0]; c.P(0, 1) = c.P(1, 0) = o[1]; c.P(0, 2) = c.P(2, 0) = o[2];
    c.P(1, 1) = o[3]; c.P(1, 2) = c.P(2, 1) = o[4]; c.P(2, 2) = o[5];
    c.v[0] = o[6]; c.v[1] = o[7]; c.v[2] = o[8]; c.N = (int)o[9];
  };
  for (int64_t a = 0; a < M; a++) {
    for (int64_t s = row_ptr[a]; s < row_ptr[a + 1]; s++) {
      PointCluster &c = p->slots[a][pose_idx[s]];
      unpack(obs10 + 10 * s, c);
      if (ccov81) for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) c.c_cov(r, q) = ccov81[(size_t)s * 81 + r * 9 + q];
    }
    if (fix10) unpack(fix10 + 10 * a, p->fix[a]);
  }
  for (int64_t a = 0; a < M; a++) p->vox.push_voxel(&p->slots[a], &p->fix[a], 0.0, 0);
  p->xs.resize(N);
  return p;
}
void ref_sim_problem_destroy(void *h) { delete static_cast<Problem *>(h); }

// VOX_HESS::left_jacobian_point(xs, beg, end, Rcov)
void ref_sim_left_jacobian_point(void *h, const double *poses12, int beg, int end, double *Rcov) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  for (int i = 0; i < p->N; i++) unpack_pose(poses12 + 12 * i, p->xs[i]);
  const int n = 6 * p->N;
  Eigen::MatrixXd R(n, n);
  p->vox.left_jacobian_point(p->xs, beg, end, R);
  for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) Rcov[(size_t)c * n + r] = R(r, c);
}
// the sim's VOX_HESS::left_evaluate_acc2 (fix cluster included in C, BAs_left.hpp:183-185)
void ref_sim_left_evaluate_acc2(void *h, const double *poses12, int head, int end, double *H, double *g, double *residual) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  for (int i = 0; i < p->N; i++) unpack_pose(poses12 + 12 * i, p->xs[i]);
  const int n = 6 * p->N;
  Eigen::MatrixXd Hess(n, n);
  Eigen::VectorXd JacT(n);
  double r = 0;
  p->vox.left_evaluate_acc2(p->xs, head, end, Hess, JacT, r);
  for (int c = 0; c < n; c++) for (int rr = 0; rr < n; rr++) H[(size_t)c * n + rr] = Hess(rr, c);
  for (int i = 0; i < n; i++) g[i] = JacT[i];
  *residual = r;
}
// PointCluster::push with POINT_NOISE (toolss.hpp:311-343): moments and c_cov of a cluster built point by point
void ref_sim_push_points(int64_t n, const double *xyz, double point_noise, double *obs10, double *ccov81) {
  pnoise = point_noise;
  PointCluster c;
  for (int64_t i = 0; i < n; i++) c.push(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
  obs10[0] = c.P(0, 0); obs10[1] = c.P(0, 1); obs10[2] = c.P(0, 2); obs10[3] = c.P(1, 1); obs10[4] = c.P(1, 2); obs10[5] = c.P(2, 2);
  obs10[6] = c.v[0]; obs10[7] = c.v[1]; obs10[8] = c.v[2]; obs10[9] = c.N;
  for (int r = 0; r < 9; r++) for (int q = 0; q < 9; q++) ccov81[r * 9 + q] = c.c_cov(r, q);
}

}  // extern "C"
