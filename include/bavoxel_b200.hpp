// bavoxel_b200.hpp -- drop-in C++ shim: the reference's VOX_HESS / BALM2 call surface on top of libbalm_b200.so.
//
// Keeps the class and method names, argument order and in/out conventions of
//   /root/reference/src/benchmark/bavoxel.hpp
//     VOX_HESS::push_voxel              :30-51      VOX_HESS::left_evaluate_acc2   :304-426
//     VOX_HESS::evaluate_only_residual  :428-470    BALM2::divide_thread_left      :1025-1059
//     BALM2::only_residual              :1061-1067  BALM2::damping_iter            :1069-1166
// so that benchmark_realworld.cpp:194-218 / consistency.cpp:154-162 and the downstream projects that copied that
// pattern compile unchanged after replacing  #include "bavoxel.hpp"  by this header for the BA part (the octree /
// association code of bavoxel.hpp:626-965,1170-1223 is untouched and keeps calling push_voxel).
//
// Two build modes:
//   * with Eigen (BALM_B200_WITH_EIGEN defined, tools.hpp included first): uses the reference's own PointCluster,
//     IMUST, Eigen::MatrixXd / VectorXd types;
//   * without Eigen (this image has none): minimal POD stand-ins with the same member names, so the shim and its
//     packing logic can be compiled and tested here (tests/test_shim.py builds tests/shim_smoke.cpp).
// The reference's mutable global `win_size` (bavoxel.hpp:17) must equal x_stats.size(); the shim asserts it.
#ifndef BAVOXEL_B200_HPP
#define BAVOXEL_B200_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>
#include "balm_b200.h"

#ifdef BALM_B200_WITH_EIGEN
// tools.hpp (PointCluster, IMUST, Eigen, PCL's PointType) must already be included by the translation unit.
#ifndef BALM_B200_PLPTR_TYPE  // element type of VOX_HESS::plptrs (bavoxel.hpp:28)
#define BALM_B200_PLPTR_TYPE pcl::PointCloud<PointType>::Ptr
#define BALM_B200_PLPTR_NEW (pcl::PointCloud<PointType>::Ptr(new pcl::PointCloud<PointType>()))
#endif
namespace balm_b200_shim {
typedef Eigen::MatrixXd DenseMat;
typedef Eigen::VectorXd DenseVec;
inline double *mat_data(DenseMat &m) { return m.data(); }
inline double *vec_data(DenseVec &v) { return v.data(); }
inline void mat_resize(DenseMat &m, int n) { m.resize(n, n); }
inline void vec_resize(DenseVec &v, int n) { v.resize(n); }
inline void cluster_pack(const PointCluster &c, double *o) {
  o[0] = c.P(0, 0); o[1] = c.P(0, 1); o[2] = c.P(0, 2); o[3] = c.P(1, 1); o[4] = c.P(1, 2); o[5] = c.P(2, 2);
  o[6] = c.v[0]; o[7] = c.v[1]; o[8] = c.v[2]; o[9] = c.N;
}
inline void pose_pack(const IMUST &x, double *o) {
  std::memcpy(o, x.R.data(), 9 * sizeof(double));  // Eigen default: column-major
  o[9] = x.p[0]; o[10] = x.p[1]; o[11] = x.p[2];
}
inline void pose_unpack(const double *o, IMUST &x) {
  std::memcpy(x.R.data(), o, 9 * sizeof(double));
  x.p[0] = o[9]; x.p[1] = o[10]; x.p[2] = o[11];
}
}  // namespace balm_b200_shim
#else
// ---- POD stand-ins with the reference's member names (include/tools.hpp:141-201, 290-349) ----
struct PointCluster {
  double P[3][3];  // P(r,c)
  double v[3];
  int N;
  PointCluster() { clear(); }
  void clear() { std::memset(P, 0, sizeof(P)); std::memset(v, 0, sizeof(v)); N = 0; }
  void push(const double p[3]) {  // tools.hpp:311-316
    N++;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) P[r][c] += p[r] * p[c];
      v[r] += p[r];
    }
  }
  PointCluster &operator+=(const PointCluster &o) {
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) P[r][c] += o.P[r][c];
      v[r] += o.v[r];
    }
    N += o.N;
    return *this;
  }
};
struct IMUST {
  double R[9];  // column-major, like Eigen::Matrix3d
  double p[3];
  IMUST() { std::memset(R, 0, sizeof(R)); R[0] = R[4] = R[8] = 1; p[0] = p[1] = p[2] = 0; }
};
namespace balm_b200_shim {
struct DenseMat {
  std::vector<double> a;
  int n = 0;
  double &operator()(int r, int c) { return a[(size_t)c * n + r]; }
};
typedef std::vector<double> DenseVec;
inline double *mat_data(DenseMat &m) { return m.a.data(); }
inline double *vec_data(DenseVec &v) { return v.data(); }
inline void mat_resize(DenseMat &m, int n) { m.n = n; m.a.assign((size_t)n * n, 0.0); }
inline void vec_resize(DenseVec &v, int n) { v.assign(n, 0.0); }
inline void cluster_pack(const PointCluster &c, double *o) {
  o[0] = c.P[0][0]; o[1] = c.P[0][1]; o[2] = c.P[0][2]; o[3] = c.P[1][1]; o[4] = c.P[1][2]; o[5] = c.P[2][2];
  o[6] = c.v[0]; o[7] = c.v[1]; o[8] = c.v[2]; o[9] = c.N;
}
inline void pose_pack(const IMUST &x, double *o) { std::memcpy(o, x.R, 9 * sizeof(double)); std::memcpy(o + 9, x.p, 3 * sizeof(double)); }
inline void pose_unpack(const double *o, IMUST &x) { std::memcpy(x.R, o, 9 * sizeof(double)); std::memcpy(x.p, o + 9, 3 * sizeof(double)); }
}  // namespace balm_b200_shim
#ifndef BALM_B200_NO_WIN_SIZE
static int win_size = 20;  // bavoxel.hpp:17
#endif
#define BALM_B200_PLPTR_TYPE std::shared_ptr<void>
#define BALM_B200_PLPTR_NEW (std::shared_ptr<void>())
#endif  // BALM_B200_WITH_EIGEN

class VOX_HESS {
 public:
  // same fields as the reference (bavoxel.hpp:24-26); pointers are borrowed, exactly as there
  std::vector<const PointCluster *> sig_vecs;
  std::vector<const std::vector<PointCluster> *> plvec_voxels;
  std::vector<double> coeffs, coeffs_back;
  std::vector<BALM_B200_PLPTR_TYPE> plptrs;  // bavoxel.hpp:28 (one empty cloud per voxel, :49-50; unused by the BA)

  int device = 0;
  int precision = BALM_PREC_TENSOR;

  // bavoxel.hpp:30-51
  void push_voxel(const std::vector<PointCluster> *vec_orig, const PointCluster *fix, double feat_eigen, int layer) {
    (void)feat_eigen; (void)layer;
    int process_size = 0;
    for (int i = 0; i < win_size; i++)
      if ((*vec_orig)[i].N != 0) process_size++;
    if (process_size < 2) return;  // :37
    double coe = 0;
    for (int j = 0; j < win_size; j++) coe += (*vec_orig)[j].N;  // :42-44
    plvec_voxels.push_back(vec_orig);
    sig_vecs.push_back(fix);
    coeffs.push_back(coe);
    plptrs.push_back(BALM_B200_PLPTR_NEW);
    dirty_ = true;
  }

  // bavoxel.hpp:304-426. A plain member function on the reference's own dense types (Eigen::MatrixXd / VectorXd with
  // Eigen), so that  &VOX_HESS::left_evaluate_acc2  binds -- the reference's divide_thread_left hands exactly that to
  // std::thread with a voxel range per thread (bavoxel.hpp:1044-1047). A context owns one CUDA stream and is not
  // re-entrant: concurrent calls are serialised here (each range is still evaluated on the whole GPU, in the order the
  // threads arrive; the caller's ordered reduction of the per-thread results is untouched).
  void left_evaluate_acc2(const std::vector<IMUST> &xs, int head, int end, balm_b200_shim::DenseMat &Hess,
                          balm_b200_shim::DenseVec &JacT, double &residual) {
    std::lock_guard<std::mutex> lock(mu_);
    balm_ctx *c = context((int)xs.size());
    const int n = 6 * (int)xs.size();
    balm_b200_shim::mat_resize(Hess, n);
    balm_b200_shim::vec_resize(JacT, n);
    std::vector<double> poses = pack(xs);
    check(balm_evaluate(c, poses.data(), head, end, /*include_fix=*/0, balm_b200_shim::mat_data(Hess),
                        balm_b200_shim::vec_data(JacT), &residual));
  }

  // bavoxel.hpp:428-470
  void evaluate_only_residual(const std::vector<IMUST> &xs, double &residual) {
    std::lock_guard<std::mutex> lock(mu_);
    balm_ctx *c = context((int)xs.size());
    std::vector<double> poses = pack(xs);
    check(balm_residual(c, poses.data(), &residual));
  }

  ~VOX_HESS() { if (ctx_) balm_destroy(ctx_); }

  // ---- shim internals ----
  std::mutex &mutex() { return mu_; }
  balm_ctx *context(int n_poses) {  // callers hold mu_
    if (n_poses != win_size) {
      std::fprintf(stderr, "balm_b200: win_size (%d) must equal x_stats.size() (%d)\n", win_size, n_poses);
      std::abort();
    }
    if (ctx_ && !dirty_ && n_poses == n_poses_) return ctx_;
    if (ctx_) { balm_destroy(ctx_); ctx_ = nullptr; }
    check(balm_create(&ctx_, n_poses, device, precision));
    // pack the pointer-of-vectors into CSR once (association is static during BA, benchmark_realworld.cpp:187-218)
    std::vector<int64_t> row_ptr(1, 0);
    std::vector<int32_t> pose_idx;
    std::vector<double> obs, fix;
    bool any_fix = false;
    for (size_t a = 0; a < plvec_voxels.size(); a++) {
      const std::vector<PointCluster> &vec = *plvec_voxels[a];
      for (int i = 0; i < win_size; i++)
        if (vec[i].N != 0) {
          pose_idx.push_back(i);
          obs.resize(obs.size() + 10);
          balm_b200_shim::cluster_pack(vec[i], &obs[obs.size() - 10]);
        }
      row_ptr.push_back((int64_t)pose_idx.size());
      fix.resize(fix.size() + 10, 0.0);
      if (sig_vecs[a]) {
        balm_b200_shim::cluster_pack(*sig_vecs[a], &fix[fix.size() - 10]);
        any_fix = any_fix || sig_vecs[a]->N != 0;
      }
    }
    check(balm_set_voxels(ctx_, (int64_t)plvec_voxels.size(), row_ptr.data(), pose_idx.data(), obs.data(),
                          any_fix ? fix.data() : nullptr, coeffs.data()));
    n_poses_ = n_poses;
    dirty_ = false;
    return ctx_;
  }
  static std::vector<double> pack(const std::vector<IMUST> &xs) {
    std::vector<double> p(12 * xs.size());
    for (size_t i = 0; i < xs.size(); i++) balm_b200_shim::pose_pack(xs[i], &p[12 * i]);
    return p;
  }
  static void check(int st) {
    if (st != BALM_OK) {
      std::fprintf(stderr, "balm_b200: status %d: %s\n", st, balm_last_error());
      std::abort();
    }
  }

 private:
  balm_ctx *ctx_ = nullptr;
  int n_poses_ = 0;
  bool dirty_ = true;
  std::mutex mu_;
};

class BALM2 {
 public:
  BALM2() {}

  // bavoxel.hpp:1025-1059. The 4-way std::thread split and ordered reduction happen inside the library
  // (CTAs / GPUs instead of threads); x_ab is unused there as well (:1096-1102).
  double divide_thread_left(std::vector<IMUST> &x_stats, VOX_HESS &voxhess, std::vector<IMUST> &x_ab,
                            balm_b200_shim::DenseMat &Hess, balm_b200_shim::DenseVec &JacT) {
    (void)x_ab;
    double residual = 0;
    voxhess.left_evaluate_acc2(x_stats, 0, (int)voxhess.plvec_voxels.size(), Hess, JacT, residual);
    return residual;
  }

  // bavoxel.hpp:1061-1067
  double only_residual(std::vector<IMUST> &x_stats, VOX_HESS &voxhess, std::vector<IMUST> &x_ab) {
    (void)x_ab;
    double residual2 = 0;
    voxhess.evaluate_only_residual(x_stats, residual2);
    return residual2;
  }

  // bavoxel.hpp:1069-1166
  void damping_iter(std::vector<IMUST> &x_stats, VOX_HESS &voxhess) {
    std::lock_guard<std::mutex> lock(voxhess.mutex());
    balm_ctx *c = voxhess.context((int)x_stats.size());
    balm_lm_opts o;
    balm_default_lm_opts(&o);  // u=0.01, v=2, 10 iterations, 1e-6, >=20 planes per pose: the reference constants
    o.verbose = 1;             // prints the reference's per-iteration line (:1132)
    std::vector<double> poses = VOX_HESS::pack(x_stats);
    int n_iters = 0;
    const int st = balm_damping_iter(c, poses.data(), &o, nullptr, &n_iters, nullptr);
    if (st == BALM_ERR_TOO_FEW_PLANES) {  // :1079-1085
      std::printf("Initial error too large.\n");
      std::printf("Please loose plane determination criteria for more planes.\n");
      std::printf("The optimization is terminated.\n");
      std::exit(0);
    }
    VOX_HESS::check(st);
    for (size_t i = 0; i < x_stats.size(); i++) balm_b200_shim::pose_unpack(&poses[12 * i], x_stats[i]);
  }
};

#endif
