// shim_eigen_branch.cpp -- compiles the BALM_B200_WITH_EIGEN branch of include/bavoxel_b200.hpp (against the Eigen
// stand-in of tests/eigen_stub: this image has no Eigen) and runs the REFERENCE'S OWN divide_thread_left on top of it:
// four std::threads, each given  &VOX_HESS::left_evaluate_acc2  and a voxel range, partial results summed in thread order
// (/root/reference/src/benchmark/bavoxel.hpp:1025-1059). The result must equal one call over the whole range.
#include <cmath>
#include <cstdio>
#include <random>
#include <thread>
#include <Eigen/Core>

// what tools.hpp provides in a real build: PointCluster / IMUST with these member names, and the global win_size
struct PointCluster {
  Eigen::Matrix3d P;
  Eigen::Vector3d v;
  int N = 0;
  void push(const double p[3]) {
    N++;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) P(r, c) += p[r] * p[c]; v[r] += p[r]; }
  }
};
struct IMUST {
  Eigen::Matrix3d R;
  Eigen::Vector3d p;
  IMUST() { R.setIdentity(); }
};
int win_size = 20;
#define BALM_B200_WITH_EIGEN
#define BALM_B200_PLPTR_TYPE std::shared_ptr<int>        // stands in for pcl::PointCloud<PointType>::Ptr
#define BALM_B200_PLPTR_NEW (std::make_shared<int>(0))
#include "bavoxel_b200.hpp"
using namespace std;

// verbatim structure of BALM2::divide_thread_left (bavoxel.hpp:1025-1059)
static double divide_thread_left_reference(vector<IMUST> &x_stats, VOX_HESS &voxhess, Eigen::MatrixXd &Hess, Eigen::VectorXd &JacT) {
  int thd_num = 4;
  double residual = 0;
  Hess.setZero(); JacT.setZero();
  vector<Eigen::MatrixXd> hessians(thd_num);
  vector<Eigen::VectorXd> jacobins(thd_num);
  for (int i = 0; i < thd_num; i++) { hessians[i].resize(6 * win_size, 6 * win_size); jacobins[i].resize(6 * win_size); }
  int tthd_num = thd_num;
  vector<double> resis(tthd_num, 0);
  int g_size = voxhess.plvec_voxels.size();
  if (g_size < tthd_num) tthd_num = 1;
  vector<thread *> mthreads(tthd_num);
  double part = 1.0 * g_size / tthd_num;
  for (int i = 0; i < tthd_num; i++)
    mthreads[i] = new thread(&VOX_HESS::left_evaluate_acc2, &voxhess, x_stats, part * i, part * (i + 1), ref(hessians[i]),
                             ref(jacobins[i]), ref(resis[i]));
  for (int i = 0; i < tthd_num; i++) {
    mthreads[i]->join();
    Hess += hessians[i];
    JacT += jacobins[i];
    residual += resis[i];
    delete mthreads[i];
  }
  return residual;
}

int main(int argc, char **argv) {
  const int N = 10, M = 90, PTS = 20;
  win_size = N;
  mt19937 e(11);
  uniform_real_distribution<double> U(-0.5, 0.5), C(-2, 2);
  normal_distribution<double> G(0.0, 0.01);
  vector<IMUST> x(N);
  for (int i = 0; i < N; i++) { x[i].p[0] = 0.04 * i; x[i].p[1] = -0.02 * i; x[i].p[2] = 0.01 * i; }
  vector<vector<PointCluster>> voxels(M, vector<PointCluster>(N));
  vector<PointCluster> fix(M);
  VOX_HESS voxhess;
  if (argc > 1) voxhess.precision = atoi(argv[1]);
  for (int a = 0; a < M; a++) {
    const double ctr[3] = {C(e), C(e), C(e)};
    const int ax = a % 3;  // axis-aligned planes
    for (int i = 0; i < N; i++) {
      if ((a + 2 * i) % 7 == 0) continue;
      for (int k = 0; k < PTS; k++) {
        double l[3] = {U(e), U(e), U(e)};
        l[ax] = G(e);
        const double b[3] = {l[0] + ctr[0] - x[i].p[0] + 0.01 * i, l[1] + ctr[1] - x[i].p[1], l[2] + ctr[2] - x[i].p[2]};
        voxels[a][i].push(b);
      }
    }
    voxhess.push_voxel(&voxels[a], &fix[a], 0.0, 0);
  }
  if (voxhess.plptrs.size() != voxhess.plvec_voxels.size()) return 2;  // bavoxel.hpp:49-50
  Eigen::MatrixXd H4, H1;
  Eigen::VectorXd g4, g1;
  H4.resize(6 * N, 6 * N); g4.resize(6 * N);
  const double r4 = divide_thread_left_reference(x, voxhess, H4, g4);
  BALM2 opt;
  vector<IMUST> x_ab(N);
  const double r1 = opt.divide_thread_left(x, voxhess, x_ab, H1, g1);
  double dH = 0, mH = 0, dg = 0, mg = 0;
  for (int r = 0; r < 6 * N; r++) {
    dg = fmax(dg, fabs(g4[r] - g1[r])); mg = fmax(mg, fabs(g1[r]));
    for (int c = 0; c < 6 * N; c++) { dH = fmax(dH, fabs(H4(r, c) - H1(r, c))); mH = fmax(mH, fabs(H1(r, c))); }
  }
  printf("shim_eigen_branch: r %.9f vs %.9f, max|dH|/max|H| %.2e, max|dg|/max|g| %.2e\n", r4, r1, dH / mH, dg / mg);
  opt.damping_iter(x, voxhess);
  return (fabs(r4 - r1) <= 1e-12 * fabs(r1) && dH <= 2e-8 * mH && dg <= 1e-12 * mg) ? 0 : 1;
}
