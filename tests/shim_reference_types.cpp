// shim_reference_types.cpp -- include/bavoxel_b200.hpp in its BALM_B200_WITH_EIGEN mode on top of the REFERENCE'S OWN
// include/tools.hpp (PointCluster, IMUST, PointType; compiled where it lies under /root/reference against the Eigen / PCL
// stand-ins of oracle/ref_stubs -- this image has neither library), i.e. the translation unit a BALM maintainer gets after
// replacing  #include "bavoxel.hpp"  by the shim. Builds the voxel set through the reference's PointCluster::push and
// VOX_HESS::push_voxel, runs divide_thread_left / damping_iter (benchmark_realworld.cpp:194-218 call pattern).
// Built only where /root/reference exists (tests/test_shim.py); the binary travels to the GPU box and runs there.
#include <cmath>
#include <cstdio>
#include <random>
#include "tools.hpp"

int win_size = 20;  // bavoxel.hpp:17
#define BALM_B200_WITH_EIGEN
#include "bavoxel_b200.hpp"
using namespace std;

int main(int argc, char **argv) {
  const int N = 12, M = 120, PTS = 20;
  win_size = N;
  mt19937 e(5);
  uniform_real_distribution<double> U(-0.5, 0.5), C(-2, 2);
  normal_distribution<double> G(0.0, 0.01);
  vector<IMUST> x(N);
  for (int i = 0; i < N; i++) x[i].p = Eigen::Vector3d(0.04 * i, -0.02 * i, 0.01 * i);
  vector<vector<PointCluster>> voxels(M, vector<PointCluster>(N));
  vector<PointCluster> fix(M);
  VOX_HESS voxhess;
  if (argc > 1) voxhess.precision = atoi(argv[1]);
  for (int a = 0; a < M; a++) {
    const Eigen::Vector3d ctr(C(e), C(e), C(e));
    const int ax = a % 3;
    for (int i = 0; i < N; i++) {
      if ((a + 2 * i) % 7 == 0) continue;
      for (int k = 0; k < PTS; k++) {
        Eigen::Vector3d l(U(e), U(e), U(e));
        l[ax] = G(e);
        Eigen::Vector3d b = l + ctr - x[i].p;   // body-frame point of a scan at a slightly wrong pose estimate
        b[0] += 0.01 * i;
        voxels[a][i].push(b);                   // PointCluster::push, tools.hpp:306-311
      }
    }
    voxhess.push_voxel(&voxels[a], &fix[a], 0.0, 0);
  }
  BALM2 opt;
  vector<IMUST> x_ab(N);
  Eigen::MatrixXd H;
  Eigen::VectorXd g;
  const double r0 = opt.divide_thread_left(x, voxhess, x_ab, H, g);
  opt.damping_iter(x, voxhess);
  const double r1 = opt.only_residual(x, voxhess, x_ab);
  printf("shim_reference_types: residual %.6f -> %.6f\n", r0, r1);
  return (std::isfinite(r1) && r1 < 0.5 * r0) ? 0 : 1;
}
