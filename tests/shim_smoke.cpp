// shim_smoke.cpp -- the reference's call pattern (benchmark_realworld.cpp:194-218) against include/bavoxel_b200.hpp
// in the Eigen-free build mode: VOX_HESS voxhess; ...push_voxel...; BALM2 opt; opt.damping_iter(x_buf, voxhess);
#include <cmath>
#include <cstdio>
#include <random>
#include "bavoxel_b200.hpp"

static void rodrigues(const double w[3], double R[9]) {  // column-major
  const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (th > 1e-11) {
    const double x = w[0] / th, y = w[1] / th, z = w[2] / th;
    double k[3][3] = {{0, -z, y}, {z, 0, -x}, {-y, x, 0}};
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[r][c] = k[r][c];
  }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double kk = 0;
      for (int t = 0; t < 3; t++) kk += K[r][t] * K[t][c];
      R[c * 3 + r] = (r == c) + std::sin(th) * K[r][c] + (1 - std::cos(th)) * kk;
    }
}

int main(int argc, char **argv) {
  const int N = 12, M = 120, PTS = 30;
  win_size = N;
  std::mt19937 e(7);
  std::uniform_real_distribution<double> U(-0.5, 0.5), C(-2, 2), A(-3.1, 3.1);
  std::normal_distribution<double> G(0.0, 0.01), NR(0.0, 0.02), NT(0.0, 0.05);
  std::vector<IMUST> gt(N), x_buf(N);
  for (int i = 0; i < N; i++) {
    double w[3] = {0.03 * i, -0.02 * i, 0.01 * i};
    rodrigues(w, gt[i].R);
    gt[i].p[0] = 0.05 * i; gt[i].p[1] = 0.02 * i; gt[i].p[2] = -0.03 * i;
    double dw[3] = {NR(e), NR(e), NR(e)}, dR[9];
    rodrigues(dw, dR);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int t = 0; t < 3; t++) s += gt[i].R[t * 3 + r] * dR[c * 3 + t];
        x_buf[i].R[c * 3 + r] = s;
      }
    for (int r = 0; r < 3; r++) x_buf[i].p[r] = gt[i].p[r] + NT(e);
  }
  std::vector<std::vector<PointCluster>> voxels(M, std::vector<PointCluster>(N));
  std::vector<PointCluster> fix(M);
  VOX_HESS voxhess;
  if (argc > 1) voxhess.precision = std::atoi(argv[1]);
  for (int a = 0; a < M; a++) {
    double w[3] = {A(e), A(e), A(e)}, Rp[9], ctr[3] = {C(e), C(e), C(e)};
    rodrigues(w, Rp);
    for (int i = 0; i < N; i++) {
      if ((a + i) % 5 == 0) continue;  // ragged co-visibility
      for (int k = 0; k < PTS; k++) {
        const double l[3] = {U(e), U(e), G(e)};
        double wp[3], b[3];
        for (int r = 0; r < 3; r++) wp[r] = Rp[0 * 3 + r] * l[0] + Rp[1 * 3 + r] * l[1] + Rp[2 * 3 + r] * l[2] + ctr[r] - gt[i].p[r];
        for (int r = 0; r < 3; r++) b[r] = gt[i].R[r * 3 + 0] * wp[0] + gt[i].R[r * 3 + 1] * wp[1] + gt[i].R[r * 3 + 2] * wp[2];
        voxels[a][i].push(b);
      }
    }
    voxhess.push_voxel(&voxels[a], &fix[a], 0.0, 0);
  }
  BALM2 opt;
  std::vector<IMUST> x_ab(N);
  const double r0 = opt.only_residual(x_buf, voxhess, x_ab);
  balm_b200_shim::DenseMat H;
  balm_b200_shim::DenseVec g;
  const double r0b = opt.divide_thread_left(x_buf, voxhess, x_ab, H, g);
  opt.damping_iter(x_buf, voxhess);
  const double r1 = opt.only_residual(x_buf, voxhess, x_ab);
  std::printf("shim_smoke: residual %.6f (eval %.6f) -> %.6f, H(0,0)=%.3f\n", r0, r0b, r1, H(0, 0));
  return (r1 < 0.2 * r0 && std::fabs(r0 - r0b) < 1e-9 * r0) ? 0 : 1;
}
