// oracle/ref_harness.cpp -- C entry points around the REFERENCE'S OWN hot-path code, compiled from where it lies:
//   #include "bavoxel.hpp"  ->  /root/reference/src/benchmark/bavoxel.hpp (+ /root/reference/include/tools.hpp)
// against oracle/ref_stubs (Eigen / PCL / ROS are absent from this image: the stand-ins restate Eigen's generic dense
// arithmetic and nothing of the reference). TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into
// oracle/_ref/libbalm_ref.so when /root/reference is present; tests/test_reference_pin.py checks the restated oracles
// (oracle/balm_oracle.c, tests/assoc_ref.py) against what this library -- the reference's code -- computes.
// Nothing here restates the reference: the functions below only move data in and out of its classes
//   VOX_HESS::{push_voxel, left_evaluate_acc2, evaluate_only_residual}   bavoxel.hpp:30-51, 304-426, 428-470
//   BALM2::{divide_thread_left, only_residual, damping_iter}              bavoxel.hpp:1025-1166
//   cut_voxel, OCTO_TREE_ROOT/NODE::{recut, tras_opt, marginalize}        bavoxel.hpp:1170-1223, 737-776, 908-929, 948-963
// and walk the octree (read-only) to label every pushed voxel with its 63-bit node key.
#include <ros/ros.h>
#include <stdint.h>
#include <map>
#include "bavoxel.hpp"

namespace {
void unpack_pose(const double *p12, IMUST &x) {
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) x.R(r, c) = p12[c * 3 + r];
  for (int r = 0; r < 3; r++) x.p[r] = p12[9 + r];
}
void pack_pose(const IMUST &x, double *p12) {
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) p12[c * 3 + r] = x.R(r, c);
  for (int r = 0; r < 3; r++) p12[9 + r] = x.p[r];
}
void unpack_cluster(const double *o, PointCluster &c) {
  c.P(0, 0) = o[0]; c.P(0, 1) = c.P(1, 0) = o[1]; c.P(0, 2) = c.P(2, 0) = o[2];
  c.P(1, 1) = o[3]; c.P(1, 2) = c.P(2, 1) = o[4]; c.P(2, 2) = o[5];
  c.v[0] = o[6]; c.v[1] = o[7]; c.v[2] = o[8]; c.N = (int)o[9];
}
void pack_cluster(const PointCluster &c, double *o) {
  o[0] = c.P(0, 0); o[1] = c.P(0, 1); o[2] = c.P(0, 2); o[3] = c.P(1, 1); o[4] = c.P(1, 2); o[5] = c.P(2, 2);
  o[6] = c.v[0]; o[7] = c.v[1]; o[8] = c.v[2]; o[9] = c.N;
}

// A registered problem in the reference's own containers: per voxel a vector<PointCluster> of win_size slots + a fix cluster
struct Problem {
  int N;
  std::vector<std::vector<PointCluster>> slots;
  std::vector<PointCluster> fix;
  VOX_HESS vox;
  std::vector<IMUST> xs;
};
Problem *make_problem(int N, int64_t M, const int64_t *row_ptr, const int32_t *pose_idx, const double *obs10,
                      const double *fix10) {
  win_size = N;  // the reference's mutable global (bavoxel.hpp:17): every evaluator sizes its buffers from it
  Problem *p = new Problem();
  p->N = N;
  p->slots.assign((size_t)M, std::vector<PointCluster>((size_t)N));
  p->fix.assign((size_t)M, PointCluster());
  for (int64_t a = 0; a < M; a++) {
    for (int64_t s = row_ptr[a]; s < row_ptr[a + 1]; s++) unpack_cluster(obs10 + 10 * s, p->slots[a][pose_idx[s]]);
    if (fix10) unpack_cluster(fix10 + 10 * a, p->fix[a]);
  }
  for (int64_t a = 0; a < M; a++) p->vox.push_voxel(&p->slots[a], &p->fix[a], 0.0, 0);
  p->xs.resize(N);
  return p;
}
void set_poses(Problem *p, const double *poses12) { for (int i = 0; i < p->N; i++) unpack_pose(poses12 + 12 * i, p->xs[i]); }
}  // namespace

extern "C" {

void *ref_problem_create(int N, int64_t M, const int64_t *row_ptr, const int32_t *pose_idx, const double *obs10,
                         const double *fix10) { return make_problem(N, M, row_ptr, pose_idx, obs10, fix10); }
void ref_problem_destroy(void *h) { delete static_cast<Problem *>(h); }
int64_t ref_problem_pushed(void *h) { return (int64_t) static_cast<Problem *>(h)->vox.plvec_voxels.size(); }
void ref_problem_coeffs(void *h, double *coe) { Problem *p = static_cast<Problem *>(h); for (size_t i = 0; i < p->vox.coeffs.size(); i++) coe[i] = p->vox.coeffs[i]; }

// VOX_HESS::left_evaluate_acc2(xs, head, end, Hess, JacT, residual)
void ref_left_evaluate_acc2(void *h, const double *poses12, int head, int end, double *H, double *g, double *residual) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  set_poses(p, poses12);
  const int n = 6 * p->N;
  Eigen::MatrixXd Hess(n, n);
  Eigen::VectorXd JacT(n);
  double r = 0;
  p->vox.left_evaluate_acc2(p->xs, head, end, Hess, JacT, r);
  for (int c = 0; c < n; c++) for (int rr = 0; rr < n; rr++) H[(size_t)c * n + rr] = Hess(rr, c);
  for (int i = 0; i < n; i++) g[i] = JacT[i];
  *residual = r;
}
// VOX_HESS::acc_evaluate2(xs, head, end, Hess, JacT, residual): the right-update evaluator (bavoxel.hpp:53-158; dead code in
// the reference -- its only call site is commented out, :1108 -- but an independent derivation of residual and gradient)
void ref_acc_evaluate2(void *h, const double *poses12, int head, int end, double *H, double *g, double *residual) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  set_poses(p, poses12);
  const int n = 6 * p->N;
  Eigen::MatrixXd Hess(n, n);
  Eigen::VectorXd JacT(n);
  double r = 0;
  p->vox.acc_evaluate2(p->xs, head, end, Hess, JacT, r);
  for (int c = 0; c < n; c++) for (int rr = 0; rr < n; rr++) H[(size_t)c * n + rr] = Hess(rr, c);
  for (int i = 0; i < n; i++) g[i] = JacT[i];
  *residual = r;
}
// VOX_HESS::evaluate_only_residual
void ref_evaluate_only_residual(void *h, const double *poses12, double *residual) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  set_poses(p, poses12);
  double r = 0;
  p->vox.evaluate_only_residual(p->xs, r);
  *residual = r;
}
// BALM2::divide_thread_left (4 std::threads, ordered reduction)
double ref_divide_thread_left(void *h, const double *poses12, double *H, double *g) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  set_poses(p, poses12);
  const int n = 6 * p->N;
  Eigen::MatrixXd Hess(n, n);
  Eigen::VectorXd JacT(n);
  BALM2 opt;
  std::vector<IMUST> x_ab(p->N);
  const double r = opt.divide_thread_left(p->xs, p->vox, x_ab, Hess, JacT);
  for (int c = 0; c < n; c++) for (int rr = 0; rr < n; rr++) H[(size_t)c * n + rr] = Hess(rr, c);
  for (int i = 0; i < n; i++) g[i] = JacT[i];
  return r;
}
// BALM2::damping_iter(x_stats, voxhess): poses updated in place (the reference prints its trace and exit(0)s on too few planes)
void ref_damping_iter(void *h, double *poses12) {
  Problem *p = static_cast<Problem *>(h);
  win_size = p->N;
  set_poses(p, poses12);
  BALM2 opt;
  opt.damping_iter(p->xs, p->vox);
  for (int i = 0; i < p->N; i++) pack_pose(p->xs[i], poses12 + 12 * i);
}
// the dense solve of one LM step exactly as the reference writes it (bavoxel.hpp:1113-1114)
void ref_lm_solve(int n, const double *H, const double *g, double u, double *dx) {
  Eigen::MatrixXd Hess(n, n), D(n, n);
  Eigen::VectorXd JacT(n);
  for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) Hess(r, c) = H[(size_t)c * n + r];
  for (int i = 0; i < n; i++) JacT[i] = g[i];
  D.setIdentity();
  D.diagonal() = Hess.diagonal();
  Eigen::VectorXd dxi = (Hess + u * D).ldlt().solve(-JacT);
  for (int i = 0; i < n; i++) dx[i] = dxi[i];
}
void ref_exp(const double *phi, double *Rcm) {
  Eigen::Matrix3d R = Exp(Eigen::Vector3d(phi[0], phi[1], phi[2]));
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) Rcm[c * 3 + r] = R(r, c);
}
void ref_log(const double *Rcm, double *phi) {
  Eigen::Matrix3d R;
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) R(r, c) = Rcm[c * 3 + r];
  Eigen::Vector3d v = Log(R);
  for (int r = 0; r < 3; r++) phi[r] = v[r];
}

// ---------------- association: the reference's octree, kept alive between calls ----------------
struct Session {
  int N;
  unordered_map<VOXEL_LOC, OCTO_TREE_ROOT *> surf_map;
  std::vector<IMUST> x_buf;
  // last export
  std::vector<uint64_t> keys;
  std::vector<int32_t> layers;
  std::vector<int64_t> row_ptr;
  std::vector<int32_t> pose_idx;
  std::vector<double> obs10, fix10, coe;
};
void *ref_session_create(int N, double vsize, int llimit, int minps, const double *eig3) {
  win_size = N; voxel_size = vsize; layer_limit = llimit; min_ps = minps;
  for (int i = 0; i < 3; i++) eigen_value_array[i] = (float)eig3[i];  // `float eigen_value_array[]` (bavoxel.hpp:11)
  Session *s = new Session();
  s->N = N;
  s->x_buf.resize(N);
  return s;
}
void ref_session_destroy(void *h) {
  Session *s = static_cast<Session *>(h);
  for (auto &kv : s->surf_map) delete kv.second;
  delete s;
}
// cut_voxel(surf_map, pl_feat, x_key, fnum)  (benchmark_realworld.cpp:187-188)
void ref_session_cut_voxel(void *h, int64_t n, const float *xyz, const double *pose12, int fnum) {
  Session *s = static_cast<Session *>(h);
  win_size = s->N;
  pcl::PointCloud<PointType> pl;
  pl.reserve(n);
  for (int64_t i = 0; i < n; i++) {
    PointType p;
    p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
    pl.push_back(p);
  }
  unpack_pose(pose12, s->x_buf[fnum]);
  cut_voxel(s->surf_map, pl, s->x_buf[fnum], fnum);
}
// for every root: recut(win_count)  (benchmark_realworld.cpp:196-197)
void ref_session_recut(void *h, int win_count) {
  Session *s = static_cast<Session *>(h);
  win_size = s->N;
  for (auto &kv : s->surf_map) kv.second->recut(win_count);
}
// for every root: marginalize(mg_size, x_poses, win_count)  (consistency.cpp:131-135); poses12 may be NULL (x_poses empty)
void ref_session_marginalize(void *h, int mg_size, const double *poses12, int win_count) {
  Session *s = static_cast<Session *>(h);
  win_size = s->N;
  std::vector<IMUST> xp;
  if (poses12) { xp.resize(s->N); for (int i = 0; i < s->N; i++) unpack_pose(poses12 + 12 * i, xp[i]); }
  for (auto &kv : s->surf_map) kv.second->marginalize(mg_size, xp, win_count);
}
// key = [root | octant at layer 1 | octant at layer 2], 7 = not split; digit 7 is also octant 7, so the node's layer
// travels with the key
typedef std::pair<uint64_t, int> KeyLayer;
static void walk(OCTO_TREE_NODE *nd, uint64_t root_bits, int o1, int o2, std::map<const void *, KeyLayer> &out) {
  if (nd->octo_state != 1) {
    out[(const void *)&nd->sig_orig] = KeyLayer((root_bits << 6) | ((uint64_t)o1 << 3) | (uint64_t)o2, nd->layer);
    return;
  }
  for (int i = 0; i < 8; i++)
    if (nd->leaves[i] != nullptr) {
      if (nd->layer == 0) walk(nd->leaves[i], root_bits, i, 7, out);
      else walk(nd->leaves[i], root_bits, o1, i, out);
    }
}
// for every root: tras_opt(voxhess, win_count)  (benchmark_realworld.cpp:198) -> the pushed voxels, labelled with their node
// key and sorted by it. Returns the number of voxels; arrays are fetched with ref_session_fetch.
int64_t ref_session_export(void *h, int win_count) {
  Session *s = static_cast<Session *>(h);
  win_size = s->N;
  VOX_HESS vox;
  std::map<const void *, KeyLayer> key_of;
  for (auto &kv : s->surf_map) {
    kv.second->tras_opt(vox, win_count);
    const int64_t b = 1ll << 18;
    const uint64_t rb = ((uint64_t)(kv.first.x + b) << 38) | ((uint64_t)(kv.first.y + b) << 19) | (uint64_t)(kv.first.z + b);
    walk(kv.second, rb, 7, 7, key_of);
  }
  std::vector<std::pair<KeyLayer, size_t>> order;
  for (size_t k = 0; k < vox.plvec_voxels.size(); k++) order.push_back({key_of.at((const void *)vox.plvec_voxels[k]), k});
  std::sort(order.begin(), order.end());
  s->keys.clear(); s->layers.clear(); s->row_ptr.assign(1, 0); s->pose_idx.clear(); s->obs10.clear(); s->fix10.clear(); s->coe.clear();
  for (auto &e : order) {
    const std::vector<PointCluster> &vec = *vox.plvec_voxels[e.second];
    for (int i = 0; i < s->N; i++)
      if (vec[i].N != 0) {
        s->pose_idx.push_back(i);
        s->obs10.resize(s->obs10.size() + 10);
        pack_cluster(vec[i], &s->obs10[s->obs10.size() - 10]);
      }
    s->row_ptr.push_back((int64_t)s->pose_idx.size());
    s->fix10.resize(s->fix10.size() + 10);
    pack_cluster(*vox.sig_vecs[e.second], &s->fix10[s->fix10.size() - 10]);
    s->coe.push_back(vox.coeffs[e.second]);
    s->keys.push_back(e.first.first);
    s->layers.push_back((int32_t)e.first.second);
  }
  return (int64_t)s->keys.size();
}
int64_t ref_session_num_obs(void *h) { return (int64_t) static_cast<Session *>(h)->pose_idx.size(); }
void ref_session_fetch_layers(void *h, int32_t *layers) {
  Session *s = static_cast<Session *>(h);
  std::copy(s->layers.begin(), s->layers.end(), layers);
}
void ref_session_fetch(void *h, uint64_t *keys, int64_t *row_ptr, int32_t *pose_idx, double *obs10, double *fix10, double *coe) {
  Session *s = static_cast<Session *>(h);
  std::copy(s->keys.begin(), s->keys.end(), keys);
  std::copy(s->row_ptr.begin(), s->row_ptr.end(), row_ptr);
  std::copy(s->pose_idx.begin(), s->pose_idx.end(), pose_idx);
  std::copy(s->obs10.begin(), s->obs10.end(), obs10);
  std::copy(s->fix10.begin(), s->fix10.end(), fix10);
  std::copy(s->coe.begin(), s->coe.end(), coe);
}

}  // extern "C"
