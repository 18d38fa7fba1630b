// assoc.cu -- GPU association: point clouds + poses -> plane voxels (CSR of body-frame point clusters) directly in HBM.
//
// Restates the step that FEEDS the hot path:
//   cut_voxel                                  /root/reference/src/benchmark/bavoxel.hpp:1170-1223
//   OCTO_TREE_NODE::recut / judge_eigen / cut_func                     bavoxel.hpp:654-776
//   OCTO_TREE_NODE::tras_opt -> VOX_HESS::push_voxel                   bavoxel.hpp:908-929, 30-51
// The reference grows a pointer octree per root voxel on the CPU. Here every point gets a hierarchical 63-bit key
//   [root x | root y | root z | octant at layer 1 | octant at layer 2]     (19+19+19+3+3 bits, roots biased by 2^18)
// computed with the reference's own arithmetic (float32 voxel coordinate with the "-1 for negatives" rule, float
// voxel centres, child centre = centre +- quater_length), and the tree is evaluated level by level with radix sorts
// and segmented reductions: for each layer L the points are sorted by the key truncated to L digits (stable, frames
// stay ascending), (node, frame) segments are reduced to body-frame clusters (what push_voxel registers) and world-frame
// moments (what judge_eigen tests), nodes are judged planar / not planar, and a node becomes a plane voxel iff it is
// planar, has more than min_ps points and at least two observing frames, and every ancestor was large enough and NOT
// planar. All reductions run in a fixed order: the result is deterministic. Output order: ascending node key
// (digit 7 = "not split at this layer"), the order tests/assoc_ref.py uses.
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>
#include "internal.cuh"

namespace {

constexpr int ROOT_BITS = 19;
constexpr long long ROOT_BIAS = 1ll << 18;

struct AssocParams {
  double voxel_size;
  int layer_limit, min_ps;
  double eig[3];
};

__device__ __forceinline__ long long ref_voxel_index(double w, double vs) {
  float loc = (float)(w / vs);           // float loc_xyz[j] = pvec_tran[j] / voxel_size     (bavoxel.hpp:1180)
  if (loc < 0) loc -= 1.0f;              // :1181
  return (long long)loc;                 // (int64_t) cast truncates toward zero              (:1184)
}

// 63-bit key of a world point: root voxel (cut_voxel) + the octants cut_func would choose at layers 1 and 2
__device__ __forceinline__ unsigned long long world_point_key(const double *w, const AssocParams &p, int *bad);

__global__ void point_key_kernel(const float *xyz, const int *frame, const double *poses, int n_poses, int64_t n,
                                 AssocParams p, unsigned long long *key, int *bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = frame[i];
  if (f < 0 || f >= n_poses) { atomicOr(bad, 2); f = 0; }
  double r[9], t[3];
  load_pose(poses + 12 * f, r, t);
  const double b[3] = {(double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]};
  double w[3];
#pragma unroll
  for (int a = 0; a < 3; a++) w[a] = r[a * 3] * b[0] + r[a * 3 + 1] * b[1] + r[a * 3 + 2] * b[2] + t[a];
  key[i] = world_point_key(w, p, bad);
}

__device__ __forceinline__ unsigned long long world_point_key(const double *w, const AssocParams &p, int *bad) {
  unsigned long long k = 0;
  int oct1 = 0, oct2 = 0;
  const float quater = (float)(p.voxel_size / 4.0);  // ot->quater_length (:1216)
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const long long idx = ref_voxel_index(w[a], p.voxel_size);
    if (idx < -ROOT_BIAS || idx >= ROOT_BIAS) atomicOr(bad, 1);
    k = (k << ROOT_BITS) | (unsigned long long)((idx + ROOT_BIAS) & ((1ll << ROOT_BITS) - 1));
    const float c0 = (float)((0.5 + (double)idx) * p.voxel_size);  // float voxel_center (:1213-1215)
    const int b1 = w[a] > (double)c0 ? 1 : 0;                       // cut_func :711
    const float c1 = c0 + (float)(2 * b1 - 1) * quater;             // :717-719
    const int b2 = w[a] > (double)c1 ? 1 : 0;
    oct1 = (oct1 << 1) | b1;
    oct2 = (oct2 << 1) | b2;
  }
  return (k << 6) | (unsigned long long)(oct1 << 3) | (unsigned long long)oct2;
}

__global__ void iota_kernel(int *v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int)i;
}
__global__ void gather_frame_kernel(const int *frame, const int *idx, unsigned *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned)frame[idx[i]];
}
// key of layer L: digits below L forced to 7 ("not split")
__global__ void level_key_kernel(const unsigned long long *key, const int *idx, unsigned long long *out, int64_t n, int L) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = key[idx[i]];
  if (L == 0) k |= 63ull;
  else if (L == 1) k |= 7ull;
  out[i] = k;
}
__global__ void seg_flag_kernel(const unsigned long long *keyL, const int *idx, const int *frame, int64_t n, int *flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keyL[i] != keyL[i - 1] || frame[idx[i]] != frame[idx[i - 1]]) ? 1 : 0;
}
__global__ void seg_start_kernel(const int *flag, const int *scan, int64_t n, int *seg_start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) seg_start[scan[i]] = (int)i;
}

// One thread per (node, frame) segment: body-frame cluster (PointCluster::push, tools.hpp:311-316) and world-frame
// moments, accumulated in sorted order.
__global__ void seg_reduce_kernel(const float *xyz, const int *frame, const double *poses, const int *idx,
                                  const unsigned long long *keyL, const int *seg_start, int nseg, int64_t n,
                                  double *segB, double *segW, int *seg_frame, unsigned long long *seg_key) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int64_t a = seg_start[s], b = (s + 1 < nseg) ? seg_start[s + 1] : n;
  const int f = frame[idx[a]];
  double r[9], t[3];
  load_pose(poses + 12 * f, r, t);
  double B[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, W[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = a; i < b; i++) {
    const int j = idx[i];
    const double x = xyz[3 * j], y = xyz[3 * j + 1], z = xyz[3 * j + 2];
    B[0] += x * x; B[1] += x * y; B[2] += x * z; B[3] += y * y; B[4] += y * z; B[5] += z * z;
    B[6] += x; B[7] += y; B[8] += z; B[9] += 1.0;
    const double wx = r[0] * x + r[1] * y + r[2] * z + t[0];
    const double wy = r[3] * x + r[4] * y + r[5] * z + t[1];
    const double wz = r[6] * x + r[7] * y + r[8] * z + t[2];
    W[0] += wx * wx; W[1] += wx * wy; W[2] += wx * wz; W[3] += wy * wy; W[4] += wy * wz; W[5] += wz * wz;
    W[6] += wx; W[7] += wy; W[8] += wz; W[9] += 1.0;
  }
#pragma unroll
  for (int c = 0; c < 10; c++) { segB[(size_t)c * nseg + s] = B[c]; segW[(size_t)c * nseg + s] = W[c]; }
  seg_frame[s] = f;
  seg_key[s] = keyL[a];
}
__global__ void node_flag_kernel(const unsigned long long *seg_key, int nseg, int *flag) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) flag[s] = (s == 0 || seg_key[s] != seg_key[s - 1]) ? 1 : 0;
}
// One thread per node: total world moments over its frames -> judge_eigen (bavoxel.hpp:654-699)
__global__ void node_judge_kernel(const double *segW, const unsigned long long *seg_key, const int *node_start, int nnode,
                                  int nseg, double eig_thr, unsigned long long *node_key, int *node_size,
                                  int *node_planar, int *node_nframes) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnode) return;
  const int a = node_start[v], b = (v + 1 < nnode) ? node_start[v + 1] : nseg;
  double W[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = a; s < b; s++)
#pragma unroll
    for (int c = 0; c < 10; c++) W[c] += segW[(size_t)c * nseg + s];
  const double inv = 1.0 / W[9];
  const double c0 = W[6] * inv, c1 = W[7] * inv, c2 = W[8] * inv;
  double lam[3], u0[3], u1[3], u2[3];
  eig3_jacobi(W[0] * inv - c0 * c0, W[1] * inv - c0 * c1, W[2] * inv - c0 * c2, W[3] * inv - c1 * c1,
              W[4] * inv - c1 * c2, W[5] * inv - c2 * c2, lam, u0, u1, u2);
  node_key[v] = seg_key[a];
  node_size[v] = (int)(W[9] + 0.5);
  node_planar[v] = (lam[0] / lam[1] < eig_thr) ? 1 : 0;  // decision < eigen_value_array[layer]  (:665,697)
  node_nframes[v] = b - a;
}

struct LevelTables {  // device arrays of one layer
  int nseg = 0, nnode = 0;
  double *segB = nullptr;          // [10][nseg] body-frame clusters
  int *seg_frame = nullptr;
  int *node_start = nullptr;       // [nnode] first segment
  unsigned long long *node_key = nullptr;
  int *node_size = nullptr, *node_planar = nullptr, *node_nframes = nullptr, *node_leaf = nullptr;
};

__device__ __forceinline__ int find_key(const unsigned long long *keys, int n, unsigned long long k) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long v = keys[mid];
    if (v == k) return mid;
    if (v < k) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// recut (bavoxel.hpp:737-776) evaluated per node: a node is reached iff every ancestor had more than min_ps points and
// failed the planarity test; it becomes a plane voxel iff it is planar with more than min_ps points (push_state = 1)
// and at least two frames observe it (push_voxel :37).
__global__ void leaf_decide_kernel(int L, int nnode, const unsigned long long *key, const int *size, const int *planar,
                                   const int *nframes, const unsigned long long *key0, const int *size0,
                                   const int *planar0, int n0, const unsigned long long *key1, const int *size1,
                                   const int *planar1, int n1, int min_ps, int layer_limit, int *leaf) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnode) return;
  bool reached = L <= layer_limit;
  if (reached && L >= 1) {
    const int r = find_key(key0, n0, key[v] | 63ull);
    reached = r >= 0 && size0[r] > min_ps && !planar0[r];
  }
  if (reached && L >= 2) {
    const int r = find_key(key1, n1, key[v] | 7ull);
    reached = r >= 0 && size1[r] > min_ps && !planar1[r];
  }
  leaf[v] = (reached && size[v] > min_ps && planar[v] && nframes[v] >= 2) ? 1 : 0;
}

__global__ void leaf_collect_kernel(int L, int nnode, const int *leaf, const int *leaf_scan, const unsigned long long *key,
                                    unsigned long long *out_key, int *out_ref, int base) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnode || !leaf[v]) return;
  const int o = base + leaf_scan[v];
  out_key[o] = key[v];
  out_ref[o] = (L << 28) | v;  // layer and node id
}
__global__ void leaf_count_kernel(const int *ref, int nleaf, const int *nf0, const int *nf1, const int *nf2, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nleaf) return;
  const int L = ref[i] >> 28, v = ref[i] & 0x0FFFFFFF;
  cnt[i] = (L == 0 ? nf0 : (L == 1 ? nf1 : nf2))[v];
}
__global__ void leaf_layer_kernel(const int *ref, int nleaf, int *layer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nleaf) layer[i] = ref[i] >> 28;
}
struct EmitLevel {
  const double *segB;
  const int *seg_frame, *node_start, *node_size;
  int nseg;
};
__global__ void emit_kernel(const int *ref, const int *obs_start, int nleaf, EmitLevel e0, EmitLevel e1, EmitLevel e2,
                            const int *cnt, double *obs, int64_t Kp, int *pose_idx, long long *row_ptr, double *coe) {
  const int i = blockIdx.x;  // one CTA per plane voxel
  if (i >= nleaf) return;
  const int L = ref[i] >> 28, v = ref[i] & 0x0FFFFFFF;
  const EmitLevel e = L == 0 ? e0 : (L == 1 ? e1 : e2);
  const int s0 = e.node_start[v], k = cnt[i], o0 = obs_start[i];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 10; c++) obs[(size_t)c * Kp + o0 + j] = e.segB[(size_t)c * e.nseg + s0 + j];
    pose_idx[o0 + j] = e.seg_frame[s0 + j];
  }
  if (threadIdx.x == 0) {
    row_ptr[i] = o0;
    coe[i] = (double)e.node_size[v];  // coe = sum of N over the window (bavoxel.hpp:42-44)
    if (i == nleaf - 1) row_ptr[nleaf] = o0 + k;
  }
}

// Every scratch buffer of assoc_build lives in a pool that frees what is left on ANY exit path (the early returns of
// ATRY / CUDA_TRY, "no plane voxels found", a bad frame index ...), so a failed call leaks nothing.
struct ScratchPool {
  std::vector<void *> ptrs;
  ~ScratchPool() { for (void *q : ptrs) cudaFree(q); }
  template <typename T>
  int get(T **p, size_t n) {
    CUDA_TRY(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
    ptrs.push_back((void *)*p);
    return BALM_OK;
  }
  void drop(void *q) {
    if (!q) return;
    for (size_t i = 0; i < ptrs.size(); i++)
      if (ptrs[i] == q) { ptrs[i] = ptrs.back(); ptrs.pop_back(); break; }
    cudaFree(q);
  }
};
#define ATRY(x) do { int _s = (x); if (_s != BALM_OK) return _s; } while (0)

}  // namespace

// Builds the plane voxels of `ctx` from raw scans. xyz (n x 3 float32, body frame) and frame (n) are HOST arrays,
// poses12 the initial poses. On success the voxels are registered exactly as balm_set_voxels would have done.
int assoc_build(balm_ctx *c, int64_t n, const float *xyz_h, const int *frame_h, const double *poses12_h,
                double voxel_size, int layer_limit, int min_ps, const double *eig3, int64_t *M_out, int64_t *K_out,
                int (*register_csr)(balm_ctx *, int64_t, int64_t)) {
  if (n < 1 || n >= (1ll << 31) || layer_limit < 0 || layer_limit > 2) {
    balm_set_error("balm_cut_voxels: bad arguments (n < 2^31, layer_limit in 0..2)");
    return BALM_ERR_INVALID;
  }
  cudaStream_t st = c->stream;
  ScratchPool pool;
  AssocParams P{voxel_size, layer_limit, min_ps, {eig3[0], eig3[1], eig3[2]}};
  float *xyz = nullptr;
  int *frame = nullptr, *idx = nullptr, *idx2 = nullptr, *flag = nullptr, *scan = nullptr, *seg_start = nullptr, *bad = nullptr;
  unsigned *fkey = nullptr, *fkey2 = nullptr;
  unsigned long long *key = nullptr, *keyL = nullptr, *keyL2 = nullptr;
  double *poses = nullptr;
  ATRY(pool.get(&xyz, (size_t)3 * n)); ATRY(pool.get(&frame, (size_t)n)); ATRY(pool.get(&idx, (size_t)n)); ATRY(pool.get(&idx2, (size_t)n));
  ATRY(pool.get(&flag, (size_t)n)); ATRY(pool.get(&scan, (size_t)n)); ATRY(pool.get(&seg_start, (size_t)n)); ATRY(pool.get(&bad, 1));
  ATRY(pool.get(&fkey, (size_t)n)); ATRY(pool.get(&fkey2, (size_t)n));
  ATRY(pool.get(&key, (size_t)n)); ATRY(pool.get(&keyL, (size_t)n)); ATRY(pool.get(&keyL2, (size_t)n));
  ATRY(pool.get(&poses, (size_t)12 * c->N));
  CUDA_TRY(cudaMemcpyAsync(xyz, xyz_h, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(frame, frame_h, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(poses, poses12_h, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned gb = (unsigned)((n + 255) / 256);
  point_key_kernel<<<gb, 256, 0, st>>>(xyz, frame, poses, c->N, n, P, key, bad);
  // stable order by frame first, so that every later stable sort keeps the frames ascending inside a node
  iota_kernel<<<gb, 256, 0, st>>>(idx, n);
  gather_frame_kernel<<<gb, 256, 0, st>>>(frame, idx, fkey, n);
  size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, fkey, fkey2, idx, idx2, (int)n, 0, 32, st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, keyL, keyL2, idx, idx2, (int)n, 0, 63, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, flag, scan, (int)n, st);
  tmp_bytes = std::max(tmp_bytes, std::max(tb2, tb3));
  char *tmp = nullptr;
  ATRY(pool.get(&tmp, tmp_bytes));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, fkey, fkey2, idx, idx2, (int)n, 0, 32, st));
  int *order = idx2;   // frame-sorted point order
  int *work = idx;     // per-level sorted order
  c->launches += 4;

  LevelTables T[3];
  for (int L = 0; L <= layer_limit; L++) {
    level_key_kernel<<<gb, 256, 0, st>>>(key, order, keyL, n, L);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keyL, keyL2, order, work, (int)n, 0, 63, st));
    seg_flag_kernel<<<gb, 256, 0, st>>>(keyL2, work, frame, n, flag);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, scan, (int)n, st));
    int last_flag = 0, last_scan = 0;
    CUDA_TRY(cudaMemcpyAsync(&last_flag, flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&last_scan, scan + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int nseg = last_scan + last_flag;
    seg_start_kernel<<<gb, 256, 0, st>>>(flag, scan, n, seg_start);
    LevelTables &t = T[L];
    t.nseg = nseg;
    double *segW = nullptr;
    unsigned long long *seg_key = nullptr;
    int *nflag = nullptr, *nscan = nullptr;
    ATRY(pool.get(&t.segB, (size_t)10 * nseg)); ATRY(pool.get(&segW, (size_t)10 * nseg)); ATRY(pool.get(&t.seg_frame, (size_t)nseg));
    ATRY(pool.get(&seg_key, (size_t)nseg)); ATRY(pool.get(&nflag, (size_t)nseg)); ATRY(pool.get(&nscan, (size_t)nseg));
    const unsigned gs = (unsigned)((nseg + 127) / 128);
    seg_reduce_kernel<<<gs, 128, 0, st>>>(xyz, frame, poses, work, keyL2, seg_start, nseg, n, t.segB, segW, t.seg_frame, seg_key);
    node_flag_kernel<<<gs, 128, 0, st>>>(seg_key, nseg, nflag);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, nflag, nscan, nseg, st);
    if (tb > tmp_bytes) { pool.drop(tmp); tmp = nullptr; tmp_bytes = tb; ATRY(pool.get(&tmp, tmp_bytes)); }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nflag, nscan, nseg, st));
    CUDA_TRY(cudaMemcpyAsync(&last_flag, nflag + nseg - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&last_scan, nscan + nseg - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    t.nnode = last_scan + last_flag;
    ATRY(pool.get(&t.node_start, (size_t)t.nnode)); ATRY(pool.get(&t.node_key, (size_t)t.nnode));
    ATRY(pool.get(&t.node_size, (size_t)t.nnode)); ATRY(pool.get(&t.node_planar, (size_t)t.nnode));
    ATRY(pool.get(&t.node_nframes, (size_t)t.nnode)); ATRY(pool.get(&t.node_leaf, (size_t)t.nnode));
    seg_start_kernel<<<gs, 128, 0, st>>>(nflag, nscan, nseg, t.node_start);
    node_judge_kernel<<<(unsigned)((t.nnode + 127) / 128), 128, 0, st>>>(segW, seg_key, t.node_start, t.nnode, nseg, P.eig[L],
                                                                        t.node_key, t.node_size, t.node_planar, t.node_nframes);
    CUDA_TRY(cudaStreamSynchronize(st));
    pool.drop(segW); pool.drop(seg_key); pool.drop(nflag); pool.drop(nscan);
    c->launches += 9;
  }
  int h_bad = 0;
  CUDA_TRY(cudaMemcpy(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_bad) {
    balm_set_error(h_bad & 2 ? "balm_cut_voxels: frame index outside [0, n_poses)"
                             : "balm_cut_voxels: a point lies outside the +-2^18 root-voxel range");
    return BALM_ERR_INVALID;
  }

  // ---- leaves of every layer, merged in key order ----
  int nleaf_L[3] = {0, 0, 0};
  int *lscan[3] = {nullptr, nullptr, nullptr};
  for (int L = 0; L <= layer_limit; L++) {
    LevelTables &t = T[L];
    leaf_decide_kernel<<<(unsigned)((t.nnode + 127) / 128), 128, 0, st>>>(
        L, t.nnode, t.node_key, t.node_size, t.node_planar, t.node_nframes, T[0].node_key, T[0].node_size, T[0].node_planar,
        T[0].nnode, T[1].node_key, T[1].node_size, T[1].node_planar, T[1].nnode, min_ps, layer_limit, t.node_leaf);
    ATRY(pool.get(&lscan[L], (size_t)t.nnode));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, t.node_leaf, lscan[L], t.nnode, st);
    if (tb > tmp_bytes) { pool.drop(tmp); tmp = nullptr; tmp_bytes = tb; ATRY(pool.get(&tmp, tmp_bytes)); }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, t.node_leaf, lscan[L], t.nnode, st));
    int a = 0, b = 0;
    CUDA_TRY(cudaMemcpyAsync(&a, t.node_leaf + t.nnode - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&b, lscan[L] + t.nnode - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    nleaf_L[L] = a + b;
  }
  const int nleaf = nleaf_L[0] + nleaf_L[1] + nleaf_L[2];
  if (nleaf < 1) { balm_set_error("balm_cut_voxels: no plane voxels found"); return BALM_ERR_INVALID; }
  unsigned long long *lkey = nullptr, *lkey2 = nullptr;
  int *lref = nullptr, *lref2 = nullptr, *lcnt = nullptr, *lobs = nullptr;
  ATRY(pool.get(&lkey, (size_t)nleaf)); ATRY(pool.get(&lkey2, (size_t)nleaf)); ATRY(pool.get(&lref, (size_t)nleaf));
  ATRY(pool.get(&lref2, (size_t)nleaf)); ATRY(pool.get(&lcnt, (size_t)nleaf)); ATRY(pool.get(&lobs, (size_t)nleaf));
  int base = 0;
  for (int L = 0; L <= layer_limit; L++) {
    leaf_collect_kernel<<<(unsigned)((T[L].nnode + 127) / 128), 128, 0, st>>>(L, T[L].nnode, T[L].node_leaf, lscan[L],
                                                                             T[L].node_key, lkey, lref, base);
    base += nleaf_L[L];
  }
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, lkey, lkey2, lref, lref2, nleaf, 0, 63, st);
    if (tb > tmp_bytes) { pool.drop(tmp); tmp = nullptr; tmp_bytes = tb; ATRY(pool.get(&tmp, tmp_bytes)); }
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, lkey, lkey2, lref, lref2, nleaf, 0, 63, st));
  }
  leaf_count_kernel<<<(unsigned)((nleaf + 127) / 128), 128, 0, st>>>(lref2, nleaf, T[0].node_nframes, T[1].node_nframes,
                                                                    T[2].node_nframes, lcnt);
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, lcnt, lobs, nleaf, st);
    if (tb > tmp_bytes) { pool.drop(tmp); tmp = nullptr; tmp_bytes = tb; ATRY(pool.get(&tmp, tmp_bytes)); }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, lcnt, lobs, nleaf, st));
  }
  int a = 0, b = 0;
  CUDA_TRY(cudaMemcpyAsync(&a, lcnt + nleaf - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&b, lobs + nleaf - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int64_t K = (int64_t)a + b;
  // ---- register: allocate the problem arrays of the context and emit straight into them ----
  ATRY(register_csr(c, nleaf, K));
  EmitLevel e[3];
  for (int L = 0; L < 3; L++) e[L] = EmitLevel{T[L].segB, T[L].seg_frame, T[L].node_start, T[L].node_size, T[L].nseg};
  emit_kernel<<<nleaf, 64, 0, st>>>(lref2, lobs, nleaf, e[0], e[1], e[2], lcnt, c->obs, c->Kp, c->pose_idx, c->row_ptr, c->coe);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  c->launches += 12;
  // the voxel keys (ascending) and the parameters stay with the context: balm_append_scan / balm_marginalize need them
  cudaFree(c->vox_key);
  c->vox_key = nullptr;
  CUDA_TRY(cudaMalloc((void **)&c->vox_key, sizeof(unsigned long long) * (size_t)nleaf));
  CUDA_TRY(cudaMemcpy(c->vox_key, lkey2, sizeof(unsigned long long) * (size_t)nleaf, cudaMemcpyDeviceToDevice));
  cudaFree(c->vox_layer);
  c->vox_layer = nullptr;
  CUDA_TRY(cudaMalloc((void **)&c->vox_layer, sizeof(int) * (size_t)nleaf));
  leaf_layer_kernel<<<(unsigned)((nleaf + 127) / 128), 128, 0, st>>>(lref2, nleaf, c->vox_layer);
  CUDA_TRY(cudaStreamSynchronize(st));
  c->has_keys = true;
  c->assoc_voxel_size = voxel_size; c->assoc_layer_limit = layer_limit; c->assoc_min_ps = min_ps;
  for (int q = 0; q < 3; q++) c->assoc_eig[q] = eig3[q];
  // the pool's destructor frees every scratch buffer (level tables included)
  *M_out = nleaf;
  *K_out = K;
  return BALM_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Pose-major observation lists of a sparse co-visibility problem (the reference walks a voxel's win_size slots and
// skips the empty ones, bavoxel.hpp:332,365-366,404-408; here the observation sweep runs lane = pose and needs, per pose,
// the list of its observations in ascending voxel order). Built on the device: stable radix sort of the observation
// indices by pose (CSR order is voxel-major, so a stable sort leaves every pose's list voxel-ascending), plus the
// voxel of every observation. planes_host: per-pose counts (already on the host for the >= 20 planes guard).
namespace {
__global__ void obs_voxel_kernel(const long long *row_ptr, int64_t M, int *vox_of_obs) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M) return;
  for (long long s = row_ptr[v]; s < row_ptr[v + 1]; s++) vox_of_obs[s] = (int)v;
}
__global__ void gather_int_kernel(const int *src, const int *idx, int *dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// one thread per pose: [lo, hi) = the part of its list whose voxels lie in [v0, v1)
__global__ void csc_batch_kernel(const int *csc_ptr, const int *csc_vox, int N, int v0, int v1, int *lo, int *hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int b = csc_ptr[i], e = csc_ptr[i + 1];
  int l = b, h = e;
  while (l < h) { const int m = (l + h) >> 1; if (csc_vox[m] < v0) l = m + 1; else h = m; }
  lo[i] = l;
  h = e;
  while (l < h) { const int m = (l + h) >> 1; if (csc_vox[m] < v1) l = m + 1; else h = m; }
  hi[i] = l;
}
}  // namespace

int build_pose_major_lists(balm_ctx *c, const int *planes_host) {
  const int N = c->N;
  const int64_t K = c->K, M = c->M;
  cudaStream_t st = c->stream;
  ScratchPool pool;
  std::vector<int> ptr((size_t)N + 1, 0);
  for (int i = 0; i < N; i++) ptr[i + 1] = ptr[i] + planes_host[i];
  c->csc_max_len = 0;
  for (int i = 0; i < N; i++) c->csc_max_len = std::max(c->csc_max_len, planes_host[i]);
  CUDA_TRY(cudaMalloc((void **)&c->csc_ptr, sizeof(int) * ((size_t)N + 1)));
  CUDA_TRY(cudaMalloc((void **)&c->csc_obs, sizeof(int) * (size_t)K));
  CUDA_TRY(cudaMalloc((void **)&c->csc_vox, sizeof(int) * (size_t)K));
  CUDA_TRY(cudaMalloc((void **)&c->csc_lo, sizeof(int) * (size_t)N));
  CUDA_TRY(cudaMalloc((void **)&c->csc_hi, sizeof(int) * (size_t)N));
  CUDA_TRY(cudaMemcpyAsync(c->csc_ptr, ptr.data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
  int *vox = nullptr, *idx = nullptr, *keys = nullptr, *keys2 = nullptr;
  ATRY(pool.get(&vox, (size_t)K)); ATRY(pool.get(&idx, (size_t)K)); ATRY(pool.get(&keys, (size_t)K)); ATRY(pool.get(&keys2, (size_t)K));
  obs_voxel_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(c->row_ptr, M, vox);
  iota_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(idx, K);
  CUDA_TRY(cudaMemcpyAsync(keys, c->pose_idx, sizeof(int) * (size_t)K, cudaMemcpyDeviceToDevice, st));
  int bits = 1;
  while ((1 << bits) < N) bits++;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, c->csc_obs, (int)K, 0, bits, st);
  char *tmp = nullptr;
  ATRY(pool.get(&tmp, tb));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, c->csc_obs, (int)K, 0, bits, st));
  gather_int_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(vox, c->csc_obs, c->csc_vox, K);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  c->launches += 4;
  return launch_csc_batch(c, 0, M);
}

int launch_csc_batch(balm_ctx *c, int64_t v0, int64_t v1) {
  if (c->dense || !c->csc_ptr) return BALM_OK;
  csc_batch_kernel<<<(c->N + 127) / 128, 128, 0, c->stream>>>(c->csc_ptr, c->csc_vox, c->N, (int)v0, (int)v1, c->csc_lo, c->csc_hi);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Sliding-window marginalisation on the registered voxel set:
//   OCTO_TREE_NODE::to_margi (bavoxel.hpp:778-816) on every plane leaf -- the clusters of the oldest mg_size scans,
//   transformed by their (optimised) poses, are folded into the leaf's fix_point when fix_point.N < 50 (:790-795), the
//   remaining scans shift down by mg_size (:797-803) -- followed by what tras_opt / push_voxel keep of the leaf
//   (bavoxel.hpp:908-929: fewer than min_ps points left -> not pushed; :32-37: fewer than 2 observing scans -> not pushed;
//   coe = number of points left, :42-44).
// The flattened voxel set has no octree any more, so a leaf that is not pushed is dropped together with its fix cluster
// (in the reference it stays in the map for later scans); documented in DESIGN.md.
namespace {
__global__ void marg_count_kernel(const long long *row_ptr, const int *pose_idx, const double *obs, int64_t Kp, int64_t M,
                                  int mg, int min_ps, int *keep, int *cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M) return;
  int k = 0;
  double pts = 0.0;
  for (long long s = row_ptr[v]; s < row_ptr[v + 1]; s++)
    if (pose_idx[s] >= mg) { k++; pts += obs[9 * Kp + s]; }
  const int kp = ((int)pts >= min_ps && k >= 2) ? 1 : 0;
  keep[v] = kp;
  cnt[v] = kp ? k : 0;
}
__global__ void marg_emit_kernel(const long long *row_ptr, const int *pose_idx, const double *obs, int64_t Kp,
                                 const double *fix /*SoA [10][M] or null*/, int64_t M, const double *poses, int mg,
                                 const int *keep, const int *vscan, const int *oscan, long long *row_ptr_out,
                                 int *pose_idx_out, double *obs10_out, double *fix10_out, double *coe_out, int64_t Mout,
                                 int64_t Kout, const unsigned long long *key_in, unsigned long long *key_out,
                                 const int *layer_in, int *layer_out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v == 0) row_ptr_out[Mout] = Kout;
  if (v >= M || !keep[v]) return;
  const int nv = vscan[v];
  if (key_in) { key_out[nv] = key_in[v]; layer_out[nv] = layer_in[v]; }
  long long o = oscan[v];
  row_ptr_out[nv] = o;
  double f[10];
  for (int c = 0; c < 10; c++) f[c] = fix ? fix[c * M + v] : 0.0;
  const bool absorb = (int)f[9] < 50;  // fix_point.N < 50 (bavoxel.hpp:790; push_state == 1 for a registered plane leaf)
  double pts = 0.0;
  for (long long s = row_ptr[v]; s < row_ptr[v + 1]; s++) {
    double ob[10];
    for (int c = 0; c < 10; c++) ob[c] = obs[c * Kp + s];
    const int p = pose_idx[s];
    if (p < mg) {
      if (absorb) {  // fix_point += sig_tran[i]  (transform: tools.hpp:333-339)
        double r[9], t[3];
        load_pose(poses + 12 * p, r, t);
        const WC w = world_cluster(ob, r, t);
        f[0] += w.p00; f[1] += w.p01; f[2] += w.p02; f[3] += w.p11; f[4] += w.p12; f[5] += w.p22;
        f[6] += w.v0; f[7] += w.v1; f[8] += w.v2; f[9] += w.n;
      }
    } else {
      for (int c = 0; c < 10; c++) obs10_out[o * 10 + c] = ob[c];
      pose_idx_out[o] = p - mg;
      pts += ob[9];
      o++;
    }
  }
  for (int c = 0; c < 10; c++) fix10_out[(size_t)nv * 10 + c] = f[c];
  coe_out[nv] = pts;
}
}  // namespace

// after a re-registration: the compacted keys of the new voxel set (or none)
static int install_keys(balm_ctx *c, const unsigned long long *keys_dev, const int *layers_dev, int64_t M, bool have) {
  cudaFree(c->vox_key);
  cudaFree(c->vox_layer);
  c->vox_key = nullptr;
  c->vox_layer = nullptr;
  c->has_keys = false;
  if (!have) return BALM_OK;
  CUDA_TRY(cudaMalloc((void **)&c->vox_key, sizeof(unsigned long long) * (size_t)M));
  CUDA_TRY(cudaMalloc((void **)&c->vox_layer, sizeof(int) * (size_t)M));
  CUDA_TRY(cudaMemcpy(c->vox_key, keys_dev, sizeof(unsigned long long) * (size_t)M, cudaMemcpyDeviceToDevice));
  CUDA_TRY(cudaMemcpy(c->vox_layer, layers_dev, sizeof(int) * (size_t)M, cudaMemcpyDeviceToDevice));
  c->has_keys = true;
  return BALM_OK;
}

int marginalize_build(balm_ctx *c, int mg, const double *poses_dev, int min_ps, int64_t *M_out, int64_t *K_out,
                      int (*reg)(balm_ctx *, int64_t, const int64_t *, const int32_t *, const double *, const double *,
                                 const double *, int64_t)) {
  const int64_t M = c->M;
  cudaStream_t st = c->stream;
  ScratchPool pool;
  int *keep = nullptr, *cnt = nullptr, *vscan = nullptr, *oscan = nullptr;
  ATRY(pool.get(&keep, (size_t)M)); ATRY(pool.get(&cnt, (size_t)M)); ATRY(pool.get(&vscan, (size_t)M)); ATRY(pool.get(&oscan, (size_t)M));
  const unsigned gb = (unsigned)((M + 127) / 128);
  marg_count_kernel<<<gb, 128, 0, st>>>(c->row_ptr, c->pose_idx, c->obs, c->Kp, M, mg, min_ps, keep, cnt);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, keep, vscan, (int)M, st);
  char *tmp = nullptr;
  ATRY(pool.get(&tmp, tb));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tb, keep, vscan, (int)M, st));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, oscan, (int)M, st));
  int last[4] = {0, 0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(&last[0], keep + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[1], vscan + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[2], cnt + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[3], oscan + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int64_t Mout = (int64_t)last[0] + last[1], Kout = (int64_t)last[2] + last[3];
  *M_out = Mout;
  *K_out = Kout;
  long long *rp = nullptr;
  int *pi = nullptr;
  double *ob = nullptr, *fx = nullptr, *co = nullptr;
  ATRY(pool.get(&rp, (size_t)Mout + 1)); ATRY(pool.get(&pi, (size_t)Kout)); ATRY(pool.get(&ob, (size_t)Kout * 10));
  ATRY(pool.get(&fx, (size_t)Mout * 10)); ATRY(pool.get(&co, (size_t)Mout));
  unsigned long long *keys_out = nullptr;
  int *layers_out = nullptr;
  const bool had_keys = c->has_keys && c->vox_key;
  if (had_keys) { ATRY(pool.get(&keys_out, (size_t)Mout)); ATRY(pool.get(&layers_out, (size_t)Mout)); }
  marg_emit_kernel<<<gb, 128, 0, st>>>(c->row_ptr, c->pose_idx, c->obs, c->Kp, c->fix, M, poses_dev, mg, keep, vscan, oscan, rp,
                                       pi, ob, fx, co, Mout, Kout, had_keys ? c->vox_key : nullptr, keys_out, c->vox_layer,
                                       layers_out);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  c->launches += 4;
  if (Mout < 1) { balm_set_error("balm_marginalize: no voxel is left with two observing scans"); return BALM_ERR_INVALID; }
  // re-register from the device arrays (copies them, transposes the observations, rebuilds the pose-major lists)
  ATRY(reg(c, Mout, (const int64_t *)rp, pi, ob, fx, co, Kout));
  return install_keys(c, keys_out, layers_out, Mout, had_keys);
}


// ---------------------------------------------------------------------------------------------------------------------
// Appending a scan to the voxel set in HBM (the other half of the sliding window).
// Reference: cut_voxel pushes the new scan's points into the root voxels (bavoxel.hpp:1170-1223), recut hands them down
// to the leaves (cut_func(win_count-1) on split nodes, :737-776) and RE-JUDGES every leaf with its fix cluster and all
// window clusters (judge_eigen, :654-699): still planar with more than min_ps window points -> pushed again with the new
// observation; otherwise the reference splits it further or freezes it. Here: every new point finds the leaf that contains
// it through the stored keys (the same arithmetic as balm_cut_voxels), its body-frame moments become a new observation
// in pose slot `slot`, and every voxel is re-judged from fix + all its clusters transformed by the current poses; a voxel
// that fails (not planar any more, too few points, fewer than two scans) is dropped. NOT done (documented): creating new
// roots / new leaves from points that fall outside the existing planes, and re-splitting a leaf that stopped being planar
// -- both appear at the next full balm_cut_voxels.
namespace {
__global__ void append_match_kernel(const float *xyz, const double *pose12, int64_t n, AssocParams p,
                                    const unsigned long long *vox_key, const int *vox_layer, int M, unsigned *vid1,
                                    int *bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], t[3];
  load_pose(pose12, r, t);
  const double b[3] = {(double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]};
  double w[3];
#pragma unroll
  for (int a = 0; a < 3; a++) w[a] = r[a * 3] * b[0] + r[a * 3 + 1] * b[1] + r[a * 3 + 2] * b[2] + t[a];
  int dummy = 0;
  (void)bad;
  const unsigned long long k = world_point_key(w, p, &dummy);
  // a key digit 7 means "not split" AND is octant 7, so a hit only counts when the voxel's layer agrees
  int v = find_key(vox_key, M, k | 63ull);            // a layer-0 leaf (the whole root voxel is one plane)
  if (v >= 0 && vox_layer[v] != 0) v = -1;
  if (v < 0) { v = find_key(vox_key, M, k | 7ull); if (v >= 0 && vox_layer[v] != 1) v = -1; }   // a layer-1 leaf
  if (v < 0) { v = find_key(vox_key, M, k); if (v >= 0 && vox_layer[v] != 2) v = -1; }          // a layer-2 leaf
  vid1[i] = (unsigned)(v + 1);                        // 0 = the point falls into no registered plane voxel
}
// one thread per voxel: the new scan's body-frame cluster (points in their original order)
__global__ void append_cluster_kernel(const float *xyz, const unsigned *vid_sorted, const int *idx_sorted, int64_t n, int M,
                                      double *newobs /*[M][10]*/) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M) return;
  const unsigned want = (unsigned)(v + 1);
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (vid_sorted[m] < want) lo = m + 1; else hi = m; }
  double B[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = lo; i < n && vid_sorted[i] == want; i++) {
    const int j = idx_sorted[i];
    const double x = xyz[3 * j], y = xyz[3 * j + 1], z = xyz[3 * j + 2];
    B[0] += x * x; B[1] += x * y; B[2] += x * z; B[3] += y * y; B[4] += y * z; B[5] += z * z;
    B[6] += x; B[7] += y; B[8] += z; B[9] += 1.0;
  }
#pragma unroll
  for (int q = 0; q < 10; q++) newobs[(size_t)v * 10 + q] = B[q];
}
// one thread per voxel: judge_eigen on fix + every window cluster (old ones and the new one) at the current poses
__global__ void append_judge_kernel(const long long *row_ptr, const int *pose_idx, const double *obs, int64_t Kp,
                                    const double *fix, int64_t M, const double *poses, const double *newobs, int slot,
                                    const int *vox_layer, AssocParams p, int *keep, int *cnt, int *bad) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M) return;
  double W[10];
  for (int q = 0; q < 10; q++) W[q] = fix ? fix[q * M + v] : 0.0;
  double pts = 0.0;
  int k = 0;
  for (long long s = row_ptr[v]; s < row_ptr[v + 1]; s++) {
    double ob[10];
    for (int q = 0; q < 10; q++) ob[q] = obs[q * Kp + s];
    double r[9], t[3];
    load_pose(poses + 12 * pose_idx[s], r, t);
    const WC w = world_cluster(ob, r, t);
    W[0] += w.p00; W[1] += w.p01; W[2] += w.p02; W[3] += w.p11; W[4] += w.p12; W[5] += w.p22;
    W[6] += w.v0; W[7] += w.v1; W[8] += w.v2; W[9] += w.n;
    pts += ob[9];
    k++;
    if (pose_idx[s] >= slot) atomicOr(bad, 1);  // the new scan must take a slot above every scan already in the voxel
  }
  const double *nb = newobs + (size_t)v * 10;
  if (nb[9] > 0.0) {
    double r[9], t[3];
    load_pose(poses + 12 * slot, r, t);
    const WC w = world_cluster(nb, r, t);
    W[0] += w.p00; W[1] += w.p01; W[2] += w.p02; W[3] += w.p11; W[4] += w.p12; W[5] += w.p22;
    W[6] += w.v0; W[7] += w.v1; W[8] += w.v2; W[9] += w.n;
    pts += nb[9];
    k++;
  }
  const double inv = 1.0 / W[9];
  const double c0 = W[6] * inv, c1 = W[7] * inv, c2 = W[8] * inv;
  double lam[3], u0[3], u1[3], u2[3];
  eig3_jacobi(W[0] * inv - c0 * c0, W[1] * inv - c0 * c1, W[2] * inv - c0 * c2, W[3] * inv - c1 * c1,
              W[4] * inv - c1 * c2, W[5] * inv - c2 * c2, lam, u0, u1, u2);
  const int layer = vox_layer[v];
  const bool planar = lam[0] / lam[1] < p.eig[layer];                     // judge_eigen (:665,697)
  const int kp = (planar && (int)pts > p.min_ps && k >= 2) ? 1 : 0;       // recut :755-757, push_voxel :37
  keep[v] = kp;
  cnt[v] = kp ? k : 0;
}
__global__ void append_emit_kernel(const long long *row_ptr, const int *pose_idx, const double *obs, int64_t Kp,
                                   const double *fix, const double *coe, int64_t M, const double *newobs, int slot,
                                   const int *keep, const int *vscan, const int *oscan, const unsigned long long *key_in,
                                   const int *layer_in, long long *row_ptr_out, int *pose_idx_out, double *obs10_out,
                                   double *fix10_out, double *coe_out, unsigned long long *key_out, int *layer_out,
                                   int64_t Mout, int64_t Kout) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v == 0) row_ptr_out[Mout] = Kout;
  if (v >= M || !keep[v]) return;
  const int nv = vscan[v];
  long long o = oscan[v];
  row_ptr_out[nv] = o;
  key_out[nv] = key_in[v];
  layer_out[nv] = layer_in[v];
  for (long long s = row_ptr[v]; s < row_ptr[v + 1]; s++, o++) {
    for (int q = 0; q < 10; q++) obs10_out[o * 10 + q] = obs[q * Kp + s];
    pose_idx_out[o] = pose_idx[s];
  }
  const double *nb = newobs + (size_t)v * 10;
  if (nb[9] > 0.0) {
    for (int q = 0; q < 10; q++) obs10_out[o * 10 + q] = nb[q];
    pose_idx_out[o] = slot;
  }
  for (int q = 0; q < 10; q++) fix10_out[(size_t)nv * 10 + q] = fix ? fix[q * M + v] : 0.0;
  coe_out[nv] = coe[v] + nb[9];  // push_voxel: coe = sum of N over the window (:42-44)
}
}  // namespace

int append_scan_build(balm_ctx *c, int64_t n, const float *xyz_h, const double *poses_dev, int slot, int64_t *M_out,
                      int64_t *K_out, int64_t *matched_out,
                      int (*reg)(balm_ctx *, int64_t, const int64_t *, const int32_t *, const double *, const double *,
                                 const double *, int64_t)) {
  const int64_t M = c->M;
  cudaStream_t st = c->stream;
  ScratchPool pool;
  AssocParams P{c->assoc_voxel_size, c->assoc_layer_limit, c->assoc_min_ps, {c->assoc_eig[0], c->assoc_eig[1], c->assoc_eig[2]}};
  float *xyz = nullptr;
  unsigned *vid = nullptr, *vid2 = nullptr;
  int *idx = nullptr, *idx2 = nullptr, *keep = nullptr, *cnt = nullptr, *vscan = nullptr, *oscan = nullptr, *bad = nullptr;
  double *newobs = nullptr;
  ATRY(pool.get(&xyz, (size_t)3 * n)); ATRY(pool.get(&vid, (size_t)n)); ATRY(pool.get(&vid2, (size_t)n));
  ATRY(pool.get(&idx, (size_t)n)); ATRY(pool.get(&idx2, (size_t)n)); ATRY(pool.get(&newobs, (size_t)M * 10));
  ATRY(pool.get(&keep, (size_t)M)); ATRY(pool.get(&cnt, (size_t)M)); ATRY(pool.get(&vscan, (size_t)M)); ATRY(pool.get(&oscan, (size_t)M));
  ATRY(pool.get(&bad, 1));
  CUDA_TRY(cudaMemcpyAsync(xyz, xyz_h, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned gp = (unsigned)((n + 255) / 256), gv = (unsigned)((M + 127) / 128);
  append_match_kernel<<<gp, 256, 0, st>>>(xyz, poses_dev + 12 * slot, n, P, c->vox_key, c->vox_layer, (int)M, vid, bad);
  iota_kernel<<<gp, 256, 0, st>>>(idx, n);
  int bits = 1;
  while ((1ll << bits) <= M) bits++;
  size_t tb = 0, tb2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, vid, vid2, idx, idx2, (int)n, 0, bits, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb2, keep, vscan, (int)M, st);
  tb = std::max(tb, tb2);
  char *tmp = nullptr;
  ATRY(pool.get(&tmp, tb));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, vid, vid2, idx, idx2, (int)n, 0, bits, st));  // stable: point order kept
  append_cluster_kernel<<<gv, 128, 0, st>>>(xyz, vid2, idx2, n, (int)M, newobs);
  append_judge_kernel<<<gv, 128, 0, st>>>(c->row_ptr, c->pose_idx, c->obs, c->Kp, c->fix, M, poses_dev, newobs, slot, c->vox_layer, P,
                                          keep, cnt, bad);
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tb, keep, vscan, (int)M, st));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, oscan, (int)M, st));
  int last[5] = {0, 0, 0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(&last[0], keep + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[1], vscan + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[2], cnt + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[3], oscan + M - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last[4], bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (last[4]) { balm_set_error("balm_append_scan: `slot` must be above every scan already observing the voxels"); return BALM_ERR_INVALID; }
  const int64_t Mout = (int64_t)last[0] + last[1], Kout = (int64_t)last[2] + last[3];
  if (Mout < 1) { balm_set_error("balm_append_scan: no plane voxel survives the re-judgement"); return BALM_ERR_INVALID; }
  if (matched_out) {  // points that fell into a registered voxel: n minus the leading zeros of the sorted ids
    std::vector<unsigned> h((size_t)n);
    CUDA_TRY(cudaMemcpy(h.data(), vid2, sizeof(unsigned) * (size_t)n, cudaMemcpyDeviceToHost));
    *matched_out = n - (std::upper_bound(h.begin(), h.end(), 0u) - h.begin());
  }
  long long *rp = nullptr;
  int *pi = nullptr;
  double *ob = nullptr, *fx = nullptr, *co = nullptr;
  unsigned long long *keys_out = nullptr;
  int *layers_out = nullptr;
  ATRY(pool.get(&rp, (size_t)Mout + 1)); ATRY(pool.get(&pi, (size_t)Kout)); ATRY(pool.get(&ob, (size_t)Kout * 10));
  ATRY(pool.get(&fx, (size_t)Mout * 10)); ATRY(pool.get(&co, (size_t)Mout)); ATRY(pool.get(&keys_out, (size_t)Mout));
  ATRY(pool.get(&layers_out, (size_t)Mout));
  append_emit_kernel<<<gv, 128, 0, st>>>(c->row_ptr, c->pose_idx, c->obs, c->Kp, c->fix, c->coe, M, newobs, slot, keep, vscan, oscan,
                                         c->vox_key, c->vox_layer, rp, pi, ob, fx, co, keys_out, layers_out, Mout, Kout);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  c->launches += 8;
  *M_out = Mout;
  *K_out = Kout;
  const bool any_fix = c->fix != nullptr;
  ATRY(reg(c, Mout, (const int64_t *)rp, pi, ob, any_fix ? fx : nullptr, co, Kout));
  return install_keys(c, keys_out, layers_out, Mout, true);
}
