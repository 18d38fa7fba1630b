// balm_api.cu -- C ABI of libbalm_b200.so (include/balm_b200.h): context, problem registration, the
// evaluation pipeline, the damped solve and the Levenberg-Marquardt driver of BALM2::damping_iter
// (/root/reference/src/benchmark/bavoxel.hpp:1069-1166).  Host code is plain C++; all arithmetic of the hot
// path runs in the CUDA kernels of this directory.  There is no CPU fallback: without a CUDA device every
// entry point fails with BALM_ERR_CUDA.
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <time.h>
#include "internal.cuh"

void synth_host_poses(int N, uint64_t seed, double *poses_gt, double *poses_init);

static thread_local std::string g_err;
void balm_set_error(const std::string &s) { g_err = s; }
extern "C" const char *balm_last_error(void) { return g_err.c_str(); }
extern "C" int balm_version(void) { return 100; }

// ---------------- NCCL through dlopen (no link-time dependency) ----------------
struct NcclDyn {
  void *h = nullptr;
  typedef struct { char internal[128]; } uid_t;
  int (*GetUniqueId)(uid_t *) = nullptr;
  int (*CommInitRank)(void **, int, uid_t, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
static NcclDyn g_nccl;

// ---------------- helpers ----------------
template <typename T>
static int dev_alloc(T **p, size_t count) {
  CUDA_TRY(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
  return BALM_OK;
}
#define TRY(x)                \
  do {                        \
    int _s = (x);             \
    if (_s != BALM_OK) return _s; \
  } while (0)

static void free_problem(balm_ctx *c) {
  cudaFree(c->obs); cudaFree(c->pose_idx); cudaFree(c->row_ptr); cudaFree(c->coe); cudaFree(c->fix);
  cudaFree(c->csc_ptr); cudaFree(c->csc_obs); cudaFree(c->csc_vox); cudaFree(c->csc_lo); cudaFree(c->csc_hi);
  c->csc_lo = c->csc_hi = nullptr;
  c->obs = nullptr; c->pose_idx = nullptr; c->row_ptr = nullptr; c->coe = nullptr; c->fix = nullptr;
  c->csc_ptr = c->csc_obs = c->csc_vox = nullptr;
  cudaFree(c->vsums); c->vsums = nullptr;
  cudaFree(c->stats); cudaFree(c->stats_trial); cudaFree(c->G); cudaFree(c->obs_part); cudaFree(c->syrk_part);
  c->stats = c->stats_trial = c->G = c->obs_part = c->syrk_part = nullptr;
  tensor_syrk_free(c);
  c->M = c->K = c->Kp = 0;
  c->registered = false;
}
static void drop_keys(balm_ctx *c) {  // a voxel set registered from arrays carries no octree keys
  cudaFree(c->vox_key);
  cudaFree(c->vox_layer);
  c->vox_key = nullptr;
  c->vox_layer = nullptr;
  c->has_keys = false;
}

static int alloc_problem_arrays(balm_ctx *c, int64_t M, int64_t K, bool with_fix) {
  // A window of the same shape as the previous one (the common case when BA runs scan after scan, and what every
  // repetition of a benchmark does) keeps every device buffer: cudaFree/cudaMalloc of several GB cost milliseconds.
  c->reuse_ws = c->obs && c->stats && c->M == M && c->K == K && (c->fix != nullptr) == with_fix &&
                !getenv("BALM_NO_BUFFER_REUSE");
  if (c->reuse_ws) {
    cudaFree(c->csc_ptr); cudaFree(c->csc_obs); cudaFree(c->csc_vox); cudaFree(c->csc_lo); cudaFree(c->csc_hi);
    c->csc_ptr = c->csc_obs = c->csc_vox = c->csc_lo = c->csc_hi = nullptr;
    return BALM_OK;
  }
  free_problem(c);
  c->M = M; c->K = K;
  c->Kp = (K + 31) / 32 * 32;
  TRY(dev_alloc(&c->obs, (size_t)10 * c->Kp));
  TRY(dev_alloc(&c->pose_idx, (size_t)K));
  TRY(dev_alloc(&c->row_ptr, (size_t)M + 1));
  TRY(dev_alloc(&c->coe, (size_t)M));
  if (with_fix) TRY(dev_alloc(&c->fix, (size_t)10 * M));
  return BALM_OK;
}

// Workspaces that depend on the registered problem (batch size, split counts).
static int alloc_workspaces(balm_ctx *c) {
  if (c->reuse_ws) {  // same shape: batch size, split counts, tensor maps and digit-plane buffers are all unchanged
    c->reuse_ws = false;
    tensor_syrk_new_problem(c);
    return BALM_OK;
  }
  size_t g_budget = (size_t)16 << 30;  // bytes for the fp64 G' batch
  if (const char *e = getenv("BALM_G_BUDGET_MB")) g_budget = (size_t)atoll(e) << 20;  // tests: force batching
  int64_t vb = (int64_t)(g_budget / ((size_t)3 * c->ldg * sizeof(double)));
  if (vb > c->M) vb = c->M;
  if (vb < 1) vb = 1;
  c->VB = vb;
  TRY(dev_alloc(&c->stats, (size_t)vb * BALM_STATS_STRIDE));
  if (vb == c->M && !getenv("BALM_NO_STATS_CACHE")) TRY(dev_alloc(&c->stats_trial, (size_t)vb * BALM_STATS_STRIDE));
  if (!getenv("BALM_NO_STATS_SPLIT")) TRY(dev_alloc(&c->vsums, (size_t)vb * 10));
  if (c->prec == BALM_PREC_FP64) {  // the tensor path writes int8 digit planes directly and never stores fp64 G'
    TRY(dev_alloc(&c->G, (size_t)3 * vb * c->ldg));
    CUDA_TRY(cudaMemsetAsync(c->G, 0, sizeof(double) * (size_t)3 * vb * c->ldg, c->stream));  // zero the column padding
  }
  const int tiles_p = (c->N + 31) / 32;
  c->obs_chunks = std::max(1, (c->sm_count * 16 + tiles_p - 1) / tiles_p);
  if (c->obs_chunks > 4096) c->obs_chunks = 4096;
  TRY(dev_alloc(&c->obs_part, (size_t)c->obs_chunks * BALM_ACC * c->Np));
  CUDA_TRY(cudaMemsetAsync(c->obs_part, 0, sizeof(double) * (size_t)c->obs_chunks * BALM_ACC * c->Np, c->stream));
  // SYRK work decomposition: (tile, k-split) items, sized so the last wave is nearly full
  c->syrk_nb = c->ldg / BALM_SYRK_TILE;
  c->syrk_tiles = c->syrk_nb * (c->syrk_nb + 1) / 2;
  const int64_t rows = 3 * vb;
  int max_splits = (int)std::max<int64_t>(1, std::min<int64_t>(rows / 512, 64));
  const size_t part_budget = (size_t)3 << 30;
  const size_t per_split = (size_t)c->syrk_tiles * BALM_SYRK_TILE * BALM_SYRK_TILE * sizeof(double);
  max_splits = (int)std::max<size_t>(1, std::min<size_t>(max_splits, part_budget / per_split));
  int best = 1;
  double best_eff = 0;
  // every split costs one fp64 partial tile per output tile (written by the SYRK epilogue, read by the assembly):
  // the tensor path (6 ms SYRK) takes the SMALLEST split count that keeps the last wave >= 95 % full and respects
  // the int32 exactness bound (<= 32704 contraction rows per item); the fp64 path (97 ms SYRK) maximises balance.
  const int min_splits = c->prec == BALM_PREC_TENSOR ? (int)((rows + 32575) / 32576) : 1;  // per-item rows (rounded up to 64) stay <= 32640
  for (int s = std::min(min_splits, max_splits); s <= max_splits; s++) {
    const int items = c->syrk_tiles * s;
    const int waves = (items + c->sm_count - 1) / c->sm_count;
    const double eff = (double)items / ((double)waves * c->sm_count);
    if (c->prec == BALM_PREC_TENSOR) {
      if (eff > best_eff) { best_eff = eff; best = s; }
      if (eff >= 0.95) break;
    } else if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = s;
    }
  }
  if (const char *e = getenv("BALM_SYRK_SPLITS")) best = std::max(1, std::min(atoi(e), max_splits));  // tuning knob
  c->syrk_splits = best;
  TRY(dev_alloc(&c->syrk_part, (size_t)best * c->syrk_tiles * BALM_SYRK_TILE * BALM_SYRK_TILE));
  if (c->prec == BALM_PREC_TENSOR) TRY(tensor_syrk_init(c));
  return BALM_OK;
}

// ---------------- ctx ----------------
extern "C" int balm_create(balm_ctx **out, int n_poses, int device, int precision) {
  if (!out || n_poses < 1) { balm_set_error("balm_create: bad arguments"); return BALM_ERR_INVALID; }
  if (precision != BALM_PREC_FP64 && precision != BALM_PREC_TENSOR) {
    balm_set_error("balm_create: unknown precision mode");
    return BALM_ERR_INVALID;
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { balm_set_error("balm_create: no such CUDA device"); return BALM_ERR_CUDA; }
  CUDA_TRY(cudaSetDevice(device));
  balm_ctx *c = new balm_ctx();
  c->N = n_poses; c->n = 6 * n_poses; c->device = device; c->prec = precision;
  c->ldg = (c->n + BALM_SYRK_TILE - 1) / BALM_SYRK_TILE * BALM_SYRK_TILE;
  c->Np = (n_poses + 31) / 32 * 32;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  int prio_least = 0, prio_greatest = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_least));
  for (auto &e : c->sev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  TRY(factor_kernels_setup());
  TRY(syrk_f64_setup());
  TRY(ldlt_setup());
  TRY(covariance_setup());
  const size_t n = c->n;
  TRY(dev_alloc(&c->poses, 12 * (size_t)c->N));
  TRY(dev_alloc(&c->poses_trial, 12 * (size_t)c->N));
  TRY(dev_alloc(&c->H, n * n + n + 8));
  c->g = c->H + n * n;
  TRY(dev_alloc(&c->A, n * n));
  TRY(dev_alloc(&c->W, 2 * n * BALM_NB));
  TRY(dev_alloc(&c->dx, n));
  TRY(dev_alloc(&c->dvec, n));
  TRY(dev_alloc(&c->scal, 32));
  TRY(dev_alloc(&c->flags, 8));
  TRY(dev_alloc(&c->planes, 2 * (size_t)c->N));
  TRY(dev_alloc(&c->accum, (size_t)BALM_ACC * c->Np));
  TRY(dev_alloc(&c->accum_batch, (size_t)BALM_ACC * c->Np));
  c->res_blocks = c->sm_count * 24;  // residual partials: one per warp of the stats kernel (3 CTAs x 8 warps per SM)
  TRY(dev_alloc(&c->res_part, (size_t)c->res_blocks));
  CUDA_TRY(cudaMallocHost((void **)&c->h_scal, 16 * sizeof(double)));
  CUDA_TRY(cudaMallocHost((void **)&c->h_flags, 8 * sizeof(int)));
  memset(c->h_flags, 0, 8 * sizeof(int));
  for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
  *out = c;
  return BALM_OK;
}

extern "C" int balm_destroy(balm_ctx *c) {
  if (!c) return BALM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_problem(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  cudaFree(c->poses); cudaFree(c->poses_trial); cudaFree(c->H); cudaFree(c->A); cudaFree(c->W);
  cudaFree(c->dx); cudaFree(c->dvec); cudaFree(c->scal); cudaFree(c->flags); cudaFree(c->accum); cudaFree(c->accum_batch);
  cudaFree(c->vox_key); cudaFree(c->vox_layer); cudaFree(c->res_part); cudaFree(c->planes); cudaFree(c->Xinv); cudaFree(c->dinv); cudaFree(c->sol); cudaFree(c->ysol);
  cudaFree(c->Hpack); cudaFree(c->dval); cudaFree(c->rres); cudaFree(c->rdelta); cudaFree(c->dag_flags);
  if (c->solve_graph) cudaGraphExecDestroy((cudaGraphExec_t)c->solve_graph);
  cudaFreeHost(c->h_scal); cudaFreeHost(c->h_flags);
  for (auto &e : c->ev) if (e) cudaEventDestroy(e);
  for (auto &e : c->sev) if (e) cudaEventDestroy(e);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
  return BALM_OK;
}

// ---------------- problem registration ----------------
// One thread per voxel: checks the CSR invariants (pose_idx ascending, in range), finds the widest voxel and whether
// every voxel is observed by every pose; one thread per observation builds the per-pose voxel histogram
// (the ">= 20 planes per pose" precheck of bavoxel.hpp:1071-1085). Replaces a host loop over all K observations.
__global__ void csr_check_kernel(const long long *row_ptr, const int *pose_idx, int64_t M, int N, int *out /*[3]*/) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M) return;
  const long long s0 = row_ptr[v], s1 = row_ptr[v + 1];
  const int k = (int)(s1 - s0);
  bool bad = (s1 < s0);
  int prev = -1;
  for (long long s = s0; s < s1; s++) {
    const int p = pose_idx[s];
    if (p <= prev || p >= N) bad = true;
    prev = p;
  }
  if (bad) atomicOr(&out[0], 1);
  if (k != N) atomicOr(&out[1], 1);  // not dense
  atomicMax(&out[2], k);
}
__global__ void pose_hist_kernel(const int *pose_idx, int64_t K, int N, int *planes) {
  extern __shared__ int hist[];
  for (int e = threadIdx.x; e < N; e += blockDim.x) hist[e] = 0;
  __syncthreads();
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < K; s += (int64_t)gridDim.x * blockDim.x) {
    const int p = pose_idx[s];
    if (p >= 0 && p < N) atomicAdd(&hist[p], 1);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x)
    if (hist[e]) atomicAdd(&planes[e], hist[e]);
}
static int finish_registration(balm_ctx *c) {
  const int N = c->N;
  const int64_t M = c->M;
  int *d_out = nullptr, *d_planes = c->planes;  // c->planes [N]: per-pose voxel counts, kept for the collective guard
  TRY(dev_alloc(&d_out, 4));
  CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(int) * 4, c->stream));
  CUDA_TRY(cudaMemsetAsync(d_planes, 0, sizeof(int) * N, c->stream));
  csr_check_kernel<<<(unsigned)((M + 127) / 128), 128, 0, c->stream>>>(c->row_ptr, c->pose_idx, M, N, d_out);
  pose_hist_kernel<<<c->sm_count * 4, 256, sizeof(int) * N, c->stream>>>(c->pose_idx, c->K, N, d_planes);
  c->launches += 2;
  int h_out[4];
  std::vector<int> planes(N);
  CUDA_TRY(cudaMemcpyAsync(h_out, d_out, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(planes.data(), d_planes, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(d_out);
  if (h_out[0]) {
    free_problem(c);  // leave the context without a problem rather than with a malformed one
    balm_set_error("balm_set_voxels: pose_idx must be ascending and in [0,N) inside each voxel");
    return BALM_ERR_INVALID;
  }
  const bool dense = h_out[1] == 0;
  c->dense = dense;
  c->max_k = h_out[2];
  c->min_planes = *std::min_element(planes.begin(), planes.end());
  c->registered = true;
  if (!dense) TRY(build_pose_major_lists(c, planes.data()));  // pose-major lists for the observation sweep (device-built)
  return alloc_workspaces(c);
}

__global__ void aos_to_soa_kernel(const double *aos, double *soa, int64_t count, int64_t stride) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
#pragma unroll
  for (int c = 0; c < 10; c++) soa[c * stride + s] = aos[s * 10 + c];
}
__global__ void soa_to_aos_kernel(const double *soa, double *aos, int64_t count, int64_t stride) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
#pragma unroll
  for (int c = 0; c < 10; c++) aos[s * 10 + c] = soa[c * stride + s];
}

static int upload_aos(balm_ctx *c, const double *aos, bool aos_on_device, double *soa, int64_t count, int64_t stride) {
  // stage the AoS records through a device scratch buffer and transpose on the GPU
  const int64_t chunk = 1 << 22;
  double *scratch = nullptr;
  if (!aos_on_device) TRY(dev_alloc(&scratch, (size_t)std::min(chunk, count) * 10));
  for (int64_t b = 0; b < count; b += chunk) {
    const int64_t m = std::min(chunk, count - b);
    const double *src = aos + b * 10;
    if (!aos_on_device) {
      CUDA_TRY(cudaMemcpyAsync(scratch, src, sizeof(double) * m * 10, cudaMemcpyHostToDevice, c->stream));
      src = scratch;
    }
    aos_to_soa_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(src, soa + b, m, stride);
    c->launches++;
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(scratch);
  return BALM_OK;
}

// An empty voxel set: evaluations give H = 0, g = 0, r = 0 (and still take part in the all-reduce).
static int register_empty(balm_ctx *c) {
  free_problem(c);
  CUDA_TRY(cudaMemsetAsync(c->planes, 0, sizeof(int) * c->N, c->stream));
  c->dense = true; c->max_k = 0; c->min_planes = 0; c->VB = 0;
  c->registered = true;
  return BALM_OK;
}

static double host_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

extern "C" int balm_set_voxels(balm_ctx *c, int64_t M, const int64_t *row_ptr, const int32_t *pose_idx,
                               const double *obs10, const double *fix10, const double *coe) {
  const bool trace = getenv("BALM_TRACE_REG") != nullptr;  // host-side stage times of the registration
  const double t_begin = host_ms();
  if (!c || M < 0 || !row_ptr || (M > 0 && (!pose_idx || !obs10 || !coe))) {
    balm_set_error("balm_set_voxels: bad arguments");
    return BALM_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  drop_keys(c);
  if (M == 0) return register_empty(c);  // a rank of a multi-GPU job whose shard holds no voxels still joins every collective
  const int64_t K = row_ptr[M];
  if (row_ptr[0] != 0 || K < M) { balm_set_error("balm_set_voxels: bad row_ptr"); return BALM_ERR_INVALID; }
  if (K >= ((int64_t)1 << 31)) { balm_set_error("balm_set_voxels: more than 2^31 observations per GPU"); return BALM_ERR_UNSUPPORTED; }
  TRY(alloc_problem_arrays(c, M, K, fix10 != nullptr));
  const bool reused = c->reuse_ws;
  const double t_alloc = host_ms();
  CUDA_TRY(cudaMemcpyAsync(c->row_ptr, row_ptr, sizeof(int64_t) * (M + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->pose_idx, pose_idx, sizeof(int32_t) * K, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->coe, coe, sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
  if (trace) CUDA_TRY(cudaStreamSynchronize(c->stream));
  const double t_small = host_ms();
  TRY(upload_aos(c, obs10, false, c->obs, K, c->Kp));
  if (fix10) TRY(upload_aos(c, fix10, false, c->fix, M, M));
  const double t_obs = host_ms();
  const int rc = finish_registration(c);
  if (trace)
    fprintf(stderr, "[balm_set_voxels] reuse=%d alloc %.2f ms, csr+coe copies %.2f ms, observations %.2f ms, "
            "checks+workspaces %.2f ms\n", (int)reused, t_alloc - t_begin, t_small - t_alloc, t_obs - t_small,
            host_ms() - t_obs);
  return rc;
}

extern "C" int balm_set_voxels_dev(balm_ctx *c, int64_t M, const int64_t *row_ptr_dev, const int32_t *pose_idx_dev,
                                   const double *obs10_dev, const double *fix10_dev, const double *coe_dev,
                                   int64_t K) {
  if (!c || M < 1 || !row_ptr_dev || !pose_idx_dev || !obs10_dev || !coe_dev || K < M) {
    balm_set_error("balm_set_voxels_dev: bad arguments");
    return BALM_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  {  // the same checks the host path makes on row_ptr / K (the per-voxel invariants are checked by csr_check_kernel)
    int64_t ends[2] = {-1, -1};
    CUDA_TRY(cudaMemcpy(&ends[0], row_ptr_dev, sizeof(int64_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&ends[1], row_ptr_dev + M, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (ends[0] != 0 || ends[1] != K) { balm_set_error("balm_set_voxels_dev: row_ptr[0] != 0 or row_ptr[M] != n_obs"); return BALM_ERR_INVALID; }
    if (K >= ((int64_t)1 << 31)) { balm_set_error("balm_set_voxels_dev: more than 2^31 observations per GPU"); return BALM_ERR_UNSUPPORTED; }
  }
  drop_keys(c);
  TRY(alloc_problem_arrays(c, M, K, fix10_dev != nullptr));
  CUDA_TRY(cudaMemcpyAsync(c->row_ptr, row_ptr_dev, sizeof(int64_t) * (M + 1), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->pose_idx, pose_idx_dev, sizeof(int32_t) * K, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->coe, coe_dev, sizeof(double) * M, cudaMemcpyDeviceToDevice, c->stream));
  TRY(upload_aos(c, obs10_dev, true, c->obs, K, c->Kp));
  if (fix10_dev) TRY(upload_aos(c, fix10_dev, true, c->fix, M, M));
  return finish_registration(c);
}

extern "C" int64_t balm_num_obs(balm_ctx *c) { return c ? c->K : 0; }

// ---------------- association on the GPU (assoc.cu) ----------------
int assoc_build(balm_ctx *c, int64_t n, const float *xyz_h, const int *frame_h, const double *poses12_h,
                double voxel_size, int layer_limit, int min_ps, const double *eig3, int64_t *M_out, int64_t *K_out,
                int (*register_csr)(balm_ctx *, int64_t, int64_t));

static int assoc_register(balm_ctx *c, int64_t M, int64_t K) { return alloc_problem_arrays(c, M, K, false); }

extern "C" void balm_default_assoc_opts(balm_assoc_opts *o) {
  o->voxel_size = 1.0; o->layer_limit = 2; o->min_ps = 15;
  o->eigen_value_array[0] = o->eigen_value_array[1] = o->eigen_value_array[2] = 1.0 / 16;
}

extern "C" int balm_cut_voxels(balm_ctx *c, int64_t n_points, const float *xyz, const int32_t *frame,
                               const double *poses12, const balm_assoc_opts *opts, int64_t *n_voxels_out,
                               int64_t *n_obs_out) {
  if (!c || !xyz || !frame || !poses12) { balm_set_error("balm_cut_voxels: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  balm_assoc_opts o;
  if (opts) o = *opts; else balm_default_assoc_opts(&o);
  // the reference keeps the planarity thresholds in a `float eigen_value_array[]` (bavoxel.hpp:11) and compares the
  // double ratio with the promoted float (bavoxel.hpp:697): 1/9 is (double)(float)(1.0/9), not the double 1/9
  for (double &e : o.eigen_value_array) e = (double)(float)e;
  int64_t M = 0, K = 0;
  TRY(assoc_build(c, n_points, xyz, frame, poses12, o.voxel_size, o.layer_limit, o.min_ps, o.eigen_value_array, &M, &K,
                  assoc_register));
  if (n_voxels_out) *n_voxels_out = M;
  if (n_obs_out) *n_obs_out = K;
  return finish_registration(c);
}

extern "C" int balm_download_voxel_range(balm_ctx *c, int64_t head, int64_t end, int64_t *row_ptr, int32_t *pose_idx,
                                         double *obs10, double *coe, int64_t *n_obs_out) {
  if (!c || !c->obs) { balm_set_error("balm_download_voxels: no voxels registered"); return BALM_ERR_INVALID; }
  if (head < 0 || end > c->M || head > end) { balm_set_error("balm_download_voxel_range: bad voxel range"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  int64_t s0 = 0, s1 = 0;
  CUDA_TRY(cudaMemcpy(&s0, c->row_ptr + head, sizeof(int64_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(&s1, c->row_ptr + end, sizeof(int64_t), cudaMemcpyDeviceToHost));
  const int64_t K = s1 - s0;
  if (n_obs_out) *n_obs_out = K;
  if (row_ptr) {  // re-based to 0, like a freshly registered problem of end-head voxels
    CUDA_TRY(cudaMemcpy(row_ptr, c->row_ptr + head, sizeof(int64_t) * (end - head + 1), cudaMemcpyDeviceToHost));
    for (int64_t v = 0; v <= end - head; v++) row_ptr[v] -= s0;
  }
  if (pose_idx && K) CUDA_TRY(cudaMemcpy(pose_idx, c->pose_idx + s0, sizeof(int32_t) * K, cudaMemcpyDeviceToHost));
  if (coe && end > head) CUDA_TRY(cudaMemcpy(coe, c->coe + head, sizeof(double) * (end - head), cudaMemcpyDeviceToHost));
  if (obs10 && K) {
    const int64_t chunk = 1 << 22;
    double *scratch = nullptr;
    TRY(dev_alloc(&scratch, (size_t)std::min(chunk, K) * 10));
    for (int64_t b = 0; b < K; b += chunk) {
      const int64_t m = std::min(chunk, K - b);
      soa_to_aos_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->obs + s0 + b, scratch, m, c->Kp);
      CUDA_TRY(cudaMemcpyAsync(obs10 + b * 10, scratch, sizeof(double) * m * 10, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    cudaFree(scratch);
  }
  return BALM_OK;
}

extern "C" int balm_download_voxels(balm_ctx *c, int64_t *row_ptr, int32_t *pose_idx, double *obs10, double *coe) {
  if (!c || !c->obs) { balm_set_error("balm_download_voxels: no voxels registered"); return BALM_ERR_INVALID; }
  return balm_download_voxel_range(c, 0, c->M, row_ptr, pose_idx, obs10, coe, nullptr);
}

extern "C" int balm_synth_virtual(balm_ctx *c, int64_t M, int64_t first_voxel, int pts, double noise, double range,
                                  uint64_t seed, double *poses_gt, double *poses_init) {
  if (!c || M < 1 || pts < 1) { balm_set_error("balm_synth_virtual: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  const int64_t K = M * c->N;
  if (K >= ((int64_t)1 << 31)) { balm_set_error("balm_synth_virtual: more than 2^31 observations per GPU"); return BALM_ERR_UNSUPPORTED; }
  std::vector<double> gt(12 * (size_t)c->N), init(12 * (size_t)c->N);
  synth_host_poses(c->N, seed, gt.data(), init.data());
  if (poses_gt) memcpy(poses_gt, gt.data(), sizeof(double) * gt.size());
  if (poses_init) memcpy(poses_init, init.data(), sizeof(double) * init.size());
  drop_keys(c);
  TRY(alloc_problem_arrays(c, M, K, false));
  CUDA_TRY(cudaMemcpyAsync(c->poses_trial, gt.data(), sizeof(double) * gt.size(), cudaMemcpyHostToDevice, c->stream));
  TRY(launch_synth(c, M, first_voxel, pts, noise, range, seed, c->poses_trial));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->dense = true;
  c->max_k = c->N;
  c->min_planes = (int)std::min<int64_t>(M, 1 << 30);
  {
    std::vector<int> pl((size_t)c->N, c->min_planes);
    CUDA_TRY(cudaMemcpy(c->planes, pl.data(), sizeof(int) * c->N, cudaMemcpyHostToDevice));
  }
  c->registered = true;
  return alloc_workspaces(c);
}

// ---------------- evaluation ----------------
// Multi-GPU reduction of [H | g | r]: H is symmetric, so only its lower triangle travels. With a communicator attached the
// assembly kernel writes column c of the lower triangle (n - c entries) at offset c*n - c(c-1)/2 of the send buffer,
// followed by g and r, INSTEAD of the full matrix; after the all-reduce unpack_lower_kernel writes H (both halves), g and
// r. 36 MB instead of 72 MB on the wire at N = 500, and no extra pass over the matrix.
__global__ void unpack_lower_kernel(const double *pack, double *H, int n) {
  __shared__ double tile[32][33];
  const int bc = blockIdx.y, br = blockIdx.x;  // 32 x 32 block (block row br >= block column bc)
  if (br < bc) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r = br * 32 + tx, c = bc * 32 + ty;
  const size_t tri = (size_t)n * (n + 1) / 2;
  double v = 0.0;
  if (r < n && c < n && r >= c) {
    v = pack[(size_t)c * n - (size_t)c * (c - 1) / 2 + (r - c)];
    H[(size_t)c * n + r] = v;                      // lower element, coalesced in r
  }
  tile[ty][tx] = v;
  __syncthreads();
  const int r2 = br * 32 + ty, c2 = bc * 32 + tx;  // transposed read: upper element (row c2, col r2), coalesced in c2
  if (r2 < n && c2 < n && r2 > c2) H[(size_t)r2 * n + c2] = tile[tx][ty];
  if (br == 0 && bc == 0) {
    for (int e = ty * 32 + tx; e <= n; e += 1024) H[(size_t)n * n + e] = pack[tri + e];
  }
}

static int allreduce_sum(balm_ctx *c, double *buf, size_t count) {
  if (c->world <= 1 || !c->comm) return BALM_OK;
  const int rc = g_nccl.AllReduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm, c->stream);
  if (rc != 0) {
    balm_set_error(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    return BALM_ERR_NCCL;
  }
  return BALM_OK;
}

// H, g, r of voxels [head,end) at device poses `poses`; result in c->H | c->g | c->scal[0] (all-reduced).
// stats_cached: c->stats already holds the per-voxel eigen data of exactly these poses (left there by the residual
// pass of an accepted LM step, or by the evaluation before a rejected one) and c->scal[BALM_SCAL_RCUR] this rank's residual.
static int evaluate_dev(balm_ctx *c, const double *poses, int64_t head, int64_t end, bool include_fix,
                        bool stats_cached = false) {
  if (!c->registered) { balm_set_error("no voxels registered"); return BALM_ERR_INVALID; }
  if (head < 0 || end > c->M || head > end) { balm_set_error("evaluate: bad voxel range"); return BALM_ERR_INVALID; }
  double *r_dev = c->g + c->n;  // contiguous with H and g -> one all-reduce
  CUDA_TRY(cudaMemsetAsync(r_dev, 0, sizeof(double), c->stream));
  float ms;
  bool first = true;
  // defer mode (LM loop, whole problem in one batch): nothing is waited for here; collect_eval() reads the events and
  // the tensor path's flags after the iteration's single synchronisation
  const bool deferred = c->defer && head == 0 && end == c->M && c->VB >= c->M && end > head;
  c->pending_stats_cached = stats_cached;
  if (head == end) {  // empty range (or an empty shard): this rank contributes H = 0, g = 0, r = 0 -- and still enters
    // the all-reduce below, which the other ranks of a multi-GPU job are waiting in
    CUDA_TRY(cudaMemsetAsync(c->H, 0, sizeof(double) * ((size_t)c->n * c->n + c->n + 1), c->stream));
    for (int e = 0; e <= 5; e++) CUDA_TRY(cudaEventRecord(c->ev[e], c->stream));
  }
  for (int64_t v0 = head; v0 < end; v0 += c->VB) {
    const int64_t v1 = std::min(end, v0 + c->VB);
    TRY(launch_csc_batch(c, v0, v1));  // sparse problems: every pose's list restricted to this batch's voxels
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (stats_cached) CUDA_TRY(cudaMemcpyAsync(r_dev, c->scal + BALM_SCAL_RCUR, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    else TRY(launch_voxel_stats(c, poses, v0, v1, c->stats, include_fix, r_dev));
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    if (c->prec == BALM_PREC_TENSOR) {
      TRY(tensor_obs_and_syrk(c, poses, v0, v1, first));  // records ev[2] between the sweeps and the SYRK
    } else {
      TRY(launch_obs_pass(c, poses, v0, v1, first));
      CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
      TRY(launch_syrk_f64(c, 3 * (v1 - v0), first));
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    if (deferred) {
      if (c->prec == BALM_PREC_TENSOR) TRY(tensor_syrk_check_enqueue(c));
    } else {
      CUDA_TRY(cudaEventSynchronize(c->ev[3]));
      cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
      if (!stats_cached) { c->tm.ms_stats += ms; if (first) c->tm.n_stats += 1; }
      cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->tm.ms_obs += ms;
      cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->tm.ms_syrk += ms;
      if (c->prec == BALM_PREC_TENSOR) TRY(tensor_syrk_check(c));
    }
    first = false;
  }
  CUDA_TRY(cudaMemcpyAsync(c->scal + BALM_SCAL_RCUR, r_dev, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));  // local, pre all-reduce
  if (end > head) {
    CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
    TRY(launch_assemble(c));
    CUDA_TRY(cudaEventRecord(c->ev[5], c->stream));
  }
  if (c->world > 1 && c->comm) {  // the assembly wrote the lower triangle | g into c->Hpack (an empty range: zeros)
    const int n = c->n;
    const size_t tri = (size_t)n * (n + 1) / 2, cnt = tri + n + 1;
    if (end > head) CUDA_TRY(cudaMemcpyAsync(c->Hpack + tri + n, r_dev, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    else CUDA_TRY(cudaMemsetAsync(c->Hpack, 0, sizeof(double) * cnt, c->stream));
    TRY(allreduce_sum(c, c->Hpack, cnt));
    unpack_lower_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 32), 0, c->stream>>>(c->Hpack, c->H, n);
    c->launches += 1;
  }
  CUDA_TRY(cudaEventRecord(c->ev[6], c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->scal, r_dev, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (deferred) {
    c->pending_eval = true;
    return BALM_OK;
  }
  CUDA_TRY(cudaEventSynchronize(c->ev[6]));
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]); c->tm.ms_assemble += ms;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); c->tm.ms_allreduce += ms;
  c->tm.n_eval += 1;
  return BALM_OK;
}

// defer mode: the stream has been synchronised -> account the evaluation's phases, read the tensor path's flags
static int collect_eval(balm_ctx *c) {
  if (!c->pending_eval) return BALM_OK;
  c->pending_eval = false;
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  if (!c->pending_stats_cached) { c->tm.ms_stats += ms; c->tm.n_stats += 1; }
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->tm.ms_obs += ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->tm.ms_syrk += ms;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]); c->tm.ms_assemble += ms;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); c->tm.ms_allreduce += ms;
  c->tm.n_eval += 1;
  if (c->prec == BALM_PREC_TENSOR) return tensor_syrk_check_finish(c);
  return BALM_OK;
}

static int residual_dev(balm_ctx *c, const double *poses, double *host_out, bool keep_stats = false) {
  if (!c->registered) { balm_set_error("no voxels registered"); return BALM_ERR_INVALID; }
  double *r_dev = c->scal + 2;
  CUDA_TRY(cudaEventRecord(c->ev[7], c->stream));
  CUDA_TRY(cudaMemsetAsync(r_dev, 0, sizeof(double), c->stream));
  TRY(launch_voxel_stats(c, poses, 0, c->M, keep_stats ? c->stats_trial : nullptr, true, r_dev));
  CUDA_TRY(cudaMemcpyAsync(c->scal + BALM_SCAL_RTRIAL, r_dev, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));  // local, pre all-reduce
  TRY(allreduce_sum(c, r_dev, 1));
  CUDA_TRY(cudaEventRecord(c->ev[8], c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_scal + 2, r_dev, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[7], c->ev[8]);
  c->tm.ms_residual += ms;
  c->tm.n_residual += 1;
  *host_out = c->h_scal[2];
  return BALM_OK;
}

extern "C" int balm_evaluate(balm_ctx *c, const double *poses12, int64_t head, int64_t end, int include_fix,
                             double *H, double *g, double *residual) {
  if (!c || !poses12) { balm_set_error("balm_evaluate: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(c->poses, poses12, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, c->stream));
  TRY(evaluate_dev(c, c->poses, head, end, include_fix != 0));
  const size_t n = c->n;
  if (H) CUDA_TRY(cudaMemcpyAsync(H, c->H, sizeof(double) * n * n, cudaMemcpyDeviceToHost, c->stream));
  if (g) CUDA_TRY(cudaMemcpyAsync(g, c->g, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (residual) *residual = c->h_scal[0];
  return BALM_OK;
}

extern "C" int balm_residual(balm_ctx *c, const double *poses12, double *residual) {
  if (!c || !poses12 || !residual) { balm_set_error("balm_residual: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(c->poses_trial, poses12, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, c->stream));
  return residual_dev(c, c->poses_trial, residual);
}

static int collect_solve(balm_ctx *c, double *q1, int *not_pd) {
  float ms;
  cudaEventElapsedTime(&ms, c->ev[9], c->ev[10]);
  c->tm.ms_solve += ms;
  c->tm.n_solve += 1;
  *q1 = c->h_scal[1];
  *not_pd = c->h_flags[0] != 0 || !std::isfinite(*q1);
  return BALM_OK;
}

// The backward-error check of the last solve failed (flags[4]: an indefinite H + uD factored without pivoting lost
// digits): refine dx in fp64 with the same factors; if that does not bring the error down, the step counts as not_pd.
// Returns 1 in *refined when dx changed (the caller re-applies the pose update).
static int maybe_refine(balm_ctx *c, double *q1, int *not_pd, int *refined) {
  *refined = 0;
  if (c->h_flags[4] == 0 || *not_pd) return BALM_OK;
  int still_bad = 0;
  TRY(refine_solution(c, &still_bad));
  *q1 = c->h_scal[1];
  *refined = 1;
  if (still_bad || !std::isfinite(*q1)) *not_pd = 1;
  return BALM_OK;
}

static int solve_dev(balm_ctx *c, double u, double *q1, int *not_pd) {
  CUDA_TRY(cudaEventRecord(c->ev[9], c->stream));
  TRY(launch_ldlt_solve(c, u));
  CUDA_TRY(cudaEventRecord(c->ev[10], c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_scal + 1, c->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_flags, c->flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_flags + 4, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (c->defer) return BALM_OK;  // collect_solve() / maybe_refine() after the iteration's synchronisation
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  TRY(collect_solve(c, q1, not_pd));
  int refined = 0;
  return maybe_refine(c, q1, not_pd, &refined);
}

extern "C" int balm_solve(balm_ctx *c, double u, double *dx, double *q1, int *not_pd) {
  if (!c) { balm_set_error("balm_solve: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  double q = 0;
  int bad = 0;
  TRY(solve_dev(c, u, &q, &bad));
  if (dx) CUDA_TRY(cudaMemcpy(dx, c->dx, sizeof(double) * c->n, cudaMemcpyDeviceToHost));
  if (q1) *q1 = q;
  if (not_pd) *not_pd = bad;
  return BALM_OK;
}

extern "C" void balm_default_lm_opts(balm_lm_opts *o) {
  o->max_iter = 10; o->u0 = 0.01; o->v0 = 2; o->rel_tol = 1e-6; o->hess_includes_fix = 0; o->gauge_mode = 0;
  o->min_planes_per_pose = 20; o->verbose = 0; o->force_hess = 0;
}

extern "C" int balm_damping_iter(balm_ctx *c, double *poses12, const balm_lm_opts *o, balm_trace *trace, int *n_iters,
                                 double *poses_per_iter) {
  if (!c || !poses12 || !o) { balm_set_error("balm_damping_iter: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  if (n_iters) *n_iters = 0;
  if (!c->registered) { balm_set_error("no voxels registered"); return BALM_ERR_INVALID; }
  if (o->min_planes_per_pose > 0) {  // bavoxel.hpp:1071-1085 (the shim reproduces the printf + exit(0))
    // The guard is on the GLOBAL count of voxels seeing each pose. With several ranks every rank all-reduces the
    // per-pose histogram of its shard first, so all ranks take the same branch (a rank returning alone would leave
    // the others waiting in the next collective).
    long long global_min = c->min_planes;
    if (c->world > 1 && c->comm) {
      int *red = c->planes + c->N;
      const int rc = g_nccl.AllReduce(c->planes, red, (size_t)c->N, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, c->stream);
      if (rc != 0) { balm_set_error("ncclAllReduce (plane histogram) failed"); return BALM_ERR_NCCL; }
      std::vector<int> pl((size_t)c->N);
      CUDA_TRY(cudaMemcpyAsync(pl.data(), red, sizeof(int) * c->N, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      global_min = *std::min_element(pl.begin(), pl.end());
    }
    if (global_min < o->min_planes_per_pose) {
      balm_set_error("Initial error too large. Please loose plane determination criteria for more planes.");
      return BALM_ERR_TOO_FEW_PLANES;
    }
  }
  const size_t pbytes = sizeof(double) * 12 * c->N;
  CUDA_TRY(cudaMemcpyAsync(c->poses, poses12, pbytes, cudaMemcpyHostToDevice, c->stream));
  double u = o->u0, v = o->v0, r1 = 0, r2 = 0;
  bool calc_hess = true;
  // The residual pass at the trial poses computes the same per-voxel eigen data the next evaluation starts with:
  // when the step is accepted they are handed over (same kernel, same inputs -> identical values). Needs the whole
  // problem in one batch and the same treatment of the fix cluster in both passes.
  const bool hand_over = c->stats_trial != nullptr && (c->fix == nullptr || o->hess_includes_fix != 0);
  bool stats_cached = false;
  tensor_syrk_new_problem(c);  // nothing speculative is carried over from earlier calls
  struct LmScope {             // lm_active / defer are cleared on every exit path
    balm_ctx *c;
    ~LmScope() { c->lm_active = false; c->defer = false; c->pending_eval = false; }
  } lm_scope{c};
  c->lm_active = true;
  for (int it = 0; it < o->max_iter; it++) {
    // evaluation -> solve -> pose update -> trial residual are enqueued back to back; the host waits once, at the
    // end of residual_dev, and then reads r1, q1, the pivot flag and the phase timers
    c->defer = !getenv("BALM_SYNC_PHASES");
    int rc = BALM_OK;
    if (calc_hess) {
      rc = evaluate_dev(c, c->poses, 0, c->M, o->hess_includes_fix != 0, stats_cached);
      if (rc == BALM_OK && cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = BALM_ERR_CUDA;
    }
    double q1 = 0;
    int not_pd = 0;
    if (rc == BALM_OK) rc = solve_dev(c, u, &q1, &not_pd);
    if (rc == BALM_OK) rc = launch_pose_update(c, c->poses, c->dx, c->poses_trial);
    if (rc == BALM_OK) rc = residual_dev(c, c->poses_trial, &r2, hand_over);  // synchronises the stream
    const bool was_deferred = c->defer;
    c->defer = false;
    if (rc != BALM_OK) { c->pending_eval = false; return rc; }
    if (was_deferred) {
      TRY(collect_eval(c));
      TRY(collect_solve(c, &q1, &not_pd));
      int refined = 0;
      TRY(maybe_refine(c, &q1, &not_pd, &refined));
      if (refined && !not_pd) {  // dx changed: the trial poses and their residual are recomputed (rare path, synchronous)
        TRY(launch_pose_update(c, c->poses, c->dx, c->poses_trial));
        TRY(residual_dev(c, c->poses_trial, &r2, hand_over));
      }
    }
    if (calc_hess) r1 = c->h_scal[0];
    double q = r1 - r2;
    if (not_pd || !std::isfinite(r2)) q = -1.0;  // unusable step -> rejected, u *= v
    if (o->verbose)  // the reference's trace line (bavoxel.hpp:1132)
      printf("iter%d: (%lf %lf) u: %lf v: %.1lf q: %.3lf %lf %lf\n", it, r1, r2, u, v, q / q1, q1, q);
    balm_trace t{r1, r2, u, v, q, q1, 0, calc_hess ? 1 : 0, not_pd};
    if (q > 0) {  // bavoxel.hpp:1134-1143
      std::swap(c->poses, c->poses_trial);
      if (hand_over) {
        std::swap(c->stats, c->stats_trial);
        CUDA_TRY(cudaMemcpyAsync(c->scal + BALM_SCAL_RCUR, c->scal + BALM_SCAL_RTRIAL, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      }
      stats_cached = hand_over;
      const double rho = q / q1;
      v = 2;
      const double f = 1 - pow(2 * rho - 1, 3);
      u *= (f < (1.0 / 3.0) ? (1.0 / 3.0) : f);
      calc_hess = true;
      t.accepted = 1;
    } else {  // bavoxel.hpp:1144-1149
      u = u * v;
      v = 2 * v;
      calc_hess = o->force_hess != 0;
      stats_cached = c->stats_trial != nullptr;  // c->stats still describes c->poses (single batch)
    }
    if (trace) trace[it] = t;
    if (poses_per_iter)
      CUDA_TRY(cudaMemcpyAsync(poses_per_iter + (size_t)it * 12 * c->N, c->poses, pbytes, cudaMemcpyDeviceToHost, c->stream));
    if (n_iters) *n_iters = it + 1;
    if (o->rel_tol >= 0 && fabs(r1 - r2) / r1 < o->rel_tol) break;  // bavoxel.hpp:1155
  }
  if (o->gauge_mode == 0 || o->gauge_mode == 1) TRY(launch_gauge(c, c->poses, o->gauge_mode));
  CUDA_TRY(cudaMemcpyAsync(poses12, c->poses, pbytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return BALM_OK;
}

// ---------------- sliding-window marginalisation (assoc.cu) ----------------
int marginalize_build(balm_ctx *c, int mg, const double *poses_dev, int min_ps, int64_t *M_out, int64_t *K_out,
                      int (*reg)(balm_ctx *, int64_t, const int64_t *, const int32_t *, const double *, const double *,
                                 const double *, int64_t));
extern "C" int balm_marginalize(balm_ctx *c, int mg_size, const double *poses12, int min_ps, int64_t *n_voxels_out,
                                int64_t *n_obs_out) {
  if (!c || !poses12 || mg_size < 1 || mg_size >= c->N) { balm_set_error("balm_marginalize: bad arguments"); return BALM_ERR_INVALID; }
  if (!c->registered || c->M < 1) { balm_set_error("balm_marginalize: no voxels registered"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(c->poses, poses12, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, c->stream));
  int64_t M = 0, K = 0;
  TRY(marginalize_build(c, mg_size, c->poses, min_ps, &M, &K, balm_set_voxels_dev));
  if (n_voxels_out) *n_voxels_out = M;
  if (n_obs_out) *n_obs_out = K;
  return BALM_OK;
}
int append_scan_build(balm_ctx *c, int64_t n, const float *xyz_h, const double *poses_dev, int slot, int64_t *M_out,
                      int64_t *K_out, int64_t *matched_out,
                      int (*reg)(balm_ctx *, int64_t, const int64_t *, const int32_t *, const double *, const double *,
                                 const double *, int64_t));
extern "C" int balm_append_scan(balm_ctx *c, int64_t n_points, const float *xyz, const double *poses12, int slot,
                                int64_t *n_voxels_out, int64_t *n_obs_out, int64_t *n_matched_out) {
  if (!c || !xyz || !poses12 || n_points < 1 || n_points >= (1ll << 31) || slot < 0 || slot >= c->N) {
    balm_set_error("balm_append_scan: bad arguments");
    return BALM_ERR_INVALID;
  }
  if (!c->registered || c->M < 1 || !c->has_keys) {
    balm_set_error("balm_append_scan: needs a voxel set produced by balm_cut_voxels (octree keys)");
    return BALM_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(c->poses, poses12, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, c->stream));
  int64_t M = 0, K = 0;
  TRY(append_scan_build(c, n_points, xyz, c->poses, slot, &M, &K, n_matched_out, balm_set_voxels_dev));
  if (n_voxels_out) *n_voxels_out = M;
  if (n_obs_out) *n_obs_out = K;
  return BALM_OK;
}
extern "C" int balm_download_keys(balm_ctx *c, uint64_t *keys, int32_t *layers) {
  if (!c || !c->has_keys) { balm_set_error("balm_download_keys: the voxel set carries no octree keys"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  if (keys) CUDA_TRY(cudaMemcpy(keys, c->vox_key, sizeof(uint64_t) * (size_t)c->M, cudaMemcpyDeviceToHost));
  if (layers) CUDA_TRY(cudaMemcpy(layers, c->vox_layer, sizeof(int32_t) * (size_t)c->M, cudaMemcpyDeviceToHost));
  return BALM_OK;
}
extern "C" int balm_download_fix(balm_ctx *c, double *fix10) {
  if (!c || !fix10 || !c->registered) { balm_set_error("balm_download_fix: bad arguments"); return BALM_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(c->device));
  std::vector<double> soa((size_t)10 * c->M, 0.0);
  if (c->fix) CUDA_TRY(cudaMemcpy(soa.data(), c->fix, sizeof(double) * 10 * c->M, cudaMemcpyDeviceToHost));
  for (int64_t v = 0; v < c->M; v++)
    for (int q = 0; q < 10; q++) fix10[v * 10 + q] = soa[(size_t)q * c->M + v];
  return BALM_OK;
}

// ---------------- pose covariance (covariance.cu) ----------------
int pose_covariance_dev(balm_ctx *c, const double *poses_dev, const double *ccov_host, double point_noise, bool include_fix,
                        double *Rraw_host, double *Rcov_host);
int balm_cov_evaluate_hessian(balm_ctx *c, const double *poses_dev, bool include_fix) {
  return evaluate_dev(c, poses_dev, 0, c->M, include_fix);
}
extern "C" int balm_pose_covariance(balm_ctx *c, const double *poses12, const double *c_cov81, double point_noise,
                                    int include_fix, double *Rcov_raw, double *Rcov) {
  if (!c || !poses12 || (!Rcov_raw && !Rcov)) { balm_set_error("balm_pose_covariance: bad arguments"); return BALM_ERR_INVALID; }
  if (!c->registered || c->M < 1) { balm_set_error("balm_pose_covariance: no voxels registered"); return BALM_ERR_INVALID; }
  if (c->world > 1) { balm_set_error("balm_pose_covariance: single-GPU contexts only"); return BALM_ERR_UNSUPPORTED; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(c->poses, poses12, sizeof(double) * 12 * c->N, cudaMemcpyHostToDevice, c->stream));
  return pose_covariance_dev(c, c->poses, c_cov81, point_noise, include_fix != 0, Rcov_raw, Rcov);
}

// ---------------- multi-GPU ----------------
extern "C" int balm_comm_unique_id(void *out128) {
  if (!out128) return BALM_ERR_INVALID;
  if (!g_nccl.load()) { balm_set_error("libnccl.so.2 not found"); return BALM_ERR_NCCL; }
  NcclDyn::uid_t id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) { balm_set_error("ncclGetUniqueId failed"); return BALM_ERR_NCCL; }
  memcpy(out128, &id, 128);
  return BALM_OK;
}

extern "C" int balm_comm_init(balm_ctx *c, int rank, int world, const void *unique_id128) {
  if (!c || world < 1 || rank < 0 || rank >= world) { balm_set_error("balm_comm_init: bad arguments"); return BALM_ERR_INVALID; }
  c->rank = rank; c->world = world;
  if (world == 1) return BALM_OK;
  if (!unique_id128) { balm_set_error("balm_comm_init: unique id required"); return BALM_ERR_INVALID; }
  if (!g_nccl.load()) { balm_set_error("libnccl.so.2 not found"); return BALM_ERR_NCCL; }
  CUDA_TRY(cudaSetDevice(c->device));
  NcclDyn::uid_t id;
  memcpy(&id, unique_id128, 128);
  if (!c->Hpack) TRY(dev_alloc(&c->Hpack, (size_t)c->n * (c->n + 1) / 2 + c->n + 1));
  const int rc = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (rc != 0) {
    balm_set_error(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    return BALM_ERR_NCCL;
  }
  return BALM_OK;
}

// ---------------- instrumentation ----------------
extern "C" int balm_get_timings(balm_ctx *c, balm_timings *out) {
  if (!c || !out) return BALM_ERR_INVALID;
  *out = c->tm;
  out->launches = (int)c->launches;
  return BALM_OK;
}
extern "C" int balm_reset_counters(balm_ctx *c) {
  if (!c) return BALM_ERR_INVALID;
  c->launches = 0;
  c->tm = balm_timings{};
  return BALM_OK;
}
extern "C" int balm_sync(balm_ctx *c) {
  if (!c) return BALM_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return BALM_OK;
}
extern "C" int balm_timer_begin(balm_ctx *c) {
  if (!c) return BALM_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaEventRecord(c->ev[11], c->stream));
  return BALM_OK;
}
extern "C" int balm_timer_end(balm_ctx *c, float *ms) {
  if (!c || !ms) return BALM_ERR_INVALID;
  cudaEvent_t e;
  CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(cudaEventRecord(e, c->stream));
  CUDA_TRY(cudaEventSynchronize(e));
  CUDA_TRY(cudaEventElapsedTime(ms, c->ev[11], e));
  cudaEventDestroy(e);
  return BALM_OK;
}
// Debug aid (BALM_DAG_TRACE=1): timestamps of the last tile-DAG factorisation. out: 4*nt chain entries (step start,
// mini-panel inputs ready, diagonal block ready, step published; ns) followed by 3 per CTA (start, end, tasks).
extern "C" int balm_debug_dag_trace(balm_ctx *c, unsigned long long *out, int max_entries) {
  if (!c || !out || !c->dag_trace) return BALM_ERR_INVALID;
  const int nt = (c->n + BALM_NB - 1) / BALM_NB;
  const int total = std::min(max_entries, 4 * nt + 8 * 1024);
  CUDA_TRY(cudaMemcpy(out, c->dag_trace, sizeof(unsigned long long) * total, cudaMemcpyDeviceToHost));
  return BALM_OK;
}

int tensor_overlap_probe(balm_ctx *c, const double *poses, int reps, float *out);
extern "C" int balm_debug_overlap_probe(balm_ctx *c, int reps, float *out5) {
  if (!c || !out5 || c->prec != BALM_PREC_TENSOR) return BALM_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(c->device));
  return tensor_overlap_probe(c, c->poses, reps, out5);
}

extern "C" int balm_device_views(balm_ctx *c, double **H_dev, double **g_dev) {
  if (!c) return BALM_ERR_INVALID;
  if (H_dev) *H_dev = c->H;
  if (g_dev) *g_dev = c->g;
  return BALM_OK;
}
