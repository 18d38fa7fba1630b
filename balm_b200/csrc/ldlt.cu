// ldlt.cu -- dense damped solve (kernel K6) and the scalar side of the LM step (K7).
//
//   D = diag(H);  A = H + u*D;  A = L diag(d) L^T  (blocked right-looking, fp64, no pivoting);
//   dx = A^-1 (-g);  q1 = 0.5 * dx . (u*D*dx - g)
// replaces `D.diagonal() = Hess.diagonal(); dxi = (Hess + u*D).ldlt().solve(-JacT)` and the q1 line of
// BALM2::damping_iter (bavoxel.hpp:1113-1114,1127).  Eigen's LDLT pivots on the diagonal; for the damped
// normal matrix the factorisation exists without pivoting whenever the leading minors are non-singular
// (always when H + uD is positive definite, the accepted-step regime).  A zero / non-finite pivot raises the
// not_pd flag and the LM loop treats the step as rejected (u *= v), which is what happens to a bad step in
// the reference as well (the trial residual does not decrease).  The solution of a non-singular system does
// not depend on the pivot order beyond rounding.
//
// A is column-major n x n; only the lower triangle is read/written. Panel width NB = 64.
#include "internal.cuh"

namespace {

constexpr int NB = BALM_NB;

__global__ void damp_copy_kernel(const double *H, double *A, double *dvec, int n, double u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // row (contiguous)
  const int j = blockIdx.y;
  if (i >= n) return;
  double h = H[(size_t)j * n + i];
  if (i == j) {
    dvec[i] = h;
    h += u * h;
  }
  A[(size_t)j * n + i] = h;
}

// Factor the nbw x nbw diagonal block at j0 in shared memory: unit-lower L11 (strict lower part) and d.
__global__ void __launch_bounds__(256) ldl_diag_kernel(double *A, int n, int j0, int nbw, int *flags) {
  __shared__ double S[NB][NB + 1];
  __shared__ double colj[NB];
  const int tid = threadIdx.x;
  for (int e = tid; e < nbw * nbw; e += 256) {
    const int r = e % nbw, c = e / nbw;
    S[r][c] = (r >= c) ? A[(size_t)(j0 + c) * n + j0 + r] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nbw; j++) {
    const double d = S[j][j];
    if (tid == 0 && !(fabs(d) > 1e-290 && fabs(d) < 1e300)) atomicOr(&flags[0], 1);
    for (int i = j + 1 + tid; i < nbw; i += 256) colj[i] = S[i][j];
    __syncthreads();
    const double dinv = 1.0 / d;
    const int m = nbw - j - 1;
    for (int e = tid; e < m * m; e += 256) {
      const int r = j + 1 + e % m, c = j + 1 + e / m;
      if (r >= c) S[r][c] -= colj[r] * colj[c] * dinv;
    }
    for (int i = j + 1 + tid; i < nbw; i += 256) S[i][j] = colj[i] * dinv;
    __syncthreads();
  }
  for (int e = tid; e < nbw * nbw; e += 256) {
    const int r = e % nbw, c = e / nbw;
    if (r >= c) A[(size_t)(j0 + c) * n + j0 + r] = S[r][c];
  }
}

// Rows below the diagonal block: W = A21 * L11^-T (unit), L21 = W * d^-1. One thread per row.
__global__ void __launch_bounds__(NB) ldl_panel_kernel(double *A, double *W, int n, int j0, int nbw) {
  extern __shared__ double dyn_smem[];
  double (*L11)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);
  double (*rowbuf)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));
  double *dinv = dyn_smem + 2 * NB * (NB + 1);
  const int tid = threadIdx.x;
  const int i0 = j0 + nbw + blockIdx.x * NB;
  for (int e = tid; e < nbw * nbw; e += NB) {
    const int r = e % nbw, c = e / nbw;
    L11[r][c] = A[(size_t)(j0 + c) * n + j0 + r];
  }
  __syncthreads();
  if (tid < nbw) dinv[tid] = 1.0 / L11[tid][tid];
  const int i = i0 + tid;
  const bool ok = i < n;
  if (ok)
    for (int c = 0; c < nbw; c++) rowbuf[tid][c] = A[(size_t)(j0 + c) * n + i];
  __syncthreads();
  if (ok) {
    for (int c = 0; c < nbw; c++) {
      double w = rowbuf[tid][c];
      for (int t = 0; t < c; t++) w -= rowbuf[tid][t] * L11[c][t];
      rowbuf[tid][c] = w;
    }
    for (int c = 0; c < nbw; c++) {
      const double w = rowbuf[tid][c];
      W[(size_t)c * n + i] = w;
      A[(size_t)(j0 + c) * n + i] = w * dinv[c];
    }
  }
}

// Trailing update A22 -= L21 * W^T on 64x64 tiles of the lower triangle. 256 threads, 4x4 per thread.
__global__ void __launch_bounds__(256) ldl_update_kernel(double *A, const double *W, int n, int j0, int nbw) {
  extern __shared__ double dyn_smem[];
  double (*sL)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);                  // [c][row]
  double (*sW)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));  // [c][col]
  const int base = j0 + nbw;
  // tile enumeration over the lower triangle of the trailing matrix
  int t = blockIdx.x, tr = 0;
  while (t >= tr + 1) { t -= tr + 1; tr++; }
  const int tc = t;  // tc <= tr
  const int r0 = base + tr * NB, c0 = base + tc * NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < nbw * NB; e += 256) {
    const int rr = e % NB, c = e / NB;
    sL[c][rr] = (r0 + rr < n) ? A[(size_t)(j0 + c) * n + r0 + rr] : 0.0;
    sW[c][rr] = (c0 + rr < n) ? W[(size_t)c * n + c0 + rr] : 0.0;
  }
  __syncthreads();
  const int tx = tid % 16, ty = tid / 16;  // rows tx*4.., cols ty*4..
  double acc[4][4] = {};
  for (int c = 0; c < nbw; c++) {
    double lv[4], wv[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { lv[q] = sL[c][tx * 4 + q]; wv[q] = sW[c][ty * 4 + q]; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] += lv[a] * wv[b];
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int col = c0 + ty * 4 + b;
    if (col >= n) continue;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int row = r0 + tx * 4 + a;
      if (row < n && row >= col) A[(size_t)col * n + row] -= acc[a][b];
    }
  }
}

// Single-CTA triangular solves: x = L^-T diag(d)^-1 L^-1 (-g); also q1. 1024 threads.
__global__ void __launch_bounds__(1024) ldl_solve_kernel(const double *A, const double *g, const double *dvec,
                                                        double *x, int n, double u, double *scal) {
  extern __shared__ double sx[];  // n doubles
  __shared__ double sblk[NB];
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += 1024) sx[i] = -g[i];
  __syncthreads();
  // forward: L y = b (unit lower), column-oriented by blocks
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nbw = (n - j0 < NB) ? n - j0 : NB;
    if (warp == 0) {
      for (int c = 0; c < nbw; c++) {
        const double yc = sx[j0 + c];
        for (int r = c + 1 + lane; r < nbw; r += 32) sx[j0 + r] -= A[(size_t)(j0 + c) * n + j0 + r] * yc;
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < nbw) sblk[tid] = sx[j0 + tid];
    __syncthreads();
    for (int i = j0 + nbw + tid; i < n; i += 1024) {
      double s = sx[i];
      for (int c = 0; c < nbw; c++) s -= A[(size_t)(j0 + c) * n + i] * sblk[c];
      sx[i] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += 1024) sx[i] /= A[(size_t)i * n + i];
  __syncthreads();
  // backward: L^T x = z
  const int nblk = (n + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; b--) {
    const int j0 = b * NB;
    const int nbw = (n - j0 < NB) ? n - j0 : NB;
    // s_c = sum_{i >= j0+nbw} L[i][j0+c] x_i : one warp per column (32 warps, two rounds for 64 columns)
    for (int c = warp; c < nbw; c += 32) {
      double s = 0.0;
      const double *col = A + (size_t)(j0 + c) * n;
      for (int i = j0 + nbw + lane; i < n; i += 32) s += col[i] * sx[i];
      s = warp_sum(s);
      if (lane == 0) sblk[c] = s;
    }
    __syncthreads();
    if (warp == 0) {
      for (int c = lane; c < nbw; c += 32) sx[j0 + c] -= sblk[c];
      __syncwarp();
      for (int c = nbw - 1; c >= 0; c--) {
        // x_c final; eliminate it from rows above inside the block: x_r -= L[c][r] * x_c for r < c
        const double xc = sx[j0 + c];
        for (int r = lane; r < c; r += 32) sx[j0 + r] -= A[(size_t)(j0 + r) * n + j0 + c] * xc;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  double part = 0.0;
  for (int i = tid; i < n; i += 1024) {
    const double xi = sx[i];
    x[i] = xi;
    part += xi * (u * dvec[i] * xi - g[i]);
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (warp == 0) {
    double s = red[lane];
    s = warp_sum(s);
    if (lane == 0) scal[1] = 0.5 * s;  // q1
  }
}

// R <- Exp(phi) R, p <- Exp(phi) p + dt  (bavoxel.hpp:1123-1125)
__global__ void pose_update_kernel(const double *pin, const double *dx, double *pout, int N) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double dR[9], r[9], p[3];
  exp_so3(dx + 6 * j, dR);
  load_pose(pin + 12 * j, r, p);
  double rn[9], pn[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int b = 0; b < 3; b++) rn[a * 3 + b] = dR[a * 3] * r[b] + dR[a * 3 + 1] * r[3 + b] + dR[a * 3 + 2] * r[6 + b];
    pn[a] = dR[a * 3] * p[0] + dR[a * 3 + 1] * p[1] + dR[a * 3 + 2] * p[2] + dx[6 * j + 3 + a];
  }
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int a = 0; a < 3; a++) pout[12 * j + c * 3 + a] = rn[a * 3 + c];
  pout[12 * j + 9] = pn[0]; pout[12 * j + 10] = pn[1]; pout[12 * j + 11] = pn[2];
}

// x_i <- T_0^-1 x_i (bavoxel.hpp:1159-1164); mode 1 also forces pose 0 to the exact identity
// (benchmark_virtual.cpp:472-479). pose 0 is read from a snapshot so the in-place update is race-free.
__global__ void gauge_kernel(double *poses, const double *pose0_snapshot, int N, int mode) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double r0[9], p0[3], r[9], p[3];
  load_pose(pose0_snapshot, r0, p0);
  load_pose(poses + 12 * j, r, p);
  double rn[9], pn[3];
  const double d[3] = {p[0] - p0[0], p[1] - p0[1], p[2] - p0[2]};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    pn[a] = r0[a] * d[0] + r0[3 + a] * d[1] + r0[6 + a] * d[2];
#pragma unroll
    for (int b = 0; b < 3; b++) rn[a * 3 + b] = r0[a] * r[b] + r0[3 + a] * r[3 + b] + r0[6 + a] * r[6 + b];
  }
  if (mode == 1 && j == 0) {
    rn[0] = 1; rn[1] = 0; rn[2] = 0; rn[3] = 0; rn[4] = 1; rn[5] = 0; rn[6] = 0; rn[7] = 0; rn[8] = 1;
    pn[0] = pn[1] = pn[2] = 0;
  }
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int a = 0; a < 3; a++) poses[12 * j + c * 3 + a] = rn[a * 3 + c];
  poses[12 * j + 9] = pn[0]; poses[12 * j + 10] = pn[1]; poses[12 * j + 11] = pn[2];
}

}  // namespace

int launch_ldlt_solve(balm_ctx *c, double u) {
  const int n = c->n;
  CUDA_TRY(cudaMemsetAsync(c->flags, 0, sizeof(int) * 4, c->stream));
  dim3 g1((n + 255) / 256, n);
  damp_copy_kernel<<<g1, 256, 0, c->stream>>>(c->H, c->A, c->dvec, n, u);
  c->launches += 1;
  const int panel_smem = (2 * NB * (NB + 1) + NB) * (int)sizeof(double);
  const int update_smem = 2 * NB * (NB + 1) * (int)sizeof(double);
  static bool attr_set2 = false;
  if (!attr_set2) {
    CUDA_TRY(cudaFuncSetAttribute(ldl_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem));
    CUDA_TRY(cudaFuncSetAttribute(ldl_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, update_smem));
    attr_set2 = true;
  }
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nbw = (n - j0 < NB) ? n - j0 : NB;
    ldl_diag_kernel<<<1, 256, 0, c->stream>>>(c->A, n, j0, nbw, c->flags);
    c->launches += 1;
    const int m = n - j0 - nbw;
    if (m > 0) {
      const int mt = (m + NB - 1) / NB;
      ldl_panel_kernel<<<mt, NB, panel_smem, c->stream>>>(c->A, c->W, n, j0, nbw);
      ldl_update_kernel<<<mt * (mt + 1) / 2, 256, update_smem, c->stream>>>(c->A, c->W, n, j0, nbw);
      c->launches += 2;
    }
  }
  static bool attr_set = false;
  const int smem = n * (int)sizeof(double);
  if (!attr_set && smem > 48 * 1024) {
    CUDA_TRY(cudaFuncSetAttribute(ldl_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (smem > 200 * 1024) {
    balm_set_error("n too large for the single-CTA triangular solve (n <= 25600)");
    return BALM_ERR_UNSUPPORTED;
  }
  ldl_solve_kernel<<<1, 1024, smem, c->stream>>>(c->A, c->g, c->dvec, c->dx, n, u, c->scal);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}

int launch_pose_update(balm_ctx *c, const double *poses_in, const double *dx, double *poses_out) {
  pose_update_kernel<<<(c->N + 127) / 128, 128, 0, c->stream>>>(poses_in, dx, poses_out, c->N);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}

int launch_gauge(balm_ctx *c, double *poses, int mode) {
  // snapshot pose 0 into scal[4..15]
  CUDA_TRY(cudaMemcpyAsync(c->scal + 4, poses, 12 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  gauge_kernel<<<(c->N + 127) / 128, 128, 0, c->stream>>>(poses, c->scal + 4, c->N, mode);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}
