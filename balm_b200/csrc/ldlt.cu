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
#include <stdlib.h>
#include "internal.cuh"
#define TRY_LAUNCH(x) do { int _s = (x); if (_s != BALM_OK) return _s; } while (0)

namespace {

constexpr int NB = BALM_NB;

__global__ void damp_copy_kernel(const double *H, double *A, double *dvec, int n, const double *u_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // row (contiguous)
  const int j = blockIdx.y;
  if (i >= n) return;
  const double u = *u_dev;
  double h = H[(size_t)j * n + i];
  if (i == j) {
    dvec[i] = h;
    h += u * h;
  }
  A[(size_t)j * n + i] = h;
}

// ---- panel step j0: three kernels ------------------------------------------------------------------------
// (1) ldl_diag_kernel: one CTA factors the 64x64 diagonal block in REGISTERS (16x16 threads, 4x4 each), one
//     __syncthreads per column (double-buffered column broadcast). The same sweep builds X = L11^-1 (forward
//     substitution on the identity), so the panel below needs a GEMM instead of a sequential TRSM, and applies
//     it to the right-hand side: y_j = L11^-1 b_j (forward substitution fused into the factorisation).
// (2) ldl_panel_kernel: W = A21 X^T (64x64x64 GEMM per 64-row tile), L21 = W d^-1, b_rows -= L21 y_j.
// (3) ldl_update_kernel: A22 -= L21 W^T on 128x128 tiles of the lower triangle.
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// The 64x64 diagonal block is factored in four 16-column sub-panels. A sub-panel is factored by TWO warps (one
// thread per row, its 16 panel entries in registers) with a 64-thread named barrier per column -- the serial chain
// of the whole solve is n = 6N of these column steps, so each step is kept to: post column -> barrier -> reciprocal
// -> <=15 FMAs. The rank-16 trailing update and the assembly of X = L11^-1 from 16x16 inverses use all 256 threads.
// In: S = the block (lower triangle), Xs = 0, all threads synchronised. Out: S = L (below the diagonal) and d (on
// it), Xs = L11^-1 (lower), all threads synchronised. Returns true if a pivot was zero / non-finite.
__device__ __forceinline__ bool diag_factor(double (*S)[NB + 1], double (*Xs)[NB + 1], double (*Wp)[17],
                                            double (*col16)[16]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  bool bad = false;
#pragma unroll 1
  for (int cb = 0; cb < NB; cb += 16) {
    if (tid < NB) {  // ---- sub-panel factorisation: rows r >= cb, columns cb .. cb+15 ----
      const int r = tid;
      double a[16];
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = S[r][cb + k];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int jc = cb + j;
        if (r >= cb && r < cb + 16) col16[j & 1][r - cb] = a[j];  // unscaled column j at the panel's own rows
        bar_sync_named(1, NB);
        const double d = col16[j & 1][j];
        if (!(fabs(d) > 1e-290 && fabs(d) < 1e300)) bad = true;
        const double dinv = __drcp_rn(d);
        if (r > jc) {
          const double l = a[j] * dinv;
#pragma unroll
          for (int k = j + 1; k < 16; k++) a[k] -= l * col16[j & 1][k];
          a[j] = l;
        }
      }
      if (r >= cb) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          if (r >= cb + k) S[r][cb + k] = a[k];        // L (r > cb+k) and d (r == cb+k)
        }
      }
    }
    __syncthreads();
    if (tid < NB) {  // W = L_panel * d for the trailing update
      const int r = tid;
#pragma unroll
      for (int k = 0; k < 16; k++) Wp[r][k] = (r > cb + k) ? S[r][cb + k] * S[cb + k][cb + k] : 0.0;
    }
    __syncthreads();
    {  // trailing block in 16x16 sub-blocks: thread (ri, ci) owns one element of each lower sub-block
      const int ri = tid >> 4, ci = tid & 15;
      for (int rb2 = cb + 16; rb2 < NB; rb2 += 16)
        for (int cb2 = cb + 16; cb2 <= rb2; cb2 += 16) {
          const int r = rb2 + ri, c = cb2 + ci;
          if (r >= c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 16; k++) s += Wp[r][k] * S[c][cb + k];
            S[r][c] -= s;
          }
        }
    }
    __syncthreads();
  }

  // ---- X = L11^-1: 16x16 diagonal-block inverses (one warp each, shuffles only) ----
  if (warp < 4) {
    const int cb = 16 * warp, i = lane & 15;
    double xr[16];
#pragma unroll
    for (int k = 0; k < 16; k++) xr[k] = (k == i) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < 15; j++) {
      const double l = (i > j) ? S[cb + i][cb + j] : 0.0;
#pragma unroll
      for (int k = 0; k <= j; k++) {  // row j of the inverse is final and has entries k <= j only
        const double xj = __shfl_sync(0xffffffffu, xr[k], j, 16);
        xr[k] -= l * xj;
      }
    }
    if (lane < 16) {
#pragma unroll
      for (int k = 0; k < 16; k++) Xs[cb + i][cb + k] = xr[k];
    }
  }
  __syncthreads();
  // ---- off-diagonal blocks by block rows: X_bc = -X_bb * sum_{m=c}^{b-1} L_bm X_mc ----
#pragma unroll 1
  for (int b = 1; b < 4; b++) {
    const int rb = 16 * b, ri = tid >> 4, ci = tid & 15;
    for (int cq = 0; cq < b; cq++) {  // T_bc = sum_{m in blocks c..b-1} L_bm X_mc, parked in the X_bc positions
      const int c = 16 * cq + ci;
      double sacc = 0.0;
      for (int m = 16 * cq; m < rb; m++) sacc += S[rb + ri][m] * Xs[m][c];
      Xs[rb + ri][c] = sacc;
    }
    __syncthreads();
    double xn[3];
#pragma unroll
    for (int cq = 0; cq < 3; cq++) {
      xn[cq] = 0.0;
      if (cq < b) {
        const int c = 16 * cq + ci;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k++) sacc += Xs[rb + ri][rb + k] * Xs[rb + k][c];  // X_bb (lower) times T
        xn[cq] = -sacc;
      }
    }
    __syncthreads();  // every thread has read the T values it needs
#pragma unroll
    for (int cq = 0; cq < 3; cq++)
      if (cq < b) Xs[rb + ri][16 * cq + ci] = xn[cq];
    __syncthreads();
  }
  return bad;
}

// Diagonal step: factor, write L11/d back, X, 1/d, y_j = X b_j (y goes to `ysol`; `sol` keeps the running rhs).
// (Factoring the block redundantly inside every panel CTA, to drop this kernel boundary from the critical path, was
// measured and is NOT faster: 2.81 vs 2.72 ms at n = 3000.)
__global__ void __launch_bounds__(256) ldl_diag_kernel(double *A, int n, int j0, int nbw, double *Xcm, double *dinv_all,
                                                       const double *sol, double *ysol, int *flags) {
  extern __shared__ double dyn_smem[];
  double (*S)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);                   // the block (lower) -> L, d
  double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));  // L11^-1
  double (*Wp)[17] = reinterpret_cast<double (*)[17]>(dyn_smem + 2 * NB * (NB + 1));     // current sub-panel times d
  double (*col16)[16] = reinterpret_cast<double (*)[16]>(dyn_smem + 2 * NB * (NB + 1) + NB * 17);
  double *bvec = dyn_smem + 2 * NB * (NB + 1) + NB * 17 + 32;
  const int tid = threadIdx.x;
  {
    double ld[16];  // all 16 global loads of this thread in flight together (the kernel is pure latency)
#pragma unroll
    for (int it = 0; it < 16; it++) {
      const int e = tid + it * 256, r = e & (NB - 1), c = e >> 6;
      double v = (r == c) ? 1.0 : 0.0;  // rows/cols of a short last panel act as identity
      if (r < nbw && c < nbw && r >= c) v = A[(size_t)(j0 + c) * n + j0 + r];
      ld[it] = v;
    }
    const double bv = (tid < nbw) ? sol[j0 + tid] : 0.0;
#pragma unroll
    for (int it = 0; it < 16; it++) {
      const int e = tid + it * 256, r = e & (NB - 1), c = e >> 6;
      S[r][c] = ld[it];
      Xs[r][c] = 0.0;
    }
    if (tid < NB) bvec[tid] = bv;
  }
  __syncthreads();
  const bool bad = diag_factor(S, Xs, Wp, col16);
  if (bad && tid == 0) atomicOr(&flags[0], 1);
  // ---- write back L11 / d, X (column-major), 1/d, and y_j = X b_j ----
#pragma unroll 4
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 6;
    if (r < nbw && c < nbw && r >= c) A[(size_t)(j0 + c) * n + j0 + r] = S[r][c];
    Xcm[c * NB + r] = (r >= c) ? Xs[r][c] : 0.0;
  }
  if (tid < nbw) dinv_all[j0 + tid] = 1.0 / S[tid][tid];
  if (tid < NB) {
    double sacc = 0.0;
    for (int c = 0; c <= tid; c++) sacc += Xs[tid][c] * bvec[c];
    if (tid < nbw) ysol[j0 + tid] = sacc;
  }
}

__global__ void __launch_bounds__(256) ldl_panel_kernel(double *A, double *W, int n, int j0, int nbw, const double *Xcm,
                                                        const double *dinv_all, double *sol, const double *ysol) {
  extern __shared__ double dyn_smem[];
  double (*sA)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem);                  // [k][row]
  double (*sX)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem + NB * (NB + 4));  // [k][col] = X[col][k]
  __shared__ double yj[NB], dv[NB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = j0 + nbw + blockIdx.x * NB;
  {
    // all 32 loads of this thread are issued before the first shared-memory store (latency paid once)
    double ra[16], rx[16];
#pragma unroll
    for (int it = 0; it < 16; it++) {
      const int e = tid + it * 256, r = e & (NB - 1), k = e >> 6;
      ra[it] = (k < nbw && i0 + r < n) ? A[(size_t)(j0 + k) * n + i0 + r] : 0.0;
      rx[it] = Xcm[k * NB + r];  // X[r][k]
    }
#pragma unroll
    for (int it = 0; it < 16; it++) {
      const int e = tid + it * 256, r = e & (NB - 1), k = e >> 6;
      sA[k][r] = ra[it];
      sX[k][r] = rx[it];
    }
  }
  if (tid < NB) {
    yj[tid] = tid < nbw ? ysol[j0 + tid] : 0.0;
    dv[tid] = tid < nbw ? dinv_all[j0 + tid] : 0.0;
  }
  __syncthreads();
  double acc[4][4] = {};
#pragma unroll 8
  for (int k = 0; k < NB; k++) {
    double av[4], xv[4];
#pragma unroll
    for (int p = 0; p < 4; p++) { av[p] = sA[k][4 * tx + p]; xv[p] = sX[k][4 * ty + p]; }
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[p][q] += av[p] * xv[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int c = 4 * ty + q;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int r = 4 * tx + p, i = i0 + r;
      const double l = acc[p][q] * dv[c];
      sA[c][r] = l;  // reuse as L21 tile [c][r] for the right-hand-side update
      if (c < nbw && i < n) {
        W[(size_t)c * n + i] = acc[p][q];
        A[(size_t)(j0 + c) * n + i] = l;
      }
    }
  }
  __syncthreads();
  if (tid < NB && i0 + tid < n) {
    double s = 0.0;
#pragma unroll 8
    for (int c = 0; c < NB; c++) s += sA[c][tid] * yj[c];
    sol[i0 + tid] -= s;
  }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

constexpr int UT = 128;   // update tile
constexpr int UKB = 64;   // k rows per stage: the whole panel width in one shot (one global round trip)
constexpr int ULDS = UT + 4;  // padded smem row (doubles): conflict-free DMMA fragment loads (see syrk_f64.cu)
constexpr int USTAGE = 2 * UKB * ULDS;  // 135 KB: one CTA per SM

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Trailing update A22 -= L21 * W^T on 128x128 tiles of the lower triangle with fp64 tensor-core MMAs
// (mma.sync.m8n8k4.f64, 8 warps x (64x32) warp tiles, the whole 64-wide k slab staged by one cp.async round).
// The update of a panel step is issued in two parts (look-ahead): the STRIP (strip = 1: the NB columns of the next
// panel, one tile per 128 rows, stores masked to col < base + NB) stays on the critical path; the REST (strip = 0,
// base moved past the strip) runs on the side stream while the next diagonal block and panel are factored.
__global__ void __launch_bounds__(256) ldl_update_kernel(double *A, const double *W, int n, int j0, int nbw, int base,
                                                         int strip) {
  extern __shared__ __align__(16) double dyn_smem[];
  int t = blockIdx.x, tr = 0;
  if (strip) { tr = t; t = 0; }
  else while (t >= tr + 1) { t -= tr + 1; tr++; }
  const int tc = t;  // tc <= tr
  const int r0 = base + tr * UT, c0 = base + tc * UT;
  const int climit = strip ? base + NB : n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3, fr = lane >> 2, fk = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nsteps = (nbw + UKB - 1) / UKB;
  auto load_stage = [&](int stage, int step) {
    double *sL = dyn_smem + stage * USTAGE;
    double *sW = sL + UKB * ULDS;
    const int k0 = step * UKB;
    // 16-byte pieces; n, r0, c0 are even so a pair is entirely inside or outside the matrix
#pragma unroll 4
    for (int e = tid; e < UKB * (UT / 2); e += 256) {
      const int k = e >> 6, rr = (e & 63) * 2;
      const bool kv = k0 + k < nbw;
      if (kv && r0 + rr < n) cp_async16(sL + k * ULDS + rr, A + (size_t)(j0 + k0 + k) * n + r0 + rr);
      else *reinterpret_cast<double2 *>(sL + k * ULDS + rr) = make_double2(0.0, 0.0);
      if (kv && c0 + rr < climit) cp_async16(sW + k * ULDS + rr, W + (size_t)(k0 + k) * n + c0 + rr);
      else *reinterpret_cast<double2 *>(sW + k * ULDS + rr) = make_double2(0.0, 0.0);
    }
  };
  load_stage(0, 0);
  asm volatile("cp.async.commit_group;\n" ::);
  for (int step = 0; step < nsteps; step++) {
    if (step + 1 < nsteps) {
      load_stage((step + 1) & 1, step + 1);
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const double *sL = dyn_smem + (step & 1) * USTAGE;
    const double *sW = sL + UKB * ULDS;
#pragma unroll
    for (int kk = 0; kk < UKB; kk += 4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; i++) af[i] = sL[(kk + fk) * ULDS + wm * 64 + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; j++) bf[j] = sW[(kk + fk) * ULDS + wn * 32 + j * 8 + fr];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
  // read-modify-write of the C tile: all loads of a row group are issued before the first store, otherwise every
  // "-=" serialises a full L2 round trip (64 of them per thread)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int row = r0 + wm * 64 + i * 8 + fr;
    double cv[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = c0 + wn * 32 + j * 8 + 2 * fk;
      cv[j][0] = (row < n && col < climit && row >= col) ? A[(size_t)col * n + row] : 0.0;
      cv[j][1] = (row < n && col + 1 < climit && row >= col + 1) ? A[(size_t)(col + 1) * n + row] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = c0 + wn * 32 + j * 8 + 2 * fk;
      if (row < n && col < climit && row >= col) A[(size_t)col * n + row] = cv[j][0] - acc[i][j][0];
      if (row < n && col + 1 < climit && row >= col + 1) A[(size_t)(col + 1) * n + row] = cv[j][1] - acc[i][j][1];
    }
  }
}

// rhs = -g (start of the forward substitution, fused into the factorisation kernels)
__global__ void rhs_init_kernel(const double *g, double *sol, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sol[i] = -g[i];
}

// Backward substitution L^T x = w (w = d^-1 y, scaled beforehand), block groups descending. One launch handles up to
// four 64-row blocks [g_lo, g_lo+G): every CTA first solves the group's own triangle redundantly (x_b = X_b^T w_b,
// then w_c -= L[b rows][c] x_b for the group's lower columns; a warp per column, rows contiguous -> coalesced), which
// replaces three kernel boundaries of the serial chain by block barriers, and then applies the group's G rows to its
// share of the columns left of the group (w_c -= sum_r L[g_lo+r][c] x[r], up to 2 KB contiguous per column).
constexpr int BACK_GROUP = 4;
__global__ void __launch_bounds__(256) ldl_back_kernel(const double *A, int n, int g_lo, int G, const double *Xcm,
                                                       double *sol, double *x) {
  __shared__ double w[BACK_GROUP * NB], xs[BACK_GROUP * NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  w[tid] = tid < G ? sol[g_lo + tid] : 0.0;
  xs[tid] = 0.0;
  __syncthreads();
  const int nblk = (G + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; b--) {
    const int j = NB * b;
    const int bw = G - j < NB ? G - j : NB;
    const double *Xb = Xcm + (size_t)b * NB * NB;  // column-major X = L_bb^-1 (identity-padded when bw < NB)
#pragma unroll
    for (int q = 0; q < 8; q++) {  // x_b[c] = sum_r X[r][c] w_b[r]
      const int c = warp * 8 + q;
      double sacc = Xb[c * NB + lane] * w[j + lane] + Xb[c * NB + lane + 32] * w[j + lane + 32];
      sacc = warp_sum(sacc);
      if (lane == 0) xs[j + c] = c < bw ? sacc : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int cl = warp; cl < j; cl += 8) {  // the group's columns below block b
      const double *col = A + (size_t)(g_lo + cl) * n + g_lo + j;
      double sacc = 0.0;
      if (lane < bw) sacc = col[lane] * xs[j + lane];
      if (lane + 32 < bw) sacc += col[lane + 32] * xs[j + lane + 32];
      sacc = warp_sum(sacc);
      if (lane == 0) w[cl] -= sacc;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid < G) x[g_lo + tid] = xs[tid];
#pragma unroll
  for (int q = 0; q < 4; q++) {  // columns left of the group: 8 warps x 4 columns per CTA
    const int c = (blockIdx.x * 8 + warp) * 4 + q;
    if (c >= g_lo) break;
    const double *col = A + (size_t)c * n + g_lo;
    double sacc = 0.0;
#pragma unroll
    for (int r = lane; r < BACK_GROUP * NB; r += 32)
      if (r < G) sacc += col[r] * xs[r];
    sacc = warp_sum(sacc);
    if (lane == 0) sol[c] -= sacc;
  }
}

__global__ void scale_rhs_kernel(double *sol, const double *ysol, const double *dinv_all, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sol[i] = ysol[i] * dinv_all[i];  // w = d^-1 y: start of the backward substitution
}

__global__ void __launch_bounds__(1024) q1_kernel(const double *x, const double *g, const double *dvec, int n,
                                                  const double *u_dev, double *scal) {
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double u = *u_dev;
  double part = 0.0;
  for (int i = tid; i < n; i += 1024) {
    const double xi = x[i];
    part += xi * (u * dvec[i] * xi - g[i]);
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (warp == 0) {
    double s = red[lane];
    s = warp_sum(s);
    if (lane == 0) scal[1] = 0.5 * s;  // q1 = 0.5 dx.(u D dx - g)  (bavoxel.hpp:1127)
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

constexpr int DIAG_SMEM = (2 * NB * (NB + 1) + NB * 17 + 32 + NB) * (int)sizeof(double);

// The factorisation loop with look-ahead. Main stream (high priority): diag(k) -> panel(k) -> strip(k) -> diag(k+1);
// side stream: rest(k) after panel(k). Hazards: strip(k) and rest(k-1) both update the columns of panel k+1, and
// panel(k+1) overwrites the W buffer rest(k-1) reads (W is double-buffered by step parity) -> strip(k) waits for
// rest(k-1). rest(k) touches only columns right of panel k+1, which diag/panel(k+1) never access.
static int enqueue_solve(balm_ctx *c) {
  const int n = c->n;
  cudaStream_t st = c->stream, side = c->stream2;
  dim3 g1((n + 255) / 256, n);
  damp_copy_kernel<<<g1, 256, 0, st>>>(c->H, c->A, c->dvec, n, c->scal + 3);
  rhs_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(c->g, c->sol, n);
  const int panel_smem = 2 * NB * (NB + 4) * (int)sizeof(double);
  const int update_smem = USTAGE * (int)sizeof(double);
  int launches = 2;
  int pending_rest = -1;  // parity of the side-stream event still to be waited for
  for (int j0 = 0, pi = 0; j0 < n; j0 += NB, pi++) {
    const int nbw = (n - j0 < NB) ? n - j0 : NB;
    double *X = c->Xinv + (size_t)pi * NB * NB;
    double *W = c->W + (size_t)(pi & 1) * n * NB;
    const int m = n - j0 - nbw;
    const int mt = (m + NB - 1) / NB, mu = (m + UT - 1) / UT;
    ldl_diag_kernel<<<1, 256, DIAG_SMEM, st>>>(c->A, n, j0, nbw, X, c->dinv, c->sol, c->ysol, c->flags);
    launches++;
    if (m > 0) {
      ldl_panel_kernel<<<mt, 256, panel_smem, st>>>(c->A, W, n, j0, nbw, X, c->dinv, c->sol, c->ysol);
      launches++;
      const int base = j0 + nbw;
      if (!c->solve_lookahead) {
        ldl_update_kernel<<<mu * (mu + 1) / 2, 256, update_smem, st>>>(c->A, W, n, j0, nbw, base, 0);
        launches++;
        continue;
      }
      CUDA_TRY(cudaEventRecord(c->sev[0], st));
      if (pending_rest >= 0) CUDA_TRY(cudaStreamWaitEvent(st, c->sev[1 + pending_rest], 0));
      pending_rest = -1;
      ldl_update_kernel<<<mu, 256, update_smem, st>>>(c->A, W, n, j0, nbw, base, 1);
      launches++;
      const int m2 = m - NB;
      if (m2 > 0) {
        const int mu2 = (m2 + UT - 1) / UT;
        CUDA_TRY(cudaStreamWaitEvent(side, c->sev[0], 0));
        ldl_update_kernel<<<mu2 * (mu2 + 1) / 2, 256, update_smem, side>>>(c->A, W, n, j0, nbw, base + NB, 0);
        launches++;
        CUDA_TRY(cudaEventRecord(c->sev[1 + (pi & 1)], side));
        pending_rest = pi & 1;
      }
    }
  }
  if (pending_rest >= 0) CUDA_TRY(cudaStreamWaitEvent(st, c->sev[1 + pending_rest], 0));
  scale_rhs_kernel<<<(n + 255) / 256, 256, 0, st>>>(c->sol, c->ysol, c->dinv, n);
  launches++;
  const int npan = (n + NB - 1) / NB;
  for (int hi = npan - 1; hi >= 0; hi -= BACK_GROUP) {
    const int lo = hi - BACK_GROUP + 1 > 0 ? hi - BACK_GROUP + 1 : 0;
    const int g_lo = lo * NB;
    const int g_hi = (hi + 1) * NB < n ? (hi + 1) * NB : n;
    const int blocks = g_lo > 0 ? (g_lo + 31) / 32 : 1;
    ldl_back_kernel<<<blocks, 256, 0, st>>>(c->A, n, g_lo, g_hi - g_lo, c->Xinv + (size_t)lo * NB * NB, c->sol, c->dx);
    launches++;
  }
  q1_kernel<<<1, 1024, 0, st>>>(c->dx, c->g, c->dvec, n, c->scal + 3, c->scal);
  launches++;
  c->solve_launches = launches;
  return BALM_OK;
}

int launch_ldlt_solve(balm_ctx *c, double u) {
  const int n = c->n;
  if (!c->Xinv) {
    const int npan = (n + NB - 1) / NB;
    CUDA_TRY(cudaMalloc((void **)&c->Xinv, sizeof(double) * (size_t)npan * NB * NB));
    CUDA_TRY(cudaMalloc((void **)&c->dinv, sizeof(double) * n));
    CUDA_TRY(cudaMalloc((void **)&c->sol, sizeof(double) * n));
    CUDA_TRY(cudaMalloc((void **)&c->ysol, sizeof(double) * n));
    c->solve_lookahead = getenv("BALM_NO_LOOKAHEAD") == nullptr;
  }
  c->h_scal[3] = u;  // pinned; the kernels read the damping factor from device memory so the graph is reusable
  CUDA_TRY(cudaMemcpyAsync(c->scal + 3, c->h_scal + 3, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(c->flags, 0, sizeof(int) * 4, c->stream));
  static const bool use_graph = getenv("BALM_NO_GRAPH") == nullptr;
  if (!use_graph) {
    TRY_LAUNCH(enqueue_solve(c));
    c->launches += c->solve_launches;
    CUDA_TRY(cudaGetLastError());
    return BALM_OK;
  }
  if (!c->solve_graph) {  // capture the ~190-launch sequence once, replay it with one launch afterwards
    cudaGraph_t graph;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_solve(c);
    CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
    if (rc != BALM_OK) return rc;
    CUDA_TRY(cudaGraphInstantiate((cudaGraphExec_t *)&c->solve_graph, graph, 0));
    CUDA_TRY(cudaGraphDestroy(graph));
  }
  CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)c->solve_graph, c->stream));
  c->launches += c->solve_launches;
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

int ldlt_setup() {
  CUDA_TRY(cudaFuncSetAttribute(ldl_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                2 * NB * (NB + 4) * (int)sizeof(double)));
  CUDA_TRY(cudaFuncSetAttribute(ldl_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                USTAGE * (int)sizeof(double)));
  CUDA_TRY(cudaFuncSetAttribute(ldl_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  return BALM_OK;
}
