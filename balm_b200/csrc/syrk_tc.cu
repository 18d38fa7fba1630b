// syrk_tc.cu -- tcgen05 (5th-gen tensor core) path of the rank-3M symmetric update S = G'^T G' (kernel K4).
//
// tcgen05 has no fp64 MMA kind, so the fp64 product is computed EXACTLY in integers (Ozaki-style split):
//   per column j of G' (= one pose DOF) a power-of-two scale sc_j = 2^p_j with |G'[c][j]| * sc_j < 2^(8S-2);
//   X[c][j] = rint(G'[c][j] * sc_j) is split into S balanced base-256 digits q_s in [-128,127] (int8 planes);
//   T_d = sum_{s+t=d} sum_c q_s[c][i] * q_t[c][j]  accumulates in int32 TMEM accumulators with NO rounding
//   (kind::i8, |product| <= 2^14, at most S*rows < 2^17 terms per accumulator -> |T_d| < 2^31);
//   S_ij = (256^(S-1)/sc_i)(256^(S-1)/sc_j) * sum_{d<S} 256^-d T_d   evaluated in fp64 by the epilogue.
// The only approximation is the fixed-point rounding of G' to 8S-2 = 30 bits below each column's maximum
// (S = 4) and the dropped digit pairs s+t >= S (same order); integer accumulation is associative, so the
// result does not depend on tile order, k-splits or the number of GPUs.
//
// Kernel anatomy of the 1-SM kernel (syrk_tc_kernel; one CTA per SM, persistent over (tile, k-split) items, 320
// threads). The default is syrk_tc_2sm_kernel further down: the same pipeline with 2-CTA clusters and
// tcgen05.mma.cta_group::2 (M = 256 across the pair, each CTA holding half of the shared B block).
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.3d of the S int8 planes of the A (rows bi) and B (rows bj)
//              128-column blocks, 128B-swizzled, into a 192 KB shared-memory ring of 3-6 stages (mbarrier full/empty)
//   warp 1   : MMA issuer    -- one elected lane issues tcgen05.mma.cta_group::1.kind::i8, M=128 N=128 K=32,
//              both operands MN-major (pose index contiguous), S(S+1)/2 = 10 digit-pair products per K step
//              into S accumulators (S*128 = 512 TMEM columns); tcgen05.commit releases the smem stage
//   warps 2-9: epilogue      -- tcgen05.ld 32x32b of the S accumulators, fp64 Horner combine, column scales,
//              fp64 partial tile to global (same partial layout as the fp64 path -> same assemble kernel)
#include <cuda.h>
#include <vector>
#include <algorithm>
#include "internal.cuh"

namespace {

constexpr int TILE = BALM_SYRK_TILE;  // 128
constexpr int SMAX = 4;               // digit planes (TMEM: SMAX * 128 columns = 512)
constexpr int KS = 64;                // contraction rows per pipeline stage
constexpr int UMMA_K = 32;            // int8
constexpr int PLANE_TILE_BYTES = KS * TILE;             // 8192: one plane, one operand, one stage
constexpr int RING_BYTES = 3 * 2 * SMAX * PLANE_TILE_BYTES;  // 196608 bytes of stage ring
constexpr int MAX_STAGES = 6;
// a stage holds S planes of the A block and S planes of the B block: 3 stages for S = 4, 4 for S = 3, 6 for S = 2
__host__ __device__ constexpr int stage_bytes_for(int S) { return 2 * S * PLANE_TILE_BYTES; }
__host__ __device__ constexpr int stages_for(int S) { return RING_BYTES / (2 * S * PLANE_TILE_BYTES); }
constexpr int TC_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int MAX_ROWS_PER_ITEM = 32576;                // S * rows * 2^14 < 2^31 with the 64-row rounding of the k-split

// ---------------- PTX wrappers ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a pipeline bug must not hang the GPU box. On timeout set the error flag and bail out.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err) {
  const uint32_t a = smem_u32(bar);
  for (long long it = 0; it < (1ll << 22); it++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((it & 1023) == 1023 && *((volatile int *)err) != 0) return false;
  }
  atomicExch(err, 2);
  return false;
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// multicast variant: the box lands at the same shared-memory offset in every CTA of cta_mask and completes the
// transaction bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// commit that arrives on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// USAGE: collector behaviour of the A operand (consecutive MMAs sharing A keep it in the collector buffer
// instead of re-reading 4 KB of shared memory; SASS A_KEEP / A_REUSE). 0 none, 1 fill, 2 use, 3 lastuse.
#define BALM_MMA_I8(QUAL)                                                                    \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                            \
               "tcgen05.mma.cta_group::1.kind::i8" QUAL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), \
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)                         \
               : "memory")
template <int USAGE>
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (USAGE == 1) BALM_MMA_I8(".collector::a::fill");
  else if constexpr (USAGE == 2) BALM_MMA_I8(".collector::a::use");
  else if constexpr (USAGE == 3) BALM_MMA_I8(".collector::a::lastuse");
  else BALM_MMA_I8("");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   canonical layout (16-byte units) Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)):
//   128 contiguous bytes along MN per k row, 8 k rows per 1024-byte swizzle atom, SBO = next group of 8 k rows,
//   LBO = next 128-byte block along MN (unused: M = N = 128 int8 = one block).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);           // start address  [0,14)
  d |= (uint64_t)((PLANE_TILE_BYTES >> 4) & 0x3FFF) << 16;  // LBO            [16,30)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;           // SBO            [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, signed int8 A and B, both MN-major.
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
  int64_t rows;       // contraction rows of this batch (3 per voxel)
  int nb, tiles, splits;
  const int *S_dev;   // digit-plane count chosen by tc_scale_kernel
  const double *isc;  // [ldq]  256^(S-1) / sc_j
  double *part;       // [splits][tiles][128*128]
  int accumulate;
  int *err;
  int collector;
  // 2-CTA cluster variant: work list of tile pairs (x: bi0 | bj0<<8 | bi1<<16 | bj1<<24, y: flags 1=share B, 2=dup)
  const int2 *pairs;
  int n_pairs;
};

struct ItemInfo {
  int bi, bj, sp, t;
  bool a_from_b;   // A operand = the (shared) B block (diagonal tile of a B-sharing pair): A is not loaded
  bool b_from_a;   // B operand = own A block (diagonal tile, not sharing)
  bool share;      // B block arrives half from this CTA, half from the peer (TMA multicast)
  bool store;      // false for the duplicate of an unpaired tile
};

__device__ __forceinline__ void tile_coords_tc(int t, int nb, int &bi, int &bj) {
  int r = 0, rem = t;
  while (rem >= nb - r) { rem -= nb - r; r++; }
  bi = r;
  bj = r + rem;
}

template <bool CLUSTER>
__device__ __forceinline__ ItemInfo decode_item(const TcArgs &a, int work, uint32_t rank) {
  ItemInfo it;
  if (!CLUSTER) {
    it.t = work % a.tiles;
    it.sp = work / a.tiles;
    tile_coords_tc(it.t, a.nb, it.bi, it.bj);
    it.a_from_b = false;
    it.b_from_a = (it.bi == it.bj);
    it.share = false;
    it.store = true;
  } else {
    const int2 pr = a.pairs[work % a.n_pairs];
    it.sp = work / a.n_pairs;
    it.bi = (pr.x >> (rank ? 16 : 0)) & 255;
    it.bj = (pr.x >> (rank ? 24 : 8)) & 255;
    it.t = it.bi * a.nb - it.bi * (it.bi - 1) / 2 + (it.bj - it.bi);
    it.share = (pr.y & 1) != 0;
    it.a_from_b = it.share && it.bi == it.bj;
    it.b_from_a = !it.share && it.bi == it.bj;
    it.store = !((pr.y & 2) && rank == 1);
  }
  return it;
}

// The MMA issue loop runs in ONE thread, so its scalar overhead per UTCIMMA bounds the tensor pipe: everything
// that can be a compile-time constant is (digit-plane count, the (s,t) pair list, descriptor offsets), and a
// descriptor is the stage's base word plus a constant (the 14-bit start-address field never carries).
template <int S, bool COLLECT, bool CLUSTER>
__device__ __forceinline__ void mma_issue_loop(const TcArgs &a, uint8_t *stage_base, uint64_t *full_bar,
                                               uint64_t *empty_bar, uint64_t *tmem_full, uint64_t *tmem_empty,
                                               uint32_t tmem_base, int n_items, int64_t per, int first, int stride,
                                               uint32_t rank) {
  constexpr uint32_t idesc = make_idesc_i8(TILE, TILE);
  constexpr uint32_t PLANE_U = PLANE_TILE_BYTES >> 4;     // descriptor units (16 B) between digit planes
  constexpr uint32_t KK_U = (UMMA_K * TILE) >> 4;         // ... between K=32 sub-steps
  const uint64_t desc_hi = make_desc_mn_sw128(0) & 0xFFFFFFFF00000000ull;
  const uint32_t desc_lo_fixed = (uint32_t)(make_desc_mn_sw128(0) & 0xFFFFFFFFull);  // LBO field
  constexpr int nst = stages_for(S), sbytes = stage_bytes_for(S);
  int st = 0;
  uint32_t ph = 0, n_done = 0;
  bool alive = true;
  for (int item = first; alive && item < n_items; item += stride, n_done++) {
    const ItemInfo info = decode_item<CLUSTER>(a, item, rank);
    const int64_t k_begin = (int64_t)info.sp * per;
    int64_t k_end = k_begin + per;
    if (k_end > a.rows) k_end = a.rows;
    const int nsteps = k_end > k_begin ? (int)((k_end - k_begin + KS - 1) / KS) : 0;
    const uint32_t a_off = info.a_from_b ? S * PLANE_U : 0u;
    const uint32_t b_off = info.b_from_a ? 0u : S * PLANE_U;
    // accumulators must have been drained by the epilogue of the previous item
    if (!__all_sync(0xffffffffu, mbar_wait(tmem_empty, (n_done & 1) ^ 1, a.err))) break;
    tc_fence_after();
    for (int ks = 0; ks < nsteps; ks++) {
      if (!__all_sync(0xffffffffu, mbar_wait(&full_bar[st], ph, a.err))) { alive = false; break; }
      tc_fence_after();
      const uint32_t lo_0 = desc_lo_fixed + ((smem_u32(stage_base + st * sbytes) & 0x3FFFF) >> 4);
      const uint32_t lo_a = lo_0 + a_off;
      const uint32_t lo_b = lo_0 + b_off;
      const uint32_t acc0 = ks > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KS / UMMA_K; kk++) {
#pragma unroll
          for (int s = 0; s < S; s++) {
            const uint64_t da = desc_hi | (uint64_t)(lo_a + s * PLANE_U + kk * KK_U);
#pragma unroll
            for (int tt = 0; tt + s < S; tt++) {
              const uint64_t db = desc_hi | (uint64_t)(lo_b + tt * PLANE_U + kk * KK_U);
              // first contribution to accumulator d = s + tt of this item: s == 0 at ks == 0, kk == 0
              const uint32_t acc = (s > 0 || kk > 0) ? 1u : acc0;
              constexpr int group = S;  // placeholder to keep the expression below readable
              (void)group;
              if (!COLLECT || S - s == 1) tc_mma_i8<0>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else if (tt == 0) tc_mma_i8<1>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else if (tt == S - s - 1) tc_mma_i8<3>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else tc_mma_i8<2>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
            }
          }
        }
        // frees the smem stage once these MMAs have read it (in both CTAs of a pair: the peer multicasts into it)
        if (CLUSTER) tc_commit_mc(&empty_bar[st], 3);
        else tc_commit(&empty_bar[st]);
      }
      __syncwarp();
      if (++st == nst) { st = 0; ph ^= 1; }
    }
    if (alive && elect_one()) tc_commit(tmem_full);  // accumulators complete
    __syncwarp();
  }
}

template <bool CLUSTER>
__device__ __forceinline__ void syrk_tc_body(const CUtensorMap &tmap, const CUtensorMap &tmap_half, const TcArgs &a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned stage ring, then barriers
  uint8_t *stage_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_base + RING_BYTES);
  uint64_t *empty_bar = full_bar + MAX_STAGES;
  uint64_t *tmem_full = empty_bar + MAX_STAGES;
  uint64_t *tmem_empty = tmem_full + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = *a.S_dev;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CLUSTER ? 2 : 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t rank = CLUSTER ? cluster_ctarank() : 0u;
  if (CLUSTER) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into them

  const int n_items = (CLUSTER ? a.n_pairs : a.tiles) * a.splits;
  const int first = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int stride = CLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int64_t per = ((a.rows + a.splits - 1) / a.splits + KS - 1) / KS * KS;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmap) : "memory");
      const int nst = stages_for(S), sbytes = stage_bytes_for(S);
      int st = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int item = first; alive && item < n_items; item += stride) {
        const ItemInfo info = decode_item<CLUSTER>(a, item, rank);
        const int64_t k_begin = (int64_t)info.sp * per;
        int64_t k_end = k_begin + per;
        if (k_end > a.rows) k_end = a.rows;
        const int nsteps = k_end > k_begin ? (int)((k_end - k_begin + KS - 1) / KS) : 0;
        const bool load_a = !info.a_from_b, load_b = !info.b_from_a;
        for (int ks = 0; ks < nsteps; ks++) {
          if (!mbar_wait(&empty_bar[st], ph ^ 1, a.err)) { alive = false; break; }
          uint8_t *sb = stage_base + st * sbytes;
          mbar_expect_tx(&full_bar[st], ((load_a ? 1 : 0) + (load_b ? 1 : 0)) * S * PLANE_TILE_BYTES);
          const int krow = (int)(k_begin + (int64_t)ks * KS);
          for (int s = 0; s < S; s++) {
            if (load_a) tma_load_3d(sb + s * PLANE_TILE_BYTES, &tmap, &full_bar[st], info.bi * TILE, krow, s);
            if (load_b) {
              if (CLUSTER && info.share) {
                // half of the K rows of the shared B block from each CTA of the pair, multicast to both
                tma_load_3d_mc(sb + (S + s) * PLANE_TILE_BYTES + rank * (KS / 2) * TILE, &tmap_half, &full_bar[st],
                               info.bj * TILE, krow + (int)rank * (KS / 2), s, 3);
              } else {
                tma_load_3d(sb + (S + s) * PLANE_TILE_BYTES, &tmap, &full_bar[st], info.bj * TILE, krow, s);
              }
            }
          }
          if (++st == nst) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
#define BALM_ISSUE(SS, CC) mma_issue_loop<SS, CC, CLUSTER>(a, stage_base, full_bar, empty_bar, tmem_full, tmem_empty, tmem_base, n_items, per, first, stride, rank)
    if (a.collector) {
      if (S == 4) BALM_ISSUE(4, true);
      else if (S == 3) BALM_ISSUE(3, true);
      else BALM_ISSUE(2, true);
    } else {
      if (S == 4) BALM_ISSUE(4, false);
      else if (S == 3) BALM_ISSUE(3, false);
      else BALM_ISSUE(2, false);
    }
#undef BALM_ISSUE
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;         // which 64-column half of the tile this warp drains
    const int row = quad * 32 + lane;         // tile row held by this thread
    uint32_t n_done = 0;
    const double inv256 = 1.0 / 256.0;
    for (int item = first; item < n_items; item += stride, n_done++) {
      const ItemInfo info = decode_item<CLUSTER>(a, item, rank);
      const int t = info.t, sp = info.sp, bi = info.bi, bj = info.bj;
      const int64_t k_begin = (int64_t)sp * per;
      const bool empty_item = k_begin >= a.rows;
      if (!mbar_wait(tmem_full, n_done & 1, a.err)) break;
      tc_fence_after();
      double *tile_out = a.part + ((size_t)sp * a.tiles + t) * (TILE * TILE);
      double *out = tile_out + (size_t)row * TILE;
      const double isc_row = __ldg(a.isc + bi * TILE + row);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (TILE / 2); c0 < (half + 1) * (TILE / 2); c0 += 16) {
        double val[16];
        if (empty_item) {
#pragma unroll
          for (int q = 0; q < 16; q++) val[q] = 0.0;
        } else {
          int32_t v[16];
          // Horner over the digit-pair order d = S-1 .. 0
          tc_ld16(lane_addr + (S - 1) * TILE + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; q++) val[q] = (double)v[q];
          for (int d = S - 2; d >= 0; d--) {
            tc_ld16(lane_addr + d * TILE + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 16; q++) val[q] = val[q] * inv256 + (double)v[q];
          }
        }
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          const double2 sc = *reinterpret_cast<const double2 *>(a.isc + bj * TILE + c0 + q);
          double2 w = make_double2(val[q] * isc_row * sc.x, val[q + 1] * isc_row * sc.y);
          double2 *p = reinterpret_cast<double2 *>(out + c0 + q);
          if (a.accumulate) {
            const double2 old = *p;
            w.x += old.x;
            w.y += old.y;
          }
          if (info.store) *p = w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();  // neither CTA leaves while the peer may still multicast into it / commit to it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512));
  }
}

// ====================================================================================================
// cta_group::2 variant: the two SMs of a cluster issue ONE 256 x 128 x 32 MMA per digit pair. CTA r holds the 128
// rows of its own A block (tile (bi_r, bj)) and HALF of the shared B block (64 of its 128 columns), so each SM reads
// 6 KB of operands per MMA instead of 8 KB and receives 25 % fewer TMA bytes -- shared-memory bandwidth is what bounds
// the 1-SM kernel. The leader (rank 0) issues all MMAs; both CTAs load (signalling the leader's full barrier), both
// drain their own 128 TMEM lanes.
// ====================================================================================================
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0
constexpr int BHALF_TILE_BYTES = KS * (TILE / 2);  // 4096: one plane of this CTA's half of the B block, one stage
__host__ __device__ constexpr int stage2_bytes_for(int S) { return S * (PLANE_TILE_BYTES + BHALF_TILE_BYTES); }
__host__ __device__ constexpr int stages2_for(int S) { return RING_BYTES / stage2_bytes_for(S) > MAX_STAGES ? MAX_STAGES : RING_BYTES / stage2_bytes_for(S); }

__device__ __forceinline__ void tma_load_3d_2sm(void *dst, const CUtensorMap *map, uint64_t *leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {  // arrive on the barrier at this offset in CTA rank 0
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
#define BALM_MMA_I8_2SM(QUAL)                                                                \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                            \
               "tcgen05.mma.cta_group::2.kind::i8" QUAL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), \
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)                         \
               : "memory")
template <int USAGE>
__device__ __forceinline__ void tc_mma_i8_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  if constexpr (USAGE == 1) BALM_MMA_I8_2SM(".collector::a::fill");
  else if constexpr (USAGE == 2) BALM_MMA_I8_2SM(".collector::a::use");
  else if constexpr (USAGE == 3) BALM_MMA_I8_2SM(".collector::a::lastuse");
  else BALM_MMA_I8_2SM("");
}
// MN-major, SWIZZLE_64B: 64 contiguous bytes along N per k row, 8 k rows per 512-byte atom, SBO = next 8 k rows
__device__ __forceinline__ uint64_t make_desc_mn_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((BHALF_TILE_BYTES >> 4) & 0x3FFF) << 16;  // LBO (unused: one 64-byte block along N)
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;               // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                                   // SWIZZLE_64B
  return d;
}

template <int S>
__device__ __forceinline__ void mma_issue_loop_2sm(const TcArgs &a, uint8_t *stage_base, uint64_t *full_bar,
                                                   uint64_t *empty_bar, uint64_t *tmem_full, uint64_t *tmem_empty,
                                                   uint32_t tmem_base, int n_items, int64_t per, int first, int stride) {
  constexpr uint32_t idesc = make_idesc_i8(2 * TILE, TILE);  // M = 256 across the pair, N = 128
  constexpr uint32_t PLANE_U = PLANE_TILE_BYTES >> 4, BPLANE_U = BHALF_TILE_BYTES >> 4;
  constexpr uint32_t KK_A = (UMMA_K * TILE) >> 4, KK_B = (UMMA_K * (TILE / 2)) >> 4;
  const uint64_t hi_a = make_desc_mn_sw128(0) & 0xFFFFFFFF00000000ull, hi_b = make_desc_mn_sw64(0) & 0xFFFFFFFF00000000ull;
  const uint32_t fx_a = (uint32_t)(make_desc_mn_sw128(0) & 0xFFFFFFFFull), fx_b = (uint32_t)(make_desc_mn_sw64(0) & 0xFFFFFFFFull);
  constexpr int nst = stages2_for(S), sbytes = stage2_bytes_for(S);
  int st = 0;
  uint32_t ph = 0, n_done = 0;
  bool alive = true;
  for (int item = first; alive && item < n_items; item += stride, n_done++) {
    const int sp = item / a.n_pairs;
    const int64_t k_begin = (int64_t)sp * per;
    int64_t k_end = k_begin + per;
    if (k_end > a.rows) k_end = a.rows;
    const int nsteps = k_end > k_begin ? (int)((k_end - k_begin + KS - 1) / KS) : 0;
    if (!__all_sync(0xffffffffu, mbar_wait(tmem_empty, (n_done & 1) ^ 1, a.err))) break;  // both CTAs drained
    tc_fence_after();
    for (int ks = 0; ks < nsteps; ks++) {
      if (!__all_sync(0xffffffffu, mbar_wait(&full_bar[st], ph, a.err))) { alive = false; break; }
      tc_fence_after();
      const uint32_t base = (smem_u32(stage_base + st * sbytes) & 0x3FFFF) >> 4;
      const uint32_t lo_a = fx_a + base, lo_b = fx_b + base + S * PLANE_U;
      const uint32_t acc0 = ks > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KS / UMMA_K; kk++) {
#pragma unroll
          for (int s = 0; s < S; s++) {
            const uint64_t da = hi_a | (uint64_t)(lo_a + s * PLANE_U + kk * KK_A);
#pragma unroll
            for (int tt = 0; tt + s < S; tt++) {
              const uint64_t db = hi_b | (uint64_t)(lo_b + tt * BPLANE_U + kk * KK_B);
              const uint32_t acc = (s > 0 || kk > 0) ? 1u : acc0;
              if (S - s == 1) tc_mma_i8_2sm<0>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else if (tt == 0) tc_mma_i8_2sm<1>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else if (tt == S - s - 1) tc_mma_i8_2sm<3>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
              else tc_mma_i8_2sm<2>(tmem_base + (s + tt) * TILE, da, db, idesc, acc);
            }
          }
        }
        tc_commit_2sm_mc(&empty_bar[st], 3);  // frees this stage in both CTAs
      }
      __syncwarp();
      if (++st == nst) { st = 0; ph ^= 1; }
    }
    if (alive && elect_one()) tc_commit_2sm_mc(tmem_full, 3);  // accumulators complete in both CTAs
    __syncwarp();
  }
}

__device__ __forceinline__ void syrk_tc_2sm_body(const CUtensorMap &tmap, const CUtensorMap &tmap_b64, const TcArgs &a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *stage_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_base + RING_BYTES);
  uint64_t *empty_bar = full_bar + MAX_STAGES;
  uint64_t *tmem_full = empty_bar + MAX_STAGES;
  uint64_t *tmem_empty = tmem_full + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = *a.S_dev;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 16);  // 8 epilogue warps of each CTA arrive on the leader's barrier
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  cluster_sync_all();

  const int n_items = a.n_pairs * a.splits;
  const int first = (int)(blockIdx.x >> 1), stride = (int)(gridDim.x >> 1);
  const int64_t per = ((a.rows + a.splits - 1) / a.splits + KS - 1) / KS * KS;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; transactions complete on the leader's full barrier) ==========
    if (lane == 0) {
      const int nst = stages2_for(S), sbytes = stage2_bytes_for(S);
      int st = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int item = first; alive && item < n_items; item += stride) {
        const int2 pr = a.pairs[item % a.n_pairs];
        const int sp = item / a.n_pairs;
        const int bi = (pr.x >> (rank ? 16 : 0)) & 255, bj = (pr.x >> 8) & 255;
        const int64_t k_begin = (int64_t)sp * per;
        int64_t k_end = k_begin + per;
        if (k_end > a.rows) k_end = a.rows;
        const int nsteps = k_end > k_begin ? (int)((k_end - k_begin + KS - 1) / KS) : 0;
        for (int ks = 0; ks < nsteps; ks++) {
          if (!mbar_wait(&empty_bar[st], ph ^ 1, a.err)) { alive = false; break; }
          uint8_t *sb = stage_base + st * sbytes;
          if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * sbytes);  // bytes of BOTH CTAs land on the leader's barrier
          const int krow = (int)(k_begin + (int64_t)ks * KS);
          for (int s = 0; s < S; s++) {
            tma_load_3d_2sm(sb + s * PLANE_TILE_BYTES, &tmap, &full_bar[st], bi * TILE, krow, s);
            tma_load_3d_2sm(sb + S * PLANE_TILE_BYTES + s * BHALF_TILE_BYTES, &tmap_b64, &full_bar[st],
                            bj * TILE + (int)rank * (TILE / 2), krow, s);
          }
          if (++st == nst) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (rank == 0) {
      if (S == 4) mma_issue_loop_2sm<4>(a, stage_base, full_bar, empty_bar, tmem_full, tmem_empty, tmem_base, n_items, per, first, stride);
      else if (S == 3) mma_issue_loop_2sm<3>(a, stage_base, full_bar, empty_bar, tmem_full, tmem_empty, tmem_base, n_items, per, first, stride);
      else mma_issue_loop_2sm<2>(a, stage_base, full_bar, empty_bar, tmem_full, tmem_empty, tmem_base, n_items, per, first, stride);
    }
  } else {
    // ===================== epilogue: each CTA drains its own 128 TMEM lanes =====================
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane;
    uint32_t n_done = 0;
    const double inv256 = 1.0 / 256.0;
    for (int item = first; item < n_items; item += stride, n_done++) {
      const int2 pr = a.pairs[item % a.n_pairs];
      const int sp = item / a.n_pairs;
      const int bi = (pr.x >> (rank ? 16 : 0)) & 255, bj = (pr.x >> 8) & 255;
      const bool store = !((pr.y & 2) && rank == 1);
      const bool flipped = bi > bj;  // computed as (A = bi, B = bj) with bi > bj: the transpose of tile (bj, bi)
      const int ti = flipped ? bj : bi, tj = flipped ? bi : bj;
      const int t = ti * a.nb - ti * (ti - 1) / 2 + (tj - ti);
      const int64_t k_begin = (int64_t)sp * per;
      const bool empty_item = k_begin >= a.rows;
      if (!mbar_wait(tmem_full, n_done & 1, a.err)) break;
      tc_fence_after();
      double *tile_out = a.part + ((size_t)sp * a.tiles + t) * (TILE * TILE);
      double *out = tile_out + (size_t)row * TILE;
      const double isc_row = __ldg(a.isc + bi * TILE + row);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (TILE / 2); c0 < (half + 1) * (TILE / 2); c0 += 16) {
        double val[16];
        if (empty_item) {
#pragma unroll
          for (int q = 0; q < 16; q++) val[q] = 0.0;
        } else {
          int32_t v[16];
          tc_ld16(lane_addr + (S - 1) * TILE + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; q++) val[q] = (double)v[q];
          for (int d = S - 2; d >= 0; d--) {
            tc_ld16(lane_addr + d * TILE + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 16; q++) val[q] = val[q] * inv256 + (double)v[q];
          }
        }
        if (store && !flipped) {
#pragma unroll
          for (int q = 0; q < 16; q += 2) {
            const double2 sc = *reinterpret_cast<const double2 *>(a.isc + bj * TILE + c0 + q);
            double2 w = make_double2(val[q] * isc_row * sc.x, val[q + 1] * isc_row * sc.y);
            double2 *p = reinterpret_cast<double2 *>(out + c0 + q);
            if (a.accumulate) {
              const double2 old = *p;
              w.x += old.x;
              w.y += old.y;
            }
            *p = w;
          }
        } else if (store) {  // element (row, c0+q) of the computed product is element (c0+q, row) of the stored tile
#pragma unroll
          for (int q = 0; q < 16; q++) {
            double w = val[q] * isc_row * __ldg(a.isc + bj * TILE + c0 + q);
            double *p = tile_out + (size_t)(c0 + q) * TILE + row;
            if (a.accumulate) w += *p;
            *p = w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    syrk_tc_2sm_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b64, TcArgs a) {
  syrk_tc_2sm_body(tmap, tmap_b64, a);
}
// the same kernel inside 128 registers per thread (the bound of a 512-thread CTA): 320 x 128 = 40 960 registers leave room
// for a 96-thread CTA of the observation sweep (248 registers) on the same SM
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
    syrk_tc_2sm_r128_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b64, TcArgs a) {
  syrk_tc_2sm_body(tmap, tmap_b64, a);
}

__global__ void __launch_bounds__(TC_THREADS, 1) syrk_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs a) {
  syrk_tc_body<false>(tmap, tmap, a);
}

// 2-CTA cluster variant: the two CTAs of a cluster compute two tiles of the same block column and share the B block
// (each loads half of its K rows and TMA-multicasts them to both), which cuts the L2->SM operand traffic by 25 %.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    syrk_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_half, TcArgs a) {
  syrk_tc_body<true>(tmap, tmap_half, a);
}

// Column scales and digit-plane count of one batch, decided ON THE DEVICE from the column maxima and sums of squares
// of G' (one block):
//   e_j = max|G'_j| / sqrt(sum G'_j^2)  is the concentration of column j (1 = a single dominant entry, ~1/sqrt(rows)
//   for evenly spread columns); rounding G' to 2^-(8S-2) of the column maximum perturbs H_ij by about
//   2^-(8S-2) * e * sqrt(H_ii H_jj).  S = 3 is used when that stays below 5e-9 for every column, else S = 4
//   (forced_S > 0 overrides).  Then sc_j = 2^p with colmax_j * sc_j in [2^(8S-3), 2^(8S-2)), isc_j = 256^(S-1)/sc_j.
//   fresh values (scale 2^p with |G'_j| 2^p in [2^(8S-3), 2^(8S-2)), S by the concentration rule) always go to
//   sc_next/isc_next/S_next (when given): they are what the NEXT evaluation of the same voxels speculates with.
//   spec = 0: they are also the scales of this batch (sc/isc/S_dev), *spec_ok = 0 -> the OBS_INT8 sweep runs.
//   spec = 1: the digit planes were ALREADY written by the fused sweep with sc/S_dev. They stand (*spec_ok = 1)
//             iff S is still what the rule asks for and every column (a) stayed inside the balanced-digit range,
//             |G'_j| sc_j < 0.99 * 2^(8S-1), and (b) kept a rounding unit 1/sc_j <= 1e-8 sqrt(sum G'_j^2) -- the bound
//             the S rule itself guarantees. Otherwise the fresh values are adopted and sweep 2 is re-armed.
__global__ void __launch_bounds__(1024) tc_scale_kernel(const unsigned long long *colmax_bits, const double *accum_batch,
                                                        int Np, int n, double *sc, double *isc, int ldq, int forced_S,
                                                        int *S_dev, int spec, double *sc_next, double *isc_next,
                                                        int *S_next, int *spec_ok) {
  __shared__ double red[32];
  __shared__ int S_sh, bad_sh;
  if (threadIdx.x == 0) bad_sh = 0;
  double emax = 0.0;
  for (int j = threadIdx.x; j < n; j += 1024) {
    const double m = __longlong_as_double((long long)colmax_bits[j]);
    const double h = accum_batch[(size_t)(27 + j % 6) * Np + j / 6];
    if (m > 0.0 && h > 0.0) emax = fmax(emax, m / sqrt(h));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = emax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int w = 0; w < 32; w++) e = fmax(e, red[w]);
    int S = (e * (1.0 / 4194304.0) <= 5e-9) ? 3 : 4;  // 2^-22 * e
    if (forced_S > 0) S = forced_S;
    S_sh = S;
    if (spec && *S_dev != S) bad_sh = 1;
  }
  __syncthreads();
  const int S = S_sh;
  double *fsc = sc_next ? sc_next : sc, *fisc = sc_next ? isc_next : isc;  // where the fresh values go first
  int bad = 0;
  for (int j = threadIdx.x; j < ldq; j += 1024) {
    const double m = __longlong_as_double((long long)colmax_bits[j]);
    const bool live = (m > 0.0) && (m < 1e300);
    if (spec && live && j < n) {
      const double used = sc[j];
      const double h = accum_batch[(size_t)(27 + j % 6) * Np + j / 6];
      if (!(used > 0.0) || !(m * used < 0.99 * ldexp(1.0, 8 * S - 1)) || !(1.0 <= 1e-8 * sqrt(h) * used)) bad = 1;
    }
    if (!live) {
      fsc[j] = 0.0;
      fisc[j] = 0.0;
      continue;
    }
    int e;
    frexp(m, &e);  // m = f * 2^e, f in [0.5, 1)  ->  m < 2^e
    const int p = 8 * S - 2 - e;
    fsc[j] = ldexp(1.0, p);
    fisc[j] = ldexp(1.0, 8 * (S - 1) - p);
  }
  if (bad) atomicOr(&bad_sh, 1);
  __syncthreads();
  const bool keep = spec && !bad_sh;
  if (sc_next && !keep) {
    for (int j = threadIdx.x; j < ldq; j += 1024) {
      sc[j] = sc_next[j];
      isc[j] = isc_next[j];
    }
  }
  if (threadIdx.x == 0) {
    if (!keep) *S_dev = S;
    if (S_next) *S_next = S;
    *spec_ok = keep ? 1 : 0;
  }
}

// start of a speculative evaluation: the scales found by the previous one become the ones in use
// (skew: test hook BALM_TC_SPEC_SKEW -- the adopted scales are multiplied by 2^skew so that the device-side check
//  has something to reject; 0 in production)
__global__ void tc_adopt_kernel(double *sc, double *isc, int *S_dev, const double *sc_next, const double *isc_next,
                                const int *S_next, int ldq, int skew) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ldq) {
    sc[j] = ldexp(sc_next[j], skew);
    isc[j] = ldexp(isc_next[j], -skew);
  }
  if (j == 0) *S_dev = *S_next;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  CUtensorMap map;
  CUtensorMap map_half;     // box {128, KS/2, 1}: half-stage loads of the pair kernel
  int2 *pairs = nullptr;    // device work list of tile pairs
  int n_pairs = 0;
  bool use_pairs = false;
  CUtensorMap map_b64;      // box {64, KS, 1}, 64-byte swizzle: this CTA's half of the B block (2-SM kernel)
  int2 *pairs2 = nullptr;   // work list of the 2-SM kernel (every pair shares its B block)
  int n_pairs2 = 0;
  bool use_2sm = false;
  double *sc = nullptr, *isc = nullptr;
  double *sc_next = nullptr, *isc_next = nullptr;  // fresh scales of the last full-range evaluation (speculation)
  int *S_next = nullptr, *spec_ok = nullptr;
  bool spec_ready = false;                         // sc_next/S_next describe the currently registered voxels
  bool spec_enabled = true;                        // BALM_NO_SPEC=1 keeps the two-sweep path
  bool last_spec = false;                          // the batch in flight used the fused sweep
  int spec_skew = 0;                               // BALM_TC_SPEC_SKEW (tests)
  unsigned long long *colmax = nullptr;
  int *err = nullptr;
  int64_t rows_alloc = 0;
  int forced_S = 0;      // BALM_TC_SLICES override (0 = decided on the device per batch)
  int *S_dev = nullptr;  // [1]
  int last_S = 0;        // read back with the error flag (instrumentation)
};

}  // namespace

int tensor_syrk_init(balm_ctx *c) {
  tensor_syrk_free(c);
  TcState *st = new TcState();
  c->tmap = st;
  // Digit planes (3 or 4) are chosen per batch on the device (tc_scale_kernel); BALM_TC_SLICES=2..4 forces a count.
  if (const char *e = getenv("BALM_TC_SLICES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= SMAX) st->forced_S = v;
  }
  const int ldq = c->ldg;
  st->rows_alloc = (3 * c->VB + KS - 1) / KS * KS;
  CUDA_TRY(cudaMalloc((void **)&c->Gq, (size_t)SMAX * st->rows_alloc * ldq));
  CUDA_TRY(cudaMemset(c->Gq, 0, (size_t)SMAX * st->rows_alloc * ldq));  // column padding [6N, ldg) stays zero
  CUDA_TRY(cudaMalloc((void **)&st->S_dev, sizeof(int)));
  CUDA_TRY(cudaMalloc((void **)&st->sc, sizeof(double) * ldq));
  CUDA_TRY(cudaMalloc((void **)&st->isc, sizeof(double) * ldq));
  CUDA_TRY(cudaMalloc((void **)&st->sc_next, sizeof(double) * ldq));
  CUDA_TRY(cudaMalloc((void **)&st->isc_next, sizeof(double) * ldq));
  CUDA_TRY(cudaMalloc((void **)&st->S_next, sizeof(int)));
  CUDA_TRY(cudaMalloc((void **)&st->spec_ok, sizeof(int)));
  CUDA_TRY(cudaMemset(st->spec_ok, 0, sizeof(int)));
  st->spec_enabled = getenv("BALM_NO_SPEC") == nullptr;
  if (const char *e = getenv("BALM_TC_SPEC_SKEW")) st->spec_skew = atoi(e);
  CUDA_TRY(cudaMalloc((void **)&st->colmax, sizeof(unsigned long long) * ldq));
  CUDA_TRY(cudaMalloc((void **)&st->err, sizeof(int)));
  CUDA_TRY(cudaMemset(st->err, 0, sizeof(int)));
  c->colmax = st->colmax;  // the observation pass writes the column maxima here

  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) {
    balm_set_error("cuTensorMapEncodeTiled not available from the driver");
    return BALM_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)ldq, (cuuint64_t)st->rows_alloc, (cuuint64_t)SMAX};
  const cuuint64_t strides[2] = {(cuuint64_t)ldq, (cuuint64_t)ldq * st->rows_alloc};
  const cuuint32_t box[3] = {TILE, KS, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(&st->map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->Gq, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    balm_set_error("cuTensorMapEncodeTiled failed");
    return BALM_ERR_CUDA;
  }
  const cuuint32_t box_half[3] = {TILE, KS / 2, 1};
  if (encode(&st->map_half, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->Gq, dims, strides, box_half, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    balm_set_error("cuTensorMapEncodeTiled (half box) failed");
    return BALM_ERR_CUDA;
  }
  {  // work list of the pair kernel: per block column bj, consecutive rows (bi, bi+1) share the B block bj;
     // the leftover tile of an odd column is paired with another leftover (no sharing) or duplicated
    const int nb = c->syrk_nb;
    std::vector<int2> pairs;
    std::vector<int> solo;
    for (int bj = 0; bj < nb; bj++) {
      int bi = 0;
      for (; bi + 1 <= bj; bi += 2) pairs.push_back(make_int2(bi | (bj << 8) | ((bi + 1) << 16) | (bj << 24), 1));
      if (bi <= bj) solo.push_back(bi | (bj << 8));
    }
    for (size_t i = 0; i + 1 < solo.size(); i += 2) pairs.push_back(make_int2(solo[i] | (solo[i + 1] << 16), 0));
    if (solo.size() & 1) pairs.push_back(make_int2(solo.back() | (solo.back() << 16), 2));
    st->n_pairs = (int)pairs.size();
    CUDA_TRY(cudaMalloc((void **)&st->pairs, sizeof(int2) * pairs.size()));
    CUDA_TRY(cudaMemcpy(st->pairs, pairs.data(), sizeof(int2) * pairs.size(), cudaMemcpyHostToDevice));
    const char *e = getenv("BALM_TC_PAIR");
    st->use_pairs = (e ? atoi(e) != 0 : false) && nb < 256 && (c->sm_count % 2 == 0);
    // 2-SM list: every pair is two tiles that share their B block. Tile (bi, bj), bi <= bj, normally uses B = bj, so block
    // column bj holds bj+1 tiles -- an odd number for even bj, which used to cost one wasted slot per even column (312
    // slots for the 300 tiles of n = 3000). The product is symmetric, so a tile can just as well be computed in the other
    // orientation (A = block bj, B = block bi: the transposed tile, stored transposed by the epilogue): for consecutive
    // even columns j1 < j2 the tile (j1, j2) moves from column j2 to column j1, which makes both counts even. Only when the
    // number of even columns is odd does one column keep a leftover, paired with a discarded duplicate (flag 2).
    std::vector<std::vector<int>> rows_of(nb);  // A-block indices of the tiles computed with B = bj
    for (int bj = 0; bj < nb; bj++)
      for (int bi = 0; bi <= bj; bi++) rows_of[bj].push_back(bi);
    if (!getenv("BALM_TC_NO_FLIP")) {
      for (int j1 = 0; j1 + 2 < nb; j1 += 4) {
        const int j2 = j1 + 2;
        rows_of[j2].erase(std::find(rows_of[j2].begin(), rows_of[j2].end(), j1));  // tile (j1, j2) leaves column j2 ...
        rows_of[j1].push_back(j2);                                                  // ... and joins column j1 as A = j2 > B = j1
      }
    }
    std::vector<int2> p2;
    for (int bj = 0; bj < nb; bj++) {
      const std::vector<int> &r = rows_of[bj];
      size_t q = 0;
      for (; q + 1 < r.size(); q += 2) p2.push_back(make_int2(r[q] | (bj << 8) | (r[q + 1] << 16) | (bj << 24), 1));
      if (q < r.size()) {  // leftover: duplicate partner, result discarded
        const int other = r[q] + 1 < nb ? r[q] + 1 : r[q];
        p2.push_back(make_int2(r[q] | (bj << 8) | (other << 16) | (bj << 24), 1 | 2));
      }
    }
    st->n_pairs2 = (int)p2.size();
    CUDA_TRY(cudaMalloc((void **)&st->pairs2, sizeof(int2) * p2.size()));
    CUDA_TRY(cudaMemcpy(st->pairs2, p2.data(), sizeof(int2) * p2.size(), cudaMemcpyHostToDevice));
    const cuuint32_t box_b64[3] = {TILE / 2, KS, 1};
    if (encode(&st->map_b64, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->Gq, dims, strides, box_b64, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      balm_set_error("cuTensorMapEncodeTiled (64-byte box) failed");
      return BALM_ERR_CUDA;
    }
    const char *e2 = getenv("BALM_TC_2SM");
    st->use_2sm = (e2 ? atoi(e2) != 0 : true) && nb < 256 && (c->sm_count % 2 == 0);
  }
  // int32 exactness bound: every (tile, split) item contracts at most MAX_ROWS_PER_ITEM rows
  const int64_t rows = 3 * c->VB;
  const int min_splits = (int)((rows + MAX_ROWS_PER_ITEM - 1) / MAX_ROWS_PER_ITEM);
  if (c->syrk_splits < min_splits) {
    cudaFree(c->syrk_part);
    c->syrk_splits = min_splits;
    CUDA_TRY(cudaMalloc((void **)&c->syrk_part,
                        sizeof(double) * (size_t)c->syrk_splits * c->syrk_tiles * TILE * TILE));
  }
  const int smem = RING_BYTES + 1024 + 256;
  CUDA_TRY(cudaFuncSetAttribute(syrk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(syrk_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(syrk_tc_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(syrk_tc_2sm_r128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return BALM_OK;
}

void tensor_syrk_free(balm_ctx *c) {
  if (c->tmap) {
    TcState *st = static_cast<TcState *>(c->tmap);
    cudaFree(st->sc_next); cudaFree(st->isc_next); cudaFree(st->S_next); cudaFree(st->spec_ok);
    cudaFree(st->sc); cudaFree(st->isc); cudaFree(st->colmax); cudaFree(st->err); cudaFree(st->S_dev); cudaFree(st->pairs); cudaFree(st->pairs2);
    delete st;
    c->tmap = nullptr;
  }
  cudaFree(c->Gq);
  c->Gq = nullptr;
  c->colmax = nullptr;
}

// One batch of the tensor path: (1) column maxima of G' -> power-of-two scales, (2) observation sweep writing the
// int8 digit planes directly, (3) tcgen05 SYRK. ev[2] is recorded between the observation sweeps and the SYRK.
int tensor_obs_and_syrk(balm_ctx *c, const double *poses, int64_t v0, int64_t v1, bool first_batch) {
  TcState *st = static_cast<TcState *>(c->tmap);
  if (!st) { balm_set_error("tensor path not initialised"); return BALM_ERR_INVALID; }
  const int ldq = c->ldg;
  const int64_t rows = 3 * (v1 - v0);
  const int64_t rows_padded = (rows + KS - 1) / KS * KS;
  CUDA_TRY(cudaMemsetAsync(st->colmax, 0, sizeof(unsigned long long) * ldq, c->stream));
  // Speculation lives inside one balm_damping_iter call only: its first evaluation sweeps twice, and the public
  // balm_evaluate always does -- so every entry point stays a pure function of its arguments (same inputs, same bits,
  // whatever was evaluated before).
  const bool full = c->lm_active && v0 == 0 && v1 == c->M;
  const bool spec = full && st->spec_ready && st->spec_enabled;
  const int64_t plane_stride = (int64_t)st->rows_alloc * ldq;
  int rc;
  if (spec) {  // one sweep with the previous evaluation's scales; sweep 2 below only runs if tc_scale_kernel re-arms it
    tc_adopt_kernel<<<(ldq + 255) / 256, 256, 0, c->stream>>>(st->sc, st->isc, st->S_dev, st->sc_next, st->isc_next,
                                                              st->S_next, ldq, st->spec_skew);
    c->launches += 1;
    rc = launch_obs_int8(c, poses, v0, v1, first_batch, st->sc, c->Gq, plane_stride, st->S_dev, SMAX, rows_padded, true,
                         nullptr);
  } else {
    rc = launch_obs_colmax(c, poses, v0, v1, first_batch);
  }
  if (rc != BALM_OK) return rc;
  tc_scale_kernel<<<1, 1024, 0, c->stream>>>(st->colmax, c->accum_batch, c->Np, c->n, st->sc, st->isc, ldq,
                                             st->forced_S, st->S_dev, spec ? 1 : 0, full ? st->sc_next : nullptr,
                                             full ? st->isc_next : nullptr, full ? st->S_next : nullptr, st->spec_ok);
  rc = launch_obs_int8(c, poses, v0, v1, first_batch, st->sc, c->Gq, plane_stride, st->S_dev, SMAX, rows_padded, false,
                       spec ? st->spec_ok : nullptr);
  if (rc != BALM_OK) return rc;
  if (full) st->spec_ready = true;
  st->last_spec = spec;
  CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
  TcArgs a{rows_padded, c->syrk_nb, c->syrk_tiles, c->syrk_splits, st->S_dev, st->isc, c->syrk_part,
           first_batch ? 0 : 1, st->err, getenv("BALM_TC_NO_COLLECTOR") ? 0 : 1, st->pairs, st->n_pairs};
  const int smem = RING_BYTES + 1024 + 256;
  if (st->use_2sm) {
    a.pairs = st->pairs2;
    a.n_pairs = st->n_pairs2;
    const int items = a.n_pairs * a.splits;
    int clusters = c->sm_count / 2;
    if (items < clusters) clusters = items;
    syrk_tc_2sm_kernel<<<2 * clusters, TC_THREADS, smem, c->stream>>>(st->map, st->map_b64, a);
    if (cudaGetLastError() != cudaSuccess) {  // cluster launch not possible here: fall back to the 1-SM kernel for good
      st->use_2sm = false;
      a.pairs = st->pairs;
      a.n_pairs = st->n_pairs;
      const int items1 = a.tiles * a.splits;
      syrk_tc_kernel<<<items1 < c->sm_count ? items1 : c->sm_count, TC_THREADS, smem, c->stream>>>(st->map, a);
    }
  } else if (st->use_pairs) {
    const int items = a.n_pairs * a.splits;
    int clusters = c->sm_count / 2;
    if (items < clusters) clusters = items;
    syrk_tc_pair_kernel<<<2 * clusters, TC_THREADS, smem, c->stream>>>(st->map, st->map_half, a);
  } else {
    const int items = a.tiles * a.splits;
    const int grid = items < c->sm_count ? items : c->sm_count;
    syrk_tc_kernel<<<grid, TC_THREADS, smem, c->stream>>>(st->map, a);
  }
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return BALM_OK;
}

// Experiment (scripts/overlap_probe.py): can the observation sweep run NEXT TO the SYRK on the same SMs? Needs a
// tensor-mode, dense, single-batch context whose last damping_iter left valid scales, stats and digit planes. Times
//   out[0] the 2-SM SYRK alone (168 registers)          out[1] the same inside 128 registers
//   out[2] the fused sweep alone with 128-thread CTAs   out[3] with 96-thread CTAs
//   out[4] SYRK (128 registers, high-priority stream) and the 96-thread sweep launched together
// The sweep rewrites the digit planes with the values they already hold (same poses, same scales).
int tensor_overlap_probe(balm_ctx *c, const double *poses, int reps, float *out) {
  TcState *st = static_cast<TcState *>(c->tmap);
  if (!st || !st->use_2sm || !c->dense || c->VB < c->M || !st->spec_ready) {
    balm_set_error("overlap probe: needs the 2-SM tensor path, a dense single-batch problem and a finished damping_iter");
    return BALM_ERR_INVALID;
  }
  const int64_t rows = 3 * c->M, rows_padded = (rows + KS - 1) / KS * KS;
  const int64_t plane_stride = (int64_t)st->rows_alloc * c->ldg;
  TcArgs a{rows_padded, c->syrk_nb, c->syrk_tiles, c->syrk_splits, st->S_dev, st->isc, c->syrk_part,
           0, st->err, 1, st->pairs2, st->n_pairs2};
  const int smem = RING_BYTES + 1024 + 256;
  const int items = a.n_pairs * a.splits;
  int clusters = c->sm_count / 2;
  if (items < clusters) clusters = items;
  cudaStream_t hi;
  int lo_p, hi_p;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
  CUDA_TRY(cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, hi_p));
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  auto syrk = [&](bool r128, cudaStream_t s) {
    if (r128) syrk_tc_2sm_r128_kernel<<<2 * clusters, TC_THREADS, smem, s>>>(st->map, st->map_b64, a);
    else syrk_tc_2sm_kernel<<<2 * clusters, TC_THREADS, smem, s>>>(st->map, st->map_b64, a);
  };
  auto sweep = [&](int wpc) {
    return launch_obs_int8(c, poses, 0, c->M, true, st->sc, c->Gq, plane_stride, st->S_dev, SMAX, rows_padded, true, nullptr, wpc);
  };
  for (int mode = 0; mode < 5; mode++) {
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      CUDA_TRY(cudaStreamSynchronize(hi));
      CUDA_TRY(cudaEventRecord(e0, c->stream));
      if (mode == 0) syrk(false, c->stream);
      if (mode == 1) syrk(true, c->stream);
      if (mode == 2) { int rc = sweep(4); if (rc) return rc; }
      if (mode == 3) { int rc = sweep(3); if (rc) return rc; }
      if (mode == 4) {
        CUDA_TRY(cudaStreamWaitEvent(hi, e0, 0));
        syrk(true, hi);
        CUDA_TRY(cudaEventRecord(e2, hi));
        int rc = sweep(3);
        if (rc) return rc;
        CUDA_TRY(cudaStreamWaitEvent(c->stream, e2, 0));
      }
      CUDA_TRY(cudaEventRecord(e1, c->stream));
      CUDA_TRY(cudaEventSynchronize(e1));
      CUDA_TRY(cudaGetLastError());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    out[mode] = best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  cudaStreamDestroy(hi);
  return BALM_OK;
}

// The SYRK's error flag, the digit-plane count and the speculation verdict of the batch in flight: the copies are
// enqueued (pinned host words), the host reads them after its next synchronisation of the stream.
int tensor_syrk_check_enqueue(balm_ctx *c) {
  TcState *st = static_cast<TcState *>(c->tmap);
  if (!st) return BALM_OK;
  CUDA_TRY(cudaMemcpyAsync(c->h_flags + 1, st->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_flags + 2, st->S_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_flags + 3, st->spec_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return BALM_OK;
}

int tensor_syrk_check_finish(balm_ctx *c) {  // after the stream has been synchronised
  TcState *st = static_cast<TcState *>(c->tmap);
  if (!st) return BALM_OK;
  st->last_S = c->h_flags[2];
  c->tm.digit_planes = st->last_S;
  if (st->last_spec) (c->h_flags[3] ? c->tm.single_sweeps : c->tm.redone_sweeps) += 1;
  if (c->h_flags[1] != 0) {
    balm_set_error("tcgen05 SYRK pipeline timed out (mbarrier wait bound exceeded)");
    return BALM_ERR_CUDA;
  }
  return BALM_OK;
}

int tensor_syrk_check(balm_ctx *c) {
  int rc = tensor_syrk_check_enqueue(c);
  if (rc != BALM_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return tensor_syrk_check_finish(c);
}

void tensor_syrk_new_problem(balm_ctx *c) {  // same-shape re-registration: the workspaces stay, the speculation does not
  TcState *st = static_cast<TcState *>(c->tmap);
  if (st) st->spec_ready = false;
}
