// peekvit_b200 — CTA-pair (cta_group::2) tcgen05 GEMM with fused epilogues (sm_100a).
//
//   C[M,N] = A[M,K] (bf16, row-major) x W[N,K]^T (bf16, nn.Linear layout), f32 accumulate in TMEM
//
// The dense-rows fast path of pk_gemm_bf16 (same call sites as pk_gemm.cu: reference
// models/blocks.py:94 in-proj, vit.py:49-51 out-proj + residual, blocks.py:81-83 fc1+GELU / fc2
// + residual).  Two CTAs on neighbouring SMs (one cluster) own one 256 x BN output tile:
//   * each CTA TMA-loads its own 128 rows of A and its own BN/2 rows of W per 64-wide K block
//     (the pair reads every operand byte once: half the shared-memory and L2 traffic per FLOP of
//     a 128 x BN single-CTA tile);
//   * the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256); each CTA's
//     128 x BN slice of the accumulator lands in its own TMEM;
//   * completion is multicast to both CTAs' barriers with tcgen05.commit ... multicast::cluster.
// Per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (leader only), warp 2 = TMEM allocator,
// warps 4-11 = epilogue (tcgen05.ld -> registers -> swizzled smem -> TMA store; the fp32 residual
// tile is TMA-loaded into the same smem slot two chunks ahead).  Two accumulator stages, so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Every mbarrier wait is bounded.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

#include <cstdlib>

namespace pk {

constexpr int kP_BM = 128;          // rows per CTA (256 per pair)
constexpr int kP_BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int kP_BufBytes = 4096;   // one 32-row x 128-byte staging tile

// RED: the residual epilogue is in place (out == resid), so the add is done by a TMA reduction
// (cp.reduce.async.bulk.tensor .add, executed at L2) and the residual tile never visits shared memory.
// EW epilogue warps per CTA: 8 (two per TMEM lane quarter, each owning half the tile's columns) or 16 (four per
// quarter, a quarter of the columns each: the GELU epilogue is latency-bound with two warps per scheduler).
// LN: 0 = plain; 1 = LayerNorm *producer* (residual epilogue that also emits a bf16 copy of the new residual row and
// per-row partial sums / sums of squares); 2 = LayerNorm *consumer* (bf16 epilogues whose A operand is that raw bf16 copy
// and whose weights carry the LayerNorm gain: out = rstd*acc - rstd*mean*c1[n] + c2[n]).  Together they remove the separate
// LayerNorm kernel: x -> LN -> Linear becomes one GEMM epilogue feeding the next GEMM (vit.py:48-55).
// OUTF (2-byte epilogues only): 0 = bf16, 1 = IEEE half, 2 = value split into hi + lo (both bf16) written as [lo | hi]
// planes N columns apart (the "bf16x2" arithmetic mode, include/peekvit_b200.h pk_out_format).
template <int BN, int EPI, bool RED, int EW = 8, int LN = 0, int OUTF = 0>
struct PairCfg {
  static constexpr int kEpiWarps = EW;
  static constexpr int kThreads = 128 + EW * 32;
  static constexpr int kParts = EW / 4;                                       // column parts of the tile
  static constexpr int kPartCols = BN / kParts;
  static constexpr bool kOutBf16 = (EPI == PK_EPI_BIAS_BF16 || EPI == PK_EPI_BIAS_GELU_BF16);
  static constexpr bool kResid = (EPI == PK_EPI_BIAS_RESID_F32) && !RED;
  static constexpr int kUnits = kPartCols / 32;                                // 32-column accumulator units per warp
  static_assert(kPartCols % 32 == 0, "tile width must split into 32-column units per epilogue warp");
  static constexpr int kUnitsPerStore = (kOutBf16 && kUnits % 2 == 0) ? 2 : 1; // bf16: 64 columns = one 128-byte row
  static constexpr int kChunks = kUnits / kUnitsPerStore;                     // TMA stores per warp per tile
  static constexpr int kStoreCols = 32 * kUnitsPerStore;
  static constexpr int kStoreSwizzle = kOutBf16 ? (kUnitsPerStore == 2 ? 128 : 64) : 128;
  static constexpr int kBufs = EW == 16 ? 1 : (kResid ? 3 : 2);               // staging buffers per epilogue warp
  static_assert(EW != 16 || !kResid, "the staged-residual epilogue needs its 3-deep ring");
  static_assert(OUTF == 0 || (kOutBf16 && EW == 8 && LN == 0), "fp16 / split outputs are variants of the plain 2-byte epilogues");
  static constexpr int kABytes = kP_BM * kP_BK * 2;
  static constexpr int kBBytes = (BN / 2) * kP_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlotBytes = kP_BufBytes;
  // LN producer: two 32x32 bf16 tiles per warp for the bf16 row copy.  They sit behind the residual ring instead of widening
  // its three slots: 64 KB of staging instead of 72 KB is what lets the main loop keep 5 operand stages (fc2 streams its
  // 620 MB A operand from DRAM; with 4 stages the tensor pipe idled a third of the time).
  static constexpr int kXbBytes = LN == 1 ? 2 * 2048 : 0;
  static constexpr int kWarpStagingBytes = kBufs * kSlotBytes + kXbBytes;
  static constexpr int kStagingBytes = EW * kWarpStagingBytes;
  static_assert(LN != 1 || (kResid && EW == 4), "the LN producer is the staged-residual epilogue with one warp per lane quarter");
  static_assert(LN != 2 || kOutBf16, "the LN consumer is a bf16 epilogue");
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = 232448 - 1024 - kBarBytes - kStagingBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
  static_assert(kStages >= 3, "too few pipeline stages");
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

struct PairParams {
  int M, N, K;
  const int* m_dev;
  const int* row_begin_dev;
  const int* out_row_index;  // RED only: GEMM row r accumulates into out row out_row_index[r] (un-permute of expert-sorted rows)
  const int* group_offsets;  // grouped launch: row segments [off[g], off[g + 1]) with weight rows g * N .. (g + 1) * N - 1
  int n_groups;
  const float* bias;
  void* out;
  long long ldo;
  const float* rowscale;
  unsigned int* flag;
  int l2_prefetch;
  // LayerNorm producer (LN == 1)
  void* xb_out;            // bf16 copy of the result rows (partial-tile path; full tiles go through tmap_xb)
  long long ldxb;
  float* row_stats;        // [rows][stat_parts][2]: partial (sum, sum of squares) of this launch's column tile
  int stat_parts;
  // LayerNorm consumer (LN == 2)
  const float* ln_stats;   // [rows][ln_parts][2]
  const float* ln_c1;      // [N]: sum_k W'[n,k]
  int ln_parts;
  float ln_inv_dim, ln_eps;
  int debug;      // PK_GEMM_DEBUG bits (timing experiments only): 1 = no operand loads, 2 = no epilogue math/stores
  int a_wrap_k;   // two-term split A operand stored as [lo | hi]: K index c >= 2*a_wrap_k reads A column c - a_wrap_k (0 = off)
};

// ---------------------------------------------------------------- cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into this CTA's smem; the byte count is credited to `bar` (a shared::cluster address:
// the leader CTA's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, both CTAs] * B[smem, both CTAs]^T ; issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the barrier at this smem offset in BOTH CTAs once all MMAs issued so far have completed.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }

template <int BN, int EPI, bool RED, int EW, int LN, int OUTF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + EW * 32, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                      const __grid_constant__ CUtensorMap tmap_xb, const PairParams p) {
  using Cfg = PairCfg<BN, EPI, RED, EW, LN, OUTF>;
  constexpr int kP_EpiWarps = EW;
  constexpr bool kOutBf16 = Cfg::kOutBf16;
  constexpr bool kResid = Cfg::kResid;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ab = smem;
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]   (the leader's copy is the live one)
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]   per CTA, arrived by the multicast commit
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]         per CTA, arrived by the multicast commit
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]         leader's copy: 2 x kP_EpiWarps arrivals
  uint64_t* res_bar = tempty_bar + 2;              // [kP_EpiWarps][3] residual-tile loads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 3 * kP_EpiWarps);

  const int warp = warp_id();
  const int lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int row0 = p.row_begin_dev ? *p.row_begin_dev : 0;     // first row of this launch's segment (MoE expert segments)
  const int M = p.m_dev ? max(0, min(*p.m_dev, p.M - row0)) : p.M - row0;
  const int n_tiles = (p.N + BN - 1) / BN;
  int m_tiles = (M + 2 * kP_BM - 1) / (2 * kP_BM);
  if (p.group_offsets) {                           // grouped launch: the segments' row tiles, one after the other
    m_tiles = 0;
    for (int g = 0; g < p.n_groups; ++g) m_tiles += (p.group_offsets[g + 1] - p.group_offsets[g] + 2 * kP_BM - 1) / (2 * kP_BM);
  }
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / kP_BK;
  // tile t -> first row of its 256-row block, end of the rows it may write, column block, first weight / bias row
  auto tile_map = [&](int t, int& row_blk, int& row_end, int& n_blk, int& w_row0) {
    int mt = t / n_tiles;
    n_blk = t - mt * n_tiles;
    if (!p.group_offsets) {
      row_blk = row0 + mt * 2 * kP_BM;
      row_end = row0 + M;
      w_row0 = 0;
      return;
    }
    int g = 0, o0 = p.group_offsets[0], o1 = p.group_offsets[1];
    for (;;) {
      const int tg = (o1 - o0 + 2 * kP_BM - 1) / (2 * kP_BM);
      if (mt < tg || g + 1 >= p.n_groups) break;
      mt -= tg;
      ++g;
      o0 = o1;
      o1 = p.group_offsets[g + 1];
    }
    row_blk = o0 + mt * 2 * kP_BM;
    row_end = o1;
    w_row0 = g * p.N;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if constexpr (kResid) tma_prefetch_desc(&tmap_res);
    if constexpr (LN == 1) tma_prefetch_desc(&tmap_xb);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 2 * kP_EpiWarps);
    }
    for (int i = 0; i < 3 * kP_EpiWarps; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tcgen05_fence_before();
  cluster_sync_all();                 // both CTAs' barriers and TMEM are ready before any cross-CTA signal
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // The whole warp runs the loop (warp-uniform control flow keeps addresses and coordinates in
    // uniform registers); one elected lane issues.
    int s = 0;
    uint32_t ph = 0;
    const int a_row_off = static_cast<int>(rank) * kP_BM;
    const int b_row_off = static_cast<int>(rank) * (BN / 2);
    // L2 prefetch runs kPrefetch K blocks ahead of the smem ring (across tile boundaries): operands
    // streamed from HBM arrive in L2 before their TMA load is issued, so the ring only has to cover
    // L2 latency.
    constexpr int kPrefetch = Cfg::kStages + 4;
    int pf_t = pair, pf_kb = 0;
    auto prefetch_next = [&]() {        // elected lane
      if (pf_t < num_tiles) {
        int pf_col = pf_kb * kP_BK;
        if (p.a_wrap_k > 0 && pf_col >= 2 * p.a_wrap_k) pf_col -= p.a_wrap_k;
        tma_prefetch_l2_2d(&tmap_a, pf_col, row0 + (pf_t / n_tiles) * 2 * kP_BM + a_row_off);
        tma_prefetch_l2_2d(&tmap_b, pf_kb * kP_BK, (pf_t % n_tiles) * BN + b_row_off);
      }
    };
    auto prefetch_advance = [&]() {     // whole warp (keeps the iterator warp-uniform)
      if (++pf_kb == num_kb) { pf_kb = 0; pf_t += num_pairs; }
    };
    if (p.l2_prefetch) {
      for (int i = 0; i < kPrefetch; ++i) {
        if (elect_one()) prefetch_next();
        __syncwarp();
        prefetch_advance();
      }
    }
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int row_blk, row_end_unused, n_blk, w_row0;
      tile_map(t, row_blk, row_end_unused, n_blk, w_row0);
      bool ok = true;
      for (int kb = 0; kb < num_kb; ++kb) {
        ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u, p.flag, 0x1100u + s));
        if (!ok) break;
        if (elect_one()) {
          const uint32_t fb_local = smem_u32(&full_bar[s]);
          if (p.debug & 1) {
            if (rank == 0) mbar_arrive(fb_local);
          } else {
          if (rank == 0) mbar_expect_tx(fb_local, 2 * Cfg::kStageBytes);   // both CTAs' bytes land on the leader's barrier
          const uint32_t fb = mapa_u32(fb_local, 0);
          const uint32_t a_dst = smem_u32(smem_ab + s * Cfg::kStageBytes);
          int a_col = kb * kP_BK;
          if (p.a_wrap_k > 0 && a_col >= 2 * p.a_wrap_k) a_col -= p.a_wrap_k;       // [lo | hi] read as lo, hi, hi
          tma_load_2d_pair(a_dst, &tmap_a, fb, a_col, row_blk + a_row_off);
          tma_load_2d_pair(a_dst + Cfg::kABytes, &tmap_b, fb, kb * kP_BK, w_row0 + n_blk * BN + b_row_off);
          }
          if (p.l2_prefetch) prefetch_next();
        }
        __syncwarp();
        if (p.l2_prefetch) prefetch_advance();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
      }
      if (!ok) break;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // Warp-uniform loop, one elected lane issues: a divergent single-lane loop makes the compiler
    // shuttle every descriptor through ELECT / R2UR.BROADCAST sequences (~160 instructions per K
    // block, measured 760 cycles against 512 cycles of tensor work).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kP_BM, BN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&tempty_bar[as]), aph ^ 1u, p.flag, 0x1200u + as))) break;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        bool ok = true;
        for (int kb = 0; kb < num_kb; ++kb) {
          ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&full_bar[s]), ph, p.flag, 0x1300u + s));
          if (!ok) break;
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem_ab + s * Cfg::kStageBytes);
            const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
            const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kP_BK / 16; ++k) {
              umma_bf16_pair(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                             (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(smem_u32(&empty_bar[s]));          // frees the slot in both CTAs
            if (kb == num_kb - 1) umma_commit_pair(smem_u32(&tfull_bar[as]));
          }
          __syncwarp();
          if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
        }
        if (!ok) break;
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int ew = warp - 4;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;               // which part of the tile's columns (kParts of them)
    constexpr int kPartCols = Cfg::kPartCols;
    constexpr int kUnits = Cfg::kUnits;
    constexpr int kUPS = Cfg::kUnitsPerStore;
    constexpr int kChunks = Cfg::kChunks;
    constexpr int kBufs = Cfg::kBufs;
    constexpr int kSlot = Cfg::kSlotBytes;
    uint8_t* ebuf = staging + ew * Cfg::kWarpStagingBytes;               // 1024-byte aligned
    const uint32_t ebuf_u32 = smem_u32(ebuf);
    uint64_t* rbar = res_bar + 3 * ew;
    const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
    const uint32_t total_chunks = static_cast<uint32_t>(my_tiles) * kChunks;
    const int row_in_pair = static_cast<int>(rank) * kP_BM + q * 32;
    // residual tile of chunk `ch` of tile `tt` -> staging buffer g % 3   (lane 0 only)
    auto issue_resid = [&](uint32_t g, int tt, int ch) {
      const uint32_t b = g % 3u;
      const uint32_t bar = smem_u32(&rbar[b]);
      mbar_expect_tx(bar, kP_BufBytes);
      tma_load_2d(ebuf_u32 + b * kSlot, &tmap_res, bar, (tt % n_tiles) * BN + half * kPartCols + ch * Cfg::kStoreCols,
                  row0 + (tt / n_tiles) * 2 * kP_BM + row_in_pair);
    };
    if constexpr (kResid) {
      if (lane == 0 && !(p.debug & 2)) {
        if (total_chunks > 0) issue_resid(0, pair, 0);
        if (total_chunks > 1) issue_resid(1, pair, 1);
      }
    }
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
    uint32_t g = 0;                           // running chunk counter of this warp
    int as = 0;
    uint32_t aph = 0;
    bool ok = true;
    for (int t = pair; t < num_tiles && ok; t += num_pairs) {
      int row_blk, row_end, n_blk_t, w_row0;                                               // rows >= row_end stay untouched
      tile_map(t, row_blk, row_end, n_blk_t, w_row0);
      const int row_base = row_blk + row_in_pair;                                          // this warp's first row
      const int grow = row_base + lane;
      const float* bias_g = p.bias ? p.bias + w_row0 : nullptr;                            // this group's bias
      // warp-uniform; a row-indexed launch (MoE un-permute) writes every tile row by row
      const bool full_tile = row_base + 32 <= row_end && !(RED && p.out_row_index);
      float sc = 1.0f;
      if constexpr (kResid || RED) {
        if (p.rowscale && grow < row_end) sc = p.rowscale[grow];
      }
      float ln_alpha = 1.0f, ln_beta = 0.0f;       // LN consumer: out = alpha*acc + (beta*c1[n] + c2[n])
      if constexpr (LN == 2) {
        if (grow < row_end) {
          float s1 = 0.f, s2 = 0.f;
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(grow) * p.ln_parts;
          for (int i = 0; i < p.ln_parts; ++i) {
            const float2 v2 = st[i];
            s1 += v2.x;
            s2 += v2.y;
          }
          const float mean = s1 * p.ln_inv_dim;
          const float var = fmaxf(s2 * p.ln_inv_dim - mean * mean, 0.f);
          ln_alpha = 1.0f / sqrtf(var + p.ln_eps);
          ln_beta = -ln_alpha * mean;
        }
      }
      float st_sum = 0.f, st_sq = 0.f;             // LN producer: this row's partial statistics over the tile's columns
      if (!mbar_wait(smem_u32(&tfull_bar[as]), aph, p.flag, 0x1400u + as)) { ok = false; break; }
      tcgen05_fence_after();
      const uint32_t t_col0 = t_lane + static_cast<uint32_t>(as * BN + half * kPartCols);
      uint32_t va[32], vb[32];
      tmem_ld_32x32(t_col0, va);
#pragma unroll
      for (int u = 0; u < kUnits; ++u) {
        uint32_t (&v)[32] = (u & 1) ? vb : va;
        uint32_t (&vn)[32] = (u & 1) ? va : vb;
        const int ch = u / kUPS;              // store chunk inside the tile
        const int sub = u % kUPS;             // unit inside the store chunk
        // split output: the two staging buffers hold the hi and the lo tile of the SAME chunk (no double buffering)
        const uint32_t b = (kBufs == 1 || OUTF == 2) ? 0u : (kResid ? (g % 3u) : (g & 1u));
        uint8_t* bufp = ebuf + b * kSlot;
        const int col0 = (t % n_tiles) * BN + half * kPartCols + u * 32;     // first output column of this unit
        float4 bias4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bias4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias_g && col0 + 4 * j < p.N) bias4[j] = __ldg(reinterpret_cast<const float4*>(bias_g + col0 + 4 * j));
          if constexpr (LN == 2) {
            if (col0 + 4 * j < p.N) {
              const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.ln_c1 + col0 + 4 * j));
              bias4[j].x = fmaf(c1.x, ln_beta, bias4[j].x); bias4[j].y = fmaf(c1.y, ln_beta, bias4[j].y);
              bias4[j].z = fmaf(c1.z, ln_beta, bias4[j].z); bias4[j].w = fmaf(c1.w, ln_beta, bias4[j].w);
            }
          }
        }
        tmem_ld_wait();                                                     // unit u is in registers
        if (u + 1 < kUnits) {
          tmem_ld_32x32(t_col0 + static_cast<uint32_t>((u + 1) * 32), vn);  // overlaps the math below
        } else {
          // every TMEM read of this accumulator stage is done: hand it back to the leader's MMA thread
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(as == 0 ? tempty_leader0 : tempty_leader1);
        }
        if (p.debug & 2) continue;
        if (sub == 0) {
          if constexpr (kResid) {
            if (!mbar_wait(smem_u32(&rbar[b]), (g / 3u) & 1u, p.flag, 0x1500u + ew)) { ok = false; break; }
          } else {
            // the store that last used this buffer (two chunks ago; the previous one with a single buffer) has drained
            if (lane == 0) { if constexpr (kBufs == 1 || OUTF == 2) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
            __syncwarp();
          }
        }
        uint32_t xbp[LN == 1 ? 16 : 1];
        uint32_t lo2[OUTF == 2 ? 16 : 1];         // lo plane of the split output (v holds the hi plane)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = bias4[j];
          float a0, a1, a2, a3;
          if constexpr (LN == 2) {
            a0 = fmaf(__uint_as_float(v[4 * j]), ln_alpha, b4.x); a1 = fmaf(__uint_as_float(v[4 * j + 1]), ln_alpha, b4.y);
            a2 = fmaf(__uint_as_float(v[4 * j + 2]), ln_alpha, b4.z); a3 = fmaf(__uint_as_float(v[4 * j + 3]), ln_alpha, b4.w);
          } else {
            f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), f2_pack(b4.x, b4.y)), a0, a1);
            f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), f2_pack(b4.z, b4.w)), a2, a3);
          }
          if constexpr (EPI == PK_EPI_BIAS_GELU_BF16) {
            gelu_erf_x2(a0, a1);
            gelu_erf_x2(a2, a3);
          }
          if constexpr (kResid) {
            // 128-byte rows, 16-byte chunk j of row r lives at chunk j ^ (r & 7) (SWIZZLE_128B)
            const float4 r4 = *reinterpret_cast<const float4*>(bufp + lane * 128 + ((j ^ (lane & 7)) << 4));
            a0 = fmaf(a0, sc, r4.x); a1 = fmaf(a1, sc, r4.y); a2 = fmaf(a2, sc, r4.z); a3 = fmaf(a3, sc, r4.w);
          }
          if constexpr (RED) {
            a0 *= sc; a1 *= sc; a2 *= sc; a3 *= sc;
          }
          if constexpr (LN == 1) {
            if (col0 + 4 * j < p.N) {
              st_sum += (a0 + a1) + (a2 + a3);
              st_sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            }
            xbp[2 * j] = pack_bf16(a0, a1);
            xbp[2 * j + 1] = pack_bf16(a2, a3);
          }
          if constexpr (kOutBf16) {
            if constexpr (OUTF == 1) {
              v[2 * j] = pack_f16(a0, a1);
              v[2 * j + 1] = pack_f16(a2, a3);
            } else if constexpr (OUTF == 2) {
              split2_pack(a0, a1, v[2 * j], lo2[2 * j]);
              split2_pack(a2, a3, v[2 * j + 1], lo2[2 * j + 1]);
            } else {
              v[2 * j] = pack_bf16(a0, a1);
              v[2 * j + 1] = pack_bf16(a2, a3);
            }
          } else {
            v[4 * j] = __float_as_uint(a0); v[4 * j + 1] = __float_as_uint(a1);
            v[4 * j + 2] = __float_as_uint(a2); v[4 * j + 3] = __float_as_uint(a3);
          }
        }
        if (full_tile) {
          if constexpr (kOutBf16) {
            if constexpr (kUPS == 2) {
              // 128-byte rows of 64 bf16: this unit fills 16-byte chunks sub*4 .. sub*4+3 (SWIZZLE_128B)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(bufp + lane * 128 + (((sub * 4 + j) ^ (lane & 7)) << 4)) =
                    make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if constexpr (OUTF == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(bufp + kSlot + lane * 128 + (((sub * 4 + j) ^ (lane & 7)) << 4)) =
                      make_uint4(lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
              }
            } else {
              // 64-byte rows, 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3) (SWIZZLE_64B)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(bufp + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if constexpr (OUTF == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(bufp + kSlot + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                      make_uint4(lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(bufp + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if constexpr (LN == 1) {
              // bf16 copy: 64-byte rows, 16-byte chunk j of row r at chunk j ^ ((r >> 1) & 3) (SWIZZLE_64B)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(ebuf + kBufs * kSlot + (g & 1u) * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(xbp[4 * j], xbp[4 * j + 1], xbp[4 * j + 2], xbp[4 * j + 3]);
            }
          }
          if (sub == kUPS - 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if constexpr (OUTF == 2) {
                // hi plane N columns after the lo plane; a chunk that starts at or past N lies entirely outside (N % 64 == 0)
                if (col0 - sub * 32 < p.N) {
                  tma_store_2d(&tmap_out, ebuf_u32, p.N + col0 - sub * 32, row_base);
                  tma_store_2d(&tmap_out, ebuf_u32 + kSlot, col0 - sub * 32, row_base);
                }
              } else if constexpr (RED) tma_reduce_add_2d(&tmap_out, ebuf_u32 + b * kSlot, col0 - sub * 32, row_base);
              else tma_store_2d(&tmap_out, ebuf_u32 + b * kSlot, col0 - sub * 32, row_base);
              if constexpr (LN == 1) tma_store_2d(&tmap_xb, ebuf_u32 + kBufs * kSlot + (g & 1u) * 2048, col0, row_base);
              bulk_commit();
            }
          }
        } else {
          // partial row tile: rows >= M must stay untouched -> predicated row stores from registers.
          // An empty bulk group keeps the one-group-per-chunk accounting of the buffer ring.
          if (sub == kUPS - 1 && lane == 0) bulk_commit();
        }
        if (!full_tile && grow < row_end) {
          if constexpr (kOutBf16) {
            __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(grow) * p.ldo + col0 + (OUTF == 2 ? p.N : 0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col0 + 8 * j < p.N) *reinterpret_cast<uint4*>(orow + 8 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if constexpr (OUTF == 2) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + 8 * j < p.N) *reinterpret_cast<uint4*>(orow - p.N + 8 * j) = make_uint4(lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
            }
          } else {
            long long orow_i = grow;
            if constexpr (RED) {
              if (p.out_row_index) orow_i = p.out_row_index[grow];
            }
            float* orow = static_cast<float*>(p.out) + orow_i * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col0 + 4 * j < p.N) {
                float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
                if constexpr (RED) {
                  if (p.out_row_index) {
                    // fire-and-forget vector reduction at L2 (no load round trip in the epilogue warp); the index is a
                    // permutation, so no two rows collide and the sum is order-independent
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + 4 * j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                                 : "memory");
                    continue;
                  }
                  // each element belongs to exactly one thread: plain read-modify-write
                  const float4 r = *reinterpret_cast<const float4*>(orow + 4 * j);
                  o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(orow + 4 * j) = o;
              }
            if constexpr (LN == 1) {
              __nv_bfloat16* xrow = static_cast<__nv_bfloat16*>(p.xb_out) + static_cast<long long>(grow) * p.ldxb + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + 8 * j < p.N) *reinterpret_cast<uint4*>(xrow + 8 * j) = make_uint4(xbp[4 * j], xbp[4 * j + 1], xbp[4 * j + 2], xbp[4 * j + 3]);
            }
          }
        }
        if (sub == kUPS - 1) {
          if constexpr (kResid) {
            // buffer (g+2)%3 was last read by the store of chunk g-1: one chunk of work ago
            if (lane == 0 && (LN == 1 || g + 2 < total_chunks)) {
              bulk_wait_read<1>();
              if (g + 2 < total_chunks) {
                if (ch + 2 < kChunks) issue_resid(g + 2, t, ch + 2);
                else issue_resid(g + 2, t + num_pairs, ch + 2 - kChunks);
              }
            }
            // LN producer: the bf16 tile of chunk g+1 is the one chunk g-1 stored from; its drain (just waited for by lane 0)
            // must be visible to every lane before they write it
            if constexpr (LN == 1) __syncwarp();
          }
          ++g;
        }
      }
      if constexpr (LN == 1) {
        // this row's (sum, sum of squares) over the tile's BN columns -> slot n_blk of its statistics record
        if (ok && grow < row_end)
          reinterpret_cast<float2*>(p.row_stats)[static_cast<long long>(grow) * p.stat_parts + (t % n_tiles)] = make_float2(st_sum, st_sq);
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (lane == 0) bulk_wait_read<0>();     // smem must outlive the last stores' reads
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();                       // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// PK_GEMM_L2_PREFETCH=1 enables the producer's L2 prefetch (measured slower on B200: 1359 -> 1236 TFLOP/s on the
// in-proj shape; kept for experiments).
static int pair_l2_prefetch() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_L2_PREFETCH");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}
// PK_GEMM_TMA_REDUCE=0 keeps the staged-residual epilogue for in-place residual GEMMs (A/B experiments).
static int pair_tma_reduce() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_TMA_REDUCE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
// PK_GEMM_GELU_WARPS=16 gives the GELU epilogue four warps per scheduler (measured slower on B200: fc1 1234 -> 1166 TFLOP/s;
// kept for experiments).
static int pair_gelu_warps() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_GELU_WARPS");
    v = (e && atoi(e) == 16) ? 16 : 8;
  }
  return v;
}
static int pair_debug() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int BN, int EPI, bool RED = false, int EW = 8, int LN = 0, int OUTF = 0>
static int launch_pair(const pk_gemm_args* a, cudaStream_t stream) {
  using Cfg = PairCfg<BN, EPI, RED, EW, LN, OUTF>;
  static bool attr_set = false;
  auto kfn = gemm_bf16_pair_kernel<BN, EPI, RED, EW, LN, OUTF>;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap ta, tb, tout, tres;
  const int a_cols = a->a_wrap_k > 0 ? 2 * a->a_wrap_k : a->K;          // [lo | hi] is stored once, read as lo, hi, hi
  int rc = make_tmap_bf16_2d(&ta, a->A, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a_cols), static_cast<uint64_t>(a->lda),
                             kP_BM, kP_BK);
  if (rc != PK_OK) return rc;
  const uint64_t w_rows = static_cast<uint64_t>(a->N) * (a->group_offsets ? a->n_groups : 1);       // stacked group weights
  rc = make_tmap_bf16_2d(&tb, a->W, w_rows, static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->ldw), BN / 2, kP_BK);
  if (rc != PK_OK) return rc;
  rc = make_tmap_2d(&tout, a->out, Cfg::kOutBf16 ? 2 : 4, static_cast<uint64_t>(a->M), static_cast<uint64_t>(OUTF == 2 ? 2 * a->N : a->N),
                    static_cast<uint64_t>(a->ldo), 32, Cfg::kStoreCols, Cfg::kStoreSwizzle);
  if (rc != PK_OK) return rc;
  tres = tout;
  if (Cfg::kResid) {
    rc = make_tmap_2d(&tres, a->resid, 4, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->ldr), 32,
                      32, 128);
    if (rc != PK_OK) return rc;
  }
  CUtensorMap txb = tout;
  if (LN == 1) {
    rc = make_tmap_2d(&txb, a->xb_out, 2, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->ldxb), 32,
                      32, 64);
    if (rc != PK_OK) return rc;
  }
  PairParams p;
  p.xb_out = a->xb_out; p.ldxb = a->ldxb; p.row_stats = a->row_stats;
  p.stat_parts = (a->N + BN - 1) / BN;
  p.ln_stats = a->ln_stats; p.ln_c1 = a->ln_c1; p.ln_parts = a->ln_parts;
  p.ln_inv_dim = a->ln_dim > 0 ? 1.0f / static_cast<float>(a->ln_dim) : 0.f;
  p.ln_eps = a->ln_eps;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_dev = a->m_dev;
  p.row_begin_dev = a->row_begin_dev;
  p.out_row_index = a->out_row_index;
  p.group_offsets = a->group_offsets;
  p.n_groups = a->n_groups;
  p.bias = a->bias;
  p.out = a->out; p.ldo = a->ldo;
  p.rowscale = a->rowscale;
  p.flag = device_flag_ptr();
  p.l2_prefetch = a->group_offsets ? 0 : pair_l2_prefetch();
  p.debug = pair_debug();
  p.a_wrap_k = a->a_wrap_k;
  // grouped launch: every segment may end in a partial row tile
  const int m_tiles = (a->M + 2 * kP_BM - 1) / (2 * kP_BM) + (a->group_offsets ? a->n_groups : 0), n_tiles = (a->N + BN - 1) / BN;
  int pairs = m_tiles * n_tiles;
  const int sms = a->max_ctas > 0 ? a->max_ctas : num_sms();
  if (pairs > sms / 2) pairs = sms / 2;
  if (pairs < 1) pairs = 1;
  kfn<<<2 * pairs, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tout, tres, txb, p);
  return check_cuda(cudaGetLastError(), "gemm_bf16_pair_kernel launch");
}

template <int BN>
static int dispatch_pair_epi(const pk_gemm_args* a, cudaStream_t stream) {
  switch (a->epilogue) {
    case PK_EPI_BIAS_BF16:
      if (a->out_format == PK_OUT_F16) return launch_pair<BN, PK_EPI_BIAS_BF16, false, 8, 0, 1>(a, stream);
      if (a->out_format == PK_OUT_BF16X2) return launch_pair<BN, PK_EPI_BIAS_BF16, false, 8, 0, 2>(a, stream);
      if (a->ln_stats) return launch_pair<BN, PK_EPI_BIAS_BF16, false, 8, 2>(a, stream);
      return launch_pair<BN, PK_EPI_BIAS_BF16>(a, stream);
    case PK_EPI_BIAS_GELU_BF16:
      if (a->out_format == PK_OUT_BF16X2) return launch_pair<BN, PK_EPI_BIAS_GELU_BF16, false, 8, 0, 2>(a, stream);
      if (a->ln_stats) return launch_pair<BN, PK_EPI_BIAS_GELU_BF16, false, 8, 2>(a, stream);
      if constexpr (BN != 192) {
        if (pair_gelu_warps() == 16) return launch_pair<BN, PK_EPI_BIAS_GELU_BF16, false, 16>(a, stream);
      }
      return launch_pair<BN, PK_EPI_BIAS_GELU_BF16>(a, stream);
    case PK_EPI_BIAS_RESID_F32:
      if (a->xb_out) return launch_pair<BN, PK_EPI_BIAS_RESID_F32, false, 4, 1>(a, stream);
      // in-place residual (x += ...): let the TMA reduction do the add at L2
      if (a->resid == a->out && a->ldr == a->ldo && pair_tma_reduce()) return launch_pair<BN, PK_EPI_BIAS_RESID_F32, true>(a, stream);
      return launch_pair<BN, PK_EPI_BIAS_RESID_F32>(a, stream);
    case PK_EPI_BIAS_F32: return launch_pair<BN, PK_EPI_BIAS_F32>(a, stream);
  }
  set_last_error("pk_gemm_bf16: unknown epilogue %d", a->epilogue);
  return PK_ERR_INVALID;
}

// Pair tile width: the widest of {256, 192, 128} that wastes the fewest padded columns.
static int pick_pair_block_n(int N) {
  int best = 128;
  long long best_cost = -1;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long padded = static_cast<long long>((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || padded < best_cost) { best_cost = padded; best = bn; }
  }
  return best;
}

// The pair kernel needs plain row mapping (TMA tile stores), 16-byte aligned rows and N % 8 == 0.
bool pair_gemm_eligible(const pk_gemm_args* a) {
  if (a->epilogue_mode == 2) return false;
  if (a->rows_per_group > 0) return false;
  if (a->group_offsets) {
    // grouped launch: no staged residual / LayerNorm variants, and a group's weight rows must tile by the column block
    const bool staged = a->epilogue == PK_EPI_BIAS_RESID_F32 && !(a->resid == a->out && a->ldr == a->ldo && pair_tma_reduce());
    if (a->n_groups < 1 || a->n_groups > 64 || staged || a->xb_out || a->ln_stats || a->rowscale) return false;
  }
  // row-indexed output: only as the in-place accumulate x[idx[r]] += ... (the TMA-reduce variant's row-store path)
  if (a->out_row_index && !(a->epilogue == PK_EPI_BIAS_RESID_F32 && a->resid == a->out && a->ldr == a->ldo && !a->xb_out &&
                            pair_tma_reduce()))
    return false;
  const bool bf = a->epilogue == PK_EPI_BIAS_BF16 || a->epilogue == PK_EPI_BIAS_GELU_BF16;
  const long long eb = bf ? 2 : 4;
  if ((a->ldo * eb) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) return false;
  if (a->N % 8 != 0) return false;
  if (a->epilogue == PK_EPI_BIAS_RESID_F32 && ((a->ldr * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->resid) & 15) != 0)) return false;
  return true;
}

// Number of per-row statistics slots the LN-producer epilogue writes for an N-column output (= column tiles).
int pair_row_stat_parts(int N) {
  const int bn = pick_pair_block_n(N);
  return (N + bn - 1) / bn;
}

int launch_pair_gemm(const pk_gemm_args* a, cudaStream_t stream) {
  const int bn = a->block_n > 0 ? a->block_n : pick_pair_block_n(a->N);
  switch (bn) {
    case 128: return dispatch_pair_epi<128>(a, stream);
    case 192: return dispatch_pair_epi<192>(a, stream);
    case 256: return dispatch_pair_epi<256>(a, stream);
  }
  set_last_error("pk_gemm_bf16: unsupported block_n %d", bn);
  return PK_ERR_INVALID;
}

}  // namespace pk
