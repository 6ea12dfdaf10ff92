// peekvit_b200 — tcgen05/TMEM/TMA GEMM with fused epilogues (sm_100a).
//
//   C[M,N] = A[M,K] (bf16, row-major) x W[N,K]^T (bf16, row-major = nn.Linear layout), f32 accumulate
//
// replaces the reference's cuBLAS/cuDNN call sites on the encoder forward path:
//   K1 conv_proj patch GEMM + bias + pos_embedding      (reference models/vit.py:212, :92)
//   K3 MHA packed in-projection + bias                   (models/blocks.py:94 -> in_proj)
//   K5 MHA out-projection + bias + residual              (models/vit.py:49-51)
//   K6 fc1 + bias + exact GELU                           (models/blocks.py:81-82)
//   K7 fc2 + bias + residual                             (models/blocks.py:83, vit.py:55)
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: A tile 128x64 and W tile BNx64 (128-byte swizzle) into a smem ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma 128xBNx16, accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue, two flavours:
//               TMA  (default): tcgen05.ld (thread = row) -> bias / GELU / row-scale / residual in
//                    registers -> 128B/64B-swizzled smem tile -> cp.async.bulk.tensor store; the fp32
//                    residual tile is TMA-loaded into the same smem one chunk ahead (no register or
//                    scoreboard pressure, fully coalesced, asynchronous both ways);
//               SIMT (row remap / row segments / scattered rows): padded smem transpose -> 16-byte
//                    coalesced global accesses with per-row predication
// Barriers: full[s]/empty[s] (TMA <-> MMA), tmem_full[a]/tmem_empty[a] (MMA <-> epilogue).
// Every mbarrier wait is bounded (pk_common.cuh) so a protocol bug cannot hang the GPU.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

namespace pk {

constexpr int kBM = 128;
constexpr int kBK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
constexpr int kStagePitch = 36;   // floats per staged row (32 + 4 pad): conflict-free v4 access

constexpr int kEpiBufBytes = 4096;  // one 32-row x 128-byte staging tile
template <int BN, bool TMA_EPI>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 3 : (BN == 192 ? 4 : 5);
  static constexpr int kStagingBytes = TMA_EPI ? kEpiWarps * 2 * kEpiBufBytes : kEpiWarps * 32 * kStagePitch * 4;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

struct GemmKernelParams {
  int M, N, K;
  const int* m_dev;
  const float* bias;
  void* out;
  long long ldo;
  const float* resid;
  long long ldr;
  const float* rowscale;
  int rows_per_group, group_stride, group_offset, resid_is_pos, pos_offset;
  const int* row_begin_dev;
  const int* out_row_index;
  unsigned int* flag;
};

template <int BN, int EPI, bool TMA_EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                         const GemmKernelParams p) {
  using Cfg = GemmCfg<BN, TMA_EPI>;
  constexpr bool kOutBf16 = (EPI == PK_EPI_BIAS_BF16 || EPI == PK_EPI_BIAS_GELU_BF16);
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ab = smem;
  float* staging = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* res_bar = tempty_bar + 2;              // [kEpiWarps][2] residual-tile loads (TMA epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

  const int warp = warp_id();
  const int lane = lane_id();
  const int row0 = p.row_begin_dev ? *p.row_begin_dev : 0;     // first row of this launch's segment
  const int M = p.m_dev ? max(0, min(*p.m_dev, p.M - row0)) : p.M;
  const int m_tiles = (M + kBM - 1) / kBM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (TMA_EPI) {
      tma_prefetch_desc(&tmap_out);
      if constexpr (EPI == PK_EPI_BIAS_RESID_F32) tma_prefetch_desc(&tmap_res);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), kEpiWarps);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (warp-uniform loop, one elected lane issues: see pk_gemm2.cu for why a lane-0 loop is slow)
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / n_tiles, n_blk = t % n_tiles;
      bool ok = true;
      for (int kb = 0; kb < num_kb; ++kb) {
        ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u, p.flag, 0x100u + s));
        if (!ok) break;
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, Cfg::kStageBytes);
          const uint32_t a_dst = smem_u32(smem_ab + s * Cfg::kStageBytes);
          tma_load_2d(a_dst, &tmap_a, fb, kb * kBK, row0 + m_blk * kBM);
          tma_load_2d(a_dst + Cfg::kABytes, &tmap_b, fb, kb * kBK, n_blk * BN);
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
      }
      if (!ok) break;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane)
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&tempty_bar[as]), aph ^ 1u, p.flag, 0x200u + as))) break;
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      bool ok = true;
      for (int kb = 0; kb < num_kb; ++kb) {
        ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&full_bar[s]), ph, p.flag, 0x300u + s));
        if (!ok) break;
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_ab + s * Cfg::kStageBytes);
          const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
          const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));            // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) umma_commit(smem_u32(&tfull_bar[as]));
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
      }
      if (!ok) break;
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 4;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;               // which half of the tile's columns
    constexpr int kChunksPerWarp = BN / 64; // 32-column chunks per warp
    if constexpr (TMA_EPI) {
      // ---------------- TMA epilogue: registers -> swizzled smem tile -> bulk tensor store
      uint8_t* ebuf = reinterpret_cast<uint8_t*>(staging) + ew * 2 * kEpiBufBytes;     // 1024-byte aligned
      const uint32_t ebuf_u32 = smem_u32(ebuf);
      uint64_t* rbar = res_bar + 2 * ew;
      uint32_t g = 0;                       // running chunk counter: buffer g&1, residual barrier parity (g>>1)&1
      int as = 0;
      uint32_t aph = 0;
      auto chunk_col = [&](int t, int ch) { return (t % n_tiles) * BN + half * (BN / 2) + ch * 32; };
      auto chunk_row = [&](int t) { return row0 + (t / n_tiles) * kBM + q * 32; };
      if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
        if (lane == 0 && static_cast<int>(blockIdx.x) < num_tiles) {
          mbar_expect_tx(smem_u32(&rbar[0]), kEpiBufBytes);
          tma_load_2d(ebuf_u32, &tmap_res, smem_u32(&rbar[0]), chunk_col(blockIdx.x, 0), chunk_row(blockIdx.x));
        }
      }
      bool ok = true;
      for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x) {
        const int grow = chunk_row(t) + lane;                 // this thread's output row
        float sc = 1.0f;
        if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
          if (p.rowscale && grow - row0 < M) sc = p.rowscale[grow];
        }
#pragma unroll 1
        for (int ch = 0; ch < kChunksPerWarp; ++ch) {
          const uint32_t buf = ebuf_u32 + (g & 1u) * kEpiBufBytes;
          uint8_t* bufp = ebuf + (g & 1u) * kEpiBufBytes;
          const int col0 = chunk_col(t, ch);
          if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
            // prefetch the next chunk's residual tile into the other buffer
            int nt = t, nch = ch + 1;
            if (nch == kChunksPerWarp) { nt = t + gridDim.x; nch = 0; }
            if (lane == 0 && nt < num_tiles) {
              bulk_wait_read<0>();          // the store that last used that buffer has drained its smem reads
              const uint32_t nb = smem_u32(&rbar[(g + 1u) & 1u]);
              mbar_expect_tx(nb, kEpiBufBytes);
              tma_load_2d(ebuf_u32 + ((g + 1u) & 1u) * kEpiBufBytes, &tmap_res, nb, chunk_col(nt, nch), chunk_row(nt));
            }
          } else {
            if (lane == 0) bulk_wait_read<1>();   // the store issued two chunks ago no longer reads this buffer
            __syncwarp();
          }
          if (ch == 0) {
            if (!mbar_wait(smem_u32(&tfull_bar[as]), aph, p.flag, 0x400u + as)) { ok = false; break; }
            tcgen05_fence_after();
          }
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
          uint32_t v[32];
          tmem_ld_32x32(t_row + static_cast<uint32_t>(half * (BN / 2) + ch * 32), v);
          tmem_ld_wait();
          if (ch == kChunksPerWarp - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
          }
          if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
            if (!mbar_wait(smem_u32(&rbar[g & 1u]), (g >> 1) & 1u, p.flag, 0x500u + ew)) { ok = false; break; }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && col0 + 4 * j < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
            float a0 = __uint_as_float(v[4 * j]) + b4.x, a1 = __uint_as_float(v[4 * j + 1]) + b4.y;
            float a2 = __uint_as_float(v[4 * j + 2]) + b4.z, a3 = __uint_as_float(v[4 * j + 3]) + b4.w;
            if constexpr (EPI == PK_EPI_BIAS_GELU_BF16) {
              gelu_erf_x2(a0, a1);
              gelu_erf_x2(a2, a3);
            }
            if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
              // 128-byte rows, 16-byte chunk j of row r lives at chunk j ^ (r & 7) (SWIZZLE_128B)
              const float4 r4 = *reinterpret_cast<const float4*>(bufp + lane * 128 + ((j ^ (lane & 7)) << 4));
              a0 = fmaf(a0, sc, r4.x); a1 = fmaf(a1, sc, r4.y); a2 = fmaf(a2, sc, r4.z); a3 = fmaf(a3, sc, r4.w);
            }
            if constexpr (kOutBf16) {
              v[2 * j] = pack_bf16(a0, a1);
              v[2 * j + 1] = pack_bf16(a2, a3);
            } else {
              v[4 * j] = __float_as_uint(a0); v[4 * j + 1] = __float_as_uint(a1);
              v[4 * j + 2] = __float_as_uint(a2); v[4 * j + 3] = __float_as_uint(a3);
            }
          }
          const bool full_tile = (t / n_tiles) * kBM + kBM <= M;      // warp-uniform
          if (full_tile) {
            if constexpr (kOutBf16) {
              // 64-byte rows, 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3) (SWIZZLE_64B)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(bufp + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(bufp + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_out, buf, col0, chunk_row(t));
              bulk_commit();
            }
          } else if (grow - row0 < M) {
            // last, partial row tile: rows >= M must stay untouched -> predicated row stores from registers
            if constexpr (kOutBf16) {
              __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(grow) * p.ldo + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + 8 * j < p.N) *reinterpret_cast<uint4*>(orow + 8 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              float* orow = static_cast<float*>(p.out) + static_cast<long long>(grow) * p.ldo + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col0 + 4 * j < p.N) *reinterpret_cast<uint4*>(orow + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          ++g;
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
      if (lane == 0) bulk_wait_read<0>();   // smem must outlive the last stores' reads
    } else {
      // ---------------- SIMT epilogue (per-row predication: remapped / segmented / scattered rows)
      float* stg = staging + ew * 32 * kStagePitch;
      const int c4 = lane & 7;                // phase 2: this lane's 4-column group inside the chunk
      const int r_sub = lane >> 3;            // phase 2: row inside each 4-row group
      int as = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / n_tiles, n_blk = t % n_tiles;
        const int row_base = m_blk * kBM + q * 32 + r_sub;       // + it*4 in phase 2
        // Residual operands do not depend on the accumulators: fetch the first chunk's rows before
        // blocking on the MMA, and chunk c+1's while chunk c is processed (keeps 8 independent
        // 16-byte loads in flight per lane instead of a load->add->store chain).
        float4 rnext[8];
        auto load_resid = [&](int ch, float4 (&dst)[8]) {
          if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
            const int gcol = n_blk * BN + half * (BN / 2) + ch * 32 + c4 * 4;
  #pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int lrow = row_base + it * 4;
              dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (lrow < M && gcol < p.N) {
                const int grow = row0 + lrow;
                long long rrow = p.out_row_index ? p.out_row_index[grow] : grow;
                if (p.rows_per_group > 0) {
                  const int g = grow / p.rows_per_group, pos = grow - g * p.rows_per_group;
                  rrow = p.resid_is_pos ? static_cast<long long>(p.pos_offset + pos)
                                        : static_cast<long long>(g) * p.group_stride + p.group_offset + pos;
                }
                dst[it] = *reinterpret_cast<const float4*>(p.resid + rrow * p.ldr + gcol);
              }
            }
          }
        };
        load_resid(0, rnext);
        if (!mbar_wait(smem_u32(&tfull_bar[as]), aph, p.flag, 0x400u + as)) break;
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
        // fully unrolled only where the residual double-buffer needs register renaming; the GELU
        // body is large and thrashes the instruction cache when replicated per chunk
  #pragma unroll (EPI == PK_EPI_BIAS_RESID_F32 ? kChunksPerWarp : 1)
        for (int ch = 0; ch < kChunksPerWarp; ++ch) {
          const int col0 = half * (BN / 2) + ch * 32;
          float4 rcur[8];
          if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
  #pragma unroll
            for (int it = 0; it < 8; ++it) rcur[it] = rnext[it];
            if (ch + 1 < kChunksPerWarp) load_resid(ch + 1, rnext);
          }
          uint32_t v[32];
          tmem_ld_32x32(t_row + static_cast<uint32_t>(col0), v);
          tmem_ld_wait();
          if (ch == kChunksPerWarp - 1) {
            // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
          }
          // phase 1: thread = row, 32 consecutive f32 -> padded smem
          float4* srow = reinterpret_cast<float4*>(stg + lane * kStagePitch);
  #pragma unroll
          for (int j = 0; j < 8; ++j)
            srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          __syncwarp();
          // phase 2: thread = 4 columns, 8 lanes cover one 128-B row segment, 4 rows per instruction
          const int gcol = n_blk * BN + col0 + c4 * 4;
          const bool col_ok = gcol < p.N;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + gcol);
  #pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + r_sub;
            const int lrow = row_base + it * 4;
            if (lrow < M && col_ok) {
              const int grow = row0 + lrow;
              float4 a = *reinterpret_cast<const float4*>(stg + rr * kStagePitch + c4 * 4);
              a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
              long long orow = p.out_row_index ? p.out_row_index[grow] : grow;
              if (p.rows_per_group > 0) {
                const int g = grow / p.rows_per_group, pos = grow - g * p.rows_per_group;
                orow = static_cast<long long>(g) * p.group_stride + p.group_offset + pos;
              }
              if constexpr (EPI == PK_EPI_BIAS_GELU_BF16) {
                a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
              }
              if constexpr (EPI == PK_EPI_BIAS_RESID_F32) {
                if (p.rowscale) {
                  const float sc = p.rowscale[grow];
                  a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
                }
                a.x += rcur[it].x; a.y += rcur[it].y; a.z += rcur[it].z; a.w += rcur[it].w;
              }
              if constexpr (EPI == PK_EPI_BIAS_BF16 || EPI == PK_EPI_BIAS_GELU_BF16) {
                uint2 o = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + gcol) = o;
              } else {
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + gcol) = a;
              }
            }
          }
          __syncwarp();
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI, bool TMA_EPI>
static int launch_gemm(const pk_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, TMA_EPI>;
  static bool attr_set = false;
  auto kfn = gemm_bf16_tcgen05_kernel<BN, EPI, TMA_EPI>;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  constexpr bool kOutBf16 = (EPI == PK_EPI_BIAS_BF16 || EPI == PK_EPI_BIAS_GELU_BF16);
  CUtensorMap ta, tb, tout, tres;
  int rc = make_tmap_bf16_2d(&ta, a->A, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->K),
                             static_cast<uint64_t>(a->lda), kBM, kBK);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tb, a->W, static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->K),
                         static_cast<uint64_t>(a->ldw), BN, kBK);
  if (rc != PK_OK) return rc;
  tout = ta;
  tres = ta;
  if (TMA_EPI) {
    rc = make_tmap_2d(&tout, a->out, kOutBf16 ? 2 : 4, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->N),
                      static_cast<uint64_t>(a->ldo), 32, 32, kOutBf16 ? 64 : 128);
    if (rc != PK_OK) return rc;
    if (EPI == PK_EPI_BIAS_RESID_F32) {
      rc = make_tmap_2d(&tres, a->resid, 4, static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->N),
                        static_cast<uint64_t>(a->ldr), 32, 32, 128);
      if (rc != PK_OK) return rc;
    }
  }
  GemmKernelParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_dev = a->m_dev;
  p.bias = a->bias;
  p.out = a->out; p.ldo = a->ldo;
  p.resid = a->resid; p.ldr = a->ldr;
  p.rowscale = a->rowscale;
  p.rows_per_group = a->rows_per_group; p.group_stride = a->group_stride;
  p.group_offset = a->group_offset; p.resid_is_pos = a->resid_is_pos; p.pos_offset = a->pos_offset;
  p.row_begin_dev = a->row_begin_dev; p.out_row_index = a->out_row_index;
  p.flag = device_flag_ptr();
  const int m_tiles = (a->M + kBM - 1) / kBM, n_tiles = (a->N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  const int sms = a->max_ctas > 0 ? a->max_ctas : num_sms();
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  kfn<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tout, tres, p);
  return check_cuda(cudaGetLastError(), "gemm_bf16_tcgen05_kernel launch");
}

// The TMA epilogue writes whole 32x32 boxes clipped only by the tensor bounds, so it is used when
// output rows are the GEMM rows themselves; remapped (patch embedding), segmented or scattered
// (MoE) rows need per-row predication and take the SIMT epilogue.
static bool can_use_tma_epilogue(const pk_gemm_args* a) {
  if (a->epilogue_mode == 2) return false;
  if (a->rows_per_group > 0 || a->row_begin_dev || a->out_row_index) return false;
  const bool bf = a->epilogue == PK_EPI_BIAS_BF16 || a->epilogue == PK_EPI_BIAS_GELU_BF16;
  const long long eb = bf ? 2 : 4;
  if ((a->ldo * eb) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) return false;
  if (bf && a->N % 8 != 0) return false;          // the partial-tile fallback stores 8 bf16 at a time
  if (a->epilogue == PK_EPI_BIAS_RESID_F32 && ((a->ldr * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->resid) & 15) != 0)) return false;
  return true;
}

template <int BN, int EPI>
static int dispatch_path(const pk_gemm_args* a, cudaStream_t stream) {
  return can_use_tma_epilogue(a) ? launch_gemm<BN, EPI, true>(a, stream) : launch_gemm<BN, EPI, false>(a, stream);
}

template <int BN>
static int dispatch_epi(const pk_gemm_args* a, cudaStream_t stream) {
  switch (a->epilogue) {
    case PK_EPI_BIAS_BF16: return dispatch_path<BN, PK_EPI_BIAS_BF16>(a, stream);
    case PK_EPI_BIAS_GELU_BF16: return dispatch_path<BN, PK_EPI_BIAS_GELU_BF16>(a, stream);
    case PK_EPI_BIAS_RESID_F32: return dispatch_path<BN, PK_EPI_BIAS_RESID_F32>(a, stream);
    case PK_EPI_BIAS_F32: return dispatch_path<BN, PK_EPI_BIAS_F32>(a, stream);
  }
  set_last_error("pk_gemm_bf16: unknown epilogue %d", a->epilogue);
  return PK_ERR_INVALID;
}

// Tile width: 256 when it divides N (fewest A re-reads, best measured), else the width of
// {192, 256, 128} that wastes the fewest padded columns.
int pick_block_n(int N) {
  if (N % 256 == 0) return 256;
  int best = 128;
  long long best_cost = -1;
  const int cands[3] = {192, 256, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long padded = static_cast<long long>((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || padded < best_cost) { best_cost = padded; best = bn; }
  }
  return best;
}

// pk_gemm2.cu: CTA-pair (cta_group::2) kernel for plain row mappings
bool pair_gemm_eligible(const pk_gemm_args* a);
int launch_pair_gemm(const pk_gemm_args* a, cudaStream_t stream);
int pair_row_stat_parts(int N);

}  // namespace pk

extern "C" int pk_gemm_bf16(const pk_gemm_args* a, void* stream) {
  using namespace pk;
  PK_REQUIRE(a != nullptr, "pk_gemm_bf16: null args");
  PK_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "pk_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  PK_REQUIRE(a->K % kBK == 0, "pk_gemm_bf16: K=%d must be a multiple of %d", a->K, kBK);
  PK_REQUIRE(a->N % 4 == 0, "pk_gemm_bf16: N=%d must be a multiple of 4", a->N);
  PK_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0, "pk_gemm_bf16: lda/ldw must be multiples of 8 elements (TMA 16-B strides)");
  PK_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0,
             "pk_gemm_bf16: A/W must be 16-byte aligned");
  PK_REQUIRE(a->out != nullptr && a->ldo % 4 == 0, "pk_gemm_bf16: out null or ldo not a multiple of 4");
  if (a->epilogue == PK_EPI_BIAS_RESID_F32)
    PK_REQUIRE(a->resid != nullptr && a->ldr % 4 == 0, "pk_gemm_bf16: residual epilogue needs resid with ldr %% 4 == 0");
  if (a->M == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->group_offsets) {
    PK_REQUIRE(a->n_groups >= 1 && pair_gemm_eligible(a) && a->cta_pair != 1,
               "pk_gemm_bf16: a grouped launch needs the CTA-pair kernel (contiguous aligned rows, no staged residual / LayerNorm "
               "fusion / rowscale, 1 <= n_groups <= 64)");
    return launch_pair_gemm(a, s);
  }
  const bool ln_fused = a->xb_out != nullptr || a->ln_stats != nullptr;
  if (a->a_wrap_k > 0 || a->out_format != PK_OUT_BF16) {
    // two-term split operands / fp16 / split outputs (bf16x2 mode): CTA-pair kernel only
    PK_REQUIRE(pair_gemm_eligible(a) && a->cta_pair != 1 && !ln_fused,
               "pk_gemm_bf16: a_wrap_k / out_format need the CTA-pair kernel (contiguous rows, aligned) without fused LayerNorm");
    PK_REQUIRE(a->a_wrap_k == 0 || (a->K == 3 * a->a_wrap_k && a->a_wrap_k % kBK == 0 && a->lda >= 2ll * a->a_wrap_k),
               "pk_gemm_bf16: a_wrap_k=%d needs K == 3*a_wrap_k (K=%d), a_wrap_k %% 64 == 0 and lda >= 2*a_wrap_k", a->a_wrap_k, a->K);
    if (a->out_format != PK_OUT_BF16) {
      PK_REQUIRE(a->epilogue == PK_EPI_BIAS_BF16 || a->epilogue == PK_EPI_BIAS_GELU_BF16, "pk_gemm_bf16: out_format applies to the 2-byte epilogues");
      PK_REQUIRE(a->out_format == PK_OUT_F16 ? a->epilogue == PK_EPI_BIAS_BF16 : a->out_format == PK_OUT_BF16X2,
                 "pk_gemm_bf16: out_format %d is not available for epilogue %d", a->out_format, a->epilogue);
      if (a->out_format == PK_OUT_BF16X2)
        PK_REQUIRE(a->N % 64 == 0 && a->ldo >= 2ll * a->N, "pk_gemm_bf16: split output needs N %% 64 == 0 and ldo >= 2N (N=%d)", a->N);
    }
    return launch_pair_gemm(a, s);
  }
  if (ln_fused) {
    PK_REQUIRE(pair_gemm_eligible(a) && a->cta_pair != 1 && a->block_n == 0,
               "pk_gemm_bf16: the fused-LayerNorm epilogues need the CTA-pair kernel (contiguous rows, aligned, block_n auto)");
    if (a->xb_out) {
      PK_REQUIRE(a->epilogue == PK_EPI_BIAS_RESID_F32 && a->row_stats != nullptr && a->ldxb % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(a->xb_out) & 15) == 0,
                 "pk_gemm_bf16: xb_out needs the residual epilogue, row_stats and a 16-byte aligned bf16 buffer");
    }
    if (a->ln_stats) {
      PK_REQUIRE((a->epilogue == PK_EPI_BIAS_BF16 || a->epilogue == PK_EPI_BIAS_GELU_BF16) && a->ln_c1 != nullptr && a->ln_parts > 0 &&
                     a->ln_dim > 0,
                 "pk_gemm_bf16: ln_stats needs a bf16 epilogue, ln_c1, ln_parts and ln_dim");
    }
    return launch_pair_gemm(a, s);
  }
  if (a->cta_pair != 1 && pair_gemm_eligible(a) && (a->cta_pair == 2 || a->M > 256)) return launch_pair_gemm(a, s);
  PK_REQUIRE(a->cta_pair != 2, "pk_gemm_bf16: the CTA-pair kernel needs contiguous rows, 16-byte aligned rows and N %% 8 == 0");
  const int bn = a->block_n > 0 ? a->block_n : pick_block_n(a->N);
  switch (bn) {
    case 128: return dispatch_epi<128>(a, s);
    case 192: return dispatch_epi<192>(a, s);
    case 256: return dispatch_epi<256>(a, s);
  }
  set_last_error("pk_gemm_bf16: unsupported block_n %d", bn);
  return PK_ERR_INVALID;
}

extern "C" int pk_gemm_row_stat_parts(int N) { return N > 0 ? pk::pair_row_stat_parts(N) : 0; }
