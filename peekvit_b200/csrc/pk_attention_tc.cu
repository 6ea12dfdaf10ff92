// peekvit_b200 — tcgen05/TMEM attention for the dense ViT shape (sm_100a).
//
// Replaces nn.MultiheadAttention's core (reference models/blocks.py:93-95) for uniform-length
// samples with 64 < n <= 256 tokens and head_dim 64 (ViT-B/16 and ViT-S/16 at 224 px: n = 197/198).
// Other shapes (ragged batches, key multiplicities, n <= 128, n > 256, head_dim 32) take the
// general mma.sync kernel in pk_attention.cu.
//
// One work item = one (sample, head).  Persistent CTAs, warp-specialised:
//   warp 0      TMA producer: Q (two 128-row tiles), K and V (n_pad x 64, 128-byte swizzle) of the
//               next item into a 2-slot shared-memory ring.  The tensor map is 3-D
//               (column, token-in-sample, sample), so rows past the end of a sample are zero-filled
//               instead of leaking the next sample's tokens.
//   warp 1      MMA issuer.  Per 128-query tile r (a TMEM "region" of 256 columns):
//                 S_r = Q_r K^T      tcgen05.mma  A,B from smem (K-major), 128 x n_pad fp32 in TMEM
//                 O_r = P_r V        tcgen05.mma  A = P_r from TMEM (bf16, written by the softmax
//                                    warps over the dead S columns), B = V from smem (MN-major)
//   warps 4-7   softmax + output of tile 0;  warps 8-11 the same for tile 1.  A thread owns one query
//               row (= one TMEM lane): row max and row sum need no shuffles.  Two passes over the
//               S row in TMEM (max, then exp2/sum/pack), P stored back with tcgen05.st; after PV the
//               thread reads its O row, scales by 1/sum and writes 128 contiguous bytes.
// The two regions are software-pipelined against each other: while the softmax warps of one tile run,
// the tensor core works on the other tile / the next item.  Every mbarrier wait is bounded.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

#include <cstdlib>

namespace pk {

constexpr int kTcThreads = 512;                      // 16 warps: TMA, 2 MMA issuers, 1 idle, 8 softmax, 4 output
constexpr int kTcRegsSoftmax = 168;                 // 8 warps x 32 x 168 + 8 warps x 32 x 88 = 65536 registers
constexpr int kTcRegsOther = 88;
constexpr int kTcDH = 64;
constexpr int kTcQTileBytes = 128 * 128;            // 128 query rows x 64 bf16
constexpr int kTcMaxKeys = 256;
// Shared-memory rings.  Q (both tiles) and K of an item are dead as soon as both S = Q K^T are done, V only after
// both PV: the Q/K ring (2 slots) is released right after the QK MMAs, V has its own, deeper ring, so the ~2 us TMA
// latency of the next items' operands is off the critical path of the two staggered regions.
// SPLIT_OUT (bf16x2 mode): the output leaves as two bf16 tiles (hi and lo) per 32-row block, so the staging area doubles and the
// V ring drops to two slots to make room.
template <int NPAD, int SPLIT_OUT = 0>
struct TcSmem {
  static constexpr int kKBytes = NPAD * 128;                          // NPAD keys x 64 bf16 (multiple of 1024 since NPAD % 16 == 0 -> 2048)
  static constexpr int kQKSlotBytes = 2 * kTcQTileBytes + kKBytes;
  static constexpr int kQKSlots = 2;
  static constexpr int kVSlots = (NPAD <= 224 && !SPLIT_OUT) ? 3 : 2;
  static constexpr int kVOff = kQKSlots * kQKSlotBytes;
  static constexpr int kStageOff = kVOff + kVSlots * kKBytes;
  static constexpr int kStageBytes = (1 + SPLIT_OUT) * 4 * 4096;      // one (or two) 32-row x 128-byte tiles per output warp
  static constexpr int kBytes = kStageOff + kStageBytes + 1024 /* inv_sum */ + 256 /* barriers */ + 1024 /* alignment */;
  static_assert(kBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};
constexpr int kTcBarBytes = 256;
constexpr int kTcOutStageBytes = 4 * 4096;           // one 32-row x 128-byte output tile per output warp
constexpr int kTcRegionCols = 256;                  // TMEM columns per query tile
constexpr int kTcOCol = 128;
#ifndef PK_TC_POLY
#define PK_TC_POLY 0
#endif
// Of every 4 column pairs, how many take the FMA-pipe exp2 instead of the MUFU.  Measured on B200 (B=256, H=12,
// n=197): 0 -> 102 us, 1 -> 108 us, 2 -> 105-114 us: the packed-FMA polynomial costs as much pipe time as the MUFU op
// it replaces, so the default is 0 (all MUFU).
constexpr int kTcPoly = PK_TC_POLY;                        // O accumulator at region columns [128, 192)

struct TcAttParams {
  __nv_bfloat16* out;
  int batch, num_heads, seq_len, n_pad;             // n_pad = seq_len rounded up to 16
  float scale_log2;
  unsigned int* flag;
  unsigned long long* trace;   // PK_ATT_TRACE=1: CTA 0 records clock64 at pipeline events (tools/attn_trace.py)
  int debug;      // PK_ATT_DEBUG bits (timing experiments only): 1 = no row-max pass, 2 = no exp pass, 4 = no output store
};

// 3-D tiled store shared -> global (bulk async group); rows past the end of the sample are clipped by the tensor map.
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs per column] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major, 128-byte-swizzled operand (V: rows = keys (K dim), 64 contiguous head-dim elements = MN):
// 8-key groups 1024 B apart (SBO); a single 64-element MN atom, so LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
// The clock64 trace is compiled in only with -DPK_ATT_TRACE_BUILD (PEEKVIT_B200_NVCC_FLAGS=-DPK_ATT_TRACE_BUILD python -m
// peekvit_b200.build --force): even the untaken `if (p.trace && ...)` of ~25 call sites sat on the single-warp critical paths
// of these kernels and cost the quad-region kernel 18 % and the ragged kernel 7 % (same-box A/B, profiles/r02).
// trace slot layout: [item(0..15)][warp(0..15)][event(0..7)]
__device__ __forceinline__ void tc_trace(const TcAttParams& p, int it, int ev) {
  if (p.trace && blockIdx.x == 0 && it < 16 && (threadIdx.x & 31) == 0) p.trace[(it * 16 + (threadIdx.x >> 5)) * 8 + ev] = clock64();
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Running maximum over 16 columns of the S row (3-input max: 8 instructions); columns >= n are skipped
// when MASK (only the last 16-column group of a row can hold padding).
template <bool MASK>
__device__ __forceinline__ float row_max16(const uint32_t* s, float mx, int col0, int n) {
  if constexpr (!MASK) {
    // four independent 3-input max chains, then a 2-level combine (short dependency chains)
    float m0 = fmax3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
    float m1 = fmax3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
    float m2 = fmax3(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
    float m3 = fmax3(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
    m0 = fmax3(m0, __uint_as_float(s[12]), __uint_as_float(s[13]));
    m1 = fmax3(m1, __uint_as_float(s[14]), __uint_as_float(s[15]));
    mx = fmax3(mx, fmax3(m0, m1, m2), m3);
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) if (col0 + e < n) mx = fmaxf(mx, __uint_as_float(s[e]));
  }
  return mx;
}
// exp2 of two non-positive arguments on the FMA pipe (the MUFU pipe is the softmax bottleneck: 4 lanes/clk/SMSP).
// Cody-Waite split x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (max rel. error
// 7.5e-5, far below the bf16 rounding of P); 2^n by adding n to the exponent field.  Packed f32x2 arithmetic.
__device__ __forceinline__ void exp2_poly_x2(uint64_t x, float& r0, float& r1) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x0 = fmaxf(x0, -125.0f);                     // keep n inside the exponent range (result ~ 2^-125 = 0 for P)
  x1 = fmaxf(x1, -125.0f);
  const uint64_t xc = f2_pack(x0, x1);
  const uint64_t t = f2_add(xc, f2_pack(12582912.0f, 12582912.0f));        // 1.5 * 2^23: n lands in the low mantissa bits
  const uint64_t nn = f2_add(t, f2_pack(-12582912.0f, -12582912.0f));
  const uint64_t f = f2_fma(nn, f2_pack(-1.0f, -1.0f), xc);
  uint64_t q = f2_fma(f2_pack(0.055171288549900055f, 0.055171288549900055f), f, f2_pack(0.24261046946048737f, 0.24261046946048737f));
  q = f2_fma(q, f, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  q = f2_fma(q, f, f2_pack(0.9999281167984009f, 0.9999281167984009f));
  float q0, q1, t0, t1;
  f2_unpack(q, q0, q1);
  f2_unpack(t, t0, t1);
  r0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// One 16-column group of the S row -> 8 packed bf16x2 probabilities; returns the partial row sum.
// POLY of every 4 column pairs take the FMA-pipe exp2, the rest the MUFU.
template <bool MASK, int POLY, int F16 = 0>
__device__ __forceinline__ float softmax_group16(const uint32_t* s, uint32_t* p, float scale_log2, float neg_max_scaled, int col0, int n) {
  if constexpr (MASK) {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(s[e]), scale_log2, neg_max_scaled));
      float p1 = ex2_approx(fmaf(__uint_as_float(s[e + 1]), scale_log2, neg_max_scaled));
      if (col0 + e >= n) p0 = 0.f;
      if (col0 + e + 1 >= n) p1 = 0.f;
      sum0 += p0;
      sum1 += p1;
      p[e >> 1] = F16 ? pack_f16(p0, p1) : pack_bf16(p0, p1);
    }
    return sum0 + sum1;
  } else {
    const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(neg_max_scaled, neg_max_scaled);
    uint64_t acc = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint64_t x = f2_fma(f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
      float p0, p1;
      if ((i & 3) < POLY) {
        exp2_poly_x2(x, p0, p1);
      } else {
        float x0, x1;
        f2_unpack(x, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      acc = f2_add(acc, f2_pack(p0, p1));
      p[i] = F16 ? pack_f16(p0, p1) : pack_bf16(p0, p1);
    }
    float a0, a1;
    f2_unpack(acc, a0, a1);
    return a0 + a1;
  }
}

template <int NPAD>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q128, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_out, const TcAttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using SM = TcSmem<NPAD>;
  uint8_t* out_stage = smem + SM::kStageOff;
  float* inv_sum = reinterpret_cast<float*>(out_stage + kTcOutStageBytes);          // [2 regions][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kTcOutStageBytes + 1024);
  uint64_t* qk_full = bars;            // [2] Q/K slot loaded (tx bytes)
  uint64_t* qk_empty = bars + 2;       // [2] Q/K slot consumed (one umma commit per MMA warp, after its QK)
  uint64_t* v_full = bars + 4;         // [3] V slot loaded (tx bytes)
  uint64_t* v_empty = bars + 7;        // [3] V slot consumed (one umma commit per MMA warp, after its PV)
  uint64_t* s_full = bars + 10;        // [2] region: S ready (umma commit)
  uint64_t* p_ready = bars + 12;       // [2] region: P written, row sums published (4 softmax-warp arrivals)
  uint64_t* o_full = bars + 14;        // [2] region: O ready (umma commit)
  uint64_t* s_free = bars + 16;        // [2] region: O read out, region reusable (4 output-warp arrivals)
  uint64_t* sum_ready = bars + 18;     // [2] region: inv_sum[r] valid for the output warps (4 softmax-warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = warp_id();
  const int lane = lane_id();
  const int n = p.seq_len;
  constexpr int n_pad = NPAD;
  const int D = p.num_heads * kTcDH;
  const int num_items = p.batch * p.num_heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q128);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&qk_full[i]), 1);
      mbar_init(smem_u32(&qk_empty[i]), 2);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 4);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
      mbar_init(smem_u32(&sum_ready[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register rebalancing is the first statement of every role: the softmax warps keep half an S row resident.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kTcRegsOther>();
    int qs = 0, vs = 0;
    uint32_t qph = 0, vph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.num_heads, h = item - b * p.num_heads;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_empty[qs]), qph ^ 1u, p.flag, 0x2100u + qs))) break;
      tc_trace(p, (item - blockIdx.x) / gridDim.x, 0);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&qk_full[qs]);
        const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
        mbar_expect_tx(bar, SM::kQKSlotBytes);
        tma_load_3d(base, &tmap_q128, bar, h * kTcDH, 0, b);
        tma_load_3d(base + kTcQTileBytes, &tmap_q128, bar, h * kTcDH, 128, b);
        tma_load_3d(base + 2 * kTcQTileBytes, &tmap_kv, bar, D + h * kTcDH, 0, b);
      }
      __syncwarp();
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1u, p.flag, 0x2110u + vs))) break;
      tc_trace(p, (item - blockIdx.x) / gridDim.x, 1);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&v_full[vs]);
        mbar_expect_tx(bar, SM::kKBytes);
        tma_load_3d(smem_u32(smem + SM::kVOff + vs * SM::kKBytes), &tmap_kv, bar, 2 * D + h * kTcDH, 0, b);
      }
      __syncwarp();
      if (++qs == SM::kQKSlots) { qs = 0; qph ^= 1u; }
      if (++vs == SM::kVSlots) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> query tile 0, warp 2 -> tile 1
    // One issuer per region, so neither region ever waits behind the other's barrier.  Region 1 starts half a
    // period late (after region 0's first softmax): the two softmax groups then use the MUFU alternately instead
    // of in lockstep, and the tensor core works for one region while the other is in its softmax.
    setmaxnreg_dec<kTcRegsOther>();
    const int r = warp - 1;
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, NPAD);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kTcDH, /*b_mn_major=*/1);
    constexpr int k_steps_pv = NPAD / 16;
    const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(r * kTcRegionCols);
    int qs = 0, vs = 0;
    uint32_t qph = 0, vph = 0, rph = 0;
    bool first = true;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_full[qs]), qph, p.flag, 0x2200u + qs))) break;
      tc_trace(p, it, 4);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&s_free[r]), rph ^ 1u, p.flag, 0x2300u + r))) break;
      if (first && r == 1) {
        if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[0]), 0u, p.flag, 0x2310u))) break;
      }
      first = false;
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
      tc_trace(p, it, 0);
      if (elect_one()) {                       // S_r = Q_r K^T
        const uint64_t a_desc = umma_desc_kmajor_sw128(base + r * kTcQTileBytes);
        const uint64_t b_desc = umma_desc_kmajor_sw128(base + 2 * kTcQTileBytes);
#pragma unroll
        for (int k = 0; k < kTcDH / 16; ++k)
          umma_bf16(s_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&s_full[r]));
        umma_commit(smem_u32(&qk_empty[qs]));     // Q/K of this item are dead once both regions' QK have retired
      }
      __syncwarp();
      tc_trace(p, it, 1);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[r]), rph, p.flag, 0x2400u + r))) break;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_full[vs]), vph, p.flag, 0x2410u + vs))) break;
      tcgen05_fence_after();
      tc_trace(p, it, 2);
      if (elect_one()) {                       // O_r = P_r V
        const uint64_t v_desc = umma_desc_mnmajor_sw128(smem_u32(smem + SM::kVOff + vs * SM::kKBytes));
        const uint32_t o_tmem = s_tmem + kTcOCol;
#pragma unroll
        for (int k = 0; k < k_steps_pv; ++k)      // 16 keys per step: 8 TMEM columns of P, 2048 B of V
          umma_bf16_ts(o_tmem, s_tmem + static_cast<uint32_t>(8 * k), v_desc + static_cast<uint64_t>(128 * k), idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&o_full[r]));
        umma_commit(smem_u32(&v_empty[vs]));      // 2 arrivals (one per region) free the V slot
      }
      __syncwarp();
      tc_trace(p, it, 3);
      rph ^= 1u;
      if (++qs == SM::kQKSlots) { qs = 0; qph ^= 1u; }
      if (++vs == SM::kVSlots) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 3) {
    setmaxnreg_dec<kTcRegsOther>();
  } else if (warp < 12) {
    // ------------------------------------------------------------------ softmax warps (4 per query tile)
    setmaxnreg_inc<kTcRegsSoftmax>();
    const int r = (warp - 4) >> 2;               // query tile / TMEM region
    const int q = warp & 3;                      // TMEM lane quarter
    const bool warp_has_rows = r * 128 + q * 32 < n;          // warp-uniform: idle warps only keep the barrier protocol alive
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
    // The S row is NPAD columns = NC chunks of 32 (+ one 16-column tail).  The last KC chunks and the tail stay in
    // registers between the two passes; the first NT chunks are read from TMEM twice.
    constexpr int NC = NPAD / 32;
    constexpr bool TAIL = (NPAD % 32) != 0;
    constexpr int KC = NC < 2 ? NC : 2;
    constexpr int NT = NC - KC;
    constexpr int KEEP = KC * 32 + (TAIL ? 16 : 0);
    const float scale_log2 = p.scale_log2;
    uint32_t rph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      tc_trace(p, it, 0);
      if (!mbar_wait(smem_u32(&s_full[r]), rph, p.flag, 0x2500u + r)) break;
      tcgen05_fence_after();
      tc_trace(p, it, 1);
      if (warp_has_rows) {
        uint32_t keep[KEEP];
        uint32_t ha[16], hb[16];
        // ---- pass 1: row maximum (the kept chunks' loads overlap the transient chunks' processing)
        float mx = -INFINITY;
        if (p.debug & 1) mx = 0.f;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          uint32_t (&dst)[32] = *reinterpret_cast<uint32_t (*)[32]>(&keep[32 * c]);
          tmem_ld_32x32(t_base + static_cast<uint32_t>(32 * (NT + c)), dst);
        }
        if constexpr (TAIL) {
          uint32_t (&dst)[16] = *reinterpret_cast<uint32_t (*)[16]>(&keep[32 * KC]);
          tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(32 * NC), dst);
        }
        if (!(p.debug & 1)) {
          // transient columns in 16-column steps, the next step's TMEM load in flight while the current one is reduced
          if constexpr (NT > 0) tmem_ld_32x32_x16(t_base, ha);
          tmem_ld_wait();                              // kept chunks + first transient step
#pragma unroll
          for (int j = 0; j < 2 * NT; ++j) {
            uint32_t (&cur)[16] = (j & 1) ? hb : ha;
            uint32_t (&nxt)[16] = (j & 1) ? ha : hb;
            if (j + 1 < 2 * NT) tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * (j + 1)), nxt);
            mx = row_max16<false>(cur, mx, 0, n);
            if (j + 1 < 2 * NT) tmem_ld_wait();
          }
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            mx = row_max16<false>(&keep[32 * c], mx, 0, n);
            if (!TAIL && c == KC - 1) mx = row_max16<true>(&keep[32 * c + 16], mx, 32 * (NT + c) + 16, n);
            else mx = row_max16<false>(&keep[32 * c + 16], mx, 0, n);
          }
          if constexpr (TAIL) mx = row_max16<true>(&keep[32 * KC], mx, 32 * NC, n);
        } else {
          tmem_ld_wait();
        }
        const float neg_max_scaled = -mx * scale_log2;
        tc_trace(p, it, 2);
        // ---- pass 2: p = exp2(s*scale - max*scale), row sum, bf16 P written over the S columns (in column order:
        // P of chunk j lands on S columns that were consumed by chunk j/2)
        float sum = 0.f;
        if (p.debug & 2) sum = 1.f;
        else {
          if constexpr (NT > 0) {
            tmem_ld_32x32_x16(t_base, ha);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 2 * NT; ++j) {
            uint32_t (&cur)[16] = (j & 1) ? hb : ha;
            uint32_t (&nxt)[16] = (j & 1) ? ha : hb;
            if (j + 1 < 2 * NT) tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * (j + 1)), nxt);
            uint32_t pk[8];
            sum += softmax_group16<false, kTcPoly>(cur, pk, scale_log2, neg_max_scaled, 0, n);
            if (j + 1 < 2 * NT) tmem_ld_wait();        // the next step is in registers before its columns may be overwritten
            tmem_st_32x32_x8(t_base + static_cast<uint32_t>(8 * j), pk);
          }
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            uint32_t pk[16];
            sum += softmax_group16<false, kTcPoly>(&keep[32 * c], pk, scale_log2, neg_max_scaled, 0, n);
            if (!TAIL && c == KC - 1) sum += softmax_group16<true, kTcPoly>(&keep[32 * c + 16], pk + 8, scale_log2, neg_max_scaled, 32 * (NT + c) + 16, n);
            else sum += softmax_group16<false, kTcPoly>(&keep[32 * c + 16], pk + 8, scale_log2, neg_max_scaled, 0, n);
            tmem_st_32x32_x8(t_base + static_cast<uint32_t>(16 * (NT + c)), pk);
            tmem_st_32x32_x8(t_base + static_cast<uint32_t>(16 * (NT + c) + 8), pk + 8);
          }
          if constexpr (TAIL) {
            uint32_t pk[8];
            sum += softmax_group16<true, kTcPoly>(&keep[32 * KC], pk, scale_log2, neg_max_scaled, 32 * NC, n);
            tmem_st_32x32_x8(t_base + static_cast<uint32_t>(16 * NC), pk);
          }
        }
        inv_sum[r * 128 + q * 32 + lane] = 1.0f / sum;
        tmem_st_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&p_ready[r]));
        mbar_arrive(smem_u32(&sum_ready[r]));
      }
      tc_trace(p, it, 3);
      rph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ output warps (one per TMEM lane quarter, both regions)
    setmaxnreg_dec<kTcRegsOther>();
    const int q = warp & 3;
    uint8_t* stg = out_stage + q * 4096;
    uint32_t rph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      const int b = item / p.num_heads, h = item - b * p.num_heads;
      bool ok = true;
      for (int r = 0; r < 2; ++r) {
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
        const int row0 = r * 128 + q * 32;
        if (!mbar_wait(smem_u32(&sum_ready[r]), rph, p.flag, 0x2700u + r)) { ok = false; break; }
        if (!mbar_wait(smem_u32(&o_full[r]), rph, p.flag, 0x2600u + r)) { ok = false; break; }
        tcgen05_fence_after();
        tc_trace(p, it, r * 3 + 0);
        const bool has_rows = row0 < n;          // warp-uniform
        const float inv = inv_sum[r * 128 + q * 32 + lane];
        // O row: 64 fp32 out of TMEM first (the region is handed back as early as possible), then scaled bf16 into
        // this warp's swizzled staging tile
        uint32_t o0[32], o1[32];
        if (has_rows) {
          tmem_ld_32x32(t_base + kTcOCol, o0);
          tmem_ld_32x32(t_base + kTcOCol + 32, o1);
          tmem_ld_wait();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_free[r]));         // the region may take the next item's S
        if (has_rows && !(p.debug & 4)) {
          if (lane == 0) bulk_wait_read<0>();                     // the previous store has drained this staging tile
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                           pack_bf16(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                           pack_bf16(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                           pack_bf16(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + c) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(o1[8 * c]) * inv, __uint_as_float(o1[8 * c + 1]) * inv),
                           pack_bf16(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv),
                           pack_bf16(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv),
                           pack_bf16(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv));
        }
        tc_trace(p, it, r * 3 + 1);
        if (has_rows && !(p.debug & 4)) {
          // one TMA store of the 32-row x 64-column tile to out[b, row0 .., h*64 ..] (rows >= n clipped by the 3-D map)
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmap_out, smem_u32(stg), h * kTcDH, row0, b);
            bulk_commit();
          }
        }
        tc_trace(p, it, r * 3 + 2);
      }
      if (!ok) break;
      rph ^= 1u;
    }
    if (lane == 0) bulk_wait_read<0>();       // smem must outlive the last store's reads
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ======================================================================================================================
// Column-split variant ("tc3"): 16 softmax warps.  The per-region chain  S ready -> max -> exp -> PV -> O out -> region free
// is what bounds the kernel (6400 cycles per item with the MUFU busy half of the time), and its longest link is the softmax
// of one 128-row tile on four warps.  Here every TMEM lane quarter of a region is served by TWO warps that split the S row by
// columns (16-column groups [0, G0) and [G0, G)): partial row maxima are exchanged through shared memory behind a 64-thread
// named barrier, partial row sums are added by the output warp.  Each half packs its probabilities into the S columns it has
// already consumed itself (half 0 at column 8j, half 1 at 16*G0 + 8(j - G0)), so the halves never touch each other's columns
// and the PV MMA simply takes its A operand of k-step j from the matching address.  O moves to region columns [192, 256).
// 24 warps: TMA, 2 MMA issuers, 1 idle, 16 softmax, 4 output; registers per role 40 / 56 / 40 / 80 / 88 out of the 768 x 80 pool.
constexpr int kTc3Threads = 768;
constexpr int kTc3RegsSoftmax = 80;
constexpr int kTc3RegsOther = 40;                  // TMA producer, idle warp
constexpr int kTc3RegsMma = 56;
constexpr int kTc3RegsOut = 88;                    // both 32-column halves of the O row in flight
constexpr int kTc3OCol = 192;
template <int NPAD, int SPLIT_OUT = 0>
struct Tc3Smem {
  using B = TcSmem<NPAD, SPLIT_OUT>;
  static constexpr int kPartOff = B::kStageOff + B::kStageBytes;              // row-max and row-sum partials: [2][region][half][128] f32
  static constexpr int kBarOff = kPartOff + 4096;
  static constexpr int kBytes = kBarOff + 256 /* barriers */ + 1024 /* alignment */;
  static_assert(kBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// trace slots: warps 0-11 as they are, output warps 20-23 in slots 12-15
// Compiled in for the trace build -- and, as it always was, for NPAD >= 192: the ViT-B/16 kernel (NPAD 208) owes its register
// allocation to the scheduling points these calls are (without them ptxas spills 200 bytes and the kernel runs 184 us instead
// of 134); the smaller tiles lose nothing and run 11 % faster without (uniform 99 tokens: 89 -> 79 us).
template <int NPAD>
__device__ __forceinline__ void tc3_trace(const TcAttParams& p, int it, int ev) {
#ifndef PK_ATT_TRACE_BUILD
  if constexpr (NPAD >= 192)
#endif
  {
  if (p.trace && blockIdx.x == 0 && it < 16 && (threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    const int slot = w < 12 ? w : (w >= 20 ? w - 8 : -1);
    if (slot >= 0) p.trace[(it * 16 + slot) * 8 + ev] = clock64();
  }
  }
}

// Softmax of one item for the column groups of one half of the row.
template <int NPAD, int HALF, int POLY, int F16>
__device__ __forceinline__ void tc3_softmax_item(uint32_t t_base, int n, float scale_log2, float* mx_mine, const float* mx_other,
                                                 float* sum_mine, int bar_id, const TcAttParams& p, int it) {
  const int debug = p.debug;
  constexpr int G = NPAD / 16, G0 = (G + 1) / 2;
  constexpr int GB = HALF ? G0 : 0;                 // first 16-column group of this half
  constexpr int GN = HALF ? G - G0 : G0;            // number of groups
  constexpr int PCOL0 = HALF ? 16 * G0 : 0;         // where this half packs its probabilities
  const uint32_t s_col = t_base + static_cast<uint32_t>(16 * GB);
  uint32_t ha[16], hb[16];
  // ---- pass 1: partial row maximum.  The pass is bound by the TMEM load round trip (the 16 max instructions of a group are
  // nothing against it), so it moves 32 columns per trip
  float mx = -INFINITY;
  if (!(debug & 1)) {
    constexpr int G32 = GN / 2;                     // 32-column steps; an odd group is left for a final 16-column step
#pragma unroll
    for (int j = 0; j < G32; ++j) {
      uint32_t w[32];
      tmem_ld_32x32(s_col + static_cast<uint32_t>(32 * j), w);
      tmem_ld_wait();
      mx = row_max16<false>(w, mx, 0, n);
      if (HALF == 1 && 2 * j + 1 == GN - 1) mx = row_max16<true>(w + 16, mx, 16 * (GB + 2 * j + 1), n);   // the row's last group can hold padding
      else mx = row_max16<false>(w + 16, mx, 0, n);
    }
    if constexpr (GN % 2 == 1) {
      tmem_ld_32x32_x16(s_col + static_cast<uint32_t>(32 * G32), ha);
      tmem_ld_wait();
      if (HALF == 1) mx = row_max16<true>(ha, mx, 16 * (GB + GN - 1), n);
      else mx = row_max16<false>(ha, mx, 0, n);
    }
  } else {
    mx = 0.f;
  }
  *mx_mine = mx;
  tc3_trace<NPAD>(p, it, 4);
  named_bar_sync(bar_id, 64);
  mx = fmaxf(mx, *mx_other);
  tc3_trace<NPAD>(p, it, 2);
  const float neg_max_scaled = -mx * scale_log2;
  // ---- pass 2: p = exp2(s*scale - max*scale), partial row sum, bf16 P packed over this half's own consumed S columns
  float sum = 0.f;
  if (!(debug & 2)) {
    tmem_ld_32x32_x16(s_col, ha);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < GN; ++j) {
      uint32_t (&cur)[16] = (j & 1) ? hb : ha;
      uint32_t (&nxt)[16] = (j & 1) ? ha : hb;
      if (j + 1 < GN) tmem_ld_32x32_x16(s_col + static_cast<uint32_t>(16 * (j + 1)), nxt);
      uint32_t pk[8];
      if (HALF == 1 && j == GN - 1) sum += softmax_group16<true, POLY, F16>(cur, pk, scale_log2, neg_max_scaled, 16 * (GB + j), n);
      else sum += softmax_group16<false, POLY, F16>(cur, pk, scale_log2, neg_max_scaled, 0, n);
      if (j + 1 < GN) tmem_ld_wait();
      tmem_st_32x32_x8(t_base + static_cast<uint32_t>(PCOL0 + 8 * j), pk);   // lands on S group GB + j/2: already in registers
    }
  } else {
    sum = 0.5f;
  }
  *sum_mine = sum;
  tmem_st_wait();
}

// FMT 0: bf16 q / k / v / probabilities, bf16 output [rows, D].  FMT 1 (bf16x2 arithmetic mode): IEEE-half q / k / v and
// probabilities (11 significant bits, same tensor-core rate), output split into hi + lo (both bf16) as [lo | hi] in a
// [rows, 2D] buffer: the A operand of the split out-projection GEMM.
template <int NPAD, int POLY = 0, int FMT = 0>
__global__ void __launch_bounds__(kTc3Threads, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q128, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_out, const TcAttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using SM = TcSmem<NPAD, FMT>;
  using SM3 = Tc3Smem<NPAD, FMT>;
  uint8_t* out_stage = smem + SM::kStageOff;
  float* mx_part = reinterpret_cast<float*>(smem + SM3::kPartOff);                 // [region][half][128]
  float* sum_part = mx_part + 512;                                                  // [region][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM3::kBarOff);
  uint64_t* qk_full = bars;            // [2]
  uint64_t* qk_empty = bars + 2;       // [2]
  uint64_t* v_full = bars + 4;         // [3]
  uint64_t* v_empty = bars + 7;        // [3]
  uint64_t* s_full = bars + 10;        // [2] region: S ready (umma commit)
  uint64_t* p_ready = bars + 12;       // [2] region: P written (8 softmax-warp arrivals)
  uint64_t* o_full = bars + 14;        // [2] region: O ready (umma commit)
  uint64_t* s_free = bars + 16;        // [2] region: O read out (4 output-warp arrivals)
  uint64_t* sum_ready = bars + 18;     // [2] region: row-sum partials published (8 softmax-warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = warp_id();
  const int lane = lane_id();
  const int n = p.seq_len;
  const int D = p.num_heads * kTcDH;
  const int num_items = p.batch * p.num_heads;
  constexpr int G = NPAD / 16, G0 = (G + 1) / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q128);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&qk_full[i]), 1);
      mbar_init(smem_u32(&qk_empty[i]), 2);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 8);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
      mbar_init(smem_u32(&sum_ready[i]), 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kTc3RegsOther>();
    int qs = 0, vs = 0;
    uint32_t qph = 0, vph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.num_heads, h = item - b * p.num_heads;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_empty[qs]), qph ^ 1u, p.flag, 0x2100u + qs))) break;
      tc3_trace<NPAD>(p, (item - blockIdx.x) / gridDim.x, 0);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&qk_full[qs]);
        const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
        mbar_expect_tx(bar, SM::kQKSlotBytes);
        tma_load_3d(base, &tmap_q128, bar, h * kTcDH, 0, b);
        tma_load_3d(base + kTcQTileBytes, &tmap_q128, bar, h * kTcDH, 128, b);
        tma_load_3d(base + 2 * kTcQTileBytes, &tmap_kv, bar, D + h * kTcDH, 0, b);
      }
      __syncwarp();
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1u, p.flag, 0x2110u + vs))) break;
      tc3_trace<NPAD>(p, (item - blockIdx.x) / gridDim.x, 1);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&v_full[vs]);
        mbar_expect_tx(bar, SM::kKBytes);
        tma_load_3d(smem_u32(smem + SM::kVOff + vs * SM::kKBytes), &tmap_kv, bar, 2 * D + h * kTcDH, 0, b);
      }
      __syncwarp();
      if (++qs == SM::kQKSlots) { qs = 0; qph ^= 1u; }
      if (++vs == SM::kVSlots) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> query tile 0, warp 2 -> tile 1
    setmaxnreg_dec<kTc3RegsMma>();
    const int r = warp - 1;
    constexpr uint32_t idesc_qk = FMT ? umma_idesc_f16(128, NPAD) : umma_idesc_bf16(128, NPAD);
    constexpr uint32_t idesc_pv = FMT ? umma_idesc_f16(128, kTcDH, /*b_mn_major=*/1) : umma_idesc_bf16(128, kTcDH, /*b_mn_major=*/1);
    const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(r * kTcRegionCols);
    int qs = 0, vs = 0;
    uint32_t qph = 0, vph = 0, rph = 0;
    bool first = true;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_full[qs]), qph, p.flag, 0x2200u + qs))) break;
      tc3_trace<NPAD>(p, it, 4);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&s_free[r]), rph ^ 1u, p.flag, 0x2300u + r))) break;
      if (first && r == 1) {
        if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[0]), 0u, p.flag, 0x2310u))) break;
      }
      first = false;
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
      tc3_trace<NPAD>(p, it, 0);
      if (elect_one()) {                       // S_r = Q_r K^T
        const uint64_t a_desc = umma_desc_kmajor_sw128(base + r * kTcQTileBytes);
        const uint64_t b_desc = umma_desc_kmajor_sw128(base + 2 * kTcQTileBytes);
#pragma unroll
        for (int k = 0; k < kTcDH / 16; ++k)
          umma_bf16(s_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&s_full[r]));
        umma_commit(smem_u32(&qk_empty[qs]));
      }
      __syncwarp();
      tc3_trace<NPAD>(p, it, 1);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[r]), rph, p.flag, 0x2400u + r))) break;
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_full[vs]), vph, p.flag, 0x2410u + vs))) break;
      tcgen05_fence_after();
      tc3_trace<NPAD>(p, it, 2);
      if (elect_one()) {                       // O_r = P_r V; P of key group k sits where the half that owns it packed it
        const uint64_t v_desc = umma_desc_mnmajor_sw128(smem_u32(smem + SM::kVOff + vs * SM::kKBytes));
        const uint32_t o_tmem = s_tmem + kTc3OCol;
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const uint32_t p_col = k < G0 ? static_cast<uint32_t>(8 * k) : static_cast<uint32_t>(16 * G0 + 8 * (k - G0));
          umma_bf16_ts(o_tmem, s_tmem + p_col, v_desc + static_cast<uint64_t>(128 * k), idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&o_full[r]));
        umma_commit(smem_u32(&v_empty[vs]));
      }
      __syncwarp();
      tc3_trace<NPAD>(p, it, 3);
      rph ^= 1u;
      if (++qs == SM::kQKSlots) { qs = 0; qph ^= 1u; }
      if (++vs == SM::kVSlots) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 3) {
    setmaxnreg_dec<kTc3RegsOther>();
  } else if (warp < 20) {
    // ------------------------------------------------------------------ softmax warps: 8 per query tile = 4 lane quarters x 2 column halves
    setmaxnreg_inc<kTc3RegsSoftmax>();
    const int r = (warp - 4) >> 3;               // query tile / TMEM region
    const int half = ((warp - 4) >> 2) & 1;      // column half of the S row
    const int q = warp & 3;                      // TMEM lane quarter
    const bool warp_has_rows = r * 128 + q * 32 < n;          // warp-uniform, same for both halves of a quarter
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
    float* mx_mine = mx_part + (r * 2 + half) * 128 + q * 32 + lane;
    const float* mx_other = mx_part + (r * 2 + (half ^ 1)) * 128 + q * 32 + lane;
    float* sum_mine = sum_part + (r * 2 + half) * 128 + q * 32 + lane;
    const int bar_id = 1 + r * 4 + q;
    const float scale_log2 = p.scale_log2;
    uint32_t rph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      tc3_trace<NPAD>(p, it, 0);
      if (!mbar_wait(smem_u32(&s_full[r]), rph, p.flag, 0x2500u + r)) break;
      tcgen05_fence_after();
      tc3_trace<NPAD>(p, it, 1);
      if (warp_has_rows) {
        if (half == 0) tc3_softmax_item<NPAD, 0, POLY, FMT>(t_base, n, scale_log2, mx_mine, mx_other, sum_mine, bar_id, p, it);
        else tc3_softmax_item<NPAD, 1, POLY, FMT>(t_base, n, scale_log2, mx_mine, mx_other, sum_mine, bar_id, p, it);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&p_ready[r]));
        mbar_arrive(smem_u32(&sum_ready[r]));
      }
      tc3_trace<NPAD>(p, it, 3);
      rph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ output warps (one per TMEM lane quarter, both regions)
    setmaxnreg_inc<kTc3RegsOut>();
    const int q = warp & 3;
    uint8_t* stg = out_stage + q * (FMT ? 8192 : 4096);          // FMT 1: hi tile, then lo tile
    uint32_t rph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int it = (item - blockIdx.x) / gridDim.x;
      const int b = item / p.num_heads, h = item - b * p.num_heads;
      bool ok = true;
      for (int r = 0; r < 2; ++r) {
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
        const int row0 = r * 128 + q * 32;
        if (!mbar_wait(smem_u32(&sum_ready[r]), rph, p.flag, 0x2700u + r)) { ok = false; break; }
        if (!mbar_wait(smem_u32(&o_full[r]), rph, p.flag, 0x2600u + r)) { ok = false; break; }
        tcgen05_fence_after();
        tc3_trace<NPAD>(p, it, r * 3 + 0);
        const bool has_rows = row0 < n;          // warp-uniform
        const float inv = 1.0f / (sum_part[(r * 2 + 0) * 128 + q * 32 + lane] + sum_part[(r * 2 + 1) * 128 + q * 32 + lane]);
        // O row: both 32-column halves in flight at once, so the region is handed back after a single TMEM round trip
        uint32_t o0[32], o[32];
        if (has_rows) {
          tmem_ld_32x32(t_base + kTc3OCol, o0);
          tmem_ld_32x32(t_base + kTc3OCol + 32, o);
          tmem_ld_wait();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_free[r]));         // the region may take the next item's S
        if (has_rows && !(p.debug & 4)) {
          if (lane == 0) bulk_wait_read<0>();                     // the previous store has drained this staging tile
          __syncwarp();
          if constexpr (FMT == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                  make_uint4(pack_bf16(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                             pack_bf16(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                             pack_bf16(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                             pack_bf16(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + c) ^ (lane & 7)) << 4)) =
                  make_uint4(pack_bf16(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv),
                             pack_bf16(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv),
                             pack_bf16(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv),
                             pack_bf16(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv));
          } else {
            // hi / lo split of the scaled fp32 row: 16-byte chunk c of the hi tile and of the lo tile (4 KB further)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t* src = c < 4 ? &o0[8 * c] : &o[8 * (c - 4)];
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                split2_pack(__uint_as_float(src[2 * e]) * inv, __uint_as_float(src[2 * e + 1]) * inv, hi[e], lo[e]);
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(stg + 4096 + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          tc3_trace<NPAD>(p, it, r * 3 + 1);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if constexpr (FMT == 0) {
              tma_store_3d(&tmap_out, smem_u32(stg), h * kTcDH, row0, b);
            } else {
              tma_store_3d(&tmap_out, smem_u32(stg), D + h * kTcDH, row0, b);        // hi plane
              tma_store_3d(&tmap_out, smem_u32(stg) + 4096, h * kTcDH, row0, b);     // lo plane
            }
            bulk_commit();
          }
        }
        tc3_trace<NPAD>(p, it, r * 3 + 2);
      }
      if (!ok) break;
      rph ^= 1u;
    }
    if (lane == 0) bulk_wait_read<0>();       // smem must outlive the last store's reads
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ======================================================================================================================
// Ragged variant ("tcr"): the tcgen05/TMEM attention for packed rows of unequal sample lengths (cu_seqlens), per-key
// multiplicities and the virtual bias key -- what the compacted ResidualViT / A-ViT rows need (call sites
// reference models/residualvit.py:252-256, adavit.py:53-80 via blocks.py:93-95) -- and for short uniform sequences.
//
// One work unit = one 128-query tile of one (sample, head); units whose tile starts past the end of their sample are skipped by
// every role alike.  The k-th PROCESSED unit of a CTA runs in TMEM region k & 1, so the two regions always hold two different
// units (with samples of ~80 rows a second query tile never exists: tying both regions to one (sample, head) like the dense
// kernel would idle half the softmax warps).  Per unit the producer TMA-loads the Q tile and the sample's K / V rows through 2-D
// maps over the packed [rows, 3D] buffer: rows past the end of the sample belong to the next sample (or are zero-filled past
// the end of the buffer); they are neutralised by the per-key bias row below, never by the tensor map.
//   * key count per unit is a run-time value: n_keys = len (+1 virtual key), padded to 16 for the MMA's N and the loop bounds;
//   * warp 3 ("patch" warp) writes the virtual key (k-bias | v-bias of the head, what a zeroed token projects to) into row
//     `len` of the K and V tiles after the TMA has landed, and builds the unit's bias row  lm[j] = log2(mult_j)  (0 for a
//     plain key, log2(M_drop) for the virtual key, -inf for padding): multiplicities and masking are one FFMA in the softmax;
//   * softmax as in tc3 (two warps per TMEM lane quarter split the S row by 16-column groups), on x = s*scale*log2e + lm;
//   * output: full 32-row blocks leave through one 2-D TMA store, the block that straddles the end of the sample is written
//     row by row from registers (rows past the end belong to the next sample).
constexpr int kTcrThreads = 768;
template <int NMAX>
struct TcrSmem {
  static constexpr int kKBytes = NMAX * 128;
  static constexpr int kQKSlotBytes = kTcQTileBytes + kKBytes;            // one Q tile + K rows
  static constexpr int kQKSlots = 2;                                       // slot == region (units alternate regions)
  static constexpr int kVSlots = 3;
  static constexpr int kVOff = kQKSlots * kQKSlotBytes;
  static constexpr int kStageOff = kVOff + kVSlots * kKBytes;
  static constexpr int kLmOff = kStageOff + 4 * 4096;                      // bias rows: 4 x NMAX f32 (unit k uses row k & 3)
  static constexpr int kPartOff = kLmOff + 4 * 256 * 4;                    // row-max / row-sum partials: [2][region][half][128] f32
  static constexpr int kBarOff = kPartOff + 4096;
  static constexpr int kBytes = kBarOff + 512 + 1024;
  static_assert(NMAX % 16 == 0 && NMAX <= 256, "key tile");
  static_assert(kBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

struct TcrParams {
  const int* cu_seqlens;       // [batch + 1] or nullptr (uniform seq_len)
  const float* key_mult;       // [rows] or nullptr
  const __nv_bfloat16* extra_kv;   // [2D] or nullptr
  const float* extra_mult;     // [batch] or nullptr
  __nv_bfloat16* out;          // [rows, D]
  int batch, num_heads, seq_len, q_tiles;   // q_tiles = 128-row tiles per sample the grid is sized for
  int box_small;               // key rows of the small K / V box (the full box is NMAX)
  float scale_log2;
  unsigned int* flag;
  unsigned long long* trace;   // PK_ATT_TRACE=1: CTA 0 records clock64 at pipeline events of its first 16 units (tools/attn_trace_tcr.py)
  const int* route_rows;       // device-side routing: run only when *route_rows >= route_min_rows
  int route_min_rows;
  int min_keys, max_keys;      // unit filter: samples whose key count (rows + virtual key) lies outside are another launch's
};

// trace slot layout: [unit(0..15)][slot(0..15)][event(0..7)]; slots: warps 0-3 as they are, softmax warps 4 / 8 / 12 / 16 -> 4..7
// (region 0 half 0, region 0 half 1, region 1 half 0, region 1 half 1; lane quarter 0 each), output warps 20-23 -> 8..11
__device__ __forceinline__ void tcr_trace(const TcrParams& p, int k, int ev) {
#ifdef PK_ATT_TRACE_BUILD
  if (p.trace && blockIdx.x == 0 && k < 16 && (threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    const int slot = w < 4 ? w : (w >= 20 ? w - 12 : ((w & 3) == 0 ? 4 + ((w - 4) >> 2) : -1));
    if (slot >= 0) p.trace[(k * 16 + slot) * 8 + ev] = clock64();
  }
#endif
}

struct TcrUnit {
  int b, h, qt, row0, len, n_keys, npad;
  float extra;                 // multiplicity of the virtual key (0: none)
};
__device__ __forceinline__ bool tcr_unit(const TcrParams& p, int u, TcrUnit& t) {
  const int item = u / p.q_tiles;
  t.qt = u - item * p.q_tiles;
  t.b = item / p.num_heads;
  t.h = item - t.b * p.num_heads;
  if (p.cu_seqlens) {
    t.row0 = p.cu_seqlens[t.b];
    t.len = p.cu_seqlens[t.b + 1] - t.row0;
  } else {
    t.row0 = t.b * p.seq_len;
    t.len = p.seq_len;
  }
  if (t.qt * 128 >= t.len) return false;
  t.extra = (p.extra_kv && p.extra_mult) ? p.extra_mult[t.b] : 0.f;
  if (!(t.extra > 0.f)) t.extra = 0.f;
  t.n_keys = t.len + (t.extra > 0.f ? 1 : 0);
  t.npad = (t.n_keys + 15) & ~15;
  // per-sample split between the ragged kernels of one attention call: the quad-region kernel takes the samples of at most
  // 128 keys, the two-region / general kernel the longer ones; every role of a kernel skips the others' units alike
  return t.n_keys >= p.min_keys && t.n_keys <= p.max_keys;
}

// The groups of a unit that hold special keys (multiplicity != 1, the virtual key, padding): x = s * scale + lm[j] with the
// bias row read from shared memory by its shared-space address (the generic-pointer form costs an address computation and
// an LD.E per 4 keys).  Volatile without a memory clobber: ordered after the mbarrier wait, invisible to everything else.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float tcr_lm_max16(const uint32_t* s, uint32_t lm_addr, uint64_t sc2, float mx) {
  float xs[16];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 l = lds_f4(lm_addr + 16 * e);
    f2_unpack(f2_fma(f2_pack(__uint_as_float(s[4 * e]), __uint_as_float(s[4 * e + 1])), sc2, f2_pack(l.x, l.y)), xs[4 * e], xs[4 * e + 1]);
    f2_unpack(f2_fma(f2_pack(__uint_as_float(s[4 * e + 2]), __uint_as_float(s[4 * e + 3])), sc2, f2_pack(l.z, l.w)), xs[4 * e + 2], xs[4 * e + 3]);
  }
  const float m0 = fmax3(xs[0], xs[1], xs[2]), m1 = fmax3(xs[3], xs[4], xs[5]);
  const float m2 = fmax3(xs[6], xs[7], xs[8]), m3 = fmax3(xs[9], xs[10], xs[11]);
  return fmax3(mx, fmax3(fmax3(m0, xs[12], xs[13]), fmax3(m1, xs[14], xs[15]), m2), m3);
}
__device__ __forceinline__ float tcr_lm_group16(const uint32_t* s, uint32_t* pk, uint32_t lm_addr, uint64_t sc2, float neg_max) {
  const uint64_t nm2 = f2_pack(neg_max, neg_max);
  uint64_t acc = f2_pack(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 l = lds_f4(lm_addr + 16 * e);
    float x0, x1, x2, x3;
    f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(s[4 * e]), __uint_as_float(s[4 * e + 1])), sc2, f2_pack(l.x, l.y)), nm2), x0, x1);
    f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(s[4 * e + 2]), __uint_as_float(s[4 * e + 3])), sc2, f2_pack(l.z, l.w)), nm2), x2, x3);
    const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
    acc = f2_add(acc, f2_add(f2_pack(p0, p1), f2_pack(p2, p3)));
    pk[2 * e] = pack_bf16(p0, p1);
    pk[2 * e + 1] = pack_bf16(p2, p3);
  }
  float a0, a1;
  f2_unpack(acc, a0, a1);
  return a0 + a1;
}

template <int NMAX>
__global__ void __launch_bounds__(kTcrThreads, 1)
attention_tcr_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv_small,
                     const __grid_constant__ CUtensorMap tmap_kv_full, const __grid_constant__ CUtensorMap tmap_out, const TcrParams p) {
  if (p.route_rows && *p.route_rows < p.route_min_rows) return;      // short samples: the general kernel's launch takes this batch
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using SM = TcrSmem<NMAX>;
  uint8_t* out_stage = smem + SM::kStageOff;
  float* lm_rows = reinterpret_cast<float*>(smem + SM::kLmOff);                   // [4][256]
  float* mx_part = reinterpret_cast<float*>(smem + SM::kPartOff);                 // [region][half][128]
  float* sum_part = mx_part + 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* qk_full = bars;            // [2] TMA landed
  uint64_t* qk_ready = bars + 2;       // [2] virtual key + bias row written (patch warp)
  uint64_t* qk_empty = bars + 4;       // [2] QK MMA retired
  uint64_t* v_full = bars + 6;         // [3]
  uint64_t* v_ready = bars + 9;        // [3]
  uint64_t* v_empty = bars + 12;       // [3] PV MMA retired
  uint64_t* s_full = bars + 15;        // [2] region: S ready
  uint64_t* p_ready = bars + 17;       // [2] region: P written (8 softmax-warp arrivals)
  uint64_t* o_full = bars + 19;        // [2] region: O ready
  uint64_t* s_free = bars + 21;        // [2] region: O read out (4 output-warp arrivals)
  uint64_t* sum_ready = bars + 23;     // [2] region: row-sum partials published (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  int* plain_groups = reinterpret_cast<int*>(bars + 32);   // [4] leading 16-key groups of unit k & 3 whose bias is all zero

  const int warp = warp_id();
  const int lane = lane_id();
  const int D = p.num_heads * kTcDH;
  const int num_units = p.batch * p.num_heads * p.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv_small);
    tma_prefetch_desc(&tmap_kv_full);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_ready[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&qk_full[i]), 1);
      mbar_init(smem_u32(&qk_ready[i]), 1);
      mbar_init(smem_u32(&qk_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 8);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
      mbar_init(smem_u32(&sum_ready[i]), 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Every role walks the same unit sequence; k counts the units actually processed: region = QK slot = k & 1 (phase (k >> 1) & 1),
  // V slot = k % 3 (phase (k / 3) & 1), bias row = k & 3.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kTc3RegsOther>();
    int k = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const int qs = k & 1, vs = k % 3;
      const uint32_t qph = (k >> 1) & 1u, vph = (k / 3) & 1u;
      const bool small = t.npad <= p.box_small;
      const uint32_t kv_bytes = static_cast<uint32_t>(small ? p.box_small : NMAX) * 128u;
      const CUtensorMap* kvmap = small ? &tmap_kv_small : &tmap_kv_full;
      tcr_trace(p, k, 0);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_empty[qs]), qph ^ 1u, p.flag, 0x3100u + qs))) break;
      tcr_trace(p, k, 1);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&qk_full[qs]);
        const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
        mbar_expect_tx(bar, kTcQTileBytes + kv_bytes);
        tma_load_2d(base, &tmap_q, bar, t.h * kTcDH, t.row0 + t.qt * 128);
        tma_load_2d(base + kTcQTileBytes, kvmap, bar, D + t.h * kTcDH, t.row0);
      }
      __syncwarp();
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1u, p.flag, 0x3110u + vs))) break;
      tcr_trace(p, k, 2);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&v_full[vs]);
        mbar_expect_tx(bar, kv_bytes);
        tma_load_2d(smem_u32(smem + SM::kVOff + vs * SM::kKBytes), kvmap, bar, 2 * D + t.h * kTcDH, t.row0);
      }
      __syncwarp();
      ++k;
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> region 0 (even units), warp 2 -> region 1
    setmaxnreg_dec<kTc3RegsMma>();
    const int r = warp - 1;
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kTcDH, /*b_mn_major=*/1);
    const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(r * kTcRegionCols);
    int k = 0;
    bool first = true;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const int kk = k++;
      if ((kk & 1) != r) continue;
      const int qs = r, vs = kk % 3;
      const uint32_t qph = (kk >> 1) & 1u, vph = (kk / 3) & 1u, rph = qph;
      tcr_trace(p, kk, 0);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_ready[qs]), qph, p.flag, 0x3200u + qs))) break;
      tcr_trace(p, kk, 1);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&s_free[r]), rph ^ 1u, p.flag, 0x3300u + r))) break;
      tcr_trace(p, kk, 2);
      if (first && r == 1) {
        // start half a period after region 0 (the two softmax groups then use the MUFU alternately).  A scheduling hint, not a
        // dependency: bounded polling, because region 0 may legitimately be two phases ahead by the time this warp looks
        for (int i = 0; i < 4096 && !mbar_try_wait(smem_u32(&p_ready[0]), 0u); ++i) {}
      }
      first = false;
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + qs * SM::kQKSlotBytes);
      const uint32_t idesc_qk = umma_idesc_bf16(128, t.npad);
      if (elect_one()) {                       // S = Q K^T over the unit's padded key count
        const uint64_t a_desc = umma_desc_kmajor_sw128(base);
        const uint64_t b_desc = umma_desc_kmajor_sw128(base + kTcQTileBytes);
#pragma unroll
        for (int ks = 0; ks < kTcDH / 16; ++ks)
          umma_bf16(s_tmem, a_desc + static_cast<uint64_t>(2 * ks), b_desc + static_cast<uint64_t>(2 * ks), idesc_qk, ks != 0 ? 1u : 0u);
        umma_commit(smem_u32(&s_full[r]));
        umma_commit(smem_u32(&qk_empty[qs]));
      }
      __syncwarp();
      tcr_trace(p, kk, 3);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[r]), rph, p.flag, 0x3400u + r))) break;
      tcr_trace(p, kk, 4);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_ready[vs]), vph, p.flag, 0x3410u + vs))) break;
      tcr_trace(p, kk, 5);
      tcgen05_fence_after();
      const int G = t.npad >> 4, G0 = (G + 1) >> 1;
      if (elect_one()) {                       // O = P V; P of key group g sits where the half that owns it packed it
        const uint64_t v_desc = umma_desc_mnmajor_sw128(smem_u32(smem + SM::kVOff + vs * SM::kKBytes));
        const uint32_t o_tmem = s_tmem + kTc3OCol;
        for (int g = 0; g < G; ++g) {
          const uint32_t p_col = g < G0 ? static_cast<uint32_t>(8 * g) : static_cast<uint32_t>(16 * G0 + 8 * (g - G0));
          umma_bf16_ts(o_tmem, s_tmem + p_col, v_desc + static_cast<uint64_t>(128 * g), idesc_pv, g != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&o_full[r]));
        umma_commit(smem_u32(&v_empty[vs]));
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ patch warp: virtual key rows + bias row of every unit
    setmaxnreg_dec<kTc3RegsOther>();
    int k = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const int qs = k & 1, vs = k % 3;
      const uint32_t qph = (k >> 1) & 1u, vph = (k / 3) & 1u;
      float* lm = lm_rows + (k & 3) * 256;
      // bias row (its previous user, unit k - 4, finished its softmax before this unit's Q/K slot could even be refilled)
      int first_special = NMAX;                    // first key whose bias is not zero (multiplicity != 1, virtual key, padding)
      // all multiplicity loads in flight at once (one dependent L2 round trip per 32 keys made this warp, which every unit
      // passes through in order, the kernel's bottleneck: tools/attn_trace_tcq.py)
      constexpr int kIt = (NMAX + 31) / 32;
      float kmv[kIt];
#pragma unroll
      for (int i = 0; i < kIt; ++i) {
        const int j = lane + 32 * i;
        kmv[i] = (p.key_mult && j < t.len) ? p.key_mult[t.row0 + j] : 1.0f;
      }
#pragma unroll
      for (int i = 0; i < kIt; ++i) {
        const int j = lane + 32 * i;
        if (j < NMAX) {
          float v = -INFINITY;
          if (j < t.len) v = p.key_mult ? __log2f(kmv[i]) : 0.f;
          else if (j == t.len && t.extra > 0.f) v = __log2f(t.extra);
          lm[j] = v;
          if (v != 0.f && j < first_special) first_special = j;
        }
      }
      // Multiplicities other than 1 sit at the END of a sample (its ghost row, then the virtual key, then padding), so nearly
      // every 16-key group is "plain": the softmax takes x = s * scale there without reading the bias row.
      first_special = warp_min_i32(first_special);
      if (lane == 0) plain_groups[k & 3] = first_special >> 4;
      tcr_trace(p, k, 0);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_full[qs]), qph, p.flag, 0x3800u + qs))) break;
      tcr_trace(p, k, 1);
      if (t.extra > 0.f && lane < 8) {
        // K row `len`: 16-byte chunk c of row j lives at chunk c ^ (j & 7) of its 128-byte line (SWIZZLE_128B)
        const uint4 kb = *reinterpret_cast<const uint4*>(p.extra_kv + t.h * kTcDH + lane * 8);
        *reinterpret_cast<uint4*>(smem + qs * SM::kQKSlotBytes + kTcQTileBytes + t.len * 128 + ((lane ^ (t.len & 7)) << 4)) = kb;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&qk_ready[qs]));
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_full[vs]), vph, p.flag, 0x3810u + vs))) break;
      tcr_trace(p, k, 2);
      if (t.extra > 0.f && lane < 8) {
        const uint4 vb = *reinterpret_cast<const uint4*>(p.extra_kv + D + t.h * kTcDH + lane * 8);
        *reinterpret_cast<uint4*>(smem + SM::kVOff + vs * SM::kKBytes + t.len * 128 + ((lane ^ (t.len & 7)) << 4)) = vb;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&v_ready[vs]));
      ++k;
    }
  } else if (warp < 20) {
    // ------------------------------------------------------------------ softmax warps: 8 per region = 4 lane quarters x 2 column halves
    setmaxnreg_inc<kTc3RegsSoftmax>();
    const int r = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int q = warp & 3;
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
    float* mx_mine = mx_part + (r * 2 + half) * 128 + q * 32 + lane;
    const float* mx_other = mx_part + (r * 2 + (half ^ 1)) * 128 + q * 32 + lane;
    float* sum_mine = sum_part + (r * 2 + half) * 128 + q * 32 + lane;
    const int bar_id = 1 + r * 4 + q;
    const float scale_log2 = p.scale_log2;
    int k = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const int kk = k++;
      if ((kk & 1) != r) continue;
      const uint32_t rph = (kk >> 1) & 1u;
      const float* lm = lm_rows + (kk & 3) * 256;
      tcr_trace(p, kk, 0);
      if (!mbar_wait(smem_u32(&s_full[r]), rph, p.flag, 0x3500u + r)) break;
      tcr_trace(p, kk, 1);
      tcgen05_fence_after();
      const bool warp_has_rows = t.qt * 128 + q * 32 < t.len;           // warp-uniform, same for both halves of a quarter
      if (warp_has_rows) {
        const int G = t.npad >> 4, G0 = (G + 1) >> 1;
        const int gb = half ? G0 : 0, gn = half ? G - G0 : G0;          // this half's 16-column groups
        const uint32_t pcol0 = half ? static_cast<uint32_t>(16 * G0) : 0u;
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2);
        uint32_t ha[16], hb[16];
        // ---- pass 1: partial row maximum of x = s * scale * log2e + lm (padding columns carry lm = -inf).  Groups of plain
        // keys (lm == 0: all but the last one or two of a sample) take the maximum of the raw scores, scaled once at the end
        // (scale > 0), 32 columns per TMEM round trip like the dense kernel.
        const int n_plain = plain_groups[kk & 3];
        const int np = min(gn, max(n_plain - gb, 0));               // this half's leading plain groups
        const uint32_t lm_s = smem_u32(lm);
        float mx = -INFINITY, mraw = -INFINITY;
        int j = 0;
        for (; j + 2 <= np; j += 2) {
          uint32_t w[32];
          tmem_ld_32x32(t_base + static_cast<uint32_t>(16 * (gb + j)), w);
          tmem_ld_wait();
          mraw = row_max16<false>(w, mraw, 0, 0);
          mraw = row_max16<false>(w + 16, mraw, 0, 0);
        }
        for (; j < gn; ++j) {
          tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * (gb + j)), ha);
          tmem_ld_wait();
          if (j < np) mraw = row_max16<false>(ha, mraw, 0, 0);
          else mx = tcr_lm_max16(ha, lm_s + static_cast<uint32_t>(64 * (gb + j)), sc2, mx);
        }
        mx = fmaxf(mx, mraw * scale_log2);           // -inf * scale stays -inf
        *mx_mine = mx;
        tcr_trace(p, kk, 2);
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, *mx_other);
        tcr_trace(p, kk, 3);
        // ---- pass 2: p = exp2(x - max), partial row sum, bf16 P packed over this half's own consumed S columns
        float sum = 0.f;
        if (gn > 0) {
          tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * gb), ha);
          tmem_ld_wait();
        }
        tcr_trace(p, kk, 5);
        for (j = 0; j < gn; ++j) {
          uint32_t (&cur)[16] = (j & 1) ? hb : ha;
          uint32_t (&nxt)[16] = (j & 1) ? ha : hb;
          if (j + 1 < gn) tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * (gb + j + 1)), nxt);
          uint32_t pk[8];
          if (j < np) sum += softmax_group16<false, 0, 0>(cur, pk, scale_log2, -mx, 0, 0);
          else sum += tcr_lm_group16(cur, pk, lm_s + static_cast<uint32_t>(64 * (gb + j)), sc2, -mx);
          if (j + 1 < gn) tmem_ld_wait();        // the next group is in registers before its columns may be overwritten
          tmem_st_32x32_x8(t_base + pcol0 + static_cast<uint32_t>(8 * j), pk);
        }
        tcr_trace(p, kk, 7);
        *sum_mine = sum;
        tmem_st_wait();
      }
      tcr_trace(p, kk, 4);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&p_ready[r]));
        mbar_arrive(smem_u32(&sum_ready[r]));
      }
    }
  } else {
    // ------------------------------------------------------------------ output warps (one per TMEM lane quarter, both regions)
    setmaxnreg_inc<kTc3RegsOut>();
    const int q = warp & 3;
    uint8_t* stg = out_stage + q * 4096;
    int k = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const int kk = k++;
      const int r = kk & 1;
      const uint32_t rph = (kk >> 1) & 1u;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcRegionCols);
      tcr_trace(p, kk, 0);
      if (!mbar_wait(smem_u32(&sum_ready[r]), rph, p.flag, 0x3700u + r)) break;
      if (!mbar_wait(smem_u32(&o_full[r]), rph, p.flag, 0x3600u + r)) break;
      tcr_trace(p, kk, 1);
      tcgen05_fence_after();
      const int lrow0 = t.qt * 128 + q * 32;               // first row of this warp's block inside the sample
      const bool has_rows = lrow0 < t.len;                 // warp-uniform
      const bool full = lrow0 + 32 <= t.len;
      uint32_t o0[32], o[32];
      float inv = 0.f;
      if (has_rows) {
        inv = 1.0f / (sum_part[(r * 2 + 0) * 128 + q * 32 + lane] + sum_part[(r * 2 + 1) * 128 + q * 32 + lane]);
        tmem_ld_32x32(t_base + kTc3OCol, o0);
        tmem_ld_32x32(t_base + kTc3OCol + 32, o);
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[r]));     // the region may take its next unit's S
      tcr_trace(p, kk, 2);
      if (!has_rows) continue;
      auto chunk = [&](int c) {                             // 8 scaled bf16 values = 16-byte chunk c of this thread's row
        const uint32_t* src = c < 4 ? &o0[8 * c] : &o[8 * (c - 4)];
        return make_uint4(pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv),
                          pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv),
                          pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv),
                          pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv));
      };
      if (full) {
        if (lane == 0) bulk_wait_read<0>();                 // the previous store has drained this staging tile
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = chunk(c);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_out, smem_u32(stg), t.h * kTcDH, t.row0 + lrow0);
          bulk_commit();
        }
      } else if (lrow0 + lane < t.len) {
        // the block that straddles the end of the sample: the rows after it are the next sample's
        uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long long>(t.row0 + lrow0 + lane) * D + t.h * kTcDH);
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[c] = chunk(c);
      }
      tcr_trace(p, kk, 3);
    }
    if (lane == 0) bulk_wait_read<0>();       // smem must outlive the last store's reads
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ======================================================================================================================
// Quad-region variant ("tcq") for the samples of at most 128 keys (most of a ResidualViT batch at budgets <= ~0.5, the late
// A-ViT layers): the ragged kernel above keeps two units in flight per SM, and its per-unit chain (S ready -> row max -> exp
// -> P.V -> O read-out -> next Q.K^T, ~10 000 cycles) is the same whether a sample has 80 rows or 200 -- short samples cannot
// amortise it.  With at most 128 keys a TMEM region needs only 128 columns (S at [0, npad), bf16 P packed over the consumed
// S columns at [0, npad / 2), O at [64, 128)), so FOUR units run concurrently: unit k of a CTA uses region / Q-K slot / V slot
// k & 3.  One MMA-issuing warp per region (on four different schedulers), one softmax + output warp per TMEM lane quarter and region
// (no column split, hence no partial maxima to exchange) -- the issuing warp also writes its region's bias row and virtual
// key rows; every softmax warp also reads its rows of O out, scales them and stores them (staged in the region's V slot).
// Units = one (sample, head); q_tiles == 1.  Samples with more than 128 keys are skipped (TcrParams.max_keys) and taken by the
// two-region / general kernel launched next to this one, which in turn skip the short ones: a per-sample split on the device.
constexpr int kTcqThreads = 768;                   // warps: 0 Q/K TMA, 1-4 MMA + patch (region 0-3), 6 V TMA, 5 / 7 idle, 8-23 softmax + output
constexpr int kTcqNMax = 128;
constexpr int kTcqRegionCols = 128;
constexpr int kTcqOCol = 64;
constexpr int kTcqRegsSoftmax = 96, kTcqRegsMma = 56, kTcqRegsOther = 40;
// The pool is what the CTA was LAUNCHED with (768 threads x 80 registers = 61 440), not the SM's 65 536: a setmaxnreg.inc
// that asks for more than the other warps have released blocks for ever (no watchdog can see it).
// (16 x 96 + 4 x 56 + 4 x 40) warps x 32 = 61 440 registers
static_assert((16 * kTcqRegsSoftmax + 4 * kTcqRegsMma + 4 * kTcqRegsOther) * 32 <= kTcqThreads * 80, "setmaxnreg budget exceeds the launch allocation");
struct TcqSmem {
  static constexpr int kKBytes = kTcqNMax * 128;
  static constexpr int kQKSlotBytes = kTcQTileBytes + kKBytes;            // one Q tile + K rows: 32 KB
  static constexpr int kVOff = 4 * kQKSlotBytes;                           // 128 KB
  static constexpr int kLmOff = kVOff + 4 * kKBytes;                       // + 64 KB (the V slots double as output staging)
  static constexpr int kBarOff = kLmOff + 4 * kTcqNMax * 4;                // bias rows: 4 x 128 f32
  static constexpr int kBytes = kBarOff + 512 + 1024;
  static_assert(kBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

// trace slots: warps 0-7 as they are, softmax warps 8 / 12 / 16 / 20 (lane quarter 0 of region 0-3) -> 8..11, output warps -> 12..15
__device__ __forceinline__ void tcq_trace(const TcrParams& p, int k, int ev) {
#ifdef PK_ATT_TRACE_BUILD
  if (p.trace && blockIdx.x == 0 && k >= 8 && k < 24 && (threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    const int slot = w < 8 ? w : (w >= 24 ? w - 12 : ((w & 3) == 0 ? 8 + ((w - 8) >> 2) : -1));
    if (slot >= 0) p.trace[((k - 8) * 16 + slot) * 8 + ev] = clock64();
  }
#endif
}
__global__ void __launch_bounds__(kTcqThreads, 1)
attention_tcq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv_small,
                     const __grid_constant__ CUtensorMap tmap_kv_full, const __grid_constant__ CUtensorMap tmap_out, const TcrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using SM = TcqSmem;
  float* lm_rows = reinterpret_cast<float*>(smem + SM::kLmOff);                   // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* qk_full = bars;            // [4] TMA landed
  uint64_t* qk_ready = bars + 4;       // [4] virtual key + bias row written (patch warp)
  uint64_t* qk_empty = bars + 8;       // [4] QK MMA retired
  uint64_t* v_full = bars + 12;        // [4]
  uint64_t* v_ready = bars + 16;       // [4]
  uint64_t* v_empty = bars + 20;       // [4] PV MMA retired
  uint64_t* s_full = bars + 24;        // [4] S ready
  uint64_t* p_ready = bars + 28;       // [4] P written + row sums published (4 softmax-warp arrivals)
  uint64_t* o_full = bars + 32;        // [4] O ready
  uint64_t* s_free = bars + 36;        // [4] O read out (4 output-warp arrivals)
  uint64_t* sum_ready = bars + 40;     // [4] row sums published (4 softmax-warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
  int* plain_groups = reinterpret_cast<int*>(bars + 48);   // [4] leading 16-key groups of unit k & 3 whose bias is all zero

  const int warp = warp_id();
  const int lane = lane_id();
  const int D = p.num_heads * kTcDH;
  const int num_units = p.batch * p.num_heads;             // q_tiles == 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv_small);
    tma_prefetch_desc(&tmap_kv_full);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&qk_full[i]), 1);
      mbar_init(smem_u32(&qk_ready[i]), 1);
      mbar_init(smem_u32(&qk_empty[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_ready[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 4);             // the region's four softmax / output warps, after their stores
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 4);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Every role walks the same unit sequence; the k-th unit of this CTA uses region = slot = bias row k & 3.  Empty samples are
  // skipped by every role alike (their position still counts), so barrier parities are kept per region, not derived from k.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kTcqRegsOther>();
    int k = 0;
    uint32_t phases = 0, used = 0;        // per region: parity of its next handshake / whether it has run a unit yet
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++k) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;   // empty sample (an A-ViT sample whose class token has halted): every role skips it
      const int r = k & 3;
      const uint32_t ph = (phases >> r) & 1u;
      phases ^= 1u << r;
      const bool first_use = !((used >> r) & 1u);
      used |= 1u << r;
      const bool small = t.npad <= p.box_small;
      const uint32_t kv_bytes = static_cast<uint32_t>(small ? p.box_small : kTcqNMax) * 128u;
      const CUtensorMap* kvmap = small ? &tmap_kv_small : &tmap_kv_full;
      tcq_trace(p, k, 0);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_empty[r]), ph ^ 1u, p.flag, 0x4100u + r))) break;
      tcq_trace(p, k, 1);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&qk_full[r]);
        const uint32_t base = smem_u32(smem + r * SM::kQKSlotBytes);
        mbar_expect_tx(bar, kTcQTileBytes + kv_bytes);
        tma_load_2d(base, &tmap_q, bar, t.h * kTcDH, t.row0);
        tma_load_2d(base + kTcQTileBytes, kvmap, bar, D + t.h * kTcDH, t.row0);
      }
      __syncwarp();
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------ V producer (its own warp: a V slot frees up at the END of
    // its previous unit's chain, a Q / K slot right after that unit's Q K^T -- one in-order producer would hold the next
    // regions' Q / K loads back behind this unit's V wait)
    setmaxnreg_dec<kTcqRegsOther>();
    int k = 0;
    uint32_t phases = 0, used = 0;        // per region: parity of its next handshake / whether it has run a unit yet
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++k) {
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;   // empty sample (an A-ViT sample whose class token has halted): every role skips it
      const int r = k & 3;
      const uint32_t ph = (phases >> r) & 1u;
      phases ^= 1u << r;
      const bool first_use = !((used >> r) & 1u);
      used |= 1u << r;
      const bool small = t.npad <= p.box_small;
      const uint32_t kv_bytes = static_cast<uint32_t>(small ? p.box_small : kTcqNMax) * 128u;
      const CUtensorMap* kvmap = small ? &tmap_kv_small : &tmap_kv_full;
      tcq_trace(p, k, 0);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_empty[r]), ph ^ 1u, p.flag, 0x4110u + r))) break;
      tcq_trace(p, k, 1);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&v_full[r]);
        mbar_expect_tx(bar, kv_bytes);
        tma_load_2d(smem_u32(smem + SM::kVOff + r * SM::kKBytes), kvmap, bar, 2 * D + t.h * kTcDH, t.row0);
      }
      __syncwarp();
    }
  } else if (warp >= 1 && warp <= 4) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 + r -> region r
    setmaxnreg_dec<kTcqRegsMma>();
    const int r = warp - 1;
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kTcDH, /*b_mn_major=*/1);
    const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(r * kTcqRegionCols);
    const uint32_t qk_base = smem_u32(smem + r * SM::kQKSlotBytes);
    const uint64_t a_desc = umma_desc_kmajor_sw128(qk_base);
    const uint64_t b_desc = umma_desc_kmajor_sw128(qk_base + kTcQTileBytes);
    const uint64_t v_desc = umma_desc_mnmajor_sw128(smem_u32(smem + SM::kVOff + r * SM::kKBytes));
    uint32_t ph_next = 0;
    int k = r - 4;
    for (int u = blockIdx.x + r * gridDim.x; u < num_units; u += 4 * gridDim.x) {
      k += 4;
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const uint32_t ph = ph_next;
      ph_next ^= 1u;
      // ---- this unit's bias row and virtual key, by the region's own issuing warp (a shared in-order patch warp was the
      // bottleneck of the whole kernel: 4 500 cycles per unit, every region queueing behind it).  The region's previous
      // softmax is over (its p_ready was waited for below), so the bias row is free; all multiplicity loads go out at once.
      tcq_trace(p, k, 0);
      {
        float kmv[kTcqNMax / 32];
#pragma unroll
        for (int i = 0; i < kTcqNMax / 32; ++i) {
          const int j = lane + 32 * i;
          kmv[i] = (p.key_mult && j < t.len) ? p.key_mult[t.row0 + j] : 1.0f;
        }
        float* lm = lm_rows + r * kTcqNMax;
        int first_special = kTcqNMax;
#pragma unroll
        for (int i = 0; i < kTcqNMax / 32; ++i) {
          const int j = lane + 32 * i;
          float v = -INFINITY;
          if (j < t.len) v = p.key_mult ? __log2f(kmv[i]) : 0.f;
          else if (j == t.len && t.extra > 0.f) v = __log2f(t.extra);
          lm[j] = v;
          if (v != 0.f && j < first_special) first_special = j;
        }
        first_special = warp_min_i32(first_special);
        if (lane == 0) plain_groups[r] = first_special >> 4;
      }
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&qk_full[r]), ph, p.flag, 0x4200u + r))) break;
      tcq_trace(p, k, 1);
      if (t.extra > 0.f && lane < 8) {
        // K row `len`: 16-byte chunk c of row j lives at chunk c ^ (j & 7) of its 128-byte line (SWIZZLE_128B)
        const uint4 kb = *reinterpret_cast<const uint4*>(p.extra_kv + t.h * kTcDH + lane * 8);
        *reinterpret_cast<uint4*>(smem + r * SM::kQKSlotBytes + kTcQTileBytes + t.len * 128 + ((lane ^ (t.len & 7)) << 4)) = kb;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&qk_ready[r]));          // publishes the bias row to the softmax warps
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&s_free[r]), ph ^ 1u, p.flag, 0x4300u + r))) break;
      tcq_trace(p, k, 2);
      tcgen05_fence_after();
      const uint32_t idesc_qk = umma_idesc_bf16(128, t.npad);
      if (elect_one()) {                       // S = Q K^T over the unit's padded key count
#pragma unroll
        for (int ks = 0; ks < kTcDH / 16; ++ks)
          umma_bf16(s_tmem, a_desc + static_cast<uint64_t>(2 * ks), b_desc + static_cast<uint64_t>(2 * ks), idesc_qk, ks != 0 ? 1u : 0u);
        umma_commit(smem_u32(&s_full[r]));
        umma_commit(smem_u32(&qk_empty[r]));
      }
      __syncwarp();
      tcq_trace(p, k, 3);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&v_full[r]), ph, p.flag, 0x4410u + r))) break;
      if (t.extra > 0.f && lane < 8) {
        const uint4 vb = *reinterpret_cast<const uint4*>(p.extra_kv + D + t.h * kTcDH + lane * 8);
        *reinterpret_cast<uint4*>(smem + SM::kVOff + r * SM::kKBytes + t.len * 128 + ((lane ^ (t.len & 7)) << 4)) = vb;
      }
      fence_proxy_async_smem();
      __syncwarp();
      tcq_trace(p, k, 4);
      if (!__all_sync(0xffffffffu, mbar_wait(smem_u32(&p_ready[r]), ph, p.flag, 0x4400u + r))) break;
      tcq_trace(p, k, 5);
      tcgen05_fence_after();
      const int G = t.npad >> 4;
      if (elect_one()) {                       // O = P V; the probabilities of key group g sit at region columns [8g, 8g + 8)
        const uint32_t o_tmem = s_tmem + kTcqOCol;
#pragma unroll
        for (int g = 0; g < kTcqNMax / 16; ++g)
          if (g < G) umma_bf16_ts(o_tmem, s_tmem + static_cast<uint32_t>(8 * g), v_desc + static_cast<uint64_t>(128 * g), idesc_pv, g != 0 ? 1u : 0u);
        umma_commit(smem_u32(&o_full[r]));
      }
      __syncwarp();
      tcq_trace(p, k, 6);
    }
  } else if (warp == 5 || warp == 7) {
    setmaxnreg_dec<kTcqRegsOther>();          // idle
  } else {
    // ------------------------------------------------------------------ softmax + output warps: one per region and TMEM lane quarter
    setmaxnreg_inc<kTcqRegsSoftmax>();
    const int r = (warp - 8) >> 2;
    const int q = warp & 3;
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(r * kTcqRegionCols);
    const float scale_log2 = p.scale_log2;
    const uint32_t lm_s = smem_u32(lm_rows + r * kTcqNMax);
    uint32_t ph_next = 0;
    int k = r - 4;
    for (int u = blockIdx.x + r * gridDim.x; u < num_units; u += 4 * gridDim.x) {
      k += 4;
      TcrUnit t;
      if (!tcr_unit(p, u, t)) continue;
      const uint32_t ph = ph_next;
      ph_next ^= 1u;
      tcq_trace(p, k, 0);
      if (!mbar_wait(smem_u32(&qk_ready[r]), ph, p.flag, 0x4510u + r)) break;      // the bias row of this unit is published
      if (!mbar_wait(smem_u32(&s_full[r]), ph, p.flag, 0x4500u + r)) break;
      tcq_trace(p, k, 1);
      tcgen05_fence_after();
      const bool has_rows = q * 32 < t.len;              // warp-uniform: this lane quarter holds query rows
      float row_sum = 1.0f;
      if (has_rows) {
        const int G = t.npad >> 4;
        const int np = min(G, plain_groups[r]);           // leading groups of plain keys
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2);
        uint32_t ha[16], hb[16];
        // ---- pass 1: row maximum of x = s * scale * log2e + lm
        float mx = -INFINITY, mraw = -INFINITY;
        int j = 0;
        for (; j + 2 <= np; j += 2) {
          uint32_t w[32];
          tmem_ld_32x32(t_base + static_cast<uint32_t>(16 * j), w);
          tmem_ld_wait();
          mraw = row_max16<false>(w, mraw, 0, 0);
          mraw = row_max16<false>(w + 16, mraw, 0, 0);
        }
        for (; j < G; ++j) {
          tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * j), ha);
          tmem_ld_wait();
          if (j < np) mraw = row_max16<false>(ha, mraw, 0, 0);
          else mx = tcr_lm_max16(ha, lm_s + static_cast<uint32_t>(64 * j), sc2, mx);
        }
        mx = fmaxf(mx, mraw * scale_log2);
        tcq_trace(p, k, 2);
        // ---- pass 2: p = exp2(x - max), row sum, bf16 P packed over the S columns already consumed
        float sum = 0.f;
        tmem_ld_32x32_x16(t_base, ha);
        tmem_ld_wait();
        for (j = 0; j < G; ++j) {
          uint32_t (&cur)[16] = (j & 1) ? hb : ha;
          uint32_t (&nxt)[16] = (j & 1) ? ha : hb;
          if (j + 1 < G) tmem_ld_32x32_x16(t_base + static_cast<uint32_t>(16 * (j + 1)), nxt);
          uint32_t pk[8];
          if (j < np) sum += softmax_group16<false, 0, 0>(cur, pk, scale_log2, -mx, 0, 0);
          else sum += tcr_lm_group16(cur, pk, lm_s + static_cast<uint32_t>(64 * j), sc2, -mx);
          if (j + 1 < G) tmem_ld_wait();                 // the next group is in registers before its columns may be overwritten
          tmem_st_32x32_x8(t_base + static_cast<uint32_t>(8 * j), pk);
        }
        tmem_st_wait();
        row_sum = sum;
      }
      tcq_trace(p, k, 3);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_ready[r]));
      // ---- output of the same rows by the same warp (a shared set of output warps took ~3 000 cycles per unit, every
      // region queueing behind it): O out of TMEM, scaled by 1 / row sum, staged in the region's V slot -- dead once P.V has
      // retired, and not needed again before the next unit's softmax is over -- and stored by TMA (full 32-row blocks) or row
      // by row (the block that straddles the end of the sample)
      if (!mbar_wait(smem_u32(&o_full[r]), ph, p.flag, 0x4600u + r)) break;
      tcgen05_fence_after();
      const bool full = q * 32 + 32 <= t.len;
      uint32_t pk[32];                                       // this thread's row of O, scaled, as 32 bf16 pairs
      if (has_rows) {
        const float inv = 1.0f / row_sum;
        uint32_t raw[32];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {                  // two 32-column reads: the raw fp32 half is dead before the next
          tmem_ld_32x32(t_base + kTcqOCol + 32 * hlf, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[16 * hlf + i] = pack_bf16(__uint_as_float(raw[2 * i]) * inv, __uint_as_float(raw[2 * i + 1]) * inv);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[r]));     // the region may take its next unit's S
      if (has_rows) {
        auto chunk = [&](int c) { return make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]); };
        if (full) {
          uint8_t* stg = smem + SM::kVOff + r * SM::kKBytes + q * 4096;
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = chunk(c);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, smem_u32(stg), t.h * kTcDH, t.row0 + q * 32);
            bulk_commit();
            bulk_wait_read<0>();                            // the V slot may be refilled once the store has read it
          }
        } else if (q * 32 + lane < t.len) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long long>(t.row0 + q * 32 + lane) * D + t.h * kTcDH);
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = chunk(c);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&v_empty[r]));
      tcq_trace(p, k, 4);
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 3-D view of the packed qkv buffer: (column, token in sample, sample)
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld_elems,
                      uint32_t box_rows);

// PK_ATT_TRACE=1: a 16 x 12 x 8 table of clock64 stamps written by CTA 0 (debug; read back with pk_attention_trace()).
static unsigned long long* g_tc_trace = nullptr;
static unsigned long long* tc_trace_buffer() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PK_ATT_TRACE");
    on = (e && e[0] == '1') ? 1 : 0;
    if (on) {
      if (cudaMalloc(&g_tc_trace, 16 * 16 * 8 * sizeof(unsigned long long)) != cudaSuccess) g_tc_trace = nullptr;
      else cudaMemset(g_tc_trace, 0, 16 * 16 * 8 * sizeof(unsigned long long));
    }
  }
  return g_tc_trace;
}
int attention_trace_copy(unsigned long long* host_dst) {
  if (!g_tc_trace) return PK_ERR_INVALID;
  return check_cuda(cudaMemcpy(host_dst, g_tc_trace, 16 * 16 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "trace copy");
}

bool attention_tc_eligible(const pk_attention_args* a) {
  if (a->impl == 1 || a->impl == 3) return false;
  if (a->cu_seqlens || a->key_mult || a->extra_kv || a->extra_mult) return false;
  if (a->head_dim != kTcDH) return false;
  const bool x2 = a->qkv_format == PK_OUT_F16 || a->out_format == PK_OUT_BF16X2;
  if (x2) {
    // the bf16x2 variant takes half operands AND writes the split output (one kernel variant), on the column-split kernel
    if (a->qkv_format != PK_OUT_F16 || a->out_format != PK_OUT_BF16X2) return false;
    if (a->seq_len < 17 || a->seq_len > 224) return false;           // two staging tiles per output warp: n_pad <= 224
    return (reinterpret_cast<uintptr_t>(a->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0;
  }
  // n <= 128 runs with the second query tile empty (its region only keeps the barrier protocol alive); PK_ATT_TC_MIN_SEQ
  // (default 65: measured 92 vs 138 us at n = 99, but 68 vs 54 us at n = 50, B = 512, H = 12) is the shortest sequence routed here
  static int min_seq = -1;
  if (min_seq < 0) { const char* e = getenv("PK_ATT_TC_MIN_SEQ"); min_seq = e ? atoi(e) : 65; if (min_seq < 17) min_seq = 17; }
  if (a->impl != 2 && a->seq_len < min_seq) return false;
  if (a->seq_len < 17 || a->seq_len > kTcMaxKeys) return false;
  if ((reinterpret_cast<uintptr_t>(a->qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) return false;
  return true;
}

int launch_attention_tc(const pk_attention_args* a, cudaStream_t stream) {
  const int D = a->num_heads * kTcDH;
  const int n = a->seq_len;
  const int n_pad = (n + 15) / 16 * 16;
  CUtensorMap tq, tkv;
  int rc = make_tmap_bf16_3d(&tq, a->qkv, 3ull * D, static_cast<uint64_t>(n), static_cast<uint64_t>(a->batch), 3ull * D, 128);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_3d(&tkv, a->qkv, 3ull * D, static_cast<uint64_t>(n), static_cast<uint64_t>(a->batch), 3ull * D,
                         static_cast<uint32_t>(n_pad));
  if (rc != PK_OK) return rc;
  CUtensorMap tout;
  const bool x2 = a->out_format == PK_OUT_BF16X2;
  rc = make_tmap_bf16_3d(&tout, a->out, static_cast<uint64_t>(x2 ? 2 * D : D), static_cast<uint64_t>(n), static_cast<uint64_t>(a->batch),
                         static_cast<uint64_t>(x2 ? 2 * D : D), 32);
  if (rc != PK_OK) return rc;
  TcAttParams p;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.batch = a->batch;
  p.num_heads = a->num_heads;
  p.seq_len = n;
  p.n_pad = n_pad;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.flag = device_flag_ptr();
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("PK_ATT_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.trace = tc_trace_buffer();
  const long long items = static_cast<long long>(a->batch) * a->num_heads;
  int grid = num_sms();
  if (items < grid) grid = static_cast<int>(items);
  // PK_ATT_TC_SPLIT=0 selects the 8-softmax-warp kernel (one thread per S row) instead of the column-split one
  static int split = -1;
  if (split < 0) { const char* e = getenv("PK_ATT_TC_SPLIT"); split = e ? atoi(e) : 1; }
  static int poly = -1;        // experiment: PK_ATT_TC_POLY=1|2 of every 4 column pairs take the FMA-pipe exp2 (n_pad 208 only)
  if (poly < 0) { const char* e = getenv("PK_ATT_TC_POLY"); poly = e ? atoi(e) : 0; }
  if (x2) {
    switch (n_pad) {
#define PK_TC3X_CASE(NP)                                                                                                \
  case NP: {                                                                                                            \
    static bool attr_set = false;                                                                                       \
    if (!attr_set) {                                                                                                    \
      PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel<NP, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc3Smem<NP, 1>::kBytes)); \
      attr_set = true;                                                                                                  \
    }                                                                                                                   \
    attention_tc3_kernel<NP, 0, 1><<<grid, kTc3Threads, Tc3Smem<NP, 1>::kBytes, stream>>>(tq, tkv, tout, p);              \
    break;                                                                                                              \
  }
      PK_TC3X_CASE(32) PK_TC3X_CASE(48) PK_TC3X_CASE(64) PK_TC3X_CASE(80) PK_TC3X_CASE(96) PK_TC3X_CASE(112) PK_TC3X_CASE(128)
      PK_TC3X_CASE(144) PK_TC3X_CASE(160) PK_TC3X_CASE(176) PK_TC3X_CASE(192) PK_TC3X_CASE(208) PK_TC3X_CASE(224)
#undef PK_TC3X_CASE
      default:
        set_last_error("pk_attention_fwd: unsupported padded length %d for the half-operand variant", n_pad);
        return PK_ERR_INVALID;
    }
    return check_cuda(cudaGetLastError(), "attention_tc3_kernel<half> launch");
  }
  if (split && n_pad == 208 && (poly == 1 || poly == 2)) {
    static bool attr_set = false;
    if (!attr_set) {
      PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel<208, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc3Smem<208>::kBytes));
      PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel<208, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc3Smem<208>::kBytes));
      attr_set = true;
    }
    if (poly == 1) attention_tc3_kernel<208, 1><<<grid, kTc3Threads, Tc3Smem<208>::kBytes, stream>>>(tq, tkv, tout, p);
    else attention_tc3_kernel<208, 2><<<grid, kTc3Threads, Tc3Smem<208>::kBytes, stream>>>(tq, tkv, tout, p);
    return check_cuda(cudaGetLastError(), "attention_tc3_kernel launch");
  }
  if (split) {
    switch (n_pad) {
#define PK_TC3_CASE(NP)                                                                                                 \
  case NP: {                                                                                                            \
    static bool attr_set = false;                                                                                       \
    if (!attr_set) {                                                                                                    \
      PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc3Smem<NP>::kBytes)); \
      attr_set = true;                                                                                                  \
    }                                                                                                                   \
    attention_tc3_kernel<NP><<<grid, kTc3Threads, Tc3Smem<NP>::kBytes, stream>>>(tq, tkv, tout, p);                       \
    break;                                                                                                              \
  }
      PK_TC3_CASE(32) PK_TC3_CASE(48) PK_TC3_CASE(64) PK_TC3_CASE(80) PK_TC3_CASE(96) PK_TC3_CASE(112) PK_TC3_CASE(128)
      PK_TC3_CASE(144) PK_TC3_CASE(160) PK_TC3_CASE(176) PK_TC3_CASE(192) PK_TC3_CASE(208) PK_TC3_CASE(224) PK_TC3_CASE(240) PK_TC3_CASE(256)
#undef PK_TC3_CASE
      default:
        set_last_error("pk_attention_fwd: unsupported padded length %d", n_pad);
        return PK_ERR_INVALID;
    }
    return check_cuda(cudaGetLastError(), "attention_tc3_kernel launch");
  }
  switch (n_pad) {
#define PK_TC_CASE(NP)                                                                                                  \
  case NP: {                                                                                                            \
    static bool attr_set = false;                                                                                       \
    if (!attr_set) {                                                                                                    \
      PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<NP>::kBytes)); \
      attr_set = true;                                                                                                  \
    }                                                                                                                   \
    attention_tc_kernel<NP><<<grid, kTcThreads, TcSmem<NP>::kBytes, stream>>>(tq, tkv, tout, p);                              \
    break;                                                                                                              \
  }
    PK_TC_CASE(32) PK_TC_CASE(48) PK_TC_CASE(64) PK_TC_CASE(80) PK_TC_CASE(96) PK_TC_CASE(112) PK_TC_CASE(128)
    PK_TC_CASE(144) PK_TC_CASE(160) PK_TC_CASE(176) PK_TC_CASE(192) PK_TC_CASE(208) PK_TC_CASE(224) PK_TC_CASE(240) PK_TC_CASE(256)
#undef PK_TC_CASE
    default:
      set_last_error("pk_attention_fwd: unsupported padded length %d", n_pad);
      return PK_ERR_INVALID;
  }
  return check_cuda(cudaGetLastError(), "attention_tc_kernel launch");
}

// ---------------------------------------------------------------------------------------------------- ragged kernel launch
// Eligible: head_dim 64 and at most 256 keys per sample (longest sample + the virtual key); ragged (cu_seqlens), with key
// multiplicities / a virtual key, or uniform sequences the dense kernels do not take (seq_len <= 64).  PK_ATT_TCR=0 routes
// everything back to the general mma.sync kernel (A/B runs).
static int tcr_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PK_ATT_TCR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}
// PK_ATT_TCR_UNIFORM_MAX: longest UNIFORM plain sequence routed to the ragged kernel instead of the general one.  Default 0:
// on the pruned RankViT layers (50 / 33 / 26 / 14 tokens) the general kernel is 2x faster (profiles/r02/run16_attn_ragged.json:
// 49.9 vs 109.8 us at 50 tokens), and the dense tc3 kernel takes 65 .. 256.
static int tcr_uniform_max() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PK_ATT_TCR_UNIFORM_MAX"); v = e ? atoi(e) : 0; }
  return v;
}
bool attention_tcr_eligible(const pk_attention_args* a) {
  if (a->impl == 1 || !tcr_enabled()) return false;
  if (a->head_dim != kTcDH) return false;
  if (a->qkv_format != PK_OUT_BF16 || a->out_format != PK_OUT_BF16) return false;
  const bool ragged = a->cu_seqlens || a->key_mult || a->extra_kv;
  const int max_len = a->cu_seqlens ? a->max_seq_len : a->seq_len;
  if (max_len < 1 || max_len + (a->extra_kv ? 1 : 0) > 256) return false;
  if (!ragged && max_len > tcr_uniform_max() && a->impl != 3) return false;
  if ((reinterpret_cast<uintptr_t>(a->qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) return false;
  if (a->extra_kv && (reinterpret_cast<uintptr_t>(a->extra_kv) & 15) != 0) return false;
  return a->total_rows > 0;
}

bool attention_tcq_eligible(const pk_attention_args* a);
bool attention_split_active(const pk_attention_args* a);

template <int NMAX>
static int launch_tcr(const pk_attention_args* a, cudaStream_t stream, int max_len) {
  const int D = a->num_heads * kTcDH;
  const uint64_t rows = static_cast<uint64_t>(a->total_rows);
  int box_small = ((NMAX / 2) + 15) & ~15;
  if (box_small < 16) box_small = 16;
  CUtensorMap tq, tks, tkf, tout;
  int rc = make_tmap_bf16_2d(&tq, a->qkv, rows, 3ull * D, 3ull * D, 128, 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tks, a->qkv, rows, 3ull * D, 3ull * D, static_cast<uint32_t>(box_small), 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tkf, a->qkv, rows, 3ull * D, 3ull * D, NMAX, 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tout, a->out, rows, static_cast<uint64_t>(D), static_cast<uint64_t>(D), 32, 64);
  if (rc != PK_OK) return rc;
  TcrParams p;
  p.cu_seqlens = a->cu_seqlens;
  p.key_mult = a->key_mult;
  p.extra_kv = static_cast<const __nv_bfloat16*>(a->extra_kv);
  p.extra_mult = a->extra_mult;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.batch = a->batch;
  p.num_heads = a->num_heads;
  p.seq_len = a->seq_len;
  p.q_tiles = (max_len + 127) / 128;
  p.box_small = box_small;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.flag = device_flag_ptr();
  p.trace = tc_trace_buffer();
  p.route_rows = (a->impl == 0) ? a->route_rows : nullptr;
  p.route_min_rows = a->route_min_rows;
  p.min_keys = attention_split_active(a) ? kTcqNMax + 1 : 0;      // the quad-region launch takes the short samples
  p.max_keys = 1 << 30;
  const long long units = static_cast<long long>(a->batch) * a->num_heads * p.q_tiles;
  int grid = num_sms();
  if (units < grid) grid = static_cast<int>(units);
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tcr_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcrSmem<NMAX>::kBytes));
    attr_set = true;
  }
  attention_tcr_kernel<NMAX><<<grid, kTcrThreads, TcrSmem<NMAX>::kBytes, stream>>>(tq, tks, tkf, tout, p);
  return check_cuda(cudaGetLastError(), "attention_tcr_kernel launch");
}

// ---------------------------------------------------------------------------------------------------- quad-region kernel launch
// Eligible: what the ragged kernel takes; it computes the samples of at most 128 keys and skips the others (the caller launches
// the two-region / general kernel for those unless max_seq_len says there are none).  PK_ATT_TCQ=0 switches it off (A/B runs).
static int tcq_enabled() {                       // read per call (a test compares both dispatches in one process)
  const char* e = getenv("PK_ATT_TCQ");
  return (e && e[0] == '0') ? 0 : 1;
}
// Per-sample split of one ragged call (quad-region kernel for the samples of <= 128 keys, the two-region / general kernel for
// the longer ones): correct and tested, but OFF unless PK_ATT_SPLIT=1.  It pays when a batch mixes mid-length (40 - 125 rows)
// and long samples; on the batches the calibrated ResidualViT-S / A-ViT-S models realise (bimodal: most samples keep either
// ~3 rows or all 199) the short samples are launch-bound on either kernel and the second launch only adds its fill and drain
// (profiles/r02/run39_attn_model_lens.txt: 125 vs 104 us, 58 vs 46 us per layer).  Read per call, so a test can toggle it.
bool attention_split_active(const pk_attention_args* a) {
  if (a->impl != 0 || !attention_tcq_eligible(a)) return false;
  if (a->max_seq_len + (a->extra_kv ? 1 : 0) <= kTcqNMax) return false;      // every sample fits: the quad-region kernel alone
  const char* e = getenv("PK_ATT_SPLIT");
  return e && e[0] == '1';
}

bool attention_tcq_eligible(const pk_attention_args* a) {
  if (!tcq_enabled() || !(a->impl == 0 || a->impl == 4)) return false;
  if (a->head_dim != kTcDH || a->qkv_format != PK_OUT_BF16 || a->out_format != PK_OUT_BF16) return false;
  if (!a->cu_seqlens) return false;                                        // packed ragged rows only
  if ((reinterpret_cast<uintptr_t>(a->qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) return false;
  if (a->extra_kv && (reinterpret_cast<uintptr_t>(a->extra_kv) & 15) != 0) return false;
  return a->total_rows > 0;
}

int launch_attention_tcq(const pk_attention_args* a, cudaStream_t stream) {
  const int D = a->num_heads * kTcDH;
  const uint64_t rows = static_cast<uint64_t>(a->total_rows);
  const int box_small = 64;
  CUtensorMap tq, tks, tkf, tout;
  int rc = make_tmap_bf16_2d(&tq, a->qkv, rows, 3ull * D, 3ull * D, 128, 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tks, a->qkv, rows, 3ull * D, 3ull * D, box_small, 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tkf, a->qkv, rows, 3ull * D, 3ull * D, kTcqNMax, 64);
  if (rc != PK_OK) return rc;
  rc = make_tmap_bf16_2d(&tout, a->out, rows, static_cast<uint64_t>(D), static_cast<uint64_t>(D), 32, 64);
  if (rc != PK_OK) return rc;
  TcrParams p;
  p.cu_seqlens = a->cu_seqlens;
  p.key_mult = a->key_mult;
  p.extra_kv = static_cast<const __nv_bfloat16*>(a->extra_kv);
  p.extra_mult = a->extra_mult;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.batch = a->batch;
  p.num_heads = a->num_heads;
  p.seq_len = a->seq_len;
  p.q_tiles = 1;
  p.box_small = box_small;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.flag = device_flag_ptr();
  p.trace = tc_trace_buffer();
  p.route_rows = nullptr;
  p.route_min_rows = 0;
  p.min_keys = 0;
  p.max_keys = kTcqNMax;                              // longer samples are skipped: another launch of the same call takes them
  const long long units = static_cast<long long>(a->batch) * a->num_heads;
  int grid = num_sms();
  if (units < grid) grid = static_cast<int>(units);
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(attention_tcq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcqSmem::kBytes));
    attr_set = true;
  }
  attention_tcq_kernel<<<grid, kTcqThreads, TcqSmem::kBytes, stream>>>(tq, tks, tkf, tout, p);
  return check_cuda(cudaGetLastError(), "attention_tcq_kernel launch");
}

int launch_attention_tcr(const pk_attention_args* a, cudaStream_t stream) {
  const int max_len = a->cu_seqlens ? a->max_seq_len : a->seq_len;
  const int need = (max_len + (a->extra_kv ? 1 : 0) + 15) & ~15;
  if (need <= 32) return launch_tcr<32>(a, stream, max_len);
  if (need <= 64) return launch_tcr<64>(a, stream, max_len);
  if (need <= 128) return launch_tcr<128>(a, stream, max_len);
  if (need <= 208) return launch_tcr<208>(a, stream, max_len);
  return launch_tcr<256>(a, stream, max_len);
}

}  // namespace pk
