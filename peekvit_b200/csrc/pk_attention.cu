// peekvit_b200 — flash-style attention over packed (ragged) token rows (sm_100a).
//
// Replaces nn.MultiheadAttention's core (reference models/blocks.py:93-95; torch
// functional.py multi_head_attention_forward: q*1/sqrt(dh) -> bmm -> softmax -> bmm) without
// materialising the (B,H,N,N) weights the reference asks for (need_weights=True) and discards.
//
// This is the general path: any sequence length (online softmax over 64-key tiles), head_dim
// 32, 48 or 64, ragged batches through cu_seqlens, per-key multiplicities (+log mult on the logit)
// and one virtual "bias key" per head standing for the tokens a sparse model dropped
// (SURVEY.md Appendix A).  One CTA = 64 queries of one (sample, head); 4 warps x 16 query rows;
// K/V tiles double-buffered in shared memory with cp.async, XOR-swizzled for conflict-free
// ldmatrix; QK^T and PV on tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate); softmax
// statistics in fp32 registers (exp2 domain).
#include "pk_common.cuh"

#include <cstdlib>
#include "../../include/peekvit_b200.h"

namespace pk {

constexpr int kAttBQ = 64;
constexpr int kAttBKV = 64;
constexpr int kAttThreads = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float fmax3_att(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Byte offset of 16-byte chunk `c` of row `r` in a [rows][DH] bf16 tile, XOR-swizzled so that 8
// consecutive rows at one logical chunk hit 8 distinct 16-byte bank groups.
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
  constexpr int CPR = DH / 8;                       // chunks per row: 8 (dh 64) or 4 (dh 32)
  constexpr int RPS = 8 / CPR;                      // rows per 128-byte line: 1 or 2
  return static_cast<uint32_t>(r * (DH * 2) + ((c ^ ((r / RPS) % CPR)) * 16));
}

struct AttParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  int num_heads, seq_len;
  const int* cu_seqlens;
  float scale_log2;        // scale * log2(e)
  const float* key_mult;
  const __nv_bfloat16* extra_kv;
  const float* extra_mult;
  const int* route_rows;   // device-side routing: run only when *route_rows < route_min_rows (the tcgen05 kernel takes the rest)
  int route_min_rows;
  int skip_max_keys;       // samples of at most this many keys belong to the quad-region tcgen05 launch of the same call (0: none)
};

// OCC = resident CTAs per SM the register allocation is capped for.  3 for uniform long sequences (steady-state tiles, no
// spills); ragged batches of short samples are fill / drain bound (one or two key tiles per CTA), where more CTAs in flight
// hide the load latency better than a spill-free inner loop (PK_ATT_GENERAL_OCC, profiles/r02).
template <int DH, int OCC = 3>
__global__ void __launch_bounds__(kAttThreads, OCC)
attention_fwd_kernel(const AttParams p) {
  constexpr int CPR = DH / 8;
  constexpr int KS = DH / 16;                       // k-steps over head_dim for QK^T
  // shared-memory row pitch: the XOR swizzle needs a power-of-two chunk count, so head_dim 48 (ViT-S with 8 heads,
  // configs/model/vit_small.yaml) is stored in 64-element rows with the last two chunks unused
  constexpr int SDH = DH <= 32 ? 32 : 64;
  constexpr int TILE_BYTES = kAttBKV * SDH * 2;
  __shared__ __align__(128) uint8_t s_q[kAttBQ * SDH * 2];
  __shared__ __align__(128) uint8_t s_k[2][TILE_BYTES];
  __shared__ __align__(128) uint8_t s_v[2][TILE_BYTES];
  __shared__ float s_bias[2][kAttBKV];

  if (p.route_rows && *p.route_rows >= p.route_min_rows) return;     // the ragged tcgen05 kernel's launch takes this batch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int D = p.num_heads * DH;
  const long long ld = 3LL * D;
  const long long start = p.cu_seqlens ? p.cu_seqlens[b] : (long long)b * p.seq_len;
  const int len = p.cu_seqlens ? (p.cu_seqlens[b + 1] - p.cu_seqlens[b]) : p.seq_len;
  const int q0 = blockIdx.x * kAttBQ;
  if (q0 >= len) return;
  float extra_m = 0.f;
  if (p.extra_kv && p.extra_mult) extra_m = p.extra_mult[b];
  const bool has_extra = extra_m > 0.f;
  const int n_keys = len + (has_extra ? 1 : 0);
  if (n_keys <= p.skip_max_keys) return;             // a short sample: the quad-region tcgen05 launch of this call takes it
  const int n_tiles = (n_keys + kAttBKV - 1) / kAttBKV;
  constexpr float kLog2e = 1.4426950408889634f;

  const __nv_bfloat16* q_base = p.qkv + start * ld + h * DH;
  const __nv_bfloat16* k_base = q_base + D;
  const __nv_bfloat16* v_base = q_base + 2 * D;

  // Per-thread constants of the K/V tile loader: chunk i = tid + k * 128 of a tile always maps to the same (row, 16-byte
  // chunk), so its shared-memory offset and its source offset inside the tile are computed once (the loader's address
  // arithmetic was a fifth of the kernel's issue slots: ncu showed 62 - 70 % issue-active against 33 - 40 % tensor-active).
  constexpr int LD_ITERS = kAttBKV * CPR / kAttThreads;
  static_assert(kAttBKV * CPR % kAttThreads == 0, "loader assumes a whole number of chunks per thread");
  int ld_row[LD_ITERS];
  uint32_t ld_soff[LD_ITERS];
  int ld_goff[LD_ITERS];                            // element offset inside a tile: < 64 * 3D
#pragma unroll
  for (int k = 0; k < LD_ITERS; ++k) {
    const int i = tid + k * kAttThreads;
    const int r = i / CPR, c = i % CPR;
    ld_row[k] = r;
    ld_soff[k] = tile_off<SDH>(r, c);
    ld_goff[k] = r * static_cast<int>(ld) + c * 8;
  }
  auto load_kv_tile = [&](int t, int buf) {
    const int kv0 = t * kAttBKV;
    const uint32_t sk = smem_u32(s_k[buf]), sv = smem_u32(s_v[buf]);
    const __nv_bfloat16* kt = k_base + (long long)kv0 * ld;
    const __nv_bfloat16* vt = v_base + (long long)kv0 * ld;
#pragma unroll
    for (int k = 0; k < LD_ITERS; ++k) {
      const int key = kv0 + ld_row[k];
      const __nv_bfloat16* ksrc = k_base;
      const __nv_bfloat16* vsrc = v_base;
      int bytes = 0;
      if (key < len) {
        ksrc = kt + ld_goff[k];
        vsrc = vt + ld_goff[k];
        bytes = 16;
      } else if (has_extra && key == len) {
        const int eoff = h * DH + (ld_goff[k] - ld_row[k] * static_cast<int>(ld));      // h * DH + c * 8
        ksrc = p.extra_kv + eoff;
        vsrc = p.extra_kv + D + eoff;
        bytes = 16;
      }
      cp_async16(sk + ld_soff[k], ksrc, bytes);
      cp_async16(sv + ld_soff[k], vsrc, bytes);
    }
    if (tid < kAttBKV) {
      const int key = kv0 + tid;
      float bias = -INFINITY;
      if (key < len) bias = p.key_mult ? __log2f(p.key_mult[start + key]) : 0.f;
      else if (has_extra && key == len) bias = __log2f(extra_m);
      s_bias[buf][tid] = bias;
    }
  };

  // Q tile + first K/V tile
  {
    const uint32_t sq = smem_u32(s_q);
    for (int i = tid; i < kAttBQ * CPR; i += kAttThreads) {
      const int r = i / CPR, c = i % CPR;
      const int q = q0 + r;
      const bool ok = q < len;
      cp_async16(sq + tile_off<SDH>(r, c), ok ? q_base + (long long)q * ld + c * 8 : q_base, ok ? 16 : 0);
    }
    load_kv_tile(0, 0);
    cp_async_commit();
  }

  // ldmatrix offsets of this lane inside a 16-key group: the swizzle term depends on the row only modulo 8, so stepping to
  // the next 16 keys is a constant add
  uint32_t koff[KS], voff[DH / 16];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) koff[ks] = tile_off<SDH>((lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1));
#pragma unroll
  for (int nd = 0; nd < DH / 16; ++nd) voff[nd] = tile_off<SDH>((lane & 7) + (((lane >> 3) & 1) << 3), nd * 2 + (lane >> 4));
  const bool warp_live = q0 + warp * 16 < len;       // warp-uniform
  uint32_t qf[KS][4];
  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_kv_tile(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      const uint32_t sq = smem_u32(s_q);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(sq + tile_off<SDH>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    // a warp whose 16 query rows all lie past the end of the sample (the last query tile of n = 197 has 5 valid rows) only
    // helps with the loads: its MMAs and softmax would be issue slots taken from the CTAs that share the SM
    if (warp_live) {
    const uint32_t sk = smem_u32(s_k[buf]), sv = smem_u32(s_v[buf]);

    // ---- S = Q K^T for 64 keys: 8 n-tiles of 8 keys
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {               // 16 keys per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk + koff[ks] + np * (16 * SDH * 2), b0, b1, b2, b3);
        mma_bf16_16816(s[np * 2], qf[ks], b0, b1);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // ---- scale, key bias / mask, online softmax (rows g = lane/4 and g+8) on packed f32x2 arithmetic: every accumulator
    // pair (two adjacent keys of one row) is one FFMA2 / FADD2, the row maxima use 3-input max
    const uint64_t sc2 = f2_pack(p.scale_log2, p.scale_log2);
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int kc = n * 8 + (lane & 3) * 2;
      const float2 bb = *reinterpret_cast<const float2*>(&s_bias[buf][kc]);
      const uint64_t b2 = f2_pack(bb.x, bb.y);
      f2_unpack(f2_fma(f2_pack(s[n][0], s[n][1]), sc2, b2), s[n][0], s[n][1]);      // scaled + biased logits, in place
      f2_unpack(f2_fma(f2_pack(s[n][2], s[n][3]), sc2, b2), s[n][2], s[n][3]);
      tmax[0] = fmax3_att(tmax[0], s[n][0], s[n][1]);
      tmax[1] = fmax3_att(tmax[1], s[n][2], s[n][3]);
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float m_new = fmaxf(m_run[r], tmax[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f(m_run[r] - m_use[r]);          // m_run = -inf -> 0
      m_run[r] = m_new;
    }
    uint32_t pf[4][4];                                // P as bf16 A-fragments, 4 k-steps of 16 keys
    const uint64_t nm0 = f2_pack(-m_use[0], -m_use[0]), nm1 = f2_pack(-m_use[1], -m_use[1]);
    uint64_t rs0 = f2_pack(0.f, 0.f), rs1 = rs0;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float x0, x1, x2, x3;
      f2_unpack(f2_add(f2_pack(s[n][0], s[n][1]), nm0), x0, x1);
      f2_unpack(f2_add(f2_pack(s[n][2], s[n][3]), nm1), x2, x3);
      const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
      rs0 = f2_add(rs0, f2_pack(p0, p1));
      rs1 = f2_add(rs1, f2_pack(p2, p3));
      pf[n >> 1][(n & 1) * 2] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    {
      float a, b;
      f2_unpack(rs0, a, b);
      l_run[0] = l_run[0] * corr[0] + (a + b);       // per-thread partial; reduced at the end
      f2_unpack(rs1, a, b);
      l_run[1] = l_run[1] * corr[1] + (a + b);
    }
    {
      const uint64_t c0 = f2_pack(corr[0], corr[0]), c1 = f2_pack(corr[1], corr[1]);
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        f2_unpack(f2_mul(f2_pack(o[n][0], o[n][1]), c0), o[n][0], o[n][1]);
        f2_unpack(f2_mul(f2_pack(o[n][2], o[n][3]), c1), o[n][2], o[n][3]);
      }
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                 // 16 keys per step
#pragma unroll
      for (int nd = 0; nd < DH / 16; ++nd) {         // 16 head-dim columns per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(sv + voff[nd] + kk * (16 * SDH * 2), b0, b1, b2, b3);
        mma_bf16_16816(o[nd * 2], pf[kk], b0, b1);
        mma_bf16_16816(o[nd * 2 + 1], pf[kk], b2, b3);
      }
    }
    }
    __syncthreads();   // everyone done with buf before it is refilled two iterations later
  }

  // ---- normalise and store: stage the 64 x DH bf16 tile in s_q, then 16-byte coalesced stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.0f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.0f / l_run[1] : 0.f;
  {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const int r0 = warp * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(s_q + tile_off<SDH>(r0, n) + tq * 4) = pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
      *reinterpret_cast<uint32_t*>(s_q + tile_off<SDH>(r1, n) + tq * 4) = pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
  __syncthreads();
  for (int i = tid; i < kAttBQ * CPR; i += kAttThreads) {
    const int r = i / CPR, c = i % CPR;
    const int q = q0 + r;
    if (q < len) {
      const uint4 v = *reinterpret_cast<const uint4*>(s_q + tile_off<SDH>(r, c));
      *reinterpret_cast<uint4*>(p.out + (start + q) * (long long)D + h * DH + c * 8) = v;
    }
  }
}

// pk_attention_tc.cu: tcgen05/TMEM kernel for uniform 128 < n <= 256, head_dim 64
bool attention_tc_eligible(const pk_attention_args* a);
int launch_attention_tc(const pk_attention_args* a, cudaStream_t stream);
bool attention_tcr_eligible(const pk_attention_args* a);
int launch_attention_tcr(const pk_attention_args* a, cudaStream_t stream);
bool attention_tcq_eligible(const pk_attention_args* a);
bool attention_split_active(const pk_attention_args* a);
int launch_attention_tcq(const pk_attention_args* a, cudaStream_t stream);
int attention_trace_copy(unsigned long long* host_dst);

}  // namespace pk

extern "C" int pk_attention_fwd(const pk_attention_args* a, void* stream) {
  using namespace pk;
  PK_REQUIRE(a && a->qkv && a->out, "pk_attention_fwd: null pointer");
  PK_REQUIRE(a->head_dim == 32 || a->head_dim == 48 || a->head_dim == 64, "pk_attention_fwd: head_dim %d not in {32, 48, 64}", a->head_dim);
  PK_REQUIRE(a->batch >= 0 && a->num_heads > 0 && a->max_seq_len >= 0, "pk_attention_fwd: bad shape");
  PK_REQUIRE(a->cu_seqlens || a->seq_len > 0, "pk_attention_fwd: need cu_seqlens or seq_len");
  PK_REQUIRE((a->extra_kv == nullptr) == (a->extra_mult == nullptr), "pk_attention_fwd: extra_kv and extra_mult go together");
  if (a->batch == 0 || a->max_seq_len == 0) return PK_OK;
  if (attention_tc_eligible(a)) return launch_attention_tc(a, static_cast<cudaStream_t>(stream));
  // Quad-region tcgen05 kernel: alone when the static bound says every sample has at most 128 keys (or with impl 4); with the
  // per-sample split it takes the short samples of a mixed batch and the kernels below, which skip those, the longer ones.
  const bool quad_only = attention_tcq_eligible(a) && a->max_seq_len + (a->extra_kv ? 1 : 0) <= 128;
  PK_REQUIRE(a->impl != 4 || quad_only, "pk_attention_fwd: the quad-region tcgen05 kernel needs cu_seqlens, head_dim 64 and at most 128 keys per sample");
  const bool quad = quad_only || attention_split_active(a);      // the split itself is opt-in (PK_ATT_SPLIT=1, see pk_attention_tc.cu)
  if (quad) {
    const int rc = launch_attention_tcq(a, static_cast<cudaStream_t>(stream));
    if (rc != PK_OK || quad_only) return rc;
  }
  const bool routed = a->route_rows != nullptr && a->impl == 0 && attention_tcr_eligible(a);
  if (routed) {
    const int rc = launch_attention_tcr(a, static_cast<cudaStream_t>(stream));
    if (rc != PK_OK) return rc;
  } else if (attention_tcr_eligible(a)) {
    return launch_attention_tcr(a, static_cast<cudaStream_t>(stream));
  }
  PK_REQUIRE(a->impl != 3, "pk_attention_fwd: the ragged tcgen05 kernel needs head_dim 64, <= 256 keys per sample, 16-byte aligned buffers and total_rows");
  PK_REQUIRE(a->qkv_format == PK_OUT_BF16 && a->out_format == PK_OUT_BF16,
             "pk_attention_fwd: half operands / split output exist on the tcgen05 kernel only (uniform 16 < seq_len <= 224, head_dim 64)");
  PK_REQUIRE(a->impl != 2, "pk_attention_fwd: the tcgen05 kernel needs uniform 128 < seq_len <= 256, head_dim 64, no key multiplicities");
  AttParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.num_heads = a->num_heads;
  p.seq_len = a->seq_len;
  p.cu_seqlens = a->cu_seqlens;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_mult = a->key_mult;
  p.extra_kv = static_cast<const __nv_bfloat16*>(a->extra_kv);
  p.extra_mult = a->extra_mult;
  p.route_rows = routed ? a->route_rows : nullptr;
  p.route_min_rows = a->route_min_rows;
  p.skip_max_keys = quad ? 128 : 0;
  const int max_len = a->cu_seqlens ? a->max_seq_len : a->seq_len;
  dim3 grid((max_len + kAttBQ - 1) / kAttBQ, a->num_heads, a->batch);
  PK_REQUIRE(a->batch <= 65535 && a->num_heads <= 65535, "pk_attention_fwd: batch/heads exceed grid limits; split the batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static int occ_ragged = -1;
  if (occ_ragged < 0) { const char* e = getenv("PK_ATT_GENERAL_OCC"); occ_ragged = e ? atoi(e) : 4; }     // measured 66.2 / 59.9 / 58.2 us at 3 / 4 / 5 (mean 79 rows), 160.0 / 151.5 / 159.9 (mean 165)
  if (a->head_dim == 64 && a->cu_seqlens && occ_ragged == 4) attention_fwd_kernel<64, 4><<<grid, kAttThreads, 0, s>>>(p);
  else if (a->head_dim == 64 && a->cu_seqlens && occ_ragged == 5) attention_fwd_kernel<64, 5><<<grid, kAttThreads, 0, s>>>(p);
  else if (a->head_dim == 64) attention_fwd_kernel<64><<<grid, kAttThreads, 0, s>>>(p);
  else if (a->head_dim == 48) attention_fwd_kernel<48><<<grid, kAttThreads, 0, s>>>(p);
  else {
    const char* e32 = getenv("PK_ATT_GENERAL_OCC32");
    const int occ32 = e32 ? atoi(e32) : 4;      // config A (64 x 8 heads x 785 tokens): 221.9 / 202.7 / 205.1 / 208.6 us at 3 / 4 / 5 / 6 (profiles/r02/run47)
    if (occ32 == 4) attention_fwd_kernel<32, 4><<<grid, kAttThreads, 0, s>>>(p);
    else if (occ32 == 5) attention_fwd_kernel<32, 5><<<grid, kAttThreads, 0, s>>>(p);
    else if (occ32 == 6) attention_fwd_kernel<32, 6><<<grid, kAttThreads, 0, s>>>(p);
    else attention_fwd_kernel<32><<<grid, kAttThreads, 0, s>>>(p);
  }
  return check_cuda(cudaGetLastError(), "attention_fwd_kernel");
}

// Debug aid (not part of the product path): copies the PK_ATT_TRACE=1 event table (16 items x 16 warps x 8 events of
// clock64 stamps from CTA 0 of the last tcgen05 attention launch) to host memory.
extern "C" int pk_attention_trace(unsigned long long* host_dst) {
  PK_REQUIRE(host_dst != nullptr, "pk_attention_trace: null destination");
  PK_CHECK_CUDA(cudaDeviceSynchronize());
  return pk::attention_trace_copy(host_dst);
}
