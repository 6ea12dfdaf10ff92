// peekvit_b200 — backward kernels of the fine-tuning path.
//
// The reference fine-tunes with the backbone frozen (train/train.py:97-127: `train_only_these_params(model, ['gate', 'class',
// 'head', 'threshold', 'budget'])`, models/topology.py:128-158), i.e. the only gradients it needs are those of a few small
// parameters -- but the class tokens sit at the INPUT of the encoder, so the loss gradient has to travel back through every
// block as an activation gradient (dX), never as a weight gradient.  That is what lives here, for the dense ViT block
// (models/vit.py:45-55):
//     x1 = x + Wo . attention(Wqkv . LN1(x))          x2 = x1 + W2 . gelu(W1 . LN2(x1))
// The four dX GEMMs per block run on the same tcgen05 kernels as the forward (pk_gemm_bf16 with the transposed weight);
// this file holds what is left: LayerNorm backward, exact-erf GELU forward / backward on the stored pre-activation, the
// attention core backward (softmax recomputed per (sample, head)), cross-entropy + head backward, and the reduction of the
// class-row gradients.  Activation gradients travel as bf16 between the GEMMs and as fp32 on the residual path.
#include "peekvit_b200.h"
#include "pk_common.cuh"

#include <cstdlib>

namespace pk {

static int train_grid(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------ casts / GELU
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// hid = gelu(h_pre) (models/blocks.py:82, exact erf form) on the stored bf16 pre-activation
__global__ void gelu_bf16_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ y, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(h)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      o[k] = pack_bf16(gelu_exact(__low2float(b)), gelu_exact(__high2float(b)));
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// dh_pre = dhid * gelu'(h_pre)
__global__ void gelu_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dy,
                                     __nv_bfloat16* __restrict__ dx, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 hv = reinterpret_cast<const uint4*>(h)[i];
    const uint4 gv = reinterpret_cast<const uint4*>(dy)[i];
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hw[k]);
      const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&gw[k]);
      o[k] = pack_bf16(__low2float(gb) * gelu_exact_grad(__low2float(hb)), __high2float(gb) * gelu_exact_grad(__high2float(hb)));
    }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ LayerNorm backward (input gradient only: gamma / beta are frozen)
//   xhat = (x - mean) * rstd,  g = dy * gamma,  dx (+)= rstd * (g - mean(g) - xhat * mean(g * xhat))
// One warp per row, the row held in registers (dim <= 1024).  row_index (optional) maps the launch's row r to row
// row_index[r] of x / dx (the class rows of the final LayerNorm); dy row = r / dy_div (several class tokens share one
// feature gradient: the head reads their sum, vit.py:242-243).
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma, float eps,
                     float* __restrict__ dx, int rows, int dim, const int* __restrict__ row_index, int dy_div, int accumulate,
                     const float* __restrict__ beta, const float* __restrict__ rowscale, float* __restrict__ dot_out) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const long long xr = row_index ? row_index[r] : r;
    const float4* x4 = reinterpret_cast<const float4*>(x + xr * dim);
    const float4* g4 = reinterpret_cast<const float4*>(dy + static_cast<long long>(r / dy_div) * dim);
    float4 xv[MAXV], gv[MAXV];
    float s1 = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int c = lane + 32 * j;
      xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      gv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d4) {
        xv[j] = x4[c];
        const float4 d = g4[c], w = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        gv[j] = make_float4(d.x * w.x, d.y * w.y, d.z * w.z, d.w * w.w);
        s1 += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
        if (dot_out && beta) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
          sb += (d.x * b.x + d.y * b.y) + (d.z * b.z + d.w * b.w);
        }
      }
    }
    const float mean = warp_sum(s1) / static_cast<float>(dim);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (lane + 32 * j < d4) {
        xv[j].x -= mean; xv[j].y -= mean; xv[j].z -= mean; xv[j].w -= mean;
        s2 += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(s2) / static_cast<float>(dim) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (lane + 32 * j < d4) {
        xv[j].x *= rstd; xv[j].y *= rstd; xv[j].z *= rstd; xv[j].w *= rstd;      // xhat
        sg += (gv[j].x + gv[j].y) + (gv[j].z + gv[j].w);
        sgx += (gv[j].x * xv[j].x + gv[j].y * xv[j].y) + (gv[j].z * xv[j].z + gv[j].w * xv[j].w);
      }
    }
    const float tgx = warp_sum(sgx);
    const float mg = warp_sum(sg) / static_cast<float>(dim), mgx = tgx / static_cast<float>(dim);
    // gated blocks (residualvit.py:249-260): the LayerNorm output is multiplied by the row's mask m before it is used, so
    // d mask = dy . LN(x) = sum(dy * gamma * xhat) + sum(dy * beta) and the gradient that enters the LayerNorm is m * dy
    if (dot_out) {
      const float tb = warp_sum(sb);
      if (lane == 0) dot_out[xr] += tgx + tb;
    }
    const float rs = rowscale ? rowscale[xr] : 1.f;
    if (rowscale && rs == 0.f && accumulate) continue;
    const float rstd_s = rstd * rs;
    float4* o4 = reinterpret_cast<float4*>(dx + xr * dim);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int c = lane + 32 * j;
      if (c < d4) {
        float4 o = make_float4(rstd_s * (gv[j].x - mg - xv[j].x * mgx), rstd_s * (gv[j].y - mg - xv[j].y * mgx),
                               rstd_s * (gv[j].z - mg - xv[j].z * mgx), rstd_s * (gv[j].w - mg - xv[j].w * mgx));
        if (accumulate) {
          const float4 p = o4[c];
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        o4[c] = o;
      }
    }
  }
}

// ------------------------------------------------------------------ attention core backward
// One CTA per (sample, head), uniform sequences of n <= 256 tokens, head_dim 64 or 32.  K and V of the head sit in shared
// memory (bf16, rows padded by one word: conflict-free column walks).  Query rows are taken sixteen at a time, one per warp:
//   phase A (warp = query row i): recompute the softmax row p_i (lane = key), dP_i = dO_i V^T, D_i = dO_i . O_i,
//     dS_i = p_i * (dP_i - D_i) * scale, leave p_i / dS_i / q_i / dO_i in shared memory; lane = head dimension: dQ_i = dS_i K;
//   phase B (thread = (block of DH keys, head dimension d)): dK[j][d] += dS_i[j] q_i[d], dV[j][d] += p_i[j] dO_i[d] for the
//     tile's sixteen rows, accumulated in REGISTERS for the whole CTA (n * DH accumulators of each over 512 threads).
// The first version accumulated dK / dV with shared-memory atomics from phase A: 15 ms per ViT-B layer and 128 images, 97 %
// of a fine-tuning step; the register accumulators need no atomics.  All arithmetic fp32; inputs / outputs bf16.
constexpr int kBwdThreads = 512, kBwdWarps = kBwdThreads / 32;
template <int DH>
__global__ void __launch_bounds__(kBwdThreads)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                     __nv_bfloat16* __restrict__ dqkv, int num_heads, int n, float scale) {
  constexpr int EPL = DH / 32;                  // head dimensions per lane (phase A)
  constexpr int PITCH = DH + 2;                 // bf16 elements per shared K / V row
  constexpr int KPT = 256 / (kBwdThreads / DH); // keys per thread (phase B): kBwdThreads / DH key blocks cover 256 keys
  constexpr int ROWF = 2 * DH + 512;            // floats per warp row record: q[DH] | dO[DH] | p[256] | ds[256]
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + n * PITCH;
  float* rec = reinterpret_cast<float*>(smem + ((2 * n * PITCH * 2 + 15) & ~15));
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int D = num_heads * DH;
  const long long row0 = static_cast<long long>(b) * n;
  for (int i = tid; i < n * (DH / 2); i += blockDim.x) {
    const int j = i / (DH / 2), c = i % (DH / 2);
    const __nv_bfloat16* src = qkv + (row0 + j) * 3ll * D + h * DH + 2 * c;
    *reinterpret_cast<uint32_t*>(Ks + j * PITCH + 2 * c) = *reinterpret_cast<const uint32_t*>(src + D);
    *reinterpret_cast<uint32_t*>(Vs + j * PITCH + 2 * c) = *reinterpret_cast<const uint32_t*>(src + 2 * D);
  }
  float* qrow = rec + warp * ROWF;
  float* dorow = qrow + DH;
  float* prow = dorow + DH;
  float* dsrow = prow + 256;
  const float scale_log2 = scale * 1.4426950408889634f;
  const int bd = tid % DH, bk0 = (tid / DH) * KPT;       // phase B: this thread's head dimension and first key
  float dk[KPT], dv[KPT];
#pragma unroll
  for (int e = 0; e < KPT; ++e) dk[e] = dv[e] = 0.f;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += kBwdWarps) {
    const int i = i0 + warp;
    // ---------------- phase A
    if (i < n) {
      float dsum = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int d = lane * EPL + e;
        const float qv = __bfloat162float(qkv[(row0 + i) * 3ll * D + h * DH + d]);
        const float gv = __bfloat162float(dout[(row0 + i) * static_cast<long long>(D) + h * DH + d]);
        dsum += gv * __bfloat162float(o[(row0 + i) * static_cast<long long>(D) + h * DH + d]);
        qrow[d] = qv;
        dorow[d] = gv;
      }
      const float Di = warp_sum(dsum);
      __syncwarp();
      float s[8], dp[8];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = lane + 32 * jj;
        s[jj] = -INFINITY;
        dp[jj] = 0.f;
        if (j < n) {
          float acc0 = 0.f, acc1 = 0.f, accv0 = 0.f, accv1 = 0.f;      // two chains per dot product
#pragma unroll 8
          for (int d = 0; d < DH; d += 4) {
            const __nv_bfloat162 k2 = *reinterpret_cast<const __nv_bfloat162*>(Ks + j * PITCH + d);
            const __nv_bfloat162 k3 = *reinterpret_cast<const __nv_bfloat162*>(Ks + j * PITCH + d + 2);
            const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(Vs + j * PITCH + d);
            const __nv_bfloat162 v3 = *reinterpret_cast<const __nv_bfloat162*>(Vs + j * PITCH + d + 2);
            const float4 q4 = *reinterpret_cast<const float4*>(qrow + d), g4 = *reinterpret_cast<const float4*>(dorow + d);
            acc0 = fmaf(q4.x, __low2float(k2), fmaf(q4.y, __high2float(k2), acc0));
            acc1 = fmaf(q4.z, __low2float(k3), fmaf(q4.w, __high2float(k3), acc1));
            accv0 = fmaf(g4.x, __low2float(v2), fmaf(g4.y, __high2float(v2), accv0));
            accv1 = fmaf(g4.z, __low2float(v3), fmaf(g4.w, __high2float(v3), accv1));
          }
          s[jj] = (acc0 + acc1) * scale_log2;
          dp[jj] = accv0 + accv1;
          mx = fmaxf(mx, s[jj]);
        }
      }
      mx = warp_max(mx);
      float l = 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        s[jj] = (lane + 32 * jj < n) ? exp2f(s[jj] - mx) : 0.f;
        l += s[jj];
      }
      const float inv_l = 1.0f / warp_sum(l);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = lane + 32 * jj;
        const float p = s[jj] * inv_l;                   // 0 for j >= n
        prow[j] = p;
        dsrow[j] = p * (dp[jj] - Di) * scale;            // dS scaled: dQ = dS K, dK = dS^T Q
      }
      __syncwarp();
      float dq[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) dq[e] = 0.f;
      for (int j = 0; j < n; ++j) {
        const float ds = dsrow[j];
#pragma unroll
        for (int e = 0; e < EPL; ++e) dq[e] += ds * __bfloat162float(Ks[j * PITCH + lane * EPL + e]);
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        dqkv[(row0 + i) * 3ll * D + h * DH + lane * EPL + e] = __float2bfloat16_rn(dq[e]);
    } else {
      // rows past the end of the sample contribute nothing to phase B
      for (int j = lane; j < 256; j += 32) { prow[j] = 0.f; dsrow[j] = 0.f; }
      for (int d = lane; d < DH; d += 32) { qrow[d] = 0.f; dorow[d] = 0.f; }
    }
    __syncthreads();
    // ---------------- phase B
#pragma unroll 1
    for (int r = 0; r < kBwdWarps; ++r) {
      const float* rr = rec + r * ROWF;
      const float qd = rr[bd], gd = rr[DH + bd];
      const float4* p4 = reinterpret_cast<const float4*>(rr + 2 * DH + bk0);
      const float4* s4 = reinterpret_cast<const float4*>(rr + 2 * DH + 256 + bk0);
#pragma unroll
      for (int e = 0; e < KPT / 4; ++e) {
        const float4 pv = p4[e], sv = s4[e];
        dk[4 * e] += sv.x * qd; dk[4 * e + 1] += sv.y * qd; dk[4 * e + 2] += sv.z * qd; dk[4 * e + 3] += sv.w * qd;
        dv[4 * e] += pv.x * gd; dv[4 * e + 1] += pv.y * gd; dv[4 * e + 2] += pv.z * gd; dv[4 * e + 3] += pv.w * gd;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < KPT; ++e) {
    const int j = bk0 + e;
    if (j < n) {
      dqkv[(row0 + j) * 3ll * D + D + h * DH + bd] = __float2bfloat16_rn(dk[e]);
      dqkv[(row0 + j) * 3ll * D + 2 * D + h * DH + bd] = __float2bfloat16_rn(dv[e]);
    }
  }
}

// ------------------------------------------------------------------ attention core backward on the tensor cores (head_dim 64)
// Same contract as attention_bwd_kernel, five matrix products per (sample, head) -- S = Q K^T, dP = dO V^T, dQ = dS K,
// dK = dS^T Q, dV = P^T dO -- as mma.sync m16n8k16 bf16 tiles with fp32 accumulation (the backward is a training-only path
// next to tcgen05 forward kernels; its products are 197 x 197 x 64, far below one tcgen05 tile pipeline's start-up cost).
// Q, K, V, dO of the head sit in shared memory (bf16, 144-byte rows: conflict-free ldmatrix).  Phase 1, warp = block of 16
// queries: sweep the keys once for the row statistics (log2-sum-exp), once more for P, dP, dS and dQ (dS leaves the
// accumulator registers as the A operand of dS K).  Phase 2, warp = block of 16 keys: sweep the query blocks with the
// TRANSPOSED tiles S^T = K Q^T, dP^T = V dO^T, so P^T and dS^T come out of the accumulators in A-operand layout for
// dV += P^T dO and dK += dS^T Q.  S and dP are computed twice; nothing is reduced across warps, nothing is atomic.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kMmaPitch = 72;                   // bf16 elements per shared row (64 + 8)
// Sixteen warps per CTA: the four tiles of a 197-token head take 120 KB of shared memory, so one CTA per SM is all there is,
// and with eight warps the 13 row blocks of a phase ran as two rounds (8 + 5) of latency-bound warps.  At the 128 registers per
// thread that 512 threads leave, ptxas spills 56 bytes (phase 2 holds two A operands and two 16 x 64 accumulators).
constexpr int kBwdMmaThreads = 512, kBwdMmaWarps = kBwdMmaThreads / 32;

// C tiles (16 x 16 as two n8 tiles) of  X_blk[16 rows] . Y_blk[16 rows]^T  over head_dim 64: A fragments of X preloaded,
// B fragments of Y (rows = the product's columns) by non-transposed ldmatrix.
__device__ __forceinline__ void tile_xyT(float (&c)[2][4], const uint32_t (&ax)[4][4], uint32_t y_rows_addr, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  const uint32_t lane_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 4) << 3)) * kMmaPitch + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t b[4];
    ldsm_x4(y_rows_addr + lane_off + ks * 32, b);
    mma_bf16(c[0], ax[ks], b[0], b[1]);
    mma_bf16(c[1], ax[ks], b[2], b[3]);
  }
}
// A fragments (4 k-steps over head_dim 64) of a block of 16 rows
__device__ __forceinline__ void load_a64(uint32_t (&a)[4][4], uint32_t rows_addr, int lane) {
  const uint32_t lane_off = static_cast<uint32_t>(((lane & 15) * kMmaPitch + (lane >> 4) * 8) * 2);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(rows_addr + lane_off + ks * 32, a[ks]);
}
// acc[16 x 64] += A[16 x 16] . Z_blk[16 rows x 64]   (Z rows = the contraction index: transposed ldmatrix)
__device__ __forceinline__ void acc_a_z(float (&acc)[8][4], const uint32_t (&a)[4], uint32_t z_rows_addr, int lane) {
  const uint32_t lane_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 3) & 1) * 8) * kMmaPitch + (lane >> 4) * 8) * 2);
#pragma unroll
  for (int dt2 = 0; dt2 < 4; ++dt2) {
    uint32_t b[4];
    ldsm_x4_t(z_rows_addr + lane_off + dt2 * 32, b);
    mma_bf16(acc[2 * dt2], a, b[0], b[1]);
    mma_bf16(acc[2 * dt2 + 1], a, b[2], b[3]);
  }
}

__global__ void __launch_bounds__(kBwdMmaThreads)
attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                         __nv_bfloat16* __restrict__ dqkv, int num_heads, int n, float scale) {
  constexpr int DH = 64;
  extern __shared__ __align__(16) unsigned char smem[];
  const int NP = (n + 15) & ~15;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Ks = Qs + NP * kMmaPitch;
  __nv_bfloat16* Vs = Ks + NP * kMmaPitch;
  __nv_bfloat16* Gs = Vs + NP * kMmaPitch;
  float* lse = reinterpret_cast<float*>(Gs + NP * kMmaPitch);
  float* Dv = lse + NP;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int D = num_heads * DH;
  const long long row0 = static_cast<long long>(b) * n;
  // ---- phase 0: the head's Q, K, V, dO tiles (rows >= n zero)
  for (int i = tid; i < NP * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    uint4 q4 = make_uint4(0, 0, 0, 0), k4 = q4, v4 = q4, g4 = q4;
    if (r < n) {
      const __nv_bfloat16* src = qkv + (row0 + r) * 3ll * D + h * DH + c * 8;
      q4 = *reinterpret_cast<const uint4*>(src);
      k4 = *reinterpret_cast<const uint4*>(src + D);
      v4 = *reinterpret_cast<const uint4*>(src + 2 * D);
      g4 = *reinterpret_cast<const uint4*>(dout + (row0 + r) * static_cast<long long>(D) + h * DH + c * 8);
    }
    *reinterpret_cast<uint4*>(Qs + r * kMmaPitch + c * 8) = q4;
    *reinterpret_cast<uint4*>(Ks + r * kMmaPitch + c * 8) = k4;
    *reinterpret_cast<uint4*>(Vs + r * kMmaPitch + c * 8) = v4;
    *reinterpret_cast<uint4*>(Gs + r * kMmaPitch + c * 8) = g4;
  }
  __syncthreads();
  for (int r = tid; r < NP; r += blockDim.x) {       // D_i = dO_i . O_i
    float acc = 0.f;
    if (r < n) {
      const __nv_bfloat16* orow = o + (row0 + r) * static_cast<long long>(D) + h * DH;
      for (int c = 0; c < DH; c += 2) {
        const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(Gs + r * kMmaPitch + c);
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(orow + c);
        acc += __low2float(a2) * __low2float(b2) + __high2float(a2) * __high2float(b2);
      }
    }
    Dv[r] = acc;
  }
  __syncthreads();
  const uint32_t qs = smem_u32(Qs), ks_ = smem_u32(Ks), vs = smem_u32(Vs), gs = smem_u32(Gs);
  const float scale2 = scale * 1.4426950408889634f;
  const int nblk = NP >> 4;
  const uint32_t blk_bytes = 16 * kMmaPitch * 2;
  // ---- phase 1: warp = block of 16 queries
  for (int qb = warp; qb < nblk; qb += kBwdMmaWarps) {
    uint32_t aq[4][4];
    load_a64(aq, qs + qb * blk_bytes, lane);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;           // rows g and g + 8
    for (int kb = 0; kb < nblk; ++kb) {
      float c[2][4];
      tile_xyT(c, aq, ks_ + kb * blk_bytes, lane);
      float v0[4], v1[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = kb * 16 + nt * 8 + 2 * t;
        v0[2 * nt] = col < n ? c[nt][0] * scale2 : -INFINITY;
        v0[2 * nt + 1] = col + 1 < n ? c[nt][1] * scale2 : -INFINITY;
        v1[2 * nt] = col < n ? c[nt][2] * scale2 : -INFINITY;
        v1[2 * nt + 1] = col + 1 < n ? c[nt][3] * scale2 : -INFINITY;
      }
      const float n0 = fmaxf(fmaxf(m0, fmaxf(v0[0], v0[1])), fmaxf(v0[2], v0[3]));
      const float n1 = fmaxf(fmaxf(m1, fmaxf(v1[0], v1[1])), fmaxf(v1[2], v1[3]));
      if (n0 > -INFINITY) l0 = l0 * exp2f(m0 - n0) + (exp2f(v0[0] - n0) + exp2f(v0[1] - n0)) + (exp2f(v0[2] - n0) + exp2f(v0[3] - n0));
      if (n1 > -INFINITY) l1 = l1 * exp2f(m1 - n1) + (exp2f(v1[0] - n1) + exp2f(v1[1] - n1)) + (exp2f(v1[2] - n1) + exp2f(v1[3] - n1));
      m0 = n0;
      m1 = n1;
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {          // the four lanes of a quad hold one row
      const float om0 = __shfl_xor_sync(0xffffffffu, m0, off), ol0 = __shfl_xor_sync(0xffffffffu, l0, off);
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, off), ol1 = __shfl_xor_sync(0xffffffffu, l1, off);
      const float n0 = fmaxf(m0, om0), n1 = fmaxf(m1, om1);
      l0 = (m0 > -INFINITY ? l0 * exp2f(m0 - n0) : 0.f) + (om0 > -INFINITY ? ol0 * exp2f(om0 - n0) : 0.f);
      l1 = (m1 > -INFINITY ? l1 * exp2f(m1 - n1) : 0.f) + (om1 > -INFINITY ? ol1 * exp2f(om1 - n1) : 0.f);
      m0 = n0;
      m1 = n1;
    }
    const int r0 = qb * 16 + g, r1 = r0 + 8;
    const float lse0 = r0 < n ? m0 + log2f(l0) : INFINITY, lse1 = r1 < n ? m1 + log2f(l1) : INFINITY;
    if (t == 0) {
      lse[r0] = lse0;
      lse[r1] = lse1;
    }
    const float D0 = Dv[r0], D1 = Dv[r1];
    uint32_t ag[4][4];
    load_a64(ag, gs + qb * blk_bytes, lane);
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
    for (int kb = 0; kb < nblk; ++kb) {
      float c[2][4], e[2][4];
      tile_xyT(c, aq, ks_ + kb * blk_bytes, lane);
      tile_xyT(e, ag, vs + kb * blk_bytes, lane);
      uint32_t ads[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = kb * 16 + nt * 8 + 2 * t;
        const float p00 = col < n ? exp2f(c[nt][0] * scale2 - lse0) : 0.f, p01 = col + 1 < n ? exp2f(c[nt][1] * scale2 - lse0) : 0.f;
        const float p10 = col < n ? exp2f(c[nt][2] * scale2 - lse1) : 0.f, p11 = col + 1 < n ? exp2f(c[nt][3] * scale2 - lse1) : 0.f;
        ads[2 * nt] = pack_bf16(p00 * (e[nt][0] - D0) * scale, p01 * (e[nt][1] - D0) * scale);
        ads[2 * nt + 1] = pack_bf16(p10 * (e[nt][2] - D1) * scale, p11 * (e[nt][3] - D1) * scale);
      }
      acc_a_z(dq, ads, ks_ + kb * blk_bytes, lane);             // dQ += dS K
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int col = h * DH + dt * 8 + 2 * t;
      if (r0 < n) *reinterpret_cast<uint32_t*>(dqkv + (row0 + r0) * 3ll * D + col) = pack_bf16(dq[dt][0], dq[dt][1]);
      if (r1 < n) *reinterpret_cast<uint32_t*>(dqkv + (row0 + r1) * 3ll * D + col) = pack_bf16(dq[dt][2], dq[dt][3]);
    }
  }
  __syncthreads();
  // ---- phase 2: warp = block of 16 keys, transposed tiles
  for (int kb = warp; kb < nblk; kb += kBwdMmaWarps) {
    uint32_t ak[4][4], av[4][4];
    load_a64(ak, ks_ + kb * blk_bytes, lane);
    load_a64(av, vs + kb * blk_bytes, lane);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
    const int k0 = kb * 16 + g, k1 = k0 + 8;                     // this thread's key rows
    for (int qb = 0; qb < nblk; ++qb) {
      float c[2][4], e[2][4];
      tile_xyT(c, ak, qs + qb * blk_bytes, lane);                // S^T  [keys x queries]
      tile_xyT(e, av, gs + qb * blk_bytes, lane);                // dP^T
      uint32_t ap[4], ads[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int q = qb * 16 + nt * 8 + 2 * t;
        const float2 ls = *reinterpret_cast<const float2*>(lse + q), dd = *reinterpret_cast<const float2*>(Dv + q);
        const float p00 = k0 < n ? exp2f(c[nt][0] * scale2 - ls.x) : 0.f, p01 = k0 < n ? exp2f(c[nt][1] * scale2 - ls.y) : 0.f;
        const float p10 = k1 < n ? exp2f(c[nt][2] * scale2 - ls.x) : 0.f, p11 = k1 < n ? exp2f(c[nt][3] * scale2 - ls.y) : 0.f;
        ap[2 * nt] = pack_bf16(p00, p01);
        ap[2 * nt + 1] = pack_bf16(p10, p11);
        ads[2 * nt] = pack_bf16(p00 * (e[nt][0] - dd.x) * scale, p01 * (e[nt][1] - dd.y) * scale);
        ads[2 * nt + 1] = pack_bf16(p10 * (e[nt][2] - dd.x) * scale, p11 * (e[nt][3] - dd.y) * scale);
      }
      acc_a_z(dv, ap, gs + qb * blk_bytes, lane);                // dV += P^T dO
      acc_a_z(dk, ads, qs + qb * blk_bytes, lane);               // dK += dS^T Q
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int col = h * DH + dt * 8 + 2 * t;
      if (k0 < n) {
        *reinterpret_cast<uint32_t*>(dqkv + (row0 + k0) * 3ll * D + D + col) = pack_bf16(dk[dt][0], dk[dt][1]);
        *reinterpret_cast<uint32_t*>(dqkv + (row0 + k0) * 3ll * D + 2 * D + col) = pack_bf16(dv[dt][0], dv[dt][1]);
      }
      if (k1 < n) {
        *reinterpret_cast<uint32_t*>(dqkv + (row0 + k1) * 3ll * D + D + col) = pack_bf16(dk[dt][2], dk[dt][3]);
        *reinterpret_cast<uint32_t*>(dqkv + (row0 + k1) * 3ll * D + 2 * D + col) = pack_bf16(dv[dt][2], dv[dt][3]);
      }
    }
  }
}

// ------------------------------------------------------------------ cross-entropy (mean over `inv_count`^-1 samples) + its gradient
// One warp per sample: loss_sum += -log softmax(logits)[label]; dlogits = (softmax - onehot) * inv_count; correct += argmax == label.
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int batch, int classes, float inv_count,
                    float* __restrict__ loss_sum, float* __restrict__ dlogits, int* __restrict__ correct) {
  const int lane = lane_id();
  const int b = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (b >= batch) return;
  const float* z = logits + static_cast<long long>(b) * classes;
  float mx = -INFINITY;
  int arg = 0;
  for (int c = lane; c < classes; c += 32) {
    const float v = z[c];
    if (v > mx) { mx = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float sum = 0.f;
  for (int c = lane; c < classes; c += 32) sum += __expf(z[c] - mx);
  sum = warp_sum(sum);
  const long long lab64 = labels[b];
  const bool valid = lab64 >= 0 && lab64 < classes;          // an out-of-range label contributes nothing (no out-of-bounds read);
  const int label = valid ? static_cast<int>(lab64) : -1;    // the host wrapper reports it (torch._assert_async)
  const float inv = 1.0f / sum;
  for (int c = lane; c < classes; c += 32)
    dlogits[static_cast<long long>(b) * classes + c] = valid ? (__expf(z[c] - mx) * inv - (c == label ? 1.0f : 0.0f)) * inv_count : 0.f;
  if (lane == 0 && valid) {
    atomicAdd(loss_sum, (logf(sum) + mx - z[label]) * inv_count);
    if (correct && arg == label) atomicAdd(correct, 1);
  }
}

// ------------------------------------------------------------------ head backward (models/vit.py:246: logits = feat W^T + b)
// dW[c, :] += sum_b dlogits[b, c] * feat[b, :],  db[c] += sum_b dlogits[b, c]: one CTA per class, threads over D.
__global__ void __launch_bounds__(256)
head_bwd_weight_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat, int batch, int classes, int dim,
                       float* __restrict__ dW, float* __restrict__ db) {
  const int c = blockIdx.x;
  extern __shared__ float s_dl[];                 // [batch]
  for (int b = threadIdx.x; b < batch; b += blockDim.x) s_dl[b] = dlogits[static_cast<long long>(b) * classes + c];
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < batch; ++b) acc += s_dl[b] * feat[static_cast<long long>(b) * dim + d];
    dW[static_cast<long long>(c) * dim + d] += acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < batch; ++b) acc += s_dl[b];
    db[c] += acc;
  }
}
// dfeat[b, :] = sum_c dlogits[b, c] * W[c, :]: one CTA per sample.
__global__ void __launch_bounds__(256)
head_bwd_input_kernel(const float* __restrict__ dlogits, const float* __restrict__ W, int classes, int dim, float* __restrict__ dfeat) {
  const int b = blockIdx.x;
  extern __shared__ float s_dl[];                 // [classes]
  for (int c = threadIdx.x; c < classes; c += blockDim.x) s_dl[c] = dlogits[static_cast<long long>(b) * classes + c];
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < classes; ++c) acc += s_dl[c] * W[static_cast<long long>(c) * dim + d];
    dfeat[static_cast<long long>(b) * dim + d] = acc;
  }
}

// out[t, :] += sum_b x[(b * seq + row0 + t), :]   (gradient of the class / register token parameters, vit.py:230-236)
__global__ void __launch_bounds__(256)
sum_token_rows_kernel(const float* __restrict__ x, int batch, int seq, int row0, int n_rows, int dim, float* __restrict__ out) {
  const int total = n_rows * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i / dim, d = i % dim;
    float acc = 0.f;
    for (int b = 0; b < batch; ++b) acc += x[(static_cast<long long>(b) * seq + row0 + t) * dim + d];
    out[i] += acc;
  }
}

// Backward of RankViT's gather (rankvit.py:69-77: class token + the kept tokens, no gradient through the indices):
// x[b * seq + tok(o)] = y[b * (k + 1) + o] with tok(0) = 0, tok(o) = 1 + kept[b, o - 1]; the caller zeroes x first (dropped
// tokens receive no gradient).  One warp per row.
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ y, float* __restrict__ x, const int* __restrict__ kept, int batch, int seq_len, int k, int dim) {
  const int lane = lane_id(), d4 = dim / 4;
  const int k1 = k + 1;
  const long long total = static_cast<long long>(batch) * k1;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (long long r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < total; r += warps_total) {
    const long long b = r / k1;
    const int o = static_cast<int>(r - b * k1);
    const int tok = o == 0 ? 0 : 1 + kept[b * k + (o - 1)];
    const float4* src = reinterpret_cast<const float4*>(y + r * dim);
    float4* dst = reinterpret_cast<float4*>(x + (b * seq_len + tok) * dim);
    for (int c = lane; c < d4; c += 32) dst[c] = src[c];
  }
}


// ------------------------------------------------------------------ gate regime of ResidualViT (residualvit.py:47-74,197-260)
// Training-mode block with a soft mask m >= 0 per image token (m = 1 on the class and budget rows):
//     mi = m * x      a = m * LN1(mi)      x1 = mi + m * (Wo attention(Wqkv a) + bo)      y = m * LN2(x1)      out = x1 + mlp(y)
//     m  = relu(sigmoid((x . w_g + b_g) / temp + bias) - thr_b),   thr_b = sigmoid(x_budget . w_bt + b_bt)
// The backbone is frozen; what trains is (w_g, b_g), (w_bt, b_bt), the learnable budget token, the class token and the head.
// d m collects four row dot products (dy . LN2(x1), dx1 . proj, da . LN1(mi), dmi . x); where m == 0 the relu passes nothing.

// y_bf16[r, :] = rowscale[r] * x[r, :]
__global__ void cast_rows_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ rowscale,
                                          long long n4, int d4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float sc = rowscale[i / d4];
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x * sc, v.y * sc), pack_bf16(v.z * sc, v.w * sc));
  }
}

// out[r] (+)= alpha * sum_d a[r,d] * (b[r,d] - c[r,d]) / div[r]      (c, div optional; div[r] <= 0 -> 0).  One warp per row.
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, const float* __restrict__ div,
              float* __restrict__ out, int rows, int dim, float alpha, int accumulate) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    float dv = 1.f;
    if (div) {
      dv = div[r];
      if (dv <= 0.f) {
        if (lane == 0 && !accumulate) out[r] = 0.f;
        continue;
      }
    }
    const float4* a4 = reinterpret_cast<const float4*>(a + static_cast<long long>(r) * dim);
    const float4* b4 = reinterpret_cast<const float4*>(b + static_cast<long long>(r) * dim);
    const float4* c4 = c ? reinterpret_cast<const float4*>(c + static_cast<long long>(r) * dim) : nullptr;
    float acc = 0.f;
    for (int k = lane; k < d4; k += 32) {
      const float4 av = a4[k];
      float4 bv = b4[k];
      if (c4) { const float4 cv = c4[k]; bv.x -= cv.x; bv.y -= cv.y; bv.z -= cv.z; bv.w -= cv.w; }
      acc += (av.x * bv.x + av.y * bv.y) + (av.z * bv.z + av.w * bv.w);
    }
    acc = warp_sum(acc) * alpha / dv;
    if (lane == 0) out[r] = accumulate ? out[r] + acc : acc;
  }
}

__device__ __forceinline__ float warp_rowdot(const float* __restrict__ xrow, const float* __restrict__ w, int d4, int lane) {
  float acc = 0.f;
  for (int k = lane; k < d4; k += 32) {
    const float4 xv = reinterpret_cast<const float4*>(xrow)[k], wv = __ldg(reinterpret_cast<const float4*>(w) + k);
    acc += (xv.x * wv.x + xv.y * wv.y) + (xv.z * wv.z + xv.w * wv.w);
  }
  return warp_sum(acc);
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + __expf(-z)); }

// Forward of the gate on the dense training layout [class, budget, image tokens ...] of every sample: one warp per row.
//   rowscale[r] = 1 on the special rows, m on image rows;  mask[b, t] = m,  sig[b, t] = the sigmoid value,  thr[b].
__global__ void __launch_bounds__(256)
gate_train_fwd_kernel(const float* __restrict__ x, int batch, int seq, int n_special, int budget_pos, const float* __restrict__ gate_w,
                      const float* __restrict__ gate_b_p, float inv_temp, float gate_bias, const float* __restrict__ bt_w,
                      const float* __restrict__ bt_b_p,
                      float* __restrict__ rowscale, float* __restrict__ mask, float* __restrict__ sig, float* __restrict__ thr, int dim) {
  const int lane = lane_id(), d4 = dim / 4, n_img = seq - n_special;
  const int rows = batch * seq, warps_total = gridDim.x * (blockDim.x >> 5);
  const float gate_b = __ldg(gate_b_p), bt_b = __ldg(bt_b_p);       // live parameters: read on the device, no host sync
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const int b = r / seq, t = r - b * seq;
    const float z = warp_rowdot(x + (static_cast<long long>(b) * seq + budget_pos) * dim, bt_w, d4, lane) + bt_b;
    const float th = sigmoidf_(z);                                 // residualvit.py:212
    if (t < n_special) {
      if (lane == 0) {
        rowscale[r] = 1.f;
        if (t == budget_pos) thr[b] = th;
      }
      continue;
    }
    const float logit = warp_rowdot(x + static_cast<long long>(r) * dim, gate_w, d4, lane) + gate_b;
    const float sg = sigmoidf_(logit * inv_temp + gate_bias);      // blocks.py:62-69
    const float m = fmaxf(sg - th, 0.f);                           // residualvit.py:63-64
    if (lane == 0) {
      rowscale[r] = m;
      mask[b * n_img + (t - n_special)] = m;
      sig[b * n_img + (t - n_special)] = sg;
    }
  }
}

// Backward of the gate: one CTA per sample.  dm[r] = the block's row dot products, dmask_ext (optional) = the gradient of a
// regulariser on the published mask (utils/losses.py).  For image rows with m > 0:
//   ds = dm + dmask_ext,  d thr -= ds,  dlogit = ds * s (1 - s) / temp,  dx[r] += dlogit * w_g,  d w_g += dlogit * x[r],  d b_g += dlogit
// then  dz = d thr * thr (1 - thr),  dx[budget row] += dz * w_bt,  d w_bt += dz * x[budget row],  d b_bt += dz.
constexpr int kGateBwdMaxV = 8;      // dim <= 1024
__global__ void __launch_bounds__(256)
gate_train_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dm, const float* __restrict__ dmask_ext,
                      const float* __restrict__ mask, const float* __restrict__ sig, const float* __restrict__ thr,
                      const float* __restrict__ gate_w, float inv_temp, const float* __restrict__ bt_w, int seq, int n_special,
                      int budget_pos, int dim, float* __restrict__ dx, float* __restrict__ g_gate_w, float* __restrict__ g_gate_b,
                      float* __restrict__ g_bt_w, float* __restrict__ g_bt_b) {
  extern __shared__ float gsm[];                 // [8 warps][dim] partial d w_g, then 8 + 8 scalars
  const int lane = lane_id(), warp = warp_id(), d4 = dim / 4, n_img = seq - n_special, b = blockIdx.x;
  float4 acc[kGateBwdMaxV];
#pragma unroll
  for (int j = 0; j < kGateBwdMaxV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dthr = 0.f, dbias = 0.f;
  for (int t = warp; t < n_img; t += 8) {
    const float m = mask[b * n_img + t];
    if (m <= 0.f) continue;                      // relu: nothing passes
    const long long r = static_cast<long long>(b) * seq + n_special + t;
    const float ds = dm[r] + (dmask_ext ? dmask_ext[b * n_img + t] : 0.f);
    const float sg = sig[b * n_img + t];
    const float dlogit = ds * sg * (1.f - sg) * inv_temp;
    dthr -= ds;
    dbias += dlogit;
    const float4* x4 = reinterpret_cast<const float4*>(x + r * dim);
    float4* dx4 = reinterpret_cast<float4*>(dx + r * dim);
#pragma unroll
    for (int j = 0; j < kGateBwdMaxV; ++j) {
      const int c = lane + 32 * j;
      if (c < d4) {
        const float4 xv = x4[c], wv = __ldg(reinterpret_cast<const float4*>(gate_w) + c);
        acc[j].x += dlogit * xv.x; acc[j].y += dlogit * xv.y; acc[j].z += dlogit * xv.z; acc[j].w += dlogit * xv.w;
        float4 o = dx4[c];
        o.x += dlogit * wv.x; o.y += dlogit * wv.y; o.z += dlogit * wv.z; o.w += dlogit * wv.w;
        dx4[c] = o;
      }
    }
  }
  float* part = gsm + warp * dim;
#pragma unroll
  for (int j = 0; j < kGateBwdMaxV; ++j) {
    const int c = lane + 32 * j;
    if (c < d4) reinterpret_cast<float4*>(part)[c] = acc[j];
  }
  float* scal = gsm + 8 * dim;
  if (lane == 0) { scal[warp] = dthr; scal[8 + warp] = dbias; }      // every lane of a warp holds the same dthr / dbias
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += gsm[w * dim + d];
    if (v != 0.f) atomicAdd(g_gate_w + d, v);
  }
  float dthr_b = 0.f, dbias_b = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { dthr_b += scal[w]; dbias_b += scal[8 + w]; }
  if (threadIdx.x == 0 && dbias_b != 0.f) atomicAdd(g_gate_b, dbias_b);
  const float th = thr[b];
  const float dz = dthr_b * th * (1.f - th);
  if (dz == 0.f) return;
  const long long rb = static_cast<long long>(b) * seq + budget_pos;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    atomicAdd(g_bt_w + d, dz * x[rb * dim + d]);
    dx[rb * dim + d] += dz * __ldg(bt_w + d);
  }
  if (threadIdx.x == 0) atomicAdd(g_bt_b, dz);
}
}  // namespace pk

using namespace pk;

extern "C" int pk_scatter_rows(const float* y, float* x, const int* kept, int batch, int seq_len, int k, int dim, void* stream) {
  PK_REQUIRE(y && x && (kept || k == 0) && dim % 4 == 0 && batch >= 0 && k >= 0 && k < seq_len, "pk_scatter_rows: bad arguments");
  if (batch == 0) return PK_OK;
  scatter_rows_kernel<<<train_grid(static_cast<long long>(batch) * (k + 1), 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, x, kept, batch, seq_len, k, dim);
  return check_cuda(cudaGetLastError(), "scatter_rows_kernel");
}

extern "C" int pk_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  PK_REQUIRE(x && y && n >= 0 && n % 4 == 0, "pk_cast_f32_bf16: bad arguments (n must be a multiple of 4)");
  if (n == 0) return PK_OK;
  cast_f32_bf16_kernel<<<train_grid(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), n / 4);
  return check_cuda(cudaGetLastError(), "cast_f32_bf16_kernel");
}

extern "C" int pk_gelu_bf16(const void* h_pre, void* hid, long long n, void* stream) {
  PK_REQUIRE(h_pre && hid && n >= 0 && n % 8 == 0, "pk_gelu_bf16: bad arguments (n must be a multiple of 8)");
  if (n == 0) return PK_OK;
  gelu_bf16_kernel<<<train_grid(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(h_pre), static_cast<__nv_bfloat16*>(hid), n / 8);
  return check_cuda(cudaGetLastError(), "gelu_bf16_kernel");
}

extern "C" int pk_gelu_bwd_bf16(const void* h_pre, const void* dhid, void* dh_pre, long long n, void* stream) {
  PK_REQUIRE(h_pre && dhid && dh_pre && n >= 0 && n % 8 == 0, "pk_gelu_bwd_bf16: bad arguments (n must be a multiple of 8)");
  if (n == 0) return PK_OK;
  gelu_bwd_bf16_kernel<<<train_grid(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(h_pre), static_cast<const __nv_bfloat16*>(dhid), static_cast<__nv_bfloat16*>(dh_pre), n / 8);
  return check_cuda(cudaGetLastError(), "gelu_bwd_bf16_kernel");
}

static int launch_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* beta, float eps, float* dx, int rows,
                                int dim, const int* row_index, int dy_div, int accumulate, const float* rowscale, float* dot_out,
                                cudaStream_t s) {
  const int grid = train_grid(rows, 8);
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 2) layernorm_bwd_kernel<2><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate, beta, rowscale, dot_out);
  else if (maxv <= 3) layernorm_bwd_kernel<3><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate, beta, rowscale, dot_out);
  else if (maxv <= 6) layernorm_bwd_kernel<6><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate, beta, rowscale, dot_out);
  else layernorm_bwd_kernel<8><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate, beta, rowscale, dot_out);
  return check_cuda(cudaGetLastError(), "layernorm_bwd_kernel");
}

extern "C" int pk_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, float* dx, int rows, int dim,
                                const int* row_index, int dy_div, int accumulate, void* stream) {
  PK_REQUIRE(x && dy && gamma && dx, "pk_layernorm_bwd: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024 && rows >= 0 && dy_div >= 1, "pk_layernorm_bwd: dim %d must be a multiple of 4 in [4,1024]", dim);
  if (rows == 0) return PK_OK;
  return launch_layernorm_bwd(x, dy, gamma, nullptr, eps, dx, rows, dim, row_index, dy_div, accumulate, nullptr, nullptr,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int pk_layernorm_bwd_gated(const float* x, const float* dy, const float* gamma, const float* beta, float eps, float* dx,
                                      int rows, int dim, const float* rowscale, float* dot_out, int accumulate, void* stream) {
  PK_REQUIRE(x && dy && gamma && beta && dx && rowscale && dot_out, "pk_layernorm_bwd_gated: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024 && rows >= 0, "pk_layernorm_bwd_gated: dim %d must be a multiple of 4 in [4,1024]", dim);
  if (rows == 0) return PK_OK;
  return launch_layernorm_bwd(x, dy, gamma, beta, eps, dx, rows, dim, nullptr, 1, accumulate, rowscale, dot_out,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int pk_cast_rows_f32_bf16(const float* x, void* y, const float* rowscale, int rows, int dim, void* stream) {
  PK_REQUIRE(x && y && rowscale && rows >= 0 && dim > 0 && dim % 4 == 0, "pk_cast_rows_f32_bf16: bad arguments (dim must be a multiple of 4)");
  if (rows == 0) return PK_OK;
  const long long n4 = static_cast<long long>(rows) * (dim / 4);
  cast_rows_f32_bf16_kernel<<<train_grid(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), rowscale, n4, dim / 4);
  return check_cuda(cudaGetLastError(), "cast_rows_f32_bf16_kernel");
}

extern "C" int pk_rowdot(const float* a, const float* b, const float* c, const float* div, float* out, int rows, int dim, float alpha,
                         int accumulate, void* stream) {
  PK_REQUIRE(a && b && out && rows >= 0 && dim > 0 && dim % 4 == 0, "pk_rowdot: bad arguments (dim must be a multiple of 4)");
  if (rows == 0) return PK_OK;
  rowdot_kernel<<<train_grid(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, c, div, out, rows, dim, alpha, accumulate);
  return check_cuda(cudaGetLastError(), "rowdot_kernel");
}

extern "C" int pk_residual_gate_train_fwd(const float* x, int batch, int seq, int n_special, int budget_pos, int dim, const float* gate_w,
                                          const float* gate_b, float gate_temp, float gate_bias, const float* bt_w, const float* bt_b,
                                          float* rowscale, float* mask, float* sig, float* thr, void* stream) {
  PK_REQUIRE(x && gate_w && gate_b && bt_w && bt_b && rowscale && mask && sig && thr, "pk_residual_gate_train_fwd: null pointer");
  PK_REQUIRE(batch >= 0 && n_special >= 1 && seq > n_special && budget_pos >= 0 && budget_pos < n_special && dim % 4 == 0 && dim >= 4 &&
             gate_temp != 0.f, "pk_residual_gate_train_fwd: bad shape");
  if (batch == 0) return PK_OK;
  gate_train_fwd_kernel<<<train_grid(static_cast<long long>(batch) * seq, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, batch, seq, n_special, budget_pos, gate_w, gate_b, 1.f / gate_temp, gate_bias, bt_w, bt_b, rowscale, mask, sig, thr, dim);
  return check_cuda(cudaGetLastError(), "gate_train_fwd_kernel");
}

extern "C" int pk_residual_gate_train_bwd(const float* x, const float* dm, const float* dmask_ext, const float* mask, const float* sig,
                                          const float* thr, int batch, int seq, int n_special, int budget_pos, int dim,
                                          const float* gate_w, float gate_temp, const float* bt_w, float* dx, float* g_gate_w,
                                          float* g_gate_b, float* g_bt_w, float* g_bt_b, void* stream) {
  PK_REQUIRE(x && dm && mask && sig && thr && gate_w && bt_w && dx && g_gate_w && g_gate_b && g_bt_w && g_bt_b,
             "pk_residual_gate_train_bwd: null pointer");
  PK_REQUIRE(batch >= 0 && n_special >= 1 && seq > n_special && budget_pos >= 0 && budget_pos < n_special && dim % 4 == 0 && dim >= 4 &&
             dim <= 1024 && gate_temp != 0.f, "pk_residual_gate_train_bwd: bad shape (dim a multiple of 4, <= 1024)");
  if (batch == 0) return PK_OK;
  const size_t smem = (static_cast<size_t>(8) * dim + 16) * sizeof(float);
  gate_train_bwd_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, dm, dmask_ext, mask, sig, thr, gate_w, 1.f / gate_temp, bt_w, seq, n_special, budget_pos, dim, dx, g_gate_w, g_gate_b, g_bt_w, g_bt_b);
  return check_cuda(cudaGetLastError(), "gate_train_bwd_kernel");
}

template <int DH>
static int launch_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, int batch, int num_heads, int seq_len,
                                float scale, cudaStream_t s) {
  const size_t kv = (static_cast<size_t>(2) * seq_len * (DH + 2) * 2 + 15) & ~static_cast<size_t>(15);
  const size_t bytes = kv + static_cast<size_t>(kBwdWarps) * (2 * DH + 512) * 4;
  PK_REQUIRE(bytes <= 232448, "pk_attention_bwd: %d tokens x head_dim %d need %zu bytes of shared memory", seq_len, DH, bytes);
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  attention_bwd_kernel<DH><<<dim3(num_heads, batch), kBwdThreads, bytes, s>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout),
      static_cast<__nv_bfloat16*>(dqkv), num_heads, seq_len, scale);
  return check_cuda(cudaGetLastError(), "attention_bwd_kernel");
}

extern "C" int pk_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, int batch, int num_heads, int head_dim,
                                int seq_len, float scale, void* stream) {
  PK_REQUIRE(qkv && out && dout && dqkv, "pk_attention_bwd: null pointer");
  PK_REQUIRE(head_dim == 64 || head_dim == 32, "pk_attention_bwd: head_dim %d not in {32, 64}", head_dim);
  PK_REQUIRE(seq_len >= 1 && seq_len <= 256 && batch >= 0 && num_heads > 0, "pk_attention_bwd: uniform sequences of 1 .. 256 tokens (got %d)", seq_len);
  if (batch == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (head_dim == 64) {
    // PK_ATT_BWD_MMA=0: the CUDA-core kernel (A/B runs; it is also the head_dim 32 path)
    static int use_mma = -1;
    if (use_mma < 0) { const char* e = getenv("PK_ATT_BWD_MMA"); use_mma = (e && e[0] == '0') ? 0 : 1; }
    const bool aligned = ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0;
    if (use_mma && aligned) {
      const int NP = (seq_len + 15) & ~15;
      const size_t bytes = static_cast<size_t>(4) * NP * kMmaPitch * 2 + static_cast<size_t>(2) * NP * 4;
      static bool attr_set = false;
      if (!attr_set) {
        PK_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
      }
      attention_bwd_mma_kernel<<<dim3(num_heads, batch), kBwdMmaThreads, bytes, s>>>(
          static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout),
          static_cast<__nv_bfloat16*>(dqkv), num_heads, seq_len, scale);
      return check_cuda(cudaGetLastError(), "attention_bwd_mma_kernel");
    }
    return launch_attention_bwd<64>(qkv, out, dout, dqkv, batch, num_heads, seq_len, scale, s);
  }
  return launch_attention_bwd<32>(qkv, out, dout, dqkv, batch, num_heads, seq_len, scale, s);
}

extern "C" int pk_softmax_xent(const float* logits, const long long* labels, int batch, int classes, float inv_count, float* loss_sum,
                               float* dlogits, int* correct, void* stream) {
  PK_REQUIRE(logits && labels && loss_sum && dlogits && batch >= 0 && classes > 0, "pk_softmax_xent: bad arguments");
  if (batch == 0) return PK_OK;
  softmax_xent_kernel<<<(batch + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, labels, batch, classes, inv_count, loss_sum,
                                                                                    dlogits, correct);
  return check_cuda(cudaGetLastError(), "softmax_xent_kernel");
}

extern "C" int pk_head_bwd(const float* dlogits, const float* feat, const float* weight, int batch, int classes, int dim, float* d_weight,
                           float* d_bias, float* d_feat, void* stream) {
  PK_REQUIRE(dlogits && feat && weight && d_weight && d_bias && d_feat, "pk_head_bwd: null pointer");
  PK_REQUIRE(batch >= 0 && batch <= 8192 && classes > 0 && classes <= 8192 && dim > 0, "pk_head_bwd: batch / classes must be <= 8192");
  if (batch == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  head_bwd_weight_kernel<<<classes, 256, static_cast<size_t>(batch) * 4, s>>>(dlogits, feat, batch, classes, dim, d_weight, d_bias);
  PK_CHECK_CUDA(cudaGetLastError());
  head_bwd_input_kernel<<<batch, 256, static_cast<size_t>(classes) * 4, s>>>(dlogits, weight, classes, dim, d_feat);
  return check_cuda(cudaGetLastError(), "head_bwd kernels");
}

extern "C" int pk_sum_token_rows(const float* x, int batch, int seq, int row0, int n_rows, int dim, float* out, void* stream) {
  PK_REQUIRE(x && out && batch >= 0 && seq > 0 && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= seq && dim > 0, "pk_sum_token_rows: bad arguments");
  if (batch == 0 || n_rows == 0) return PK_OK;
  sum_token_rows_kernel<<<train_grid(static_cast<long long>(n_rows) * dim, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, batch, seq, row0, n_rows, dim, out);
  return check_cuda(cudaGetLastError(), "sum_token_rows_kernel");
}
