// peekvit_b200 — fp32-accurate ("exact") mode kernels (sm_100a).
//
// The reference runs in fp32 as shipped; BASELINE's north star asks for logits within 1e-5 in an fp32 mode next to the
// bf16 headline mode.  The GEMMs of this mode run on the SAME tcgen05 bf16 kernels: an fp32 value is split into three
// bf16 terms x = h + m + l (8 + 8 + 8 mantissa bits, each difference exact in fp32), and the six products that matter
//     m*m, l*h, h*l, m*h, h*m, h*h            (dropped: m*l, l*m, l*l ~ 2^-24)
// are laid side by side along K, so one bf16 GEMM with K' = 6K accumulates them all in the fp32 TMEM accumulator.  The
// small terms come FIRST: the tensor core aligns every partial sum to the running accumulator, so adding the 2^-16 terms
// to an already large sum loses them (measured, tools/split_probe.py, K = 3072: big-first 3.1e-5, small-first 3.6e-6 of
// max|out| against a float64 reference, where torch's own fp32 matmul has 2.2e-6).
//   activations  A' = [ m | l | h | m | h | h ]      (pk_split3_bf16, optionally after LayerNorm or exact GELU)
//   weights      W' = [ m | h | l | h | m | h ]      (prepacked on the host side once per weight version)
// Attention runs in plain fp32 on the CUDA cores (pk_attention_f32): its share of the FLOPs is 4 %.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

namespace pk {

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

__device__ __forceinline__ uint2 pack4(__nv_bfloat16 a, __nv_bfloat16 b, __nv_bfloat16 c, __nv_bfloat16 d) {
  return make_uint2(static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16),
                    static_cast<uint32_t>(__bfloat16_as_ushort(c)) | (static_cast<uint32_t>(__bfloat16_as_ushort(d)) << 16));
}

// four consecutive fp32 values of one row -> the K-wide segments of the split operand row (activation order).
// LAYOUT 6: three terms, [m | l | h | m | h | h] (fp32-accurate mode); LAYOUT 2: two terms, [lo | hi] (bf16x2 mode, read
// with a_wrap_k); LAYOUT 3: two terms, [lo | hi | hi] (bf16x2 patch operand, no wrap).
template <int LAYOUT>
__device__ __forceinline__ void store_split4(__nv_bfloat16* orow, int K, int k, float4 v) {
  if constexpr (LAYOUT == 6) {
    __nv_bfloat16 h[4], m[4], l[4];
    split3(v.x, h[0], m[0], l[0]);
    split3(v.y, h[1], m[1], l[1]);
    split3(v.z, h[2], m[2], l[2]);
    split3(v.w, h[3], m[3], l[3]);
    const uint2 ph = pack4(h[0], h[1], h[2], h[3]), pm = pack4(m[0], m[1], m[2], m[3]), pl = pack4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint2*>(orow + 0 * K + k) = pm;
    *reinterpret_cast<uint2*>(orow + 1 * K + k) = pl;
    *reinterpret_cast<uint2*>(orow + 2 * K + k) = ph;
    *reinterpret_cast<uint2*>(orow + 3 * K + k) = pm;
    *reinterpret_cast<uint2*>(orow + 4 * K + k) = ph;
    *reinterpret_cast<uint2*>(orow + 5 * K + k) = ph;
  } else {
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __float2bfloat16_rn(x[i]);
      l[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h[i]));
    }
    const uint2 ph = pack4(h[0], h[1], h[2], h[3]), pl = pack4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint2*>(orow + 0 * K + k) = pl;
    *reinterpret_cast<uint2*>(orow + 1 * K + k) = ph;
    if constexpr (LAYOUT == 3) *reinterpret_cast<uint2*>(orow + 2 * K + k) = ph;
  }
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// MODE 0: split only; MODE 1: exact (erf) GELU first (reference blocks.py:82)
template <int MODE, int LAYOUT>
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int K,
                              const float* __restrict__ rowscale, const int* __restrict__ rows_dev) {
  const int k4 = K / 4;
  if (rows_dev) rows = min(static_cast<long long>(*rows_dev), rows);
  const long long total = rows * k4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / k4;
    const int k = static_cast<int>(i - r * k4) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * K + k);
    if (MODE == 1) { v.x = gelu_exact(v.x); v.y = gelu_exact(v.y); v.z = gelu_exact(v.z); v.w = gelu_exact(v.w); }
    if (rowscale) { const float sc = rowscale[r]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
    store_split4<LAYOUT>(out + r * static_cast<long long>(LAYOUT) * K, K, k, v);
  }
}

// LayerNorm (two-pass fp32 statistics, biased variance: nn.LayerNorm) then split.  One warp per row, the row in registers.
template <int MAXV, int LAYOUT>
__global__ void __launch_bounds__(256)
split3_layernorm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, int rows, int K, const float* __restrict__ rowscale,
                        const int* __restrict__ row_index, const int* __restrict__ rows_dev) {
  const int lane = lane_id();
  const int d4 = K / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  if (rows_dev) rows = min(*rows_dev, rows);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const long long src = row_index ? row_index[r] : r;          // output row r is LN(x[row_index[r]]) (expert-sorted order)
    const float sc = rowscale ? rowscale[r] : 1.0f;
    float4 v[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d4) {
        v[i] = *reinterpret_cast<const float4*>(x + src * K + c * 4);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(sum) / static_cast<float>(K);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / static_cast<float>(K) + eps);
    __nv_bfloat16* orow = out + static_cast<long long>(r) * LAYOUT * K;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c * 4), b = *reinterpret_cast<const float4*>(beta + c * 4);
        float4 o;
        o.x = sc * ((v[i].x - mean) * rstd * g.x + b.x);
        o.y = sc * ((v[i].y - mean) * rstd * g.y + b.y);
        o.z = sc * ((v[i].z - mean) * rstd * g.z + b.z);
        o.w = sc * ((v[i].w - mean) * rstd * g.w + b.w);
        store_split4<LAYOUT>(orow, K, c * 4, o);
      }
    }
  }
}

// im2col of fp32 NCHW images straight into split rows (patch q of sample b -> row b*P + q, K order (c,i,j)).
template <int LAYOUT>
__global__ void patchify_split3_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int S, int p, int n_side,
                                       long long total_chunks) {
  const int Kp = 3 * p * p;
  const int P = n_side * n_side;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_chunks;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = idx * 4;
    const long long row = e / Kp;
    const int k = static_cast<int>(e - row * Kp);
    const int c = k / (p * p);
    const int i = (k - c * p * p) / p;
    const int j = k - c * p * p - i * p;
    const long long b = row / P;
    const int patch = static_cast<int>(row - b * P);
    const int py = patch / n_side, px = patch - py * n_side;
    const float4 v = *reinterpret_cast<const float4*>(img + ((b * 3 + c) * S + (py * p + i)) * static_cast<long long>(S) + px * p + j);
    store_split4<LAYOUT>(out + row * static_cast<long long>(LAYOUT) * Kp, Kp, k, v);
  }
}

// ------------------------------------------------------------------------------ fp32 attention (CUDA cores)
// softmax(q k^T * scale) v; qkv fp32 [rows, 3*H*DH] (q | k | v, heads = DH-column slices), out fp32 [rows, H*DH].  Samples are
// uniform (n rows each) or ragged (cu_seqlens); a key may carry a multiplicity (logit + log mult: that many identical tokens)
// and every sample one virtual key (k, v) = extra_kv with multiplicity extra_mult[b] — what the compacted ResidualViT / A-ViT
// rows need.
// A CTA of 128 threads serves 128 queries of one (sample, head) and streams K / V through shared memory in tiles of 16 keys.
// Two adjacent lanes form a pair that owns TWO queries; each lane holds one half of the head dimension of both (q and the
// output accumulator in registers), so every K / V value read from shared memory feeds two FMAs (one query per thread made
// the kernel shared-memory-bound at a 1:4 load-to-FMA ratio); the two halves of a logit meet through one shuffle.
template <int DH>
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int batch, int heads, int n_uniform, float scale,
                     const int* __restrict__ cu_seqlens, const float* __restrict__ key_mult, const float* __restrict__ extra_kv,
                     const float* __restrict__ extra_mult) {
  constexpr int KT = 16;
  constexpr int HD = DH / 2;                       // head-dim half owned by a lane
  __shared__ __align__(16) float ks[KT][DH];
  __shared__ __align__(16) float vs[KT][DH];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh - b * heads;
  const int D = heads * DH;
  const long long row0 = cu_seqlens ? cu_seqlens[b] : static_cast<long long>(b) * n_uniform;
  const int n = cu_seqlens ? cu_seqlens[b + 1] - cu_seqlens[b] : n_uniform;
  if (blockIdx.x * 128 >= n) return;               // block-uniform: ragged grids are sized for the longest sample
  const int pair = threadIdx.x >> 1, hf = threadIdx.x & 1;
  const int d0 = hf * HD;
  const int qa = blockIdx.x * 128 + pair, qb = qa + 64;
  const bool va = qa < n, vb = qb < n;
  const float* base = qkv + row0 * 3 * D + h * DH;
  // q pre-scaled like nn.MultiheadAttention, times log2(e): the softmax runs in the base-2 domain (one MUFU ex2 per exponential)
  const float qs = scale * 1.4426950408889634f;
  float qA[HD], qB[HD], oA[HD], oB[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
    if (va) ta = *reinterpret_cast<const float4*>(base + static_cast<long long>(qa) * 3 * D + d0 + d);
    if (vb) tb = *reinterpret_cast<const float4*>(base + static_cast<long long>(qb) * 3 * D + d0 + d);
    qA[d] = ta.x * qs; qA[d + 1] = ta.y * qs; qA[d + 2] = ta.z * qs; qA[d + 3] = ta.w * qs;
    qB[d] = tb.x * qs; qB[d + 1] = tb.y * qs; qB[d + 2] = tb.z * qs; qB[d + 3] = tb.w * qs;
    oA[d] = oA[d + 1] = oA[d + 2] = oA[d + 3] = 0.f;
    oB[d] = oB[d + 1] = oB[d + 2] = oB[d + 3] = 0.f;
  }
  float mxA = -INFINITY, mxB = -INFINITY, sumA = 0.f, sumB = 0.f;
  for (int j0 = 0; j0 < n; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * DH / 4; i += 128) {
      const int j = i / (DH / 4), d = (i - j * (DH / 4)) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (j0 + j < n) {
        const float* row = base + static_cast<long long>(j0 + j) * 3 * D;
        kk = *reinterpret_cast<const float4*>(row + D + d);
        vv = *reinterpret_cast<const float4*>(row + 2 * D + d);
      }
      *reinterpret_cast<float4*>(&ks[j][d]) = kk;
      *reinterpret_cast<float4*>(&vs[j][d]) = vv;
    }
    __syncthreads();
    float sA[KT], sB[KT];
    float tA = mxA, tB = mxB;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float a = 0.f, c = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&ks[j][d0 + d]);
        a = fmaf(qA[d], kk.x, a); a = fmaf(qA[d + 1], kk.y, a); a = fmaf(qA[d + 2], kk.z, a); a = fmaf(qA[d + 3], kk.w, a);
        c = fmaf(qB[d], kk.x, c); c = fmaf(qB[d + 1], kk.y, c); c = fmaf(qB[d + 2], kk.z, c); c = fmaf(qB[d + 3], kk.w, c);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);      // the other half of the head dimension (commutative: both lanes agree)
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      if (j0 + j < n) {
        if (key_mult) { const float lm = log2f(key_mult[row0 + j0 + j]); a += lm; c += lm; }
      } else {
        a = -INFINITY; c = -INFINITY;
      }
      sA[j] = a; sB[j] = c;
      tA = fmaxf(tA, a); tB = fmaxf(tB, c);
    }
    const float corrA = exp2f(mxA - tA), corrB = exp2f(mxB - tB);          // 2^(-inf) = 0 on the first tile
    mxA = tA; mxB = tB;
    sumA *= corrA; sumB *= corrB;
#pragma unroll
    for (int d = 0; d < HD; ++d) { oA[d] *= corrA; oB[d] *= corrB; }
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float pa = exp2f(sA[j] - mxA), pb = exp2f(sB[j] - mxB);        // masked keys: 2^(-inf) = 0
      sumA += pa; sumB += pb;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&vs[j][d0 + d]);
        oA[d] = fmaf(pa, vv.x, oA[d]); oA[d + 1] = fmaf(pa, vv.y, oA[d + 1]); oA[d + 2] = fmaf(pa, vv.z, oA[d + 2]); oA[d + 3] = fmaf(pa, vv.w, oA[d + 3]);
        oB[d] = fmaf(pb, vv.x, oB[d]); oB[d + 1] = fmaf(pb, vv.y, oB[d + 1]); oB[d + 2] = fmaf(pb, vv.z, oB[d + 2]); oB[d + 3] = fmaf(pb, vv.w, oB[d + 3]);
      }
    }
  }
  const float em = extra_mult ? extra_mult[b] : 0.f;
  if (extra_kv && em > 0.f) {
    // the virtual key: what em zero tokens project to (k-bias | v-bias)
    const float* ke = extra_kv + h * DH + d0;
    const float* ve = extra_kv + D + h * DH + d0;
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) { a = fmaf(qA[d], ke[d], a); c = fmaf(qB[d], ke[d], c); }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    const float lm = log2f(em);
    a += lm; c += lm;
    const float tA = fmaxf(mxA, a), tB = fmaxf(mxB, c);
    const float corrA = exp2f(mxA - tA), corrB = exp2f(mxB - tB);
    const float pa = exp2f(a - tA), pb = exp2f(c - tB);
    sumA = sumA * corrA + pa; sumB = sumB * corrB + pb;
#pragma unroll
    for (int d = 0; d < HD; ++d) { oA[d] = fmaf(pa, ve[d], oA[d] * corrA); oB[d] = fmaf(pb, ve[d], oB[d] * corrB); }
  }
  if (va) {
    const float inv = 1.0f / sumA;
    float* orow = out + (row0 + qa) * D + h * DH + d0;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(orow + d) = make_float4(oA[d] * inv, oA[d + 1] * inv, oA[d + 2] * inv, oA[d + 3] * inv);
  }
  if (vb) {
    const float inv = 1.0f / sumB;
    float* orow = out + (row0 + qb) * D + h * DH + d0;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(orow + d) = make_float4(oB[d] * inv, oB[d + 1] * inv, oB[d + 2] * inv, oB[d + 3] * inv);
  }
}

static int grid_1d(long long items, int per_block) {
  const long long blocks = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace pk

namespace pk {
template <int LAYOUT>
static int launch_split(const char* who, const float* x, void* out, int rows, int dim, int mode, const float* gamma, const float* beta,
                        float eps, const float* rowscale, const int* row_index, const int* rows_dev, void* stream) {
  PK_REQUIRE(x && out && rows >= 0 && dim > 0 && dim % 8 == 0, "%s: null pointer or dim %% 8 != 0", who);
  PK_REQUIRE(mode >= 0 && mode <= 2, "%s: mode must be 0 (none), 1 (GELU) or 2 (LayerNorm)", who);
  if (rows == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (mode == 2) {
    PK_REQUIRE(gamma && beta && dim <= 1024, "%s: LayerNorm mode needs gamma / beta and dim <= 1024", who);
    const int grid = grid_1d(rows, 8);
    const int maxv = (dim / 4 + 31) / 32;
    if (maxv <= 2) split3_layernorm_kernel<2, LAYOUT><<<grid, 256, 0, s>>>(x, o, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
    else if (maxv <= 4) split3_layernorm_kernel<4, LAYOUT><<<grid, 256, 0, s>>>(x, o, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
    else split3_layernorm_kernel<8, LAYOUT><<<grid, 256, 0, s>>>(x, o, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
    return check_cuda(cudaGetLastError(), "split_layernorm_kernel");
  }
  PK_REQUIRE(row_index == nullptr, "%s: row_index is a LayerNorm-mode argument", who);
  const long long total = static_cast<long long>(rows) * (dim / 4);
  if (mode == 1) split3_kernel<1, LAYOUT><<<grid_1d(total, 256), 256, 0, s>>>(x, o, rows, dim, rowscale, rows_dev);
  else split3_kernel<0, LAYOUT><<<grid_1d(total, 256), 256, 0, s>>>(x, o, rows, dim, rowscale, rows_dev);
  return check_cuda(cudaGetLastError(), "split_kernel");
}

template <int LAYOUT>
static int launch_patchify_split(const char* who, const float* images, void* patches, int batch, int image_size, int patch_size, void* stream) {
  PK_REQUIRE(images && patches, "%s: null pointer", who);
  PK_REQUIRE(patch_size % 8 == 0 && image_size % patch_size == 0, "%s: patch_size must be a multiple of 8 dividing image_size", who);
  if (batch == 0) return PK_OK;
  const int n_side = image_size / patch_size;
  const long long total = static_cast<long long>(batch) * n_side * n_side * 3 * patch_size * patch_size / 4;
  patchify_split3_kernel<LAYOUT><<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, static_cast<__nv_bfloat16*>(patches), image_size, patch_size, n_side, total);
  return check_cuda(cudaGetLastError(), "patchify_split_kernel");
}
}  // namespace pk

extern "C" int pk_split3_bf16(const float* x, void* out, int rows, int dim, int mode, const float* gamma, const float* beta, float eps,
                              const float* rowscale, const int* row_index, const int* rows_dev, void* stream) {
  return pk::launch_split<6>("pk_split3_bf16", x, out, rows, dim, mode, gamma, beta, eps, rowscale, row_index, rows_dev, stream);
}

extern "C" int pk_split2_bf16(const float* x, void* out, int rows, int dim, int mode, const float* gamma, const float* beta, float eps,
                              const float* rowscale, const int* row_index, const int* rows_dev, void* stream) {
  return pk::launch_split<2>("pk_split2_bf16", x, out, rows, dim, mode, gamma, beta, eps, rowscale, row_index, rows_dev, stream);
}

extern "C" int pk_patchify_split3(const float* images, void* patches6, int batch, int image_size, int patch_size, void* stream) {
  return pk::launch_patchify_split<6>("pk_patchify_split3", images, patches6, batch, image_size, patch_size, stream);
}

extern "C" int pk_patchify_split2(const float* images, void* patches3, int batch, int image_size, int patch_size, void* stream) {
  return pk::launch_patchify_split<3>("pk_patchify_split2", images, patches3, batch, image_size, patch_size, stream);
}

extern "C" int pk_attention_f32(const float* qkv, float* out, int batch, int num_heads, int head_dim, int seq_len, float scale,
                                const int* cu_seqlens, const float* key_mult, const float* extra_kv, const float* extra_mult,
                                void* stream) {
  using namespace pk;
  PK_REQUIRE(qkv && out && batch >= 0 && num_heads > 0 && seq_len > 0, "pk_attention_f32: bad arguments (seq_len = longest sample)");
  PK_REQUIRE(head_dim == 32 || head_dim == 48 || head_dim == 64, "pk_attention_f32: head_dim must be 32, 48 or 64");
  if (batch == 0) return PK_OK;
  const dim3 grid((seq_len + 127) / 128, batch * num_heads);
  PK_REQUIRE(grid.y <= 65535u, "pk_attention_f32: batch * heads = %u exceeds 65535 (use smaller micro-batches)", grid.y);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (head_dim == 64)
    attention_f32_kernel<64><<<grid, 128, 0, s>>>(qkv, out, batch, num_heads, seq_len, scale, cu_seqlens, key_mult, extra_kv, extra_mult);
  else if (head_dim == 48)
    attention_f32_kernel<48><<<grid, 128, 0, s>>>(qkv, out, batch, num_heads, seq_len, scale, cu_seqlens, key_mult, extra_kv, extra_mult);
  else
    attention_f32_kernel<32><<<grid, 128, 0, s>>>(qkv, out, batch, num_heads, seq_len, scale, cu_seqlens, key_mult, extra_kv, extra_mult);
  return check_cuda(cudaGetLastError(), "attention_f32_kernel");
}
