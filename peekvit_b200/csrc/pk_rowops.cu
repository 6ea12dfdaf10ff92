// peekvit_b200 — HBM-bound row kernels: patchify, token rows, LayerNorm, class head,
// RankViT token score / stable top-k / compaction.  128-bit coalesced accesses, one warp
// per token row, warp-shuffle reductions, fp32 statistics.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

#include <algorithm>

namespace pk {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------ patchify
// One thread = 8 consecutive K elements (16-byte bf16 store) of one patch row.
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                int S, int p, int n_side, long long total_chunks, int rows_per_sample, int row_offset) {
  const int chunks_per_prow = p / 8;           // 8-wide chunks per (c,i) row of a patch
  const int Kp = 3 * p * p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx * 8;                     // linear output element
    const long long row = e / Kp;              // b*P + patch
    const int k = static_cast<int>(e - row * Kp);
    const int c = k / (p * p);
    const int i = (k - c * p * p) / p;
    const int j = k - c * p * p - i * p;
    const int P = n_side * n_side;
    const long long b = row / P;
    const int patch = static_cast<int>(row - b * P);
    const int py = patch / n_side, px = patch - py * n_side;
    const float* src = img + ((b * 3 + c) * S + (py * p + i)) * (long long)S + px * p + j;
    const float4 a = ldg4(src), d = ldg4(src + 4);
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(d.x, d.y); o.w = pack_bf16(d.z, d.w);
    *reinterpret_cast<uint4*>(out + ((b * rows_per_sample + row_offset + patch) * (long long)Kp + k)) = o;
    (void)chunks_per_prow;
  }
}

// uint8 HWC images (what the decoder produces, before T.ToTensor/T.Normalize: reference data/imagenette.py:69-73):
// ToTensor (u/255), Normalize ((t - mean_c)/std_c) and the im2col in one pass that reads 1 byte per pixel-channel.
// Same fp32 operations in the same order as the torchvision transforms, so the bf16 patches equal patchify(float path).
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int S, int p, int n_side,
                                   long long total_chunks, float m0, float m1, float m2, float s0, float s1, float s2,
                                   int rows_per_sample, int row_offset) {
  const int Kp = 3 * p * p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_chunks;
       idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx * 8;                     // linear output element
    const long long row = e / Kp;              // b*P + patch
    const int k = static_cast<int>(e - row * Kp);
    const int c = k / (p * p);
    const int i = (k - c * p * p) / p;
    const int j = k - c * p * p - i * p;
    const int P = n_side * n_side;
    const long long b = row / P;
    const int patch = static_cast<int>(row - b * P);
    const int py = patch / n_side, px = patch - py * n_side;
    const uint8_t* src = img + ((b * S + (py * p + i)) * (long long)S + px * p + j) * 3 + c;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = (static_cast<float>(__ldg(src + 3 * t)) / 255.0f - mean) / sd;
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + ((b * rows_per_sample + row_offset + patch) * (long long)Kp + k)) = o;
  }
}

// ------------------------------------------------------------------------------ token rows
__global__ void fill_token_rows_kernel(float* __restrict__ x, int batch, int seq_stride, int row_offset, int n_tokens,
                                       int dim, const float* __restrict__ tokens, const float* __restrict__ pos, float scale) {
  const int d4 = dim / 4;
  const long long total = (long long)batch * n_tokens * d4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % d4);
    const long long r = idx / d4;
    const int t = static_cast<int>(r % n_tokens);
    const long long b = r / n_tokens;
    float4 v = make_float4(scale, scale, scale, scale);
    if (tokens) {
      const float4 tk = ldg4(tokens + (long long)t * dim + c * 4);
      v = make_float4(scale * tk.x, scale * tk.y, scale * tk.z, scale * tk.w);
    }
    if (pos) {
      const float4 pe = ldg4(pos + (long long)(row_offset + t) * dim + c * 4);
      v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    }
    *reinterpret_cast<float4*>(x + (b * seq_stride + row_offset + t) * (long long)dim + c * 4) = v;
  }
}

// ------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row lives in registers (<= MAXV float4 per lane), two-pass statistics.
template <int MAXV>
__device__ __forceinline__ void ln_row_load(const float* __restrict__ row, int d4, int lane, float4 (&v)[MAXV], float& sum) {
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < d4) {
      v[i] = *reinterpret_cast<const float4*>(row + c * 4);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
template <int MAXV>
__device__ __forceinline__ void ln_row_stats(const float4 (&v)[MAXV], int d4, int lane, int dim, float sum, float eps,
                                             float& mean, float& rstd) {
  mean = warp_sum(sum) / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + i * 32 < d4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float var = warp_sum(sq) / static_cast<float>(dim);
  rstd = 1.0f / sqrtf(var + eps);
}

template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, int rows, int dim, const float* __restrict__ rowscale,
                      const int* __restrict__ row_index, const int* __restrict__ rows_dev) {
  const int lane = lane_id();
  const int d4 = dim / 4;
  const int nrows = rows_dev ? min(*rows_dev, rows) : rows;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < nrows; r += warps_total) {
    const long long src = row_index ? row_index[r] : r;
    float4 v[MAXV];
    float sum, mean, rstd;
    ln_row_load<MAXV>(x + src * dim, d4, lane, v, sum);
    ln_row_stats<MAXV>(v, d4, lane, dim, sum, eps, mean, rstd);
    const float sc = rowscale ? rowscale[r] : 1.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) {
        const float4 g = ldg4(gamma + c * 4), b = ldg4(beta + c * 4);
        const float o0 = sc * fmaf((v[i].x - mean) * rstd, g.x, b.x);
        const float o1 = sc * fmaf((v[i].y - mean) * rstd, g.y, b.y);
        const float o2 = sc * fmaf((v[i].z - mean) * rstd, g.z, b.z);
        const float o3 = sc * fmaf((v[i].w - mean) * rstd, g.w, b.w);
        *reinterpret_cast<uint2*>(y + (long long)r * dim + c * 4) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
      }
    }
  }
}

// ------------------------------------------------------------------------------ class head
// CTA = kHeadGroup samples: LN of their class rows into smem (fp32), then every warp sweeps
// classes, streaming head_w rows (L2-resident) and dotting against the staged samples.
constexpr int kHeadGroup = 8;
template <int MAXV>
__global__ void __launch_bounds__(256)
cls_head_kernel(const float* __restrict__ x, int batch, int seq_len, const int* __restrict__ cu_seqlens, int n_cls, int dim,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                const float* __restrict__ head_w, const float* __restrict__ head_b, int num_classes, float* __restrict__ logits) {
  extern __shared__ float feat[];   // [kHeadGroup][dim]
  const int lane = lane_id(), warp = warp_id();
  const int d4 = dim / 4;
  const int b0 = blockIdx.x * kHeadGroup;
  {
    const int b = b0 + warp;        // 8 warps <-> 8 samples
    float4 acc[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < batch) {
      const long long row0 = cu_seqlens ? cu_seqlens[b] : (long long)b * seq_len;
      for (int t = 0; t < n_cls; ++t) {
        float4 v[MAXV];
        float sum, mean, rstd;
        ln_row_load<MAXV>(x + (row0 + t) * dim, d4, lane, v, sum);
        ln_row_stats<MAXV>(v, d4, lane, dim, sum, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = lane + i * 32;
          if (c < d4) {
            const float4 g = ldg4(gamma + c * 4), be = ldg4(beta + c * 4);
            acc[i].x += fmaf((v[i].x - mean) * rstd, g.x, be.x);
            acc[i].y += fmaf((v[i].y - mean) * rstd, g.y, be.y);
            acc[i].z += fmaf((v[i].z - mean) * rstd, g.z, be.z);
            acc[i].w += fmaf((v[i].w - mean) * rstd, g.w, be.w);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) *reinterpret_cast<float4*>(feat + warp * dim + c * 4) = acc[i];
    }
  }
  __syncthreads();
  // blockIdx.y splits the classes so small batches still fill the machine
  const int per = (num_classes + gridDim.y - 1) / gridDim.y;
  const int cls_begin = blockIdx.y * per, cls_end = min(num_classes, cls_begin + per);
  for (int cls = cls_begin + warp; cls < cls_end; cls += (blockDim.x >> 5)) {
    float part[kHeadGroup];
#pragma unroll
    for (int s = 0; s < kHeadGroup; ++s) part[s] = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 w = ldg4(head_w + (long long)cls * dim + c * 4);
#pragma unroll
      for (int s = 0; s < kHeadGroup; ++s) {
        const float4 f = *reinterpret_cast<const float4*>(feat + s * dim + c * 4);
        part[s] += (w.x * f.x + w.y * f.y) + (w.z * f.z + w.w * f.w);
      }
    }
    const float bias = head_b ? head_b[cls] : 0.f;
#pragma unroll
    for (int s = 0; s < kHeadGroup; ++s) {
      const float tot = warp_sum(part[s]);
      if (lane == 0 && b0 + s < batch) logits[(long long)(b0 + s) * num_classes + cls] = tot + bias;
    }
  }
}

// LN(x[class rows]) summed over the class tokens -> feat f32 [batch, dim]: the head's input (vit.py:95,242-243).  One warp per
// sample.  Used when the head itself runs as a split-operand tensor-core GEMM (large batches); cls_head_kernel above fuses the
// head in for small ones.
template <int MAXV>
__global__ void __launch_bounds__(256)
cls_features_kernel(const float* __restrict__ x, int batch, int seq_len, const int* __restrict__ cu_seqlens, int n_cls, int dim,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ feat) {
  const int lane = lane_id();
  const int d4 = dim / 4;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (b >= batch) return;
  float4 acc[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long row0 = cu_seqlens ? cu_seqlens[b] : (long long)b * seq_len;
  for (int t = 0; t < n_cls; ++t) {
    float4 v[MAXV];
    float sum, mean, rstd;
    ln_row_load<MAXV>(x + (row0 + t) * dim, d4, lane, v, sum);
    ln_row_stats<MAXV>(v, d4, lane, dim, sum, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) {
        const float4 g = ldg4(gamma + c * 4), be = ldg4(beta + c * 4);
        acc[i].x += fmaf((v[i].x - mean) * rstd, g.x, be.x);
        acc[i].y += fmaf((v[i].y - mean) * rstd, g.y, be.y);
        acc[i].z += fmaf((v[i].z - mean) * rstd, g.z, be.z);
        acc[i].w += fmaf((v[i].w - mean) * rstd, g.w, be.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < d4) *reinterpret_cast<float4*>(feat + (long long)b * dim + c * 4) = acc[i];
  }
}

// ------------------------------------------------------------------------------ RankViT
// Two rows per warp iteration with every load issued before the first reduction, and 32-bit row arithmetic.  Measured: no
// change against the one-row version (65.8 us for 512 ViT-B images = 0.73 of the measured copy bandwidth, run24) -- a pure
// read stream, not latency-bound.
__global__ void __launch_bounds__(256)
token_norm_score_kernel(const float* __restrict__ x, float* __restrict__ scores, int batch, int seq_len, int dim) {
  constexpr int kRows = 2, kMaxV = 8;       // up to 8 float4 per lane per row (dim <= 1024) on the fast path
  const int lane = lane_id();
  const unsigned n = static_cast<unsigned>(seq_len - 1);
  const unsigned total = static_cast<unsigned>(batch) * n;
  const int d4 = dim / 4;
  const unsigned warps_total = gridDim.x * (blockDim.x >> 5);
  for (unsigned r0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * kRows; r0 < total; r0 += warps_total * kRows) {
    const float* row[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const unsigned r = min(r0 + u, total - 1);
      const unsigned b = r / n, i = r - b * n;
      row[u] = x + (static_cast<long long>(b) * seq_len + 1 + i) * dim;
    }
    float sq[kRows];
    if (d4 <= 32 * kMaxV) {
      float4 v[kRows][kMaxV];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
#pragma unroll
        for (int j = 0; j < kMaxV; ++j) {
          const int c = lane + 32 * j;
          v[u][j] = c < d4 ? *reinterpret_cast<const float4*>(row[u] + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        sq[u] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxV; ++j)
          if (lane + 32 * j < d4) sq[u] += (v[u][j].x * v[u][j].x + v[u][j].y * v[u][j].y) + (v[u][j].z * v[u][j].z + v[u][j].w * v[u][j].w);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        sq[u] = 0.f;
        for (int c = lane; c < d4; c += 32) {
          const float4 v = *reinterpret_cast<const float4*>(row[u] + c * 4);
          sq[u] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const float t = warp_sum(sq[u]);
      if (lane == 0 && r0 + u < total) scores[r0 + u] = sqrtf(t);
    }
  }
}

// Exact stable descending rank by counting: rank_i = #{j : s_j > s_i or (s_j == s_i and j < i)}.
// Every token gets a distinct rank, so the first k ranks are the reference's argsort[:k] with
// ties broken to the lowest index; O(n^2) compares per row out of shared memory (n <= 4096).
// The compare runs on an integer key that is a TOTAL order: floats in their numeric order (-0 == +0), every NaN equal to
// every other and above +inf (torch's descending argsort puts NaN first) — with a float compare all NaN scores would get
// rank 0 and leave slots of `kept` unwritten.
__device__ __forceinline__ unsigned int score_key(float s) {
  unsigned int u = __float_as_uint(s);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;      // NaN
  if (u == 0x80000000u) u = 0u;                                   // -0 == +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
topk_select_kernel(const float* __restrict__ scores, int* __restrict__ kept, int n, int k) {
  extern __shared__ unsigned int s_key[];
  const float* row = scores + (long long)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_key[i] = score_key(row[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned int si = s_key[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned int sj = s_key[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    if (rank < k) kept[(long long)blockIdx.x * k + rank] = i;
  }
}

// Two output rows per warp iteration: both kept-token indices are resolved first, then all row loads are issued before the
// first store (the index -> row dependency otherwise leaves one row in flight per warp: 0.68 - 0.77 of the HBM peak).
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, float* __restrict__ y, const int* __restrict__ kept, int batch, int seq_len,
                   int k, int dim) {
  constexpr int kRows = 2, kMaxV = 8;
  const int lane = lane_id();
  const int d4 = dim / 4;
  const unsigned k1 = static_cast<unsigned>(k + 1);
  const unsigned total = static_cast<unsigned>(batch) * k1;
  const unsigned warps_total = gridDim.x * (blockDim.x >> 5);
  for (unsigned r0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * kRows; r0 < total; r0 += warps_total * kRows) {
    const float4* src[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const unsigned r = min(r0 + u, total - 1);
      const unsigned b = r / k1, o = r - b * k1;
      const int src_tok = o == 0 ? 0 : 1 + kept[static_cast<long long>(b) * k + (o - 1)];
      src[u] = reinterpret_cast<const float4*>(x + (static_cast<long long>(b) * seq_len + src_tok) * dim);
    }
    if (d4 <= 32 * kMaxV) {
      float4 v[kRows][kMaxV];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
#pragma unroll
        for (int j = 0; j < kMaxV; ++j)
          if (lane + 32 * j < d4) v[u][j] = src[u][lane + 32 * j];
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (r0 + u >= total) break;
        float4* dst = reinterpret_cast<float4*>(y + static_cast<long long>(r0 + u) * dim);
#pragma unroll
        for (int j = 0; j < kMaxV; ++j)
          if (lane + 32 * j < d4) dst[lane + 32 * j] = v[u][j];
      }
    } else {
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (r0 + u >= total) break;
        float4* dst = reinterpret_cast<float4*>(y + static_cast<long long>(r0 + u) * dim);
        for (int c = lane; c < d4; c += 32) dst[c] = src[u][c];
      }
    }
  }
}

static int grid_for(long long work_items, int per_block, int max_blocks_per_sm = 8) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace pk

using namespace pk;

extern "C" int pk_patchify(const float* images, void* patches, int batch, int image_size, int patch_size, int rows_per_sample,
                           int row_offset, void* stream) {
  PK_REQUIRE(images && patches, "pk_patchify: null pointer");
  PK_REQUIRE(patch_size % 8 == 0 && image_size % patch_size == 0, "pk_patchify: patch_size %d must be a multiple of 8 dividing image_size %d",
             patch_size, image_size);
  if (batch == 0) return PK_OK;
  const int n_side = image_size / patch_size;
  const long long total = (long long)batch * n_side * n_side * 3 * patch_size * patch_size / 8;
  if (rows_per_sample <= 0) { rows_per_sample = n_side * n_side; row_offset = 0; }
  PK_REQUIRE(row_offset >= 0 && row_offset + n_side * n_side <= rows_per_sample, "pk_patchify: patches do not fit rows_per_sample");
  patchify_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, static_cast<__nv_bfloat16*>(patches), image_size, patch_size, n_side, total, rows_per_sample, row_offset);
  return check_cuda(cudaGetLastError(), "patchify_kernel");
}

extern "C" int pk_patchify_u8(const unsigned char* images_hwc, void* patches, int batch, int image_size, int patch_size,
                              const float* mean3, const float* std3, int rows_per_sample, int row_offset, void* stream) {
  PK_REQUIRE(images_hwc && patches && mean3 && std3, "pk_patchify_u8: null pointer");
  PK_REQUIRE(patch_size % 8 == 0 && image_size % patch_size == 0, "pk_patchify_u8: patch_size must be a multiple of 8 dividing image_size");
  if (batch == 0) return PK_OK;
  const int n_side = image_size / patch_size;
  const long long total = (long long)batch * n_side * n_side * 3 * patch_size * patch_size / 8;
  if (rows_per_sample <= 0) { rows_per_sample = n_side * n_side; row_offset = 0; }
  PK_REQUIRE(row_offset >= 0 && row_offset + n_side * n_side <= rows_per_sample, "pk_patchify_u8: patches do not fit rows_per_sample");
  patchify_u8_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images_hwc, static_cast<__nv_bfloat16*>(patches), image_size, patch_size, n_side, total, mean3[0], mean3[1], mean3[2], std3[0],
      std3[1], std3[2], rows_per_sample, row_offset);
  return check_cuda(cudaGetLastError(), "patchify_u8_kernel");
}

// ------------------------------------------------------------------------------ NoiseBlock (blocks.py:100-188)
// Gaussian noise at a signal-to-noise ratio: per token row, power = mean(x^2), x += noise * sqrt(power / 10^(snr_db/10))
// (forward_snr, blocks.py:117-131).  One warp per row; the row is read twice (second pass from L1/L2).
__global__ void noise_snr_kernel(float* __restrict__ x, const float* __restrict__ noise, int rows, int dim, float inv_snr_lin) {
  const int lane = threadIdx.x & 31;
  const int d4 = dim / 4;
  for (long long r = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); r < rows; r += static_cast<long long>(gridDim.x) * (blockDim.x / 32)) {
    float4* xr = reinterpret_cast<float4*>(x + r * dim);
    const float4* nr = reinterpret_cast<const float4*>(noise + r * dim);
    float ss = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 v = xr[c];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float sd = sqrtf(ss / static_cast<float>(dim) * inv_snr_lin);
    for (int c = lane; c < d4; c += 32) {
      float4 v = xr[c];
      const float4 n = ldg4(reinterpret_cast<const float*>(nr + c));
      v.x = fmaf(n.x, sd, v.x); v.y = fmaf(n.y, sd, v.y); v.z = fmaf(n.z, sd, v.z); v.w = fmaf(n.w, sd, v.w);
      xr[c] = v;
    }
  }
}

extern "C" int pk_noise_snr(float* x, const float* noise, int rows, int dim, float snr_db, void* stream) {
  PK_REQUIRE(x && noise && rows >= 0 && dim % 4 == 0 && dim >= 4, "pk_noise_snr: bad arguments");
  if (rows == 0) return PK_OK;
  const float inv_snr_lin = 1.0f / powf(10.0f, snr_db / 10.0f);
  noise_snr_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, rows, dim, inv_snr_lin);
  return check_cuda(cudaGetLastError(), "noise_snr_kernel");
}

// Token drop: rows (b, tokens[j]) of every sample are zeroed (forward_token_drop, blocks.py:141-157).
__global__ void zero_token_rows_kernel(float* __restrict__ x, int batch, int seq, const int* __restrict__ tokens, int n_tokens, int dim) {
  const int d4 = dim / 4;
  const long long total = static_cast<long long>(batch) * n_tokens * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const long long r = i / d4;
    const int j = static_cast<int>(r % n_tokens);
    const long long b = r / n_tokens;
    reinterpret_cast<float4*>(x + (b * seq + tokens[j]) * static_cast<long long>(dim))[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

extern "C" int pk_zero_token_rows(float* x, int batch, int seq, const int* tokens, int n_tokens, int dim, void* stream) {
  PK_REQUIRE(x && (tokens || n_tokens == 0) && batch >= 0 && seq > 0 && n_tokens >= 0 && n_tokens <= seq && dim % 4 == 0,
             "pk_zero_token_rows: bad arguments");
  if (batch == 0 || n_tokens == 0) return PK_OK;
  const long long total = static_cast<long long>(batch) * n_tokens * (dim / 4);
  zero_token_rows_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, batch, seq, tokens, n_tokens, dim);
  return check_cuda(cudaGetLastError(), "zero_token_rows_kernel");
}

// Masked row arithmetic of the dense-layout ResidualViT modes (residualvit.py:130-194, :239-242):
//   out[r,:] = (accumulate ? out[r,:] : 0) + w(r) * a[r,:],   w(r) = scale[r] or 1 - scale[r]
// i.e. ``mask * img_tokens`` and the add_input term ``img_tokens * (1 - mask)``; out may alias a when !accumulate.
__global__ void row_scale_add_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ scale, int rows,
                                     int dim, int one_minus, int accumulate) {
  const int d4 = dim / 4;
  const long long total = static_cast<long long>(rows) * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d4;
    const float s = scale[r];
    const float w = one_minus ? 1.0f - s : s;
    const float4 v = reinterpret_cast<const float4*>(a)[i];
    float4 o = make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
    if (accumulate) {
      const float4 p = reinterpret_cast<const float4*>(out)[i];
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

extern "C" int pk_row_scale_add(float* out, const float* a, const float* scale, int rows, int dim, int one_minus, int accumulate,
                                void* stream) {
  PK_REQUIRE(out && a && scale && rows >= 0 && dim > 0 && dim % 4 == 0, "pk_row_scale_add: bad arguments");
  if (rows == 0) return PK_OK;
  const long long total = static_cast<long long>(rows) * (dim / 4);
  row_scale_add_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, a, scale, rows, dim, one_minus, accumulate);
  return check_cuda(cudaGetLastError(), "row_scale_add_kernel");
}

extern "C" int pk_fill_token_rows(float* x, int batch, int seq_stride, int row_offset, int n_tokens, int dim,
                                  const float* tokens, const float* pos, float scale, void* stream) {
  PK_REQUIRE(x && dim % 4 == 0, "pk_fill_token_rows: null x or dim %% 4 != 0");
  if (batch == 0 || n_tokens == 0) return PK_OK;
  const long long total = (long long)batch * n_tokens * (dim / 4);
  fill_token_rows_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, batch, seq_stride, row_offset, n_tokens, dim, tokens, pos, scale);
  return check_cuda(cudaGetLastError(), "fill_token_rows_kernel");
}

extern "C" int pk_layernorm_bf16(const float* x, void* y, const float* gamma, const float* beta, float eps, int rows, int dim,
                                 const float* rowscale, const int* row_index, const int* rows_dev, void* stream) {
  PK_REQUIRE(x && y && gamma && beta, "pk_layernorm_bf16: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024, "pk_layernorm_bf16: dim %d must be a multiple of 4 in [4,1024]", dim);
  if (rows == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(rows, 8);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 2) layernorm_bf16_kernel<2><<<grid, 256, 0, s>>>(x, yb, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
  else if (maxv <= 3) layernorm_bf16_kernel<3><<<grid, 256, 0, s>>>(x, yb, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
  else if (maxv <= 6) layernorm_bf16_kernel<6><<<grid, 256, 0, s>>>(x, yb, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
  else layernorm_bf16_kernel<8><<<grid, 256, 0, s>>>(x, yb, gamma, beta, eps, rows, dim, rowscale, row_index, rows_dev);
  return check_cuda(cudaGetLastError(), "layernorm_bf16_kernel");
}

extern "C" int pk_cls_head(const float* x, int batch, int seq_len, const int* cu_seqlens, int n_cls, int dim, const float* gamma,
                           const float* beta, float eps, const float* head_w, const float* head_b, int num_classes, float* logits,
                           void* stream) {
  PK_REQUIRE(x && gamma && beta && head_w && logits, "pk_cls_head: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim <= 1024 && n_cls >= 1, "pk_cls_head: dim %d must be a multiple of 4, <= 1024", dim);
  if (batch == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = (batch + kHeadGroup - 1) / kHeadGroup;
  // classes are split over enough CTAs to fill the machine (PK_HEAD_CTAS_PER_SM, default 8: 4 - 16 measured the same,
  // 110 - 117 us per 512 samples; the kernel is bound by the shared-memory reads of the staged features, which is why large
  // batches take the tensor-core head instead, see engine.Forward.head)
  static int per_sm = -1;
  if (per_sm < 0) { const char* e = getenv("PK_HEAD_CTAS_PER_SM"); per_sm = e ? atoi(e) : 8; if (per_sm < 1) per_sm = 1; }
  int chunks = (per_sm * num_sms() + groups - 1) / groups;
  chunks = std::max(1, std::min(chunks, (num_classes + 7) / 8));      // at least one class per warp
  const dim3 grid(groups, chunks);
  const size_t smem = (size_t)kHeadGroup * dim * sizeof(float);
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 3)
    cls_head_kernel<3><<<grid, 256, smem, s>>>(x, batch, seq_len, cu_seqlens, n_cls, dim, gamma, beta, eps, head_w, head_b, num_classes, logits);
  else
    cls_head_kernel<8><<<grid, 256, smem, s>>>(x, batch, seq_len, cu_seqlens, n_cls, dim, gamma, beta, eps, head_w, head_b, num_classes, logits);
  return check_cuda(cudaGetLastError(), "cls_head_kernel");
}

extern "C" int pk_cls_features(const float* x, int batch, int seq_len, const int* cu_seqlens, int n_cls, int dim, const float* gamma,
                               const float* beta, float eps, float* feat, void* stream) {
  using namespace pk;
  PK_REQUIRE(x && gamma && beta && feat, "pk_cls_features: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim <= 1024 && n_cls >= 1, "pk_cls_features: dim %d must be a multiple of 4, <= 1024", dim);
  if (batch == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (batch + 7) / 8;
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 3) cls_features_kernel<3><<<grid, 256, 0, s>>>(x, batch, seq_len, cu_seqlens, n_cls, dim, gamma, beta, eps, feat);
  else cls_features_kernel<8><<<grid, 256, 0, s>>>(x, batch, seq_len, cu_seqlens, n_cls, dim, gamma, beta, eps, feat);
  return check_cuda(cudaGetLastError(), "cls_features_kernel");
}

// ------------------------------------------------------------------ LayerNorm statistics + raw bf16 copy (fused-LN GEMM chain)
__global__ void __launch_bounds__(256)
row_stats_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float* __restrict__ stats, int rows, int dim, int parts) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * dim);
    uint2* dst = reinterpret_cast<uint2*>(xb + static_cast<long long>(r) * dim);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 v = src[c];
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      dst[c] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    float2* st = reinterpret_cast<float2*>(stats) + static_cast<long long>(r) * parts;
    if (lane < parts) st[lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
  }
}

extern "C" int pk_row_stats_cast(const float* x, void* xb, float* row_stats, int rows, int dim, int parts, void* stream) {
  PK_REQUIRE(x && xb && row_stats && dim % 4 == 0 && parts >= 1 && parts <= 32 && rows >= 0, "pk_row_stats_cast: bad arguments");
  if (rows == 0) return PK_OK;
  row_stats_cast_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(xb), row_stats, rows, dim, parts);
  return check_cuda(cudaGetLastError(), "row_stats_cast_kernel");
}

// ------------------------------------------------------------------ eval-loop accuracy (validate/test.py:116-129)
// One warp per sample: first-maximum arg-max over the classes (torch.argmax semantics; NaN logits never win),
// compared with the label; counts[0] += #correct, counts[1] += #samples (64-bit atomics).
__global__ void __launch_bounds__(256)
argmax_count_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int batch, int num_classes,
                    int* __restrict__ pred_out, unsigned long long* __restrict__ counts) {
  const int lane = lane_id();
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  unsigned long long correct = 0, seen = 0;
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp_id(); b < batch; b += warps_total) {
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int c = lane; c < num_classes; c += 32) {
      const float v = logits[static_cast<long long>(b) * num_classes + c];
      if (v > best) { best = v; best_i = c; }            // strict: the lowest index of a tie stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (best_i == 0x7fffffff) best_i = 0;
    if (lane == 0) {
      if (pred_out) pred_out[b] = best_i;
      if (labels) correct += (labels[b] == best_i) ? 1ull : 0ull;
      seen += 1ull;
    }
  }
  if (lane == 0 && counts && seen) {
    if (correct) atomicAdd(&counts[0], correct);
    atomicAdd(&counts[1], seen);
  }
}

extern "C" int pk_argmax_count(const float* logits, const long long* labels, int batch, int num_classes, int* pred_out,
                               long long* counts, void* stream) {
  PK_REQUIRE(logits && num_classes > 0 && batch >= 0, "pk_argmax_count: bad arguments");
  PK_REQUIRE(pred_out || counts, "pk_argmax_count: need pred_out and/or counts");
  if (batch == 0) return PK_OK;
  argmax_count_kernel<<<grid_for(batch, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, batch, num_classes, pred_out, reinterpret_cast<unsigned long long*>(counts));
  return check_cuda(cudaGetLastError(), "argmax_count_kernel");
}

extern "C" int pk_token_norm_score(const float* x, float* scores, int batch, int seq_len, int dim, void* stream) {
  PK_REQUIRE(x && scores && dim % 4 == 0 && seq_len >= 1, "pk_token_norm_score: bad arguments");
  const long long total = (long long)batch * (seq_len - 1);
  if (total == 0) return PK_OK;
  token_norm_score_kernel<<<grid_for(total, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, scores, batch, seq_len, dim);
  return check_cuda(cudaGetLastError(), "token_norm_score_kernel");
}

extern "C" int pk_topk_select(const float* scores, int* kept, int batch, int n, int k, void* stream) {
  PK_REQUIRE(scores && kept, "pk_topk_select: null pointer");
  PK_REQUIRE(n >= 0 && n <= 4096 && k >= 0 && k <= n, "pk_topk_select: need 0 <= k <= n <= 4096 (n=%d k=%d)", n, k);
  if (batch == 0 || k == 0) return PK_OK;
  topk_select_kernel<<<batch, 256, (size_t)n * sizeof(float), static_cast<cudaStream_t>(stream)>>>(scores, kept, n, k);
  return check_cuda(cudaGetLastError(), "topk_select_kernel");
}

extern "C" int pk_gather_rows(const float* x, float* y, const int* kept, int batch, int seq_len, int k, int dim, void* stream) {
  PK_REQUIRE(x && y && (kept || k == 0) && dim % 4 == 0, "pk_gather_rows: bad arguments");
  const long long total = (long long)batch * (k + 1);
  if (total == 0) return PK_OK;
  gather_rows_kernel<<<grid_for(total, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, kept, batch, seq_len, k, dim);
  return check_cuda(cudaGetLastError(), "gather_rows_kernel");
}
