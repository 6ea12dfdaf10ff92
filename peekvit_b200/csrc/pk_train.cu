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
                     float* __restrict__ dx, int rows, int dim, const int* __restrict__ row_index, int dy_div, int accumulate) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const long long xr = row_index ? row_index[r] : r;
    const float4* x4 = reinterpret_cast<const float4*>(x + xr * dim);
    const float4* g4 = reinterpret_cast<const float4*>(dy + static_cast<long long>(r / dy_div) * dim);
    float4 xv[MAXV], gv[MAXV];
    float s1 = 0.f;
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
    const float mg = warp_sum(sg) / static_cast<float>(dim), mgx = warp_sum(sgx) / static_cast<float>(dim);
    float4* o4 = reinterpret_cast<float4*>(dx + xr * dim);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int c = lane + 32 * j;
      if (c < d4) {
        float4 o = make_float4(rstd * (gv[j].x - mg - xv[j].x * mgx), rstd * (gv[j].y - mg - xv[j].y * mgx),
                               rstd * (gv[j].z - mg - xv[j].z * mgx), rstd * (gv[j].w - mg - xv[j].w * mgx));
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
  const int label = static_cast<int>(labels[b]);
  const float inv = 1.0f / sum;
  for (int c = lane; c < classes; c += 32)
    dlogits[static_cast<long long>(b) * classes + c] = (__expf(z[c] - mx) * inv - (c == label ? 1.0f : 0.0f)) * inv_count;
  if (lane == 0) {
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

}  // namespace pk

using namespace pk;

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

extern "C" int pk_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, float* dx, int rows, int dim,
                                const int* row_index, int dy_div, int accumulate, void* stream) {
  PK_REQUIRE(x && dy && gamma && dx, "pk_layernorm_bwd: null pointer");
  PK_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024 && rows >= 0 && dy_div >= 1, "pk_layernorm_bwd: dim %d must be a multiple of 4 in [4,1024]", dim);
  if (rows == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = train_grid(rows, 8);
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 2) layernorm_bwd_kernel<2><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate);
  else if (maxv <= 3) layernorm_bwd_kernel<3><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate);
  else if (maxv <= 6) layernorm_bwd_kernel<6><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate);
  else layernorm_bwd_kernel<8><<<grid, 256, 0, s>>>(x, dy, gamma, eps, dx, rows, dim, row_index, dy_div, accumulate);
  return check_cuda(cudaGetLastError(), "layernorm_bwd_kernel");
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
  if (head_dim == 64) return launch_attention_bwd<64>(qkv, out, dout, dqkv, batch, num_heads, seq_len, scale, s);
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
