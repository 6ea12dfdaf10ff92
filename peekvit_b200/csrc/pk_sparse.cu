// peekvit_b200 — token-sparsification kernels: ResidualViT budget gating, AViT halting, MoE
// routing, and the ragged-batch compaction they share.  All HBM/latency-bound: one warp per
// token row with 128-bit loads and shuffle reductions, one CTA per sample for the per-sample
// plans (prefix sums by warp ballot), no host synchronisation anywhere — row counts stay on
// the device and the GEMM / LayerNorm / attention kernels read them (m_dev / rows_dev /
// cu_seqlens).
//
// Exact reformulations of the reference's dense masked forward:
//   ResidualViT (models/residualvit.py:197-260): a dropped token is a zero row that still acts as a
//     key (k = b_k, v = b_v) and leaves the block as the constant mlp(0).  Rows carry a
//     multiplicity; dropped rows of a sample are folded into one virtual key (attention kernel)
//     and one "ghost" row mlp(0) appended after the block, which may be re-admitted later.
//   AViT (models/adavit.py:140-219): halted tokens are zero rows -> virtual key; only the class
//     row's ACT-weighted output reaches the head, so a sample retires when its class token halts.
//   MoE (models/moevit.py:49-61): one-hot arg-max routing -> only the chosen expert is computed.
#include "pk_common.cuh"
#include "../../include/peekvit_b200.h"

namespace pk {

__device__ __forceinline__ float row_dot(const float* __restrict__ row, const float* __restrict__ w, int d4, int lane) {
  float acc = 0.f;
  for (int c = lane; c < d4; c += 32) {
    const float4 a = *reinterpret_cast<const float4*>(row + c * 4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(w + c * 4));
    acc += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
  return warp_sum(acc);
}
__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// Warp-0 exclusive prefix over `len` keep flags held in shared memory; returns kept count.
__device__ __forceinline__ int plan_prefix(const unsigned char* s_keep, int* __restrict__ dst_local, long long start, int len,
                                           int lane) {
  int base = 0;
  for (int j0 = 0; j0 < len; j0 += 32) {
    const int j = j0 + lane;
    const bool k = j < len && s_keep[j];
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (j < len) dst_local[start + j] = k ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
    base += __popc(bal);
  }
  return base;
}

// ------------------------------------------------------------------ ResidualViT: threshold (fixed budget)
// thr = 1 - mean(budget token over the whole batch and D)   (residualvit.py:208 + :62)
__global__ void __launch_bounds__(1024)
budget_mean_threshold_kernel(const float* __restrict__ x, const int* __restrict__ cu, int batch, int budget_pos, int dim,
                             float* __restrict__ thr_out) {
  __shared__ float s_part[32];
  float acc = 0.f;
  const int total = batch * dim;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = i / dim, d = i - b * dim;
    acc += x[(static_cast<long long>(cu[b]) + budget_pos) * dim + d];
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) s_part[warp_id()] = acc;
  __syncthreads();
  if (warp_id() == 0) {
    float v = lane_id() < (blockDim.x >> 5) ? s_part[lane_id()] : 0.f;
    v = warp_sum(v);
    if (lane_id() == 0) thr_out[0] = 1.0f - v / static_cast<float>(total);
  }
}

// ------------------------------------------------------------------ ResidualViT: gate + per-sample plan
struct ResidualGateParams {
  const float* x;            // [rows, D] block input, packed
  const int* cu_in;          // [B+1]
  const float* mult_in;      // [rows]
  int dim, n_special, budget_pos;   // specials occupy local rows [0, n_special); budget token at budget_pos (or -1)
  const float* gate_w; float gate_b, inv_temp, gate_bias; int gate_type;     // 0 sigmoid, 1 gumbel(eval)
  int thr_mode;              // 0: sigmoid(bt_w . budget_row + bt_b); 1: *thr_dev; 2: thr_const
  const float* bt_w; float bt_b; const float* thr_dev; float thr_const;
  int gated;                 // 0 -> plain layer: every live row kept with mask 1, no ghost
  float* mask;               // [rows] soft mask per input row (specials 1)
  int* dst_local;            // [rows] index inside the compacted sample, -1 = dropped
  int* sample_of;            // [rows]
  int* new_len;              // [B]
  float* mdrop;              // [B] multiplicity folded into the virtual key / ghost row
};

// Pass 1 (gated layers), one warp per packed row across the whole grid: the gate's GEMV and sigmoid
//   g[r] = sigmoid((w.x_r + b)/temp + gate_bias)   (sigmoid gate)    or    round(sigmoid(w.x_r + b))   (gumbel gate in eval)
// left in mask[r].  Streaming the rows grid-wide instead of one CTA per sample is what makes this HBM-bound
// (the per-sample version reached 0.37 of the measured HBM peak).
__global__ void __launch_bounds__(256)
residual_gate_rows_kernel(const ResidualGateParams p, int batch) {
  const int lane = lane_id(), d4 = p.dim / 4;
  const int rows = p.cu_in[batch];
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(p.gate_w);
  // four rows per warp iteration, all loads issued before the first reduction (memory-level parallelism)
  for (int r0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * 4; r0 < rows; r0 += warps_total * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < d4; c += 32) {
      const float4 wv = __ldg(w4 + c);
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        xv[u] = r0 + u < rows ? *reinterpret_cast<const float4*>(p.x + static_cast<long long>(r0 + u) * p.dim + c * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += (xv[u].x * wv.x + xv[u].y * wv.y) + (xv[u].z * wv.z + xv[u].w * wv.w);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u >= rows) break;
      const float logit = warp_sum(acc[u]) + p.gate_b;
      float g;
      if (p.gate_type == 0) g = sigmoidf_exact(logit * p.inv_temp + p.gate_bias);
      else g = rintf(sigmoidf_exact(logit));
      if (lane == 0) p.mask[r0 + u] = g;
    }
  }
}

// Register-resident variant for dim == 128 * MAXV (ViT-S 384, ViT-B 768): ROWS rows per warp iteration with *every* row load issued before the first
// FMA.  The generic loop above keeps only ROWS float4 per lane in flight per trip and measured 4.1 TB/s on ViT-S rows (ncu,
// profiles/r01/run27_ncu_membound.csv); same accumulation order, so the gate values are bit-identical.
template <int MAXV, int ROWS>
__global__ void __launch_bounds__(256)
residual_gate_rows_reg_kernel(const ResidualGateParams p, int batch) {
  const int lane = lane_id(), d4 = p.dim / 4;
  const int rows = p.cu_in[batch];
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(p.gate_w);
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(p.x);
  float4 wv[MAXV];                       // dim == 128 * MAXV exactly (host dispatch): no column predicates
#pragma unroll
  for (int j = 0; j < MAXV; ++j) wv[j] = __ldg(w4 + lane + 32 * j);
  for (int r0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * ROWS; r0 < rows; r0 += warps_total * ROWS) {
    float4 xv[ROWS][MAXV];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const float4* row = x4 + static_cast<long long>(min(r0 + u, rows - 1)) * d4 + lane;     // clamped: unconditional loads
#pragma unroll
      for (int j = 0; j < MAXV; ++j) xv[u][j] = row[32 * j];
    }
    float acc[ROWS];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      acc[u] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXV; ++j)
        acc[u] += (xv[u][j].x * wv[j].x + xv[u][j].y * wv[j].y) + (xv[u][j].z * wv[j].z + xv[u][j].w * wv[j].w);
    }
#pragma unroll
    for (int u = 0; u < ROWS; ++u) acc[u] = warp_sum(acc[u]);
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      if (r0 + u >= rows) break;
      const float logit = acc[u] + p.gate_b;
      float g;
      if (p.gate_type == 0) g = sigmoidf_exact(logit * p.inv_temp + p.gate_bias);
      else g = rintf(sigmoidf_exact(logit));
      if (lane == 0) p.mask[r0 + u] = g;
    }
  }
}

// Pass 2, one CTA per sample: threshold from the budget token, soft mask, keep decisions, positions in the compacted
// sample, new length (+1 ghost slot when gated) and the multiplicity folded into the virtual key / ghost row.
__global__ void __launch_bounds__(256)
residual_gate_plan_kernel(const ResidualGateParams p) {
  extern __shared__ unsigned char s_keep[];
  __shared__ float s_thr, s_drop[8];
  const int b = blockIdx.x, lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const long long start = p.cu_in[b];
  const int len = p.cu_in[b + 1] - p.cu_in[b];
  const int d4 = p.dim / 4;
  if (warp == 0) {
    float thr = p.thr_const;
    if (p.thr_mode == 0) thr = sigmoidf_exact(row_dot(p.x + (start + p.budget_pos) * p.dim, p.bt_w, d4, lane) + p.bt_b);
    else if (p.thr_mode == 1) thr = p.thr_dev[0];
    if (lane == 0) s_thr = thr;
  }
  __syncthreads();
  const float thr = s_thr;
  float dropped = 0.f;
  for (int j = tid; j < len; j += blockDim.x) {
    const long long r = start + j;
    const float mult = p.mult_in[r];
    float m = 1.0f;
    bool keep = true;
    if (j >= p.n_special) {
      if (p.gated) {
        const float g = p.mask[r];                                  // pass 1
        m = p.gate_type == 0 ? fmaxf(g - thr, 0.f) : g;
      }
      keep = (m > 0.f) && (mult > 0.f);
      if (!keep) dropped += mult;
    }
    p.mask[r] = m;
    p.sample_of[r] = b;
    s_keep[j] = keep ? 1 : 0;
  }
  dropped = warp_sum(dropped);
  if (lane == 0) s_drop[warp] = dropped;
  __syncthreads();
  if (warp == 0) {
    const int kept = plan_prefix(s_keep, p.dst_local, start, len, lane);
    if (lane == 0) {
      float md = 0.f;
      for (int w = 0; w < 8; ++w) md += s_drop[w];
      p.new_len[b] = kept + (p.gated ? 1 : 0);       // + ghost slot
      p.mdrop[b] = md;
    }
  }
}

// ------------------------------------------------------------------ exclusive scan of per-sample lengths
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const int* __restrict__ len, int n, int* __restrict__ cu_out, int* __restrict__ total_out) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = lane_id(), warp = warp_id();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? len[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
    if (i < n) cu_out[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cu_out[n] = s_carry;
    if (total_out) total_out[0] = s_carry;
  }
}

// ------------------------------------------------------------------ compaction of packed rows
struct CompactParams {
  const float* x_in; float* x_out; int dim;
  const int* cu_in; const int* cu_out; int batch;
  const int* dst_local; const int* sample_of;
  const float* scale_in;                 // optional per input row (ResidualViT soft mask): out = scale * in
  float* scale_out;                      // optional per output row
  const float* a0_in; float* a0_out;     // optional per-row attributes carried along (multiplicity / c / R / token id)
  const float* a1_in; float* a1_out;
  const float* a2_in; float* a2_out;
  int ghost;                             // ResidualViT: zero-initialise each sample's last output row (a0 = 0, scale = 1)
  int* pub_tok_row; float* pub_mask; int pub_n_img;   // optional fused publish (see residual_publish_kernel)
};

// Four input rows per warp iteration: the dependent index chain (dst_local -> sample_of -> cu_out) of all four is resolved
// first, then every row load is issued before the first store, so one warp keeps 4 rows in flight instead of one
// (the one-row-per-iteration version reached 0.49 of the measured HBM peak: latency-, not bandwidth-bound).
__global__ void __launch_bounds__(256)
compact_rows_kernel(const CompactParams p) {
  constexpr int kRows = 4;
  const int lane = lane_id();
  const int d4 = p.dim / 4;
  const int rows_in = p.cu_in[p.batch];
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp_id();
  for (int it0 = gw * kRows; it0 < rows_in; it0 += warps_total * kRows) {
    int dl[kRows], bs[kRows];
    long long dst[kRows];
    float sc[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const int it = it0 + u;
      dl[u] = it < rows_in ? p.dst_local[it] : -1;
      bs[u] = it < rows_in ? p.sample_of[it] : 0;
      sc[u] = (p.scale_in && it < rows_in) ? p.scale_in[it] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) dst[u] = dl[u] >= 0 ? static_cast<long long>(p.cu_out[bs[u]]) + dl[u] : -1;
    for (int c0 = 0; c0 < d4; c0 += 32) {
      const int c = c0 + lane;
      float4 v[kRows];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
        if (dst[u] >= 0 && c < d4) v[u] = reinterpret_cast<const float4*>(p.x_in + static_cast<long long>(it0 + u) * p.dim)[c];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
        if (dst[u] >= 0 && c < d4) {
          float4 o = v[u];
          o.x *= sc[u]; o.y *= sc[u]; o.z *= sc[u]; o.w *= sc[u];
          reinterpret_cast<float4*>(p.x_out + dst[u] * p.dim)[c] = o;
        }
    }
    if (lane < kRows) {
      // lane u carries row u's attributes (the per-row arrays above are warp-uniform: pick by lane without local memory)
      long long d = -1; float s1 = 1.0f;
#pragma unroll
      for (int u = 0; u < kRows; ++u) if (lane == u) { d = dst[u]; s1 = sc[u]; }
      if (d >= 0) {
        const int it = it0 + lane;
        if (p.scale_out) p.scale_out[d] = s1;
        if (p.a0_out) p.a0_out[d] = p.a0_in[it];
        if (p.a1_out) p.a1_out[d] = p.a1_in[it];
        if (p.a2_out) p.a2_out[d] = p.a2_in[it];
      }
    }
  }
  if (p.pub_mask) {
    // fused residual_publish_kernel: reads only the plan (scale_in = soft mask, dst_local, cu_out), independent of the row copies
    const int total = p.batch * p.pub_n_img;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int b = i / p.pub_n_img;
      const int r = p.pub_tok_row[i];
      p.pub_mask[i] = p.scale_in[r];
      const int dl = p.dst_local[r];
      p.pub_tok_row[i] = dl >= 0 ? p.cu_out[b] + dl : p.cu_out[b + 1] - 1;
    }
  }
  if (p.ghost) {
    for (int b = gw; b < p.batch; b += warps_total) {
      const long long dst = static_cast<long long>(p.cu_out[b + 1]) - 1;
      float4* out = reinterpret_cast<float4*>(p.x_out + dst * p.dim);
      for (int c = lane; c < d4; c += 32) out[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane == 0) {
        if (p.scale_out) p.scale_out[dst] = 1.0f;
        if (p.a0_out) p.a0_out[dst] = 0.0f;     // multiplicity 0: inert as a key during this layer
      }
    }
  }
}

// After the block: the ghost row becomes mlp(0) standing for mdrop[b] dropped tokens (residualvit.py:258-260
// evaluated on a zero row: fc2(gelu(fc1.bias)) + fc2.bias).
__global__ void __launch_bounds__(256)
residual_ghost_kernel(float* __restrict__ x, float* __restrict__ mult, const int* __restrict__ cu, const float* __restrict__ mdrop,
                      const float* __restrict__ mlp0, int batch, int dim) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp_id(); b < batch; b += warps_total) {
    const long long dst = static_cast<long long>(cu[b + 1]) - 1;
    float4* out = reinterpret_cast<float4*>(x + dst * dim);
    for (int c = lane; c < d4; c += 32) out[c] = __ldg(reinterpret_cast<const float4*>(mlp0 + c * 4));
    if (lane == 0) mult[dst] = mdrop[b];
  }
}

// Publish the soft mask per original image token (reference block.mask, (B, N_img, 1)) and move the
// token -> packed-row map to the compacted layout (dropped tokens now point at the ghost row).
__global__ void residual_publish_kernel(const float* __restrict__ mask, const int* __restrict__ dst_local, const int* __restrict__ cu_out,
                                        int* __restrict__ tok_row, float* __restrict__ mask_pub, int batch, int n_img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * n_img) return;
  const int b = i / n_img;
  const int r = tok_row[i];
  mask_pub[i] = mask[r];
  const int dl = dst_local[r];
  tok_row[i] = dl >= 0 ? cu_out[b] + dl : cu_out[b + 1] - 1;
}

// ------------------------------------------------------------------ AViT halting + per-sample plan
struct AvitParams {
  const float* x;          // [rows, D] block output, packed active rows
  const int* cu_in; int dim, seq_total;
  float* c; float* R; const float* tokid;     // per packed row state (token id stored as float)
  float gate_scale, gate_center, eps; int last_layer, early_exit;
  float* out_acc;          // [B, D] ACT-weighted class-token output (adavit.py:197,205,212-215)
  float* rho; float* counter;                 // [B, seq_total] published per original token (may be NULL)
  int* dst_local; int* sample_of; int* new_len; float* n_halted;   // n_halted[b] = seq_total - active after this layer
};

__global__ void __launch_bounds__(256)
avit_halt_plan_kernel(const AvitParams p) {
  extern __shared__ unsigned char s_keep[];
  __shared__ float s_w;
  __shared__ int s_cls_alive;
  const int b = blockIdx.x, lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const long long start = p.cu_in[b];
  const int len = p.cu_in[b + 1] - p.cu_in[b];
  if (tid == 0) { s_w = 0.f; s_cls_alive = 0; }
  __syncthreads();
  for (int j = tid; j < len; j += blockDim.x) {
    const long long r = start + j;
    const float h = p.last_layer ? 1.0f : sigmoidf_exact(p.x[r * p.dim] * p.gate_scale - p.gate_center);   // adavit.py:74,186-187
    const float c_new = p.c[r] + h;                                                                         // :190
    const float Rv = p.R[r];
    const bool reached = c_new > 1.0f - p.eps;                                                              // :195
    const bool not_reached = c_new < 1.0f - p.eps;                                                          // :202
    const float w = reached ? Rv : (not_reached ? h : 0.f);                                                 // :197,:205
    p.c[r] = c_new;
    p.R[r] = not_reached ? Rv - h : Rv;                                                                     // :204
    const int tok = static_cast<int>(p.tokid[r]);
    if (p.rho) {
      p.rho[static_cast<long long>(b) * p.seq_total + tok] += 1.0f + (reached ? Rv : 0.f);                  // :191,:198
      if (not_reached) p.counter[static_cast<long long>(b) * p.seq_total + tok] += 1.0f;                    // :207
    }
    s_keep[j] = not_reached ? 1 : 0;                                                                        // :210
    p.sample_of[r] = b;
    if (tok == 0) { s_w = w; s_cls_alive = not_reached ? 1 : 0; }
  }
  __syncthreads();
  // class row is local row 0 while it is active
  const bool has_cls = len > 0 && static_cast<int>(p.tokid[start]) == 0;
  if (has_cls) {
    const float w = s_w;
    for (int d = tid; d < p.dim; d += blockDim.x) p.out_acc[static_cast<long long>(b) * p.dim + d] += w * p.x[start * p.dim + d];
  }
  if (p.early_exit && !(has_cls && s_cls_alive)) {
    // the class token has halted: nothing later can change this sample's logits
    for (int j = tid; j < len; j += blockDim.x) s_keep[j] = 0;
  }
  __syncthreads();
  if (warp == 0) {
    const int kept = plan_prefix(s_keep, p.dst_local, start, len, lane);
    if (lane == 0) {
      p.new_len[b] = kept;
      p.n_halted[b] = static_cast<float>(p.seq_total - kept);
    }
  }
}

// ------------------------------------------------------------------ MoE routing
// expert[r] = argmax_e ( LN(x[r]) . Wg[e] + bg[e] ), first maximum wins like torch.argmax
// (moevit.py:23-32 + blocks.py:23-25).  LN is recomputed in fp32 so routing does not see bf16 rounding.
// Two rows per warp iteration, both rows' loads issued up front and their reduction chains interleaved: one row at a time,
// the seven dependent warp reductions per row left the loads of the next row unissued (2.4 TB/s on ViT-S rows, ncu
// profiles/r01/run27_ncu_membound.csv).  Per-row arithmetic and its order are unchanged, so routing is bit-identical.
template <int MAXV, int NE>
__global__ void __launch_bounds__(256)
moe_route_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 const float* __restrict__ gate_w, const float* __restrict__ gate_b, int n_experts, int rows, int dim,
                 int* __restrict__ expert) {
  constexpr int R = 2;
  const int lane = lane_id();
  const int d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r0 = (blockIdx.x * (blockDim.x >> 5) + warp_id()) * R; r0 < rows; r0 += warps_total * R) {
    float4 v[R][MAXV];
    float sum[R], mean[R], sq[R], rstd[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int r = min(r0 + u, rows - 1);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + i * 32;
        // the column is clamped instead of predicated (unconditional loads issue back to back); the duplicate is zeroed
        const float4 t = *reinterpret_cast<const float4*>(x + static_cast<long long>(r) * dim + min(c, d4 - 1) * 4);
        v[u][i] = c < d4 ? t : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      sum[u] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) sum[u] += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
    }
#pragma unroll
    for (int u = 0; u < R; ++u) mean[u] = warp_sum(sum[u]) / dim;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      sq[u] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (lane + i * 32 < d4) {
          const float a = v[u][i].x - mean[u], b = v[u][i].y - mean[u], c = v[u][i].z - mean[u], d = v[u][i].w - mean[u];
          sq[u] += (a * a + b * b) + (c * c + d * d);
        }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) rstd[u] = 1.0f / sqrtf(warp_sum(sq[u]) / dim + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < d4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c * 4)), be = __ldg(reinterpret_cast<const float4*>(beta + c * 4));
#pragma unroll
        for (int u = 0; u < R; ++u) {
          v[u][i].x = fmaf((v[u][i].x - mean[u]) * rstd[u], g.x, be.x); v[u][i].y = fmaf((v[u][i].y - mean[u]) * rstd[u], g.y, be.y);
          v[u][i].z = fmaf((v[u][i].z - mean[u]) * rstd[u], g.z, be.z); v[u][i].w = fmaf((v[u][i].w - mean[u]) * rstd[u], g.w, be.w);
        }
      }
    }
    float best[R];
    int best_e[R];
#pragma unroll
    for (int u = 0; u < R; ++u) { best[u] = -INFINITY; best_e[u] = 0; }
    if constexpr (NE > 0) {
      // compile-time expert count: all NE x R dot products first, then their NE x R butterfly reductions interleaved
      float acc[NE][R];
#pragma unroll
      for (int e = 0; e < NE; ++e) {
#pragma unroll
        for (int u = 0; u < R; ++u) acc[e][u] = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = lane + i * 32;
          if (c < d4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(gate_w + static_cast<long long>(e) * dim + c * 4));
#pragma unroll
            for (int u = 0; u < R; ++u) acc[e][u] += (v[u][i].x * w.x + v[u][i].y * w.y) + (v[u][i].z * w.z + v[u][i].w * w.w);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int u = 0; u < R; ++u) acc[e][u] = warp_sum(acc[e][u]);
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const float gb = gate_b[e];
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const float t = acc[e][u] + gb;
          if (t > best[u]) { best[u] = t; best_e[u] = e; }
        }
      }
    } else {
    for (int e = 0; e < n_experts; ++e) {
      float acc[R];
#pragma unroll
      for (int u = 0; u < R; ++u) acc[u] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + i * 32;
        if (c < d4) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(gate_w + static_cast<long long>(e) * dim + c * 4));
#pragma unroll
          for (int u = 0; u < R; ++u) acc[u] += (v[u][i].x * w.x + v[u][i].y * w.y) + (v[u][i].z * w.z + v[u][i].w * w.w);
        }
      }
      const float gb = gate_b[e];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        acc[u] = warp_sum(acc[u]) + gb;
        if (acc[u] > best[u]) { best[u] = acc[u]; best_e[u] = e; }
      }
    }
    }
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (lane == 0 && r0 + u < rows) expert[r0 + u] = best_e[u];
  }
}

// Stable counting sort of rows by expert in three small launches (a single-CTA version took 160 us for 10^5 rows):
//   hist:    block g counts the experts of its contiguous chunk of rows        -> scratch[g][e]
//   scan:    per expert, exclusive scan over the blocks + expert offsets        -> scratch[g][e] = first slot of (g, e)
//   scatter: block g ranks its rows (warp ballots, in row order)               -> src_of[pos] = original row
// offsets[E+1], counts[E]; src_of[pos] = original row of the pos-th expert-sorted row.
constexpr int kMaxExperts = 16;
constexpr int kSortBlocks = 256;      // upper bound on chunks; scratch = kSortBlocks * kMaxExperts ints

__device__ __forceinline__ int moe_chunk_rows(int rows) {
  int chunk = (rows + kSortBlocks - 1) / kSortBlocks;
  chunk = (chunk + 1023) / 1024 * 1024;
  return chunk < 1024 ? 1024 : chunk;
}

__global__ void __launch_bounds__(1024)
moe_hist_kernel(const int* __restrict__ expert, int rows, int n_experts, int* __restrict__ scratch) {
  __shared__ int s_cnt[kMaxExperts];
  const int tid = threadIdx.x;
  if (tid < kMaxExperts) s_cnt[tid] = 0;
  __syncthreads();
  const int chunk = moe_chunk_rows(rows);
  const int begin = blockIdx.x * chunk, end = min(rows, begin + chunk);
  for (int r = begin + tid; r < end; r += blockDim.x) atomicAdd(&s_cnt[expert[r]], 1);
  __syncthreads();
  if (tid < n_experts) scratch[blockIdx.x * kMaxExperts + tid] = s_cnt[tid];
}

__global__ void __launch_bounds__(32)
moe_scan_kernel(int n_blocks, int n_experts, int* __restrict__ scratch, int* __restrict__ offsets, int* __restrict__ counts) {
  __shared__ int s_tot[kMaxExperts];
  const int e = threadIdx.x;
  if (e < n_experts) {
    int acc = 0;
    for (int g = 0; g < n_blocks; ++g) {          // exclusive scan over the blocks of this expert's counts
      const int c = scratch[g * kMaxExperts + e];
      scratch[g * kMaxExperts + e] = acc;
      acc += c;
    }
    s_tot[e] = acc;
    counts[e] = acc;
  }
  __syncwarp();
  if (e == 0) {
    int acc = 0;
    for (int x = 0; x < n_experts; ++x) { offsets[x] = acc; acc += s_tot[x]; }
    offsets[n_experts] = acc;
  }
}

__global__ void __launch_bounds__(1024)
moe_scatter_kernel(const int* __restrict__ expert, int rows, int n_experts, const int* __restrict__ scratch,
                   const int* __restrict__ offsets, int* __restrict__ src_of) {
  __shared__ int s_cursor[kMaxExperts];
  __shared__ int s_wc[32][kMaxExperts];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  if (tid < n_experts) s_cursor[tid] = offsets[tid] + scratch[blockIdx.x * kMaxExperts + tid];
  __syncthreads();
  const int chunk = moe_chunk_rows(rows);
  const int begin = blockIdx.x * chunk, end = min(rows, begin + chunk);
  for (int base = begin; base < end; base += 1024) {
    const int r = base + tid;
    const int e = r < end ? expert[r] : -1;
    int my_rank = 0;
    for (int ex = 0; ex < n_experts; ++ex) {
      const unsigned bal = __ballot_sync(0xffffffffu, e == ex);
      if (e == ex) my_rank = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) s_wc[warp][ex] = __popc(bal);
    }
    __syncthreads();
    if (e >= 0) {
      int pos = s_cursor[e] + my_rank;
      for (int w = 0; w < warp; ++w) pos += s_wc[w][e];
      src_of[pos] = r;
    }
    __syncthreads();
    if (tid < n_experts) {
      int add = 0;
      for (int w = 0; w < 32; ++w) add += s_wc[w][tid];
      s_cursor[tid] += add;
    }
    __syncthreads();
  }
}

static int grid_rows(long long items, int per_block = 8, int max_per_sm = 8) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * max_per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace pk

using namespace pk;

extern "C" int pk_budget_mean_threshold(const float* x, const int* cu_seqlens, int batch, int budget_pos, int dim, float* thr_out,
                                        void* stream) {
  PK_REQUIRE(x && cu_seqlens && thr_out && batch > 0, "pk_budget_mean_threshold: bad arguments");
  budget_mean_threshold_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, cu_seqlens, batch, budget_pos, dim, thr_out);
  return check_cuda(cudaGetLastError(), "budget_mean_threshold_kernel");
}

extern "C" int pk_residual_gate_plan(const pk_residual_gate_args* a, void* stream) {
  PK_REQUIRE(a && a->x && a->cu_in && a->mult_in && a->mask && a->dst_local && a->sample_of && a->new_len && a->mdrop,
             "pk_residual_gate_plan: null pointer");
  PK_REQUIRE(a->dim % 4 == 0 && a->max_seq_len > 0 && a->max_seq_len <= 16384, "pk_residual_gate_plan: bad dim/max_seq_len");
  PK_REQUIRE(!a->gated || a->gate_w, "pk_residual_gate_plan: gated layer needs gate_w");
  PK_REQUIRE(a->thr_mode != 0 || !a->gated || (a->bt_w && a->budget_pos >= 0), "pk_residual_gate_plan: learnable threshold needs bt_w and a budget row");
  PK_REQUIRE(a->thr_mode != 1 || a->thr_dev, "pk_residual_gate_plan: thr_mode 1 needs thr_dev");
  if (a->batch == 0) return PK_OK;
  ResidualGateParams p;
  p.x = a->x; p.cu_in = a->cu_in; p.mult_in = a->mult_in;
  p.dim = a->dim; p.n_special = a->n_special; p.budget_pos = a->budget_pos;
  p.gate_w = a->gate_w; p.gate_b = a->gate_b; p.inv_temp = 1.0f / a->gate_temp; p.gate_bias = a->gate_bias; p.gate_type = a->gate_type;
  p.thr_mode = a->gated ? a->thr_mode : 2;
  p.bt_w = a->bt_w; p.bt_b = a->bt_b; p.thr_dev = a->thr_dev; p.thr_const = a->thr_const;
  p.gated = a->gated;
  p.mask = a->mask; p.dst_local = a->dst_local; p.sample_of = a->sample_of; p.new_len = a->new_len; p.mdrop = a->mdrop;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->gated) {
    const int grid = grid_rows(static_cast<long long>(a->batch) * a->max_seq_len);
    if (a->dim == 384) residual_gate_rows_reg_kernel<3, 4><<<grid, 256, 0, s>>>(p, a->batch);
    else if (a->dim == 768) residual_gate_rows_reg_kernel<6, 2><<<grid, 256, 0, s>>>(p, a->batch);
    else residual_gate_rows_kernel<<<grid, 256, 0, s>>>(p, a->batch);
    PK_CHECK_CUDA(cudaGetLastError());
  }
  residual_gate_plan_kernel<<<a->batch, 256, static_cast<size_t>(a->max_seq_len), s>>>(p);
  return check_cuda(cudaGetLastError(), "residual_gate_plan_kernel");
}

extern "C" int pk_exclusive_scan_i32(const int* lens, int n, int* cu_out, int* total_out, void* stream) {
  PK_REQUIRE(lens && cu_out && n >= 0, "pk_exclusive_scan_i32: bad arguments");
  exclusive_scan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(lens, n, cu_out, total_out);
  return check_cuda(cudaGetLastError(), "exclusive_scan_kernel");
}

extern "C" int pk_compact_rows(const pk_compact_args* a, void* stream) {
  PK_REQUIRE(a && a->x_in && a->x_out && a->cu_in && a->cu_out && a->dst_local && a->sample_of, "pk_compact_rows: null pointer");
  PK_REQUIRE(a->dim % 4 == 0 && a->rows_in_cap >= 0, "pk_compact_rows: bad dim");
  PK_REQUIRE((a->a0_in == nullptr) == (a->a0_out == nullptr) || a->ghost, "pk_compact_rows: attribute 0 in/out mismatch");
  if (a->batch == 0) return PK_OK;
  CompactParams p;
  p.x_in = a->x_in; p.x_out = a->x_out; p.dim = a->dim;
  p.cu_in = a->cu_in; p.cu_out = a->cu_out; p.batch = a->batch;
  p.dst_local = a->dst_local; p.sample_of = a->sample_of;
  p.scale_in = a->scale_in; p.scale_out = a->scale_out;
  p.a0_in = a->a0_in; p.a0_out = a->a0_out; p.a1_in = a->a1_in; p.a1_out = a->a1_out; p.a2_in = a->a2_in; p.a2_out = a->a2_out;
  p.ghost = a->ghost;
  PK_REQUIRE(!a->pub_mask || (a->pub_tok_row && a->scale_in && a->pub_n_img >= 0), "pk_compact_rows: fused publish needs pub_tok_row and scale_in");
  p.pub_tok_row = a->pub_tok_row; p.pub_mask = a->pub_mask; p.pub_n_img = a->pub_n_img;
  compact_rows_kernel<<<grid_rows(static_cast<long long>(a->rows_in_cap) + a->batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError(), "compact_rows_kernel");
}

extern "C" int pk_residual_ghost(float* x, float* mult, const int* cu_seqlens, const float* mdrop, const float* mlp0, int batch, int dim,
                                 void* stream) {
  PK_REQUIRE(x && mult && cu_seqlens && mdrop && mlp0 && dim % 4 == 0, "pk_residual_ghost: bad arguments");
  if (batch == 0) return PK_OK;
  residual_ghost_kernel<<<grid_rows(batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mult, cu_seqlens, mdrop, mlp0, batch, dim);
  return check_cuda(cudaGetLastError(), "residual_ghost_kernel");
}

extern "C" int pk_residual_publish(const float* mask, const int* dst_local, const int* cu_out, int* tok_row, float* mask_pub, int batch,
                                   int n_img, void* stream) {
  PK_REQUIRE(mask && dst_local && cu_out && tok_row && mask_pub, "pk_residual_publish: null pointer");
  const int total = batch * n_img;
  if (total == 0) return PK_OK;
  residual_publish_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, dst_local, cu_out, tok_row, mask_pub,
                                                                                               batch, n_img);
  return check_cuda(cudaGetLastError(), "residual_publish_kernel");
}

extern "C" int pk_avit_halt_plan(const pk_avit_args* a, void* stream) {
  PK_REQUIRE(a && a->x && a->cu_in && a->c && a->R && a->tokid && a->out_acc && a->dst_local && a->sample_of && a->new_len && a->n_halted,
             "pk_avit_halt_plan: null pointer");
  PK_REQUIRE((a->rho == nullptr) == (a->counter == nullptr), "pk_avit_halt_plan: rho and counter go together");
  PK_REQUIRE(a->seq_total > 0 && a->seq_total <= 16384, "pk_avit_halt_plan: bad seq_total");
  if (a->batch == 0) return PK_OK;
  AvitParams p;
  p.x = a->x; p.cu_in = a->cu_in; p.dim = a->dim; p.seq_total = a->seq_total;
  p.c = a->c; p.R = a->R; p.tokid = a->tokid;
  p.gate_scale = a->gate_scale; p.gate_center = a->gate_center; p.eps = a->eps; p.last_layer = a->last_layer; p.early_exit = a->early_exit;
  p.out_acc = a->out_acc; p.rho = a->rho; p.counter = a->counter;
  p.dst_local = a->dst_local; p.sample_of = a->sample_of; p.new_len = a->new_len; p.n_halted = a->n_halted;
  avit_halt_plan_kernel<<<a->batch, 256, static_cast<size_t>(a->seq_total), static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError(), "avit_halt_plan_kernel");
}

// ------------------------------------------------------------------ un-permute of expert-sorted rows (moevit.py:54-61)
// x[src_of[r], :] += y[r, :]: the expert MLP outputs were computed in expert-sorted order by plain (fast-path) GEMMs;
// every destination row is written by exactly one source row, so this is a gather-add without atomics.
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ src_of, int rows, int dim) {
  const int lane = lane_id(), d4 = dim / 4;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp_id(); r < rows; r += warps_total) {
    const long long dst = src_of[r];
    const float4* src = reinterpret_cast<const float4*>(y + static_cast<long long>(r) * dim);
    float4* out = reinterpret_cast<float4*>(x + dst * dim);
    for (int c = lane; c < d4; c += 32) {
      const float4 a = src[c];
      float4 b = out[c];
      b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
      out[c] = b;
    }
  }
}

extern "C" int pk_scatter_add_rows(float* x, const float* y, const int* src_of, int rows, int dim, void* stream) {
  PK_REQUIRE(x && y && src_of && dim % 4 == 0 && rows >= 0, "pk_scatter_add_rows: bad arguments");
  if (rows == 0) return PK_OK;
  scatter_add_rows_kernel<<<grid_rows(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, src_of, rows, dim);
  return check_cuda(cudaGetLastError(), "scatter_add_rows_kernel");
}

// One-hot routing weights of every expert: onehot[e*rows + r] = (expert[r] == e).  Used as the rowscale of the out-proj
// residual epilogue of attention experts (moevit.py:89-96 combines the stacked expert outputs with the one-hot gate).
__global__ void expert_onehot_kernel(const int* __restrict__ expert, float* __restrict__ onehot, int rows, int n_experts) {
  const long long total = static_cast<long long>(rows) * n_experts;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i / rows);
    const int r = static_cast<int>(i % rows);
    onehot[i] = expert[r] == e ? 1.0f : 0.0f;
  }
}

extern "C" int pk_expert_onehot(const int* expert, float* onehot, int rows, int n_experts, void* stream) {
  PK_REQUIRE(expert && onehot && rows >= 0 && n_experts >= 1, "pk_expert_onehot: bad arguments");
  if (rows == 0) return PK_OK;
  const long long total = static_cast<long long>(rows) * n_experts;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  expert_onehot_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(expert, onehot, rows, n_experts);
  return check_cuda(cudaGetLastError(), "expert_onehot_kernel");
}

extern "C" int pk_moe_route(const float* x, const float* gamma, const float* beta, float eps, const float* gate_w, const float* gate_b,
                            int n_experts, int rows, int dim, int* expert, int* offsets, int* counts, int* src_of, int* sort_scratch,
                            void* stream) {
  PK_REQUIRE(x && gamma && beta && gate_w && gate_b && expert && offsets && counts && src_of && sort_scratch, "pk_moe_route: null pointer");
  PK_REQUIRE(n_experts >= 1 && n_experts <= kMaxExperts, "pk_moe_route: 1 <= n_experts <= %d", kMaxExperts);
  PK_REQUIRE(dim % 4 == 0 && dim <= 1024, "pk_moe_route: dim must be a multiple of 4, <= 1024");
  if (rows == 0) return PK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int maxv = (dim / 4 + 31) / 32;
  if (maxv <= 3 && n_experts == 4) moe_route_kernel<3, 4><<<grid_rows(rows), 256, 0, s>>>(x, gamma, beta, eps, gate_w, gate_b, n_experts, rows, dim, expert);
  else if (maxv <= 3) moe_route_kernel<3, 0><<<grid_rows(rows), 256, 0, s>>>(x, gamma, beta, eps, gate_w, gate_b, n_experts, rows, dim, expert);
  else moe_route_kernel<8, 0><<<grid_rows(rows), 256, 0, s>>>(x, gamma, beta, eps, gate_w, gate_b, n_experts, rows, dim, expert);
  PK_CHECK_CUDA(cudaGetLastError());
  int chunk = (rows + kSortBlocks - 1) / kSortBlocks;
  chunk = (chunk + 1023) / 1024 * 1024;
  if (chunk < 1024) chunk = 1024;
  const int n_blocks = (rows + chunk - 1) / chunk;
  moe_hist_kernel<<<n_blocks, 1024, 0, s>>>(expert, rows, n_experts, sort_scratch);
  PK_CHECK_CUDA(cudaGetLastError());
  moe_scan_kernel<<<1, 32, 0, s>>>(n_blocks, n_experts, sort_scratch, offsets, counts);
  PK_CHECK_CUDA(cudaGetLastError());
  moe_scatter_kernel<<<n_blocks, 1024, 0, s>>>(expert, rows, n_experts, sort_scratch, offsets, src_of);
  return check_cuda(cudaGetLastError(), "moe_scatter_kernel");
}
