/* peekvit_b200 — C ABI of the B200-native encoder-forward kernels.
 *
 * The reference (alessiodevoto/peekvit) has no FFI layer: its hot path is eager PyTorch
 * (`nn.Conv2d`, `nn.LayerNorm`, `nn.MultiheadAttention`, `nn.Linear`, `F.gelu`, `torch.norm`,
 * `torch.argsort`, `torch.gather`) called from the `forward()` of five `nn.Module` classes.
 * Each entry point below replaces one of those ATen call sites (cited as reference
 * file:line, relative to the reference repo root); the Python host in `peekvit_b200/`
 * binds them with `ctypes` behind the reference's own module API.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer owned by the caller (PyTorch allocates);
 *     nothing is allocated, freed or retained by the library except an immutable per-process
 *     context (TMA-descriptor cache, one watchdog word);
 *   - every call is asynchronous on `stream` (a `cudaStream_t` passed as `void*`);
 *   - return value: 0 = PK_OK, negative = error; `pk_last_error()` gives the message;
 *   - "rows" are tokens packed sample after sample; ragged batches are described by an
 *     int32 `cu_seqlens[B+1]` prefix array exactly like a varlen attention API;
 *   - activations entering a GEMM are bf16, accumulation / LayerNorm statistics / softmax /
 *     the residual stream are fp32.
 */
#ifndef PEEKVIT_B200_H
#define PEEKVIT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 2

/* ---- runtime ------------------------------------------------------------------------- */
int pk_abi_version(void);
/* Create the context of `device` (one per device of the process; idempotent).  Every other entry point works on the
 * context of the caller's CURRENT CUDA device, which must be the device its pointers live on. */
int pk_init(int device);
const char* pk_last_error(void);
int pk_num_sms(void);
/* Synchronises the device and returns the watchdog word (0 = healthy; otherwise the code of
 * the bounded mbarrier wait that expired); `reset` != 0 clears it. */
int pk_device_flag(int reset);
/* Stream-ordered copy of the watchdog word into pinned host memory (no synchronisation): lets the host notice an expired
 * wait at its next call without stalling the pipeline. */
int pk_device_flag_async(unsigned int* host_dst, void* stream);

/* ---- K1/K3/K5/K6/K7: tcgen05 GEMM with fused epilogue ------------------------------- */
enum pk_epilogue {
  PK_EPI_BIAS_BF16 = 0,      /* out_bf16 = acc + bias                       (in-proj, blocks.py:94)   */
  PK_EPI_BIAS_GELU_BF16 = 1, /* out_bf16 = gelu_erf(acc + bias)             (fc1+GELU, blocks.py:81-82) */
  PK_EPI_BIAS_RESID_F32 = 2, /* out_f32 = rowscale*(acc + bias) + resid     (out-proj/fc2 + residual, vit.py:49-55;
                                rowscale = ResidualViT forward mask, residualvit.py:254) */
  PK_EPI_BIAS_F32 = 3        /* out_f32 = acc + bias */
};

typedef struct pk_gemm_args {
  const void* A;         /* bf16 [M, K], row-major, leading dimension lda (elements) */
  const void* W;         /* bf16 [N, K], row-major (nn.Linear weight layout), leading dimension ldw */
  int M, N, K;
  long long lda, ldw;
  const float* bias;     /* [N] or NULL */
  int epilogue;          /* enum pk_epilogue */
  void* out;             /* bf16 or f32 [*, N], leading dimension ldo (elements) */
  long long ldo;
  const float* resid;    /* f32, PK_EPI_BIAS_RESID_F32 only (may alias out) */
  long long ldr;
  const float* rowscale; /* f32 [M] or NULL */
  /* Row remap for the patch-embedding GEMM (vit.py:212-236,:92): when rows_per_group > 0,
   * GEMM row m = g*rows_per_group + p is written to out row g*group_stride + group_offset + p,
   * and the residual row is (pos_offset + p) of a [seq, N] table when resid_is_pos != 0
   * (pos_embedding), else the out row. */
  int rows_per_group, group_stride, group_offset, resid_is_pos, pos_offset;
  const int* m_dev;      /* optional device-side row count (<= M): ragged batches without a host sync */
  const int* row_begin_dev; /* optional device-side first row: the launch covers rows [*row_begin_dev,
                               *row_begin_dev + *m_dev) of A / out (one expert's segment of a grouped GEMM,
                               moevit.py:54-59 computed for the routed tokens only) */
  const int* out_row_index; /* optional int32 [rows]: GEMM row r is written to (and its residual read from)
                               row out_row_index[r] (un-permute of expert-sorted tokens) */
  int block_n;           /* 0 = auto, or 128 / 192 / 256 */
  int max_ctas;          /* 0 = one CTA per SM */
  int epilogue_mode;     /* 0 = auto (TMA tile store when rows are contiguous), 2 = force the SIMT epilogue */
  int cta_pair;          /* 0 = auto (CTA-pair 256 x block_n tiles, tcgen05 cta_group::2, when rows are contiguous),
                            1 = single-CTA 128 x block_n kernel, 2 = force the CTA-pair kernel */
  /* LayerNorm fused across two GEMMs (x -> LN -> Linear, vit.py:48-55) instead of a LayerNorm kernel in between.
   * Producer (PK_EPI_BIAS_RESID_F32 only): besides out, write xb_out = bf16(out) and, per row, the partial
   * (sum, sum of squares) of each column tile into row_stats[row][part][2], part < pk_gemm_row_stat_parts(N). */
  void* xb_out;          /* bf16 [M, N], leading dimension ldxb, or NULL */
  long long ldxb;
  float* row_stats;      /* f32 [M][parts][2] */
  /* Consumer (PK_EPI_BIAS_BF16 / _GELU_BF16): A is such a raw bf16 copy and W carries the LayerNorm gain
   * (W' = bf16(gamma * W)); with mean/rstd of row r from ln_stats (summed over ln_parts slots, ln_dim elements, eps):
   *   out[r,n] = act( rstd_r * acc[r,n] - rstd_r * mean_r * ln_c1[n] + bias[n] ),
   *   ln_c1[n] = sum_k W'[n,k],  bias[n] = linear bias[n] + sum_k beta[k] * W[n,k]. */
  const float* ln_stats; /* f32 [M][ln_parts][2] or NULL */
  const float* ln_c1;    /* f32 [N] */
  int ln_parts, ln_dim;
  float ln_eps;
  /* Two-term split operands (the "bf16x2" arithmetic mode: x = hi + lo, both bf16, 16 significant bits).  The product
   * A*W^T ~ lo*Wh + hi*Wl + hi*Wh runs as ONE bf16 GEMM over K = 3k (small terms first, see pk_split2_bf16): the weight row
   * is [Wh | Wl | Wh] (3k wide) and the activation row is stored ONCE as [lo | hi] (2k wide) -- with a_wrap_k = k the
   * kernel reads A column c - k for every K index c >= 2k.  CTA-pair kernel only.  0 = off (A is K wide). */
  int a_wrap_k;
  /* Output format of the two 2-byte epilogues (PK_EPI_BIAS_BF16 / PK_EPI_BIAS_GELU_BF16), CTA-pair kernel only:
   * PK_OUT_BF16 (default); PK_OUT_F16 = IEEE half (the fp16 attention operands of the bf16x2 mode);
   * PK_OUT_BF16X2 = the value split into hi + lo (both bf16) and written as [lo | hi]: `out` is [*, 2N], lo at
   * column n, hi at column N + n (N % 64 == 0) -- the A operand of the next split GEMM without an extra pass. */
  int out_format;
  /* Grouped launch (MoE expert MLPs, moevit.py:49-61 for the routed tokens only; CTA-pair kernel, epilogues without a staged
   * residual): group_offsets != NULL makes ONE launch cover n_groups row segments [group_offsets[g], group_offsets[g + 1]) of
   * A / out (device-side int32[n_groups + 1], the expert-sorted layout of pk_moe_route), each multiplied by its own weight:
   * W is the groups' weights stacked ([n_groups * N, K]), bias their biases ([n_groups * N]); N stays one group's width.
   * The tile scheduler walks the segments' 256-row tiles; m_dev / row_begin_dev are ignored. */
  const int* group_offsets; int n_groups;
} pk_gemm_args;

enum pk_out_format { PK_OUT_BF16 = 0, PK_OUT_F16 = 1, PK_OUT_BF16X2 = 2 };

int pk_gemm_bf16(const pk_gemm_args* args, void* stream);
/* Number of statistics slots per row the LayerNorm-producer epilogue writes for an N-column output. */
int pk_gemm_row_stat_parts(int N);
/* Statistics + raw bf16 copy of rows that do not come out of a producer GEMM (the embedding output):
 * xb[r,:] = bf16(x[r,:]); row_stats[r][0] = (sum, sum of squares); row_stats[r][1..parts) = 0. */
int pk_row_stats_cast(const float* x, void* xb, float* row_stats, int rows, int dim, int parts, void* stream);

/* ---- K1: patchify (the im2col half of conv_proj, vit.py:212-220) --------------------- */
/* images f32 [B,3,S,S] NCHW -> patches bf16, K order (c,i,j) to match conv_proj.weight.reshape(D, 3*p*p).
 * Patch q of sample b goes to row b*rows_per_sample + row_offset + q (rows_per_sample <= 0: densely packed,
 * [B*(S/p)^2, 3*p*p]).  With rows_per_sample = tokens per sample the patch GEMM's rows ARE the token rows of the
 * residual stream (class / register / budget rows stay zero in the operand and are masked by rowscale = 0). */
int pk_patchify(const float* images, void* patches, int batch, int image_size, int patch_size, int rows_per_sample,
                int row_offset, void* stream);

/* Input path: uint8 HWC images [B,S,S,3] as decoded, with T.ToTensor (u/255) and T.Normalize
 * ((t - mean_c)/std_c) (reference data/imagenette.py:69-73) fused into the im2col: 1 byte per pixel-channel is read
 * (and copied host->device) instead of 4.  mean3 / std3 are HOST pointers to 3 floats.  Bit-identical patches to
 * pk_patchify on the float tensor those transforms produce. */
int pk_patchify_u8(const unsigned char* images_hwc, void* patches, int batch, int image_size, int patch_size,
                   const float* mean3, const float* std3, int rows_per_sample, int row_offset, void* stream);

/* Rows of the residual stream that do not come from the patch GEMM: class / register tokens
 * (vit.py:230-236 then + pos_embedding, vit.py:92) and the ResidualViT budget token
 * (residualvit.py:572-583; it gets no pos_embedding, :338-345).
 * x[b*seq_stride + row_offset + t, :] = scale * tokens[t, :] + (pos ? pos[row_offset + t, :] : 0)
 * for t in [0, n_tokens); if `tokens` is NULL the row is filled with `scale`. */
int pk_fill_token_rows(float* x, int batch, int seq_stride, int row_offset, int n_tokens, int dim,
                       const float* tokens, const float* pos, float scale, void* stream);

/* ---- K2: LayerNorm (vit.py:48,53,95; eps 1e-5, 1e-6 in ResidualViT blocks) ----------- */
/* y_bf16[r,:] = rowscale[r] * LN(x[src(r),:]) ; src(r) = row_index ? row_index[r] : r.
 * rows_dev (optional) overrides `rows` with a device-side count. */
int pk_layernorm_bf16(const float* x, void* y, const float* gamma, const float* beta, float eps,
                      int rows, int dim, const float* rowscale, const int* row_index, const int* rows_dev,
                      void* stream);

/* ---- K4: attention over packed rows (blocks.py:93-95 -> nn.MultiheadAttention) ------- */
typedef struct pk_attention_args {
  const void* qkv;          /* bf16 [rows, 3*D]: q | k | v, head h = columns [h*dh, (h+1)*dh) of each */
  void* out;                /* bf16 [rows, D] */
  int batch, num_heads, head_dim;
  int seq_len;              /* uniform tokens per sample when cu_seqlens == NULL */
  const int* cu_seqlens;    /* int32 [batch+1] or NULL */
  int max_seq_len;          /* upper bound on tokens per sample (grid sizing) */
  float scale;              /* 1/sqrt(dh) (torch functional.py MHA pre-scales q) */
  /* sparse-model extensions: a key row j stands for key_mult[j] identical
   * tokens (+log mult on its logit); one virtual key/value per head = the in-proj bias slices
   * (what a zeroed token projects to), weighted by extra_mult[b] identical dropped tokens. */
  const float* key_mult;    /* f32 [rows] or NULL */
  const void* extra_kv;     /* bf16 [2*D]: k-bias | v-bias, or NULL */
  const float* extra_mult;  /* f32 [batch] or NULL (<= 0 disables the virtual key for that sample) */
  int impl;                 /* 0 = auto: head_dim 64 runs on tcgen05/TMEM -- the dense kernel for uniform 64 < seq_len <= 256
                               without multiplicities, the ragged kernel for cu_seqlens / multiplicities / the virtual key /
                               short uniform sequences (<= 256 keys per sample) -- everything else on the general mma.sync
                               kernel; 1 = general kernel, 2 = dense tcgen05 kernel, 3 = ragged tcgen05 kernel, 4 = quad-region
                               ragged tcgen05 kernel (cu_seqlens, <= 128 keys per sample) */
  /* bf16x2 arithmetic mode (tcgen05 kernel only, impl 0 / 2 on an eligible shape): qkv_format PK_OUT_F16 = q, k, v are IEEE
   * half (11 significant bits; the probabilities are packed as half too); out_format PK_OUT_BF16X2 = `out` is
   * [rows, 2*D] holding the fp32 result split into lo (column d) and hi (column D + d), both bf16. */
  int qkv_format, out_format;
  /* Device-side choice between the two ragged kernels (both are launched, the one not chosen exits at once): with
   * route_rows set, the ragged tcgen05 kernel runs when *route_rows >= route_min_rows (long samples), the general mma.sync
   * kernel otherwise (the two-region TMEM pipeline does not pay off below ~130 rows per sample: profiles/r02).  Independent
   * of that, a call whose max_seq_len (+ the virtual key) is at most 128 runs on the quad-region tcgen05 kernel alone (four
   * samples in flight per SM); with PK_ATT_SPLIT=1 in the environment a mixed batch is split PER SAMPLE on the device: the
   * quad-region kernel takes the samples of at most 128 keys, the kernel chosen above only the longer ones (opt-in: it loses
   * on batches whose short samples are nearly empty). */
  const int* route_rows; int route_min_rows;
  int total_rows;           /* rows of the qkv / out buffers (the ragged tcgen05 kernel's 2-D tensor maps need the extent: a
                               key tile that starts near the end of the buffer is zero-filled past it); 0 = unknown (the ragged
                               kernel is then not used).  Rows of `qkv` past the live ones must hold finite values. */
} pk_attention_args;

int pk_attention_fwd(const pk_attention_args* args, void* stream);
/* Debug aid: with PK_ATT_TRACE=1 in the environment, CTA 0 of the tcgen05 attention kernel records clock64 stamps of its
 * pipeline events; this copies the 16 x 16 x 8 table (items x warps x events) to `host_dst` (2048 uint64). */
int pk_attention_trace(unsigned long long* host_dst);

/* ---- K2(final)+K8: final LayerNorm on class rows, class-token sum, head (vit.py:95,242-246) */
/* logits[b, c] = head_b[c] + sum_d head_w[c,d] * sum_{t<n_cls} LN(x[row(b)+t, :])[d]
 * row(b) = cu_seqlens ? cu_seqlens[b] : b*seq_len. */
int pk_cls_head(const float* x, int batch, int seq_len, const int* cu_seqlens, int n_cls, int dim,
                const float* gamma, const float* beta, float eps,
                const float* head_w, const float* head_b, int num_classes, float* logits, void* stream);
/* feat[b, :] = sum_{t<n_cls} LN(x[row(b)+t, :]) (f32 [batch, dim]): the head's input on its own, for batches large enough
 * that the head runs as a split-operand tensor-core GEMM (pk_split3_bf16 + pk_gemm_bf16) instead of inside pk_cls_head. */
int pk_cls_features(const float* x, int batch, int seq_len, const int* cu_seqlens, int n_cls, int dim, const float* gamma,
                    const float* beta, float eps, float* feat, void* stream);

/* ---- eval loop: top-1 prediction and accuracy counts on the device (validate/test.py:116-129 does
 * logits.argmax + torchmetrics on the host side).  pred_out[b] = first arg-max of logits[b, :] (int32, optional);
 * if labels (int64) and counts (int64[2]) are given: counts[0] += #(pred == label), counts[1] += batch. */
int pk_argmax_count(const float* logits, const long long* labels, int batch, int num_classes, int* pred_out,
                    long long* counts, void* stream);

/* ---- K9/K10/K11: RankViT sort_and_drop (rankvit.py:55-77) ----------------------------- */
/* scores[b, i] = || x[b, 1+i, :] ||_2 for the n = seq_len-1 non-class tokens (rankvit.py:63). */
int pk_token_norm_score(const float* x, float* scores, int batch, int seq_len, int dim, void* stream);
/* kept[b, r] = index of the r-th largest score of row b, r < k; ties -> lowest index
 * (stable descending order; rankvit.py:67 + north-star tie rule). n <= 4096. */
int pk_topk_select(const float* scores, int* kept, int batch, int n, int k, void* stream);
/* y[b, 0, :] = x[b, 0, :]; y[b, 1+r, :] = x[b, 1+kept[b,r], :]  (rankvit.py:71-77), f32 rows. */
int pk_gather_rows(const float* x, float* y, const int* kept, int batch, int seq_len, int k, int dim, void* stream);

/* ---- ragged-batch plumbing shared by the sparse models -------------------------------- */
/* cu_out[0..n] = exclusive prefix sum of lens[0..n) (cu_out[n] = total, also written to *total_out). */
int pk_exclusive_scan_i32(const int* lens, int n, int* cu_out, int* total_out, void* stream);

/* Gather the kept rows of every sample into a new packed buffer:
 * x_out[cu_out[b] + dst_local[r]] = scale_in[r] * x_in[r] for rows with dst_local[r] >= 0, carrying up to three
 * per-row float attributes along.  ghost != 0 additionally zero-fills each sample's LAST output row
 * (a0_out = 0, scale_out = 1): the ResidualViT ghost slot. */
typedef struct pk_compact_args {
  const float* x_in; float* x_out; int dim;
  const int* cu_in; const int* cu_out; int batch;
  int rows_in_cap;                 /* host-side upper bound on cu_in[batch] (grid sizing) */
  const int* dst_local; const int* sample_of;
  const float* scale_in; float* scale_out;
  const float* a0_in; float* a0_out;
  const float* a1_in; float* a1_out;
  const float* a2_in; float* a2_out;
  int ghost;
  /* optional, fused pk_residual_publish (same launch): pub_mask[b,i] = scale_in[pub_tok_row[b,i]], then pub_tok_row moves to
   * the compacted layout.  NULL pub_mask = off; needs scale_in. */
  int* pub_tok_row; float* pub_mask; int pub_n_img;
} pk_compact_args;
int pk_compact_rows(const pk_compact_args* args, void* stream);

/* ---- K12/K13: ResidualViT budget gating (residualvit.py:47-74,197-244) ----------------- */
/* thr_out[0] = 1 - mean over the batch and D of the budget-token rows (fixed budget: residualvit.py:208,:62). */
int pk_budget_mean_threshold(const float* x, const int* cu_seqlens, int batch, int budget_pos, int dim, float* thr_out,
                             void* stream);

/* Per sample: threshold from its budget token, soft mask of every live row
 *   mask = relu(sigmoid((w.x + b)/temp + gate_bias) - thr)         (sigmoid gate, blocks.py:62-69, residualvit.py:62-69)
 *   mask = round(sigmoid(w.x + b))                                 (gumbel gate in eval, blocks.py:55-57)
 * the keep decision (mask > 0 and multiplicity > 0; the first n_special rows always), the position of every
 * kept row in the compacted sample, the new length (+1 ghost slot when gated) and the multiplicity folded
 * into the virtual key / ghost row. */
typedef struct pk_residual_gate_args {
  const float* x; const int* cu_in; const float* mult_in;
  int batch, dim, max_seq_len;
  int n_special, budget_pos;       /* budget_pos = local row of the budget token, or -1 */
  int gated;                       /* 0: plain layer (skip None): keep every live row with mask 1, no ghost */
  const float* gate_w; float gate_b, gate_temp, gate_bias; int gate_type;   /* 0 sigmoid, 1 gumbel */
  int thr_mode;                    /* 0: sigmoid(bt_w.budget_row + bt_b) per sample; 1: *thr_dev; 2: thr_const */
  const float* bt_w; float bt_b; const float* thr_dev; float thr_const;
  float* mask; int* dst_local; int* sample_of; int* new_len; float* mdrop;
} pk_residual_gate_args;
int pk_residual_gate_plan(const pk_residual_gate_args* args, void* stream);

/* After the block: each sample's ghost row (its last row) becomes mlp0 = fc2(gelu(fc1.bias)) + fc2.bias with
 * multiplicity mdrop[b] (what every dropped token equals when it leaves the reference block). */
int pk_residual_ghost(float* x, float* mult, const int* cu_seqlens, const float* mdrop, const float* mlp0, int batch, int dim,
                      void* stream);

/* mask_pub[b,i] = mask[tok_row[b,i]] (the reference's block.mask, (B,N_img,1), utils/utils.py:100-122), then
 * tok_row[b,i] moves to the compacted layout (dropped tokens -> the ghost row). */
int pk_residual_publish(const float* mask, const int* dst_local, const int* cu_out, int* tok_row, float* mask_pub, int batch,
                        int n_img, void* stream);

/* ---- K14: AViT halting update + plan (adavit.py:140-219) -------------------------------- */
typedef struct pk_avit_args {
  const float* x; const int* cu_in; int batch, dim, seq_total;
  float* c; float* R; const float* tokid;
  float gate_scale, gate_center, eps; int last_layer, early_exit;
  float* out_acc; float* rho; float* counter;
  int* dst_local; int* sample_of; int* new_len; float* n_halted;
} pk_avit_args;
int pk_avit_halt_plan(const pk_avit_args* args, void* stream);

/* ---- K15: MoE routing (moevit.py:23-32,49-61) ------------------------------------------- */
/* expert[r] = argmax_e(LN(x[r]).gate_w[e] + gate_b[e]) (first maximum), then a stable counting sort:
 * offsets[E+1], counts[E], src_of[pos] = original row of the pos-th expert-sorted row.  n_experts <= 16. */
int pk_moe_route(const float* x, const float* gamma, const float* beta, float eps, const float* gate_w, const float* gate_b,
                 int n_experts, int rows, int dim, int* expert, int* offsets, int* counts, int* src_of, int* sort_scratch,
                 void* stream);
#define PK_MOE_SORT_SCRATCH_INTS 4096 /* caller-owned scratch of the multi-block counting sort (256 chunks x 16 experts) */

/* x[src_of[r], :] += y[r, :] for r < rows: un-permute + residual add of expert outputs computed in expert-sorted order
 * (moevit.py:54-61; src_of from pk_moe_route is a permutation, so no two rows collide). */
int pk_scatter_add_rows(float* x, const float* y, const int* src_of, int rows, int dim, void* stream);

/* onehot[e*rows + r] = (expert[r] == e) as f32, e < n_experts: the per-expert rowscale with which the out-proj residual
 * epilogue of attention experts accumulates only the rows routed to that expert (AttentionMoE.forward_moe, moevit.py:85-96). */
int pk_expert_onehot(const int* expert, float* onehot, int rows, int n_experts, void* stream);

/* NoiseBlock.forward_snr (models/blocks.py:117-131): x[r,:] += noise[r,:] * sqrt(mean(x[r,:]^2) / 10^(snr_db/10)).
 * The caller draws ``noise`` (torch.randn_like in the reference) and skips the call for snr_db == 0 like the reference. */
int pk_noise_snr(float* x, const float* noise, int rows, int dim, float snr_db, void* stream);

/* NoiseBlock.forward_token_drop (models/blocks.py:141-157): zero rows b*seq + tokens[j] for every sample b, j < n_tokens
 * (the same randperm prefix for the whole batch). */
int pk_zero_token_rows(float* x, int batch, int seq, const int* tokens, int n_tokens, int dim, void* stream);

/* Masked row arithmetic of the dense-layout ResidualViT modes (models/residualvit.py:130-194 forward_skip_attention /
 * forward_skip_mlp, :239-242 add_input): out[r,:] = (accumulate ? out[r,:] : 0) + w(r) * a[r,:] with w(r) = scale[r], or
 * 1 - scale[r] when one_minus -- ``mask * img_tokens`` and ``img_tokens * (1 - mask)``.  out may alias a when !accumulate. */
int pk_row_scale_add(float* out, const float* a, const float* scale, int rows, int dim, int one_minus, int accumulate, void* stream);

/* ---- fp32-accurate mode (the reference's shipped dtype; north star: logits within 1e-5).  GEMMs run on the same bf16
 * tcgen05 kernels with 3-way split operands: x = h + m + l in bf16, products {mm, lh, hl, mh, hm, hh} laid along K' = 6K,
 * small terms first.  pk_split3_bf16 writes the ACTIVATION-side row [m|l|h|m|h|h] (out bf16 [rows, 6*dim]) of x f32 [rows, dim],
 * optionally after mode 1 = exact erf GELU (blocks.py:82) or mode 2 = LayerNorm(gamma, beta, eps) (vit.py:48,53).  The matching
 * WEIGHT-side row is [m|h|l|h|m|h]. */
int pk_split3_bf16(const float* x, void* out, int rows, int dim, int mode, const float* gamma, const float* beta, float eps,
                   const float* rowscale /* optional: pre(x)[r] *= rowscale[r] (soft masks, residualvit.py:252,258) */,
                   const int* row_index /* optional, LayerNorm mode: output row r reads x[row_index[r]] */,
                   const int* rows_dev /* optional device-side row count */, void* stream);
/* im2col of f32 NCHW images into split activation rows: patches6 bf16 [B*(S/p)^2, 6*3*p*p] (vit.py:203-222, fp32 mode). */
int pk_patchify_split3(const float* images, void* patches6, int batch, int image_size, int patch_size, void* stream);
/* Two-term variants of pk_split3_bf16 / pk_patchify_split3 (bf16x2 mode): x = hi + lo.  pk_split2_bf16 writes the
 * activation row [lo | hi] (2*dim wide; read by pk_gemm_bf16 with a_wrap_k = dim); the matching weight row is
 * [Wh | Wl | Wh].  pk_patchify_split2 writes [lo | hi | hi] (3*Kp wide: the patch GEMM's row remap runs on the
 * single-CTA kernel, which has no a_wrap_k). */
int pk_split2_bf16(const float* x, void* out, int rows, int dim, int mode, const float* gamma, const float* beta, float eps,
                   const float* rowscale, const int* row_index, const int* rows_dev, void* stream);
int pk_patchify_split2(const float* images, void* patches3, int batch, int image_size, int patch_size, void* stream);

/* fp32 attention core on the CUDA cores: qkv f32 [rows, 3*H*dh] (q|k|v) -> out f32 [rows, H*dh], dh 32, 48 or 64 (blocks.py:93-95,
 * fp32 mode).  Uniform samples of seq_len rows, or ragged ones (cu_seqlens int32 [B+1]; seq_len = longest sample); optional
 * per-key multiplicities and one virtual key per sample, with the meaning they have in pk_attention_args. */
int pk_attention_f32(const float* qkv, float* out, int batch, int num_heads, int head_dim, int seq_len, float scale,
                     const int* cu_seqlens, const float* key_mult, const float* extra_kv, const float* extra_mult, void* stream);

/* ---- fine-tuning path: backward of the frozen-backbone regime ---------------------------------------
 * The reference trains with `train_only_these_params(model, ['gate', 'class', 'head', 'threshold', 'budget'])`
 * (train/train.py:97-127, models/topology.py:128-158): gradients are needed for a few small parameters only, but the class
 * tokens are an INPUT of the encoder, so the loss gradient travels back through every block as an activation gradient.
 * The four dX GEMMs of a block are pk_gemm_bf16 calls on the transposed weights; the entry points below are the rest
 * (what torch autograd runs as LayerNormBackward / GeluBackward / the MHA core backward / NllLossBackward / AddmmBackward in
 * `loss.backward()`, train/train.py:113). */
int pk_cast_f32_bf16(const float* x, void* y_bf16, long long n, void* stream);
/* hid = gelu(h_pre) and dh_pre = dhid * gelu'(h_pre), exact erf form (models/blocks.py:82), bf16 in / out; the training
 * forward keeps the fc1 pre-activation instead of fusing the GELU into the fc1 epilogue. */
int pk_gelu_bf16(const void* h_pre, void* hid, long long n, void* stream);
int pk_gelu_bwd_bf16(const void* h_pre, const void* dhid, void* dh_pre, long long n, void* stream);
/* Input gradient of LayerNorm (gamma / beta frozen): dx (+)= rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
 * row_index (optional) maps launch row r to row row_index[r] of x / dx; dy row = r / dy_div (the class rows of the final
 * LayerNorm share one feature gradient, vit.py:242-243). */
int pk_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, float* dx, int rows, int dim,
                     const int* row_index, int dy_div, int accumulate, void* stream);
/* Backward of the attention core softmax(Q K^T * scale) V for uniform sequences (<= 256 tokens, head_dim 32 / 64):
 * qkv [rows, 3D] and out = the forward result [rows, D], dout [rows, D] -> dqkv [rows, 3D], all bf16 (blocks.py:93-95). */
int pk_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, int batch, int num_heads, int head_dim,
                     int seq_len, float scale, void* stream);
/* Cross-entropy of int64 labels (train/train.py:106-107, nn.CrossEntropyLoss mean): *loss_sum += sum_b -log p[label] *
 * inv_count, dlogits = (softmax - onehot) * inv_count, *correct += #(argmax == label) (optional). */
int pk_softmax_xent(const float* logits, const long long* labels, int batch, int classes, float inv_count, float* loss_sum,
                    float* dlogits, int* correct, void* stream);
/* Linear head backward (vit.py:246): d_weight += dlogits^T feat, d_bias += sum_b dlogits, d_feat = dlogits weight. */
int pk_head_bwd(const float* dlogits, const float* feat, const float* weight, int batch, int classes, int dim, float* d_weight,
                float* d_bias, float* d_feat, void* stream);
/* Backward of pk_gather_rows (RankViTBlock.sort_and_drop, rankvit.py:55-77; no gradient through the indices):
 * x[b * seq_len + tok(o), :] = y[b * (k + 1) + o, :], tok(0) = 0, tok(o) = 1 + kept[b, o - 1]; x is zeroed by the caller. */
int pk_scatter_rows(const float* y, float* x, const int* kept, int batch, int seq_len, int k, int dim, void* stream);

/* --- gate regime of ResidualViT (train/train.py:99-100 with models/residualvit.py:47-74,197-260; sigmoid gate, learnable budget
 * token).  Training-mode block on the dense layout [class, budget, image tokens ...] with a soft mask m >= 0 per image token:
 *   mi = m*x,  a = m*LN1(mi),  x1 = mi + m*(Wo attention(Wqkv a) + bo),  y = m*LN2(x1),  out = x1 + mlp(y)
 *   m = relu(sigmoid((x.w_g + b_g)/temp + bias) - thr_b),  thr_b = sigmoid(x_budget.w_bt + b_bt)       (:212,:47-74)
 * Forward reuses the inference kernels (row-scaled LayerNorm / GEMM epilogues); these entry points are the gate itself and
 * the backward pieces the masks add. */
/* rowscale [batch*seq] (1 on the n_special leading rows), mask / sig [batch, seq - n_special], thr [batch]; gate_b / bt_b point
 * at the live one-element bias parameters on the device (an optimiser updates them every step: no host copy). */
int pk_residual_gate_train_fwd(const float* x, int batch, int seq, int n_special, int budget_pos, int dim, const float* gate_w,
                               const float* gate_b, float gate_temp, float gate_bias, const float* bt_w, const float* bt_b,
                               float* rowscale, float* mask, float* sig, float* thr, void* stream);
/* dm [batch*seq] = the block's d mask row sums; dmask_ext [batch, n_img] or NULL = gradient of a regulariser on the published
 * mask (utils/losses.py).  Adds into dx (image rows: dlogit * w_g, budget row: dz * w_bt) and into the four parameter
 * gradients (fp32 atomics). */
int pk_residual_gate_train_bwd(const float* x, const float* dm, const float* dmask_ext, const float* mask, const float* sig,
                               const float* thr, int batch, int seq, int n_special, int budget_pos, int dim, const float* gate_w,
                               float gate_temp, const float* bt_w, float* dx, float* g_gate_w, float* g_gate_b, float* g_bt_w,
                               float* g_bt_b, void* stream);
/* pk_layernorm_bwd for a LayerNorm whose OUTPUT is multiplied by rowscale[r]: dx (+)= LN_bwd(rowscale[r] * dy[r]) and
 * dot_out[r] += dy[r] . LN(x)[r]  (the mask's gradient from this site). */
int pk_layernorm_bwd_gated(const float* x, const float* dy, const float* gamma, const float* beta, float eps, float* dx, int rows,
                           int dim, const float* rowscale, float* dot_out, int accumulate, void* stream);
/* y_bf16[r, :] = rowscale[r] * x[r, :] */
int pk_cast_rows_f32_bf16(const float* x, void* y_bf16, const float* rowscale, int rows, int dim, void* stream);
/* out[r] (+)= alpha * sum_d a[r,d] * (b[r,d] - c[r,d]) / div[r]   (c, div optional; rows with div[r] <= 0 contribute 0) */
int pk_rowdot(const float* a, const float* b, const float* c, const float* div, float* out, int rows, int dim, float alpha,
              int accumulate, void* stream);
/* out[t, :] += sum_b x[b * seq + row0 + t, :]: gradient of the class / register token parameters (vit.py:230-236). */
int pk_sum_token_rows(const float* x, int batch, int seq, int row0, int n_rows, int dim, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEEKVIT_B200_H */
