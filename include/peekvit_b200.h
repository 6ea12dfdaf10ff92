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

#define PK_ABI_VERSION 1

/* ---- runtime ------------------------------------------------------------------------- */
int pk_abi_version(void);
/* Bind the calling process to `device`, create the context. Idempotent. */
int pk_init(int device);
const char* pk_last_error(void);
int pk_num_sms(void);
/* Synchronises the device and returns the watchdog word (0 = healthy; otherwise the code of
 * the bounded mbarrier wait that expired); `reset` != 0 clears it. */
int pk_device_flag(int reset);

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
   * and the residual row is (group_offset + p) of a [seq, N] table when resid_is_pos != 0
   * (pos_embedding), else the out row. */
  int rows_per_group, group_stride, group_offset, resid_is_pos;
  const int* m_dev;      /* optional device-side row count (<= M): ragged batches without a host sync */
  int block_n;           /* 0 = auto, or 128 / 192 / 256 */
  int max_ctas;          /* 0 = one CTA per SM */
} pk_gemm_args;

int pk_gemm_bf16(const pk_gemm_args* args, void* stream);

/* ---- K1: patchify (the im2col half of conv_proj, vit.py:212-220) --------------------- */
/* images f32 [B,3,S,S] NCHW -> patches bf16 [B*(S/p)^2, 3*p*p], K order (c,i,j) to match
 * conv_proj.weight.reshape(D, 3*p*p). */
int pk_patchify(const float* images, void* patches, int batch, int image_size, int patch_size, void* stream);

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
} pk_attention_args;

int pk_attention_fwd(const pk_attention_args* args, void* stream);

/* ---- K2(final)+K8: final LayerNorm on class rows, class-token sum, head (vit.py:95,242-246) */
/* logits[b, c] = head_b[c] + sum_d head_w[c,d] * sum_{t<n_cls} LN(x[row(b)+t, :])[d]
 * row(b) = cu_seqlens ? cu_seqlens[b] : b*seq_len. */
int pk_cls_head(const float* x, int batch, int seq_len, const int* cu_seqlens, int n_cls, int dim,
                const float* gamma, const float* beta, float eps,
                const float* head_w, const float* head_b, int num_classes, float* logits, void* stream);

/* ---- K9/K10/K11: RankViT sort_and_drop (rankvit.py:55-77) ----------------------------- */
/* scores[b, i] = || x[b, 1+i, :] ||_2 for the n = seq_len-1 non-class tokens (rankvit.py:63). */
int pk_token_norm_score(const float* x, float* scores, int batch, int seq_len, int dim, void* stream);
/* kept[b, r] = index of the r-th largest score of row b, r < k; ties -> lowest index
 * (stable descending order; rankvit.py:67 + north-star tie rule). n <= 4096. */
int pk_topk_select(const float* scores, int* kept, int batch, int n, int k, void* stream);
/* y[b, 0, :] = x[b, 0, :]; y[b, 1+r, :] = x[b, 1+kept[b,r], :]  (rankvit.py:71-77), f32 rows. */
int pk_gather_rows(const float* x, float* y, const int* kept, int batch, int seq_len, int k, int dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEEKVIT_B200_H */
