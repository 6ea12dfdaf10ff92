"""Fine-tuning path: forward + loss + backward for the reference's frozen-backbone regime.

The reference's training loop (train/train.py:97-127) freezes everything except the parameters whose names contain
``'gate'``, ``'class'``, ``'head'``, ``'threshold'`` or ``'budget'`` (``models/topology.py:128-158``,
``train_only_these_params``) and then runs ``out = model(batch); loss = criterion(out, labels); loss.backward();
optimizer.step()``.  For the dense ``VisionTransformer`` that regime trains the class tokens and the classification head
(what ``reinit_class_tokens`` + a new head on a pretrained backbone need, topology.py:100-116).  The class tokens are an
*input* of the encoder, so the loss gradient has to go back through all blocks as an activation gradient; no weight gradient of
the backbone is ever formed.

``FineTuner`` is the drop-in for the three middle lines of that loop::

    ft = FineTuner(model)                       # applies train_only_these_params and checks the regime is supported
    for batch, labels in loader:
        optimizer.zero_grad()
        loss, logits = ft.forward_backward(batch, labels)      # fills .grad of the trainable parameters (NCCL-averaged)
        optimizer.step()                        # any torch optimiser: the parameters are the nn.Module's own

Forward: the same tcgen05 GEMM / attention kernels as inference, with the activations a backward needs kept per layer (block
input, post-attention stream, q|k|v, attention output, fc1 pre-activation).  Backward: four dX GEMMs per block on the
transposed bf16 weights (``pk_gemm_bf16``), ``pk_layernorm_bwd``, ``pk_gelu_bwd_bf16``, ``pk_attention_bwd``,
``pk_softmax_xent`` + ``pk_head_bwd``; activation gradients are bf16 between GEMMs and fp32 on the residual path.  Under an
initialised ``torch.distributed`` process group the gradients are averaged over the ranks with ONE all-reduce of a flat
bucket (NCCL on GPUs), which is what DistributedDataParallel does for the reference.

Scope: ``VisionTransformer`` and ``RankVisionTransformer`` (its blocks rank and drop tokens in training exactly as in eval,
rankvit.py:55-97; the gather's backward is a scatter of the gradient rows, no gradient flows through the indices) with
trainable ``class_tokens`` / ``head.*``; dropout must be 0 (every shipped config).  The gate / threshold / budget parameters of the ResidualViT family train through its training-mode forward
(sampled budgets, soft masks), which is not built: constructing a FineTuner for such a model raises.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import engine, ops, runner
from ._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_RESID_F32

TRAIN_WORDS = ("gate", "class", "head", "threshold", "budget")          # train/train.py:99-100
_SUPPORTED = ("class_tokens", "head.weight", "head.bias")


def train_only_these_params(model, params_list: Sequence[str] = TRAIN_WORDS) -> List[str]:
    """models/topology.py:128-158: requires_grad = the parameter's name contains any of the words.  Returns the trainable names."""
    names = []
    for name, p in model.named_parameters():
        p.requires_grad = any(w in name for w in params_list)
        if p.requires_grad:
            names.append(name)
    return names


def flatten_grads(params: Sequence[torch.Tensor]) -> torch.Tensor:
    """One flat fp32 bucket holding every parameter's gradient (the unit of the gradient all-reduce)."""
    return torch.cat([p.grad.reshape(-1) for p in params])


def unflatten_grads(bucket: torch.Tensor, params: Sequence[torch.Tensor]) -> None:
    o = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(bucket[o:o + n].view_as(p.grad))
        o += n


def all_reduce_mean_(params: Sequence[torch.Tensor], group=None) -> int:
    """Average the gradients of ``params`` over the ranks of ``group`` with one all-reduce of a flat bucket (the fine-tuning
    path's only collective; NCCL over NVLink on GPUs, gloo in the CPU tests).  Returns the world size (1: nothing to do)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    bucket = flatten_grads(params)
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    bucket.div_(world)
    unflatten_grads(bucket, params)
    return world


class FineTuner:
    def __init__(self, model, train_words: Sequence[str] = TRAIN_WORDS, micro_batch: int = 128, process_group=None):
        if getattr(model, "_family", None) not in ("vit", "rankvit"):
            raise NotImplementedError(
                f"FineTuner: the backward path is built for VisionTransformer and RankVisionTransformer (class tokens + head "
                f"regime); {type(model).__name__} trains gates / thresholds / budget tokens through a training-mode forward that is not built")
        for blk in model.encoder.layers:
            if type(blk).__name__ not in ("ViTBlock", "RankViTBlock"):
                raise NotImplementedError(f"FineTuner: encoder.layers holds a {type(blk).__name__}")
        drops = [m.p for m in model.modules() if isinstance(m, torch.nn.Dropout)] + \
                [m.dropout for m in model.modules() if isinstance(m, torch.nn.MultiheadAttention)]
        if any(float(p) != 0.0 for p in drops):
            raise NotImplementedError("FineTuner: dropout / attention_dropout must be 0 (every shipped config)")
        self.model = model
        self.micro_batch = int(micro_batch)
        self.group = process_group
        self.names = train_only_these_params(model, train_words)
        bad = [n for n in self.names if n not in _SUPPORTED]
        if bad:
            raise NotImplementedError(f"FineTuner: no backward for {bad}; supported trainable parameters: {list(_SUPPORTED)}")
        self.params: Dict[str, torch.nn.Parameter] = {n: p for n, p in model.named_parameters() if n in self.names}
        self._wt: Dict[int, tuple] = {}             # layer -> transposed bf16 weights, keyed by the weight pack they came from
        self._wt_pack = None
        # RankViT: the kept-token indices of the last micro-batch, {layer: int32 [B, k]} (what a caller / test needs to
        # reproduce the step given identical selections; top-k is discontinuous in the bf16 scores)
        self.last_kept: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------ helpers
    def _transposed(self, pm: engine.PackedModel, l: int):
        """(W2^T [F, D], W1^T [D, F], Wo^T [D, D], Wqkv^T [D, 3D]) as bf16 'nn.Linear-layout' operands of the dX GEMMs:
        dX = dY @ W is pk_gemm_bf16(A = dY, W = W^T)."""
        if self._wt_pack is not pm:
            self._wt, self._wt_pack = {}, pm
        t = self._wt.get(l)
        if t is None:
            lw = pm.layers[l]
            aw, mw = lw.attn[0], lw.mlp[0]
            t = self._wt[l] = (mw.w_fc2.t().contiguous(), mw.w_fc1.t().contiguous(), aw.w_o.t().contiguous(), aw.w_qkv.t().contiguous())
        return t

    def _zero_grads(self) -> None:
        for p in self.params.values():
            if p.grad is None:
                p.grad = torch.zeros_like(p, dtype=torch.float32)

    # ------------------------------------------------------------------ one step
    def forward_backward(self, images: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``out = model(batch); loss = CrossEntropyLoss()(out, labels); loss.backward()`` (train/train.py:105-113) on this
        rank's batch.  Gradients ACCUMULATE into ``.grad`` like autograd's (call ``optimizer.zero_grad()`` first), then are
        averaged over the ranks.  Returns (mean loss of the local batch, logits)."""
        model = self.model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("peekvit_b200 has no CPU path: move the model to a B200 with model.cuda()")
        if images.device != dev or labels.device != dev:
            raise RuntimeError(f"images / labels must be on {dev}")
        runner._check_images(model, images)
        if images.dtype == torch.uint8:
            raise NotImplementedError("FineTuner takes float images")
        images = images.detach().to(torch.float32).contiguous()
        labels = labels.detach().to(torch.int64).contiguous()
        B = images.shape[0]
        if labels.shape != (B,):
            raise ValueError(f"labels must be a ({B},) class-index tensor, got {tuple(labels.shape)}")
        # like nn.CrossEntropyLoss: class indices in [0, C); checked on the device without a host sync
        torch._assert_async(((labels >= 0) & (labels < model.num_classes)).all())
        with torch.no_grad(), torch.cuda.device(dev):
            ops.raise_if_flagged(dev.index)
            pm = runner.packed(model)                       # rebuilt whenever optimizer.step() changed a parameter
            ws = runner.workspace(model, dev)
            self._zero_grads()
            # accumulate this step's gradients separately, add into .grad at the end (autograd semantics)
            g_head_w = torch.zeros_like(pm.head_w)
            g_head_b = torch.zeros_like(pm.head_b)
            g_cls = torch.zeros_like(pm.cls_tokens)
            loss_sum = torch.zeros(1, device=dev, dtype=torch.float32)
            logits_all = torch.empty(B, pm.num_classes, device=dev, dtype=torch.float32)
            for s in range(0, B, self.micro_batch):
                n = min(self.micro_batch, B - s)
                logits_all[s:s + n].copy_(self._micro_step(pm, ws, images[s:s + n], labels[s:s + n], 1.0 / B, loss_sum,
                                                           g_head_w, g_head_b, g_cls))
            want = {"head.weight": g_head_w, "head.bias": g_head_b, "class_tokens": g_cls.view(1, -1, pm.dim)}
            for name, p in self.params.items():
                p.grad.add_(want[name].view_as(p.grad))
            all_reduce_mean_(list(self.params.values()), self.group)
            ops.device_flag_async(dev.index)
        return loss_sum[0], logits_all

    def _micro_step(self, pm, ws, images, labels, inv_count, loss_sum, g_head_w, g_head_b, g_cls) -> torch.Tensor:
        B, seq, D, H = images.shape[0], pm.seq_len, pm.dim, pm.heads
        dh = D // H
        L = len(pm.layers)
        fwd = engine.Forward(pm, ws)
        x = fwd.embed(images)                                              # fp32 [rows, D] (workspace "x": layer 0's input)
        xs = [x]                                                            # xs[l]: what block l's LayerNorm / residual read
        saved = []
        picks = {}                                                          # l -> (kept indices, tokens per sample before the drop)
        budgets = runner._rank_budgets(self.model) if pm.family == "rankvit" else {}
        bf, f32 = torch.bfloat16, torch.float32
        for l, lw in enumerate(pm.layers):
            b_l = budgets.get(l, 1.0) if lw.kind == "rank" else 1.0
            if lw.kind == "rank" and b_l != 1 and seq > 1:
                # RankViTBlock.sort_and_drop (rankvit.py:55-77), same kernels as inference; the selection is kept for the backward
                n_tok = seq - 1
                k = min(max(math.ceil(n_tok * b_l), 0), n_tok)
                scores = ops.token_norm_score(xs[-1], B, seq, ws.get(f"ft_scores_{l}", (B, n_tok), f32))
                kept = ws.get(f"ft_kept_{l}", (B, k), torch.int32)
                if k > 0:
                    ops.topk_select(scores, k, kept)
                xs[-1] = ops.gather_rows(xs[-1], kept, B, seq, ws.get(f"ft_xg_{l}", (B * (k + 1), D), f32))
                picks[l] = (kept, seq)
                seq = k + 1
            rows = B * seq
            aw, mw = lw.attn[0], lw.mlp[0]
            F = mw.w_fc1.shape[0]
            qkv = ws.get(f"ft_qkv_{l}", (rows, 3 * D), bf, zero=True)
            att = ws.get(f"ft_att_{l}", (rows, D), bf)
            x1 = ws.get(f"ft_x1_{l}", (rows, D), f32)
            hpre = ws.get(f"ft_hpre_{l}", (rows, F), bf)
            xo = ws.get(f"ft_xo_{l}", (rows, D), f32)
            a = ops.layernorm(xs[-1], lw.ln1_w, lw.ln1_b, lw.eps, ws.get("ln", (rows, D), bf), rows=rows)
            ops.gemm(a, aw.w_qkv, aw.b_qkv, qkv, PK_EPI_BIAS_BF16)
            ops.attention(qkv, att, B, H, dh, seq_len=seq)
            ops.gemm(att, aw.w_o, aw.b_o, x1, PK_EPI_BIAS_RESID_F32, resid=xs[-1])
            a = ops.layernorm(x1, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), bf), rows=rows)
            ops.gemm(a, mw.w_fc1, mw.b_fc1, hpre, PK_EPI_BIAS_BF16)              # pre-activation kept for the backward
            hid = ops.gelu_bf16(hpre, ws.get("hid", (rows, F), bf))
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, xo, PK_EPI_BIAS_RESID_F32, resid=x1)
            saved.append((qkv, att, x1, hpre, seq))
            xs.append(xo)
        xl = xs[-1]
        n_cls = pm.n_cls
        feat = ops.cls_features(xl, B, seq, n_cls, pm.ln_w, pm.ln_b, pm.ln_eps, ws.get("ft_feat", (B, D), f32))
        logits = ops.cls_head(xl, B, seq, n_cls, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b,
                              out=ws.get("ft_logits", (B, pm.num_classes), f32))
        # ---- loss and head
        dlogits = ws.get("ft_dlogits", (B, pm.num_classes), f32)
        ops.softmax_xent(logits, labels, inv_count, loss_sum, dlogits)
        dfeat = ws.get("ft_dfeat", (B, D), f32)
        ops.head_bwd(dlogits, feat, pm.head_w, g_head_w, g_head_b, dfeat)
        # ---- final LayerNorm on the class rows (sum readout: every class row of a sample gets the same feature gradient)
        rows = B * seq
        g = ws.get("ft_g", (rows, D), f32)
        g.zero_()
        cls_rows = fwd._const(f"ft_cls_rows_{B}_{seq}_{n_cls}", lambda: (
            torch.arange(B, device=g.device, dtype=torch.int32)[:, None] * seq
            + torch.arange(n_cls, device=g.device, dtype=torch.int32)[None, :]).reshape(-1).contiguous())
        ops.layernorm_bwd(xl, dfeat, pm.ln_w, pm.ln_eps, g, B * n_cls, row_index=cls_rows, dy_div=n_cls, accumulate=False)
        # ---- blocks, last to first
        for l in range(L - 1, -1, -1):
            lw = pm.layers[l]
            F = lw.mlp[0].w_fc1.shape[0]
            qkv, att, x1, hpre, seq = saved[l]
            rows = B * seq
            gb = ws.get("ft_gb", (rows, D), bf)
            da = ws.get("ft_da", (rows, D), f32)
            w2t, w1t, wot, wqkvt = self._transposed(pm, l)
            # MLP branch: x2 = x1 + W2 gelu(W1 LN2(x1) + b1) + b2
            ops.cast_bf16(g, gb)
            dhid = ops.gemm(gb, w2t, None, ws.get("ft_dhid", (rows, F), bf), PK_EPI_BIAS_BF16)
            ops.gelu_bwd_bf16(hpre, dhid, dhid)
            ops.gemm(dhid, w1t, None, da, PK_EPI_BIAS_F32)
            ops.layernorm_bwd(x1, da, lw.ln2_w, lw.eps, g, rows)                  # g = dL/dx1
            # attention branch: x1 = x + Wo attention(Wqkv LN1(x) + b) + bo
            ops.cast_bf16(g, gb)
            datt = ops.gemm(gb, wot, None, ws.get("ft_datt", (rows, D), bf), PK_EPI_BIAS_BF16)
            dqkv = ops.attention_bwd(qkv, att, datt, ws.get("ft_dqkv", (rows, 3 * D), bf), B, H, dh, seq)
            ops.gemm(dqkv, wqkvt, None, da, PK_EPI_BIAS_F32)
            ops.layernorm_bwd(xs[l], da, lw.ln1_w, lw.eps, g, rows)               # g = dL/dx_in
            if l in picks:
                # backward of the gather in front of this block: gradient rows go back to their tokens, dropped tokens get none
                kept, seq_before = picks[l]
                g_full = ws.get("ft_g", (B * seq_before, D), f32)
                g_full.zero_()
                g = ops.scatter_rows(g, g_full, kept, B, seq_before)
                seq = seq_before
        # ---- the class tokens are rows 0 .. n_cls-1 of every sample (x0 = class token + position, vit.py:230-236,:92)
        ops.sum_token_rows(g, B, seq, 0, n_cls, g_cls)
        self.last_kept = {l: kept.clone() for l, (kept, _) in picks.items()}
        return logits
