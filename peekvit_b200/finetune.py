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
trainable ``class_tokens`` / ``head.*``; and the gate regime of ``ResidualVisionTransformer`` in its shipped configuration
(configs/model/residualdeit_s_16_224.yaml: sigmoid gates, ``'attention+mlp'`` blocks, ``add_budget_token='learnable'``, one class
token): the training-mode forward samples one budget per image (residualvit.py:541-576), every block runs on all tokens with
its soft mask (``:197-260``), and the backward adds the mask's gradient -- four row dot products per block, the gate
projection, the budget-token gate, the learnable budget token -- to the activation gradient (csrc/pk_train.cu).  A regulariser
on the published masks (utils/losses.py, ``additional_losses`` in train/train.py:108-112) is passed as ``extra_loss``: it is
evaluated with torch autograd on the ``(B, N_img, 1)`` mask tensors and its mask gradients enter the same backward.  Dropout
must be 0 (every shipped config).  Other ResidualViT configurations (gumbel gates, fixed budgets, further skip modes), A-ViT and
MoE have no backward here: constructing a FineTuner for them raises.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import engine, ops, runner
from ._lib import PK_EPI_BIAS_BF16, PK_EPI_BIAS_F32, PK_EPI_BIAS_RESID_F32

TRAIN_WORDS = ("gate", "class", "head", "threshold", "budget")          # train/train.py:99-100
_SUPPORTED = ("class_tokens", "head.weight", "head.bias")
_GATE_SUFFIXES = (".residual_gate.projection.weight", ".residual_gate.projection.bias", ".budget_token_gate.weight",
                  ".budget_token_gate.bias")


def _supported(name: str, family: str) -> bool:
    if name in _SUPPORTED:
        return True
    return family == "residualvit" and (name == "learnable_budget_token_1" or
                                        (name.startswith("encoder.layers.") and name.endswith(_GATE_SUFFIXES)))


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
    def __init__(self, model, train_words: Sequence[str] = TRAIN_WORDS, micro_batch: int = 256, process_group=None):
        # micro_batch: images per forward + backward pass (the saved activations are 46 MB per image at ViT-B/16, 29 MB at the
        # ViT-S gate regime).  ViT-B/16 step, 512 images: 6.51k / 6.88k / 6.74k img/s at 128 / 256 / 512 (profiles/r02/run53)
        family = getattr(model, "_family", None)
        if family not in ("vit", "rankvit", "residualvit"):
            raise NotImplementedError(
                f"FineTuner: the backward path is built for VisionTransformer, RankVisionTransformer (class tokens + head) and "
                f"ResidualVisionTransformer (gate regime); {type(model).__name__} has none")
        if family == "residualvit":
            self._check_residual(model)
        else:
            for blk in model.encoder.layers:
                if type(blk).__name__ not in ("ViTBlock", "RankViTBlock"):
                    raise NotImplementedError(f"FineTuner: encoder.layers holds a {type(blk).__name__}")
        self.family = family
        drops = [m.p for m in model.modules() if isinstance(m, torch.nn.Dropout)] + \
                [m.dropout for m in model.modules() if isinstance(m, torch.nn.MultiheadAttention)]
        if any(float(p) != 0.0 for p in drops):
            raise NotImplementedError("FineTuner: dropout / attention_dropout must be 0 (every shipped config)")
        self.model = model
        self.micro_batch = int(micro_batch)
        self.group = process_group
        self.names = train_only_these_params(model, train_words)
        bad = [n for n in self.names if not _supported(n, family)]
        if bad:
            raise NotImplementedError(f"FineTuner: no backward for {bad}; supported trainable parameters: {list(_SUPPORTED)}"
                                      + (" + the gate / budget-token-gate projections and learnable_budget_token_1" if family == "residualvit" else ""))
        self.params: Dict[str, torch.nn.Parameter] = {n: p for n, p in model.named_parameters() if n in self.names}
        self._wt: Dict[int, tuple] = {}             # layer -> transposed bf16 weights, keyed by the weight pack they came from
        self._wt_pack = None
        # RankViT: the kept-token indices of the last micro-batch, {layer: int32 [B, k]} (what a caller / test needs to
        # reproduce the step given identical selections; top-k is discontinuous in the bf16 scores)
        self.last_kept: Dict[int, torch.Tensor] = {}

    @staticmethod
    def _check_residual(model) -> None:
        """The gate regime as shipped (configs/model/residualdeit_s_16_224.yaml); everything else raises."""
        if type(model).__name__ != "ResidualVisionTransformer":
            raise NotImplementedError(f"FineTuner: {type(model).__name__} (early-exit heads) has no backward here")
        if model.add_budget_token != "learnable":
            raise NotImplementedError("FineTuner: ResidualViT trains with add_budget_token='learnable' (a per-image sampled budget "
                                      f"scaling the learnable token, residualvit.py:572-576); got {model.add_budget_token!r}")
        if model.num_class_tokens != 1 or int(getattr(model, "num_registers", 0) or 0) != 0:
            raise NotImplementedError("FineTuner: ResidualViT gate regime needs one class token and no registers (the reference's "
                                      "forward mask is built for exactly one leading token, residualvit.py:230-235)")
        for blk in model.encoder.layers:
            if type(blk).__name__ != "ResidualViTBlock":
                raise NotImplementedError(f"FineTuner: encoder.layers holds a {type(blk).__name__}")
            if blk.skip in (None, "none"):
                continue
            if blk.skip != "attention+mlp" or blk.add_input:
                raise NotImplementedError(f"FineTuner: skip={blk.skip!r} add_input={blk.add_input} has no backward (the shipped "
                                          "configuration gates 'attention+mlp' without add_input)")
            g = blk.residual_gate
            if g.gate_type != "sigmoid" or isinstance(g.threshold, torch.nn.Parameter):
                raise NotImplementedError("FineTuner: the gate regime is built for sigmoid gates with the threshold taken from the "
                                          "budget token (gumbel noise / a learnable scalar threshold are not)")

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
    def forward_backward(self, images: torch.Tensor, labels: torch.Tensor, *, budgets: Optional[torch.Tensor] = None,
                         extra_loss: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``out = model(batch); loss = CrossEntropyLoss()(out, labels) [+ additional losses]; loss.backward()``
        (train/train.py:105-113) on this rank's batch.  Gradients ACCUMULATE into ``.grad`` like autograd's (call
        ``optimizer.zero_grad()`` first), then are averaged over the ranks.  Returns (loss of the local batch, logits).

        ResidualViT only: ``budgets`` -- one budget per image, (B,); default: sampled like the reference's training forward
        (``model._sample_budget``, residualvit.py:541-550) -- is left in ``model.current_budget``; ``extra_loss(model)`` returns
        a scalar torch tensor computed from the blocks' published ``.mask`` tensors (and ``model.current_budget``), e.g. the
        reference's ``MSELoss`` regulariser; it couples the images of a batch, so the step then runs as one micro-batch."""
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
            if self.family == "residualvit":
                return self._residual_step(pm, ws, images, labels, budgets, extra_loss)
            if budgets is not None or extra_loss is not None:
                raise ValueError("budgets / extra_loss belong to the ResidualViT gate regime")
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

    # ------------------------------------------------------------------ ResidualViT, gate regime
    def _residual_step(self, pm, ws, images, labels, budgets, extra_loss):
        model, dev = self.model, images.device
        B, D = images.shape[0], pm.dim
        if budgets is None:
            budgets = model._sample_budget(B)
        budgets = torch.as_tensor(budgets, dtype=torch.float32).to(dev).reshape(-1)
        if budgets.numel() == 1:
            budgets = budgets.expand(B)
        if budgets.numel() != B:
            raise ValueError(f"budgets must hold one value per image ({B}), got {budgets.numel()}")
        budgets = budgets.contiguous()
        model.current_budget = budgets                                   # residualvit.py:565-566
        self._zero_grads()
        f32 = torch.float32
        acc = {"head.weight": torch.zeros_like(pm.head_w), "head.bias": torch.zeros_like(pm.head_b),
               "class_tokens": torch.zeros_like(pm.cls_tokens), "learnable_budget_token_1": torch.zeros(D, device=dev, dtype=f32)}
        for l, blk in enumerate(model.encoder.layers):
            if blk.skip == "attention+mlp":
                for suf, n in ((".residual_gate.projection.weight", D), (".residual_gate.projection.bias", 1),
                               (".budget_token_gate.weight", D), (".budget_token_gate.bias", 1)):
                    acc[f"encoder.layers.{l}{suf}"] = torch.zeros(n, device=dev, dtype=f32)
        loss_sum = torch.zeros(1, device=dev, dtype=f32)
        logits_all = torch.empty(B, pm.num_classes, device=dev, dtype=f32)
        mb = B if extra_loss is not None else self.micro_batch
        for s in range(0, B, mb):
            n = min(mb, B - s)
            logits_all[s:s + n].copy_(self._micro_step_residual(pm, ws, images[s:s + n], labels[s:s + n], budgets[s:s + n], 1.0 / B,
                                                                loss_sum, acc, extra_loss))
        for name, p in self.params.items():
            if name in acc:                     # e.g. the budget_token_gate of an ungated block takes no part in the forward
                p.grad.add_(acc[name].view_as(p.grad))
        all_reduce_mean_(list(self.params.values()), self.group)
        ops.device_flag_async(dev.index)
        return loss_sum[0], logits_all

    def _micro_step_residual(self, pm, ws, images, labels, budgets, inv_count, loss_sum, acc, extra_loss) -> torch.Tensor:
        """Training-mode forward of ResidualVisionTransformer (residualvit.py:587-616; block :197-260) on the dense layout
        [class, budget token, image tokens] + its backward.  See csrc/pk_train.cu for the mask algebra."""
        model = self.model
        B, D, H = images.shape[0], pm.dim, pm.heads
        dh, L = D // H, len(pm.layers)
        n_special, bpos = 2, 1
        seq = pm.seq_len + 1
        n_img = seq - n_special
        rows = B * seq
        bf, f32 = torch.bfloat16, torch.float32
        fwd = engine.Forward(pm, ws)
        x0 = fwd.embed(images, shift=1)
        tok = model.learnable_budget_token_1.detach().reshape(1, D).to(f32)
        x0.view(B, seq, D)[:, bpos] = tok * budgets.view(B, 1)           # residualvit.py:572-576
        xs, saved = [x0], []
        for l, lw in enumerate(pm.layers):
            blk = lw.module
            xin = xs[-1]
            aw, mw = lw.attn[0], lw.mlp[0]
            F = mw.w_fc1.shape[0]
            gate = None
            m, mi = None, xin
            if blk.skip == "attention+mlp":
                g = blk.residual_gate
                gate = dict(w=g.projection.weight.detach().reshape(-1), b=g.projection.bias.detach().reshape(-1),
                            bt_w=blk.budget_token_gate.weight.detach().reshape(-1), bt_b=blk.budget_token_gate.bias.detach().reshape(-1),
                            temp=float(g.temp), bias=float(g.sigmoid_bias),
                            mask=ws.get(f"ft_mask_{l}", (B, n_img), f32), sig=ws.get(f"ft_sig_{l}", (B, n_img), f32),
                            thr=ws.get(f"ft_thr_{l}", (B,), f32))
                m = ws.get(f"ft_m_{l}", (rows,), f32)
                ops.residual_gate_train_fwd(xin, B, seq, n_special, bpos, gate["w"], gate["b"], gate["temp"], gate["bias"], gate["bt_w"],
                                            gate["bt_b"], m, gate["mask"], gate["sig"], gate["thr"])
                mi = ops.row_scale_add(ws.get(f"ft_mi_{l}", (rows, D), f32), xin, m, rows)          # masked input (:221-228)
            qkv = ws.get(f"ft_qkv_{l}", (rows, 3 * D), bf, zero=True)
            att = ws.get(f"ft_att_{l}", (rows, D), bf)
            x1 = ws.get(f"ft_x1_{l}", (rows, D), f32)
            hpre = ws.get(f"ft_hpre_{l}", (rows, F), bf)
            xo = ws.get(f"ft_xo_{l}", (rows, D), f32)
            a = ops.layernorm(mi, lw.ln1_w, lw.ln1_b, lw.eps, ws.get("ln", (rows, D), bf), rows=rows, rowscale=m)
            ops.gemm(a, aw.w_qkv, aw.b_qkv, qkv, PK_EPI_BIAS_BF16)
            ops.attention(qkv, att, B, H, dh, seq_len=seq)
            ops.gemm(att, aw.w_o, aw.b_o, x1, PK_EPI_BIAS_RESID_F32, resid=mi, rowscale=m)
            a = ops.layernorm(x1, lw.ln2_w, lw.ln2_b, lw.eps, ws.get("ln", (rows, D), bf), rows=rows, rowscale=m)
            ops.gemm(a, mw.w_fc1, mw.b_fc1, hpre, PK_EPI_BIAS_BF16)
            hid = ops.gelu_bf16(hpre, ws.get("hid", (rows, F), bf))
            ops.gemm(hid, mw.w_fc2, mw.b_fc2, xo, PK_EPI_BIAS_RESID_F32, resid=x1)
            saved.append((qkv, att, x1, hpre, m, mi, gate))
            xs.append(xo)
        xl = xs[-1]
        feat = ops.cls_features(xl, B, seq, 1, pm.ln_w, pm.ln_b, pm.ln_eps, ws.get("ft_feat", (B, D), f32))
        logits = ops.cls_head(xl, B, seq, 1, pm.ln_w, pm.ln_b, pm.ln_eps, pm.head_w, pm.head_b,
                              out=ws.get("ft_logits", (B, pm.num_classes), f32))
        # ---- the regulariser on the published masks (torch autograd on (B, N_img, 1) tensors; train/train.py:108-112)
        dmask_ext: Dict[int, torch.Tensor] = {}
        leaves = {}
        for l, (_, _, _, _, m, _, gate) in enumerate(saved):
            if gate is not None:
                leaf = gate["mask"].clone().view(B, n_img, 1)
                leaf.requires_grad_(extra_loss is not None)
                pm.layers[l].module.mask = leaf
                leaves[l] = leaf
        if extra_loss is not None and leaves:
            with torch.enable_grad():
                val = torch.as_tensor(extra_loss(model), device=images.device)
                if val.numel() != 1:
                    raise ValueError(f"extra_loss must return a scalar, got shape {tuple(val.shape)}")
                # a term that does not depend on the masks (e.g. weight 0) adds to the loss and has no gradient here
                grads = (torch.autograd.grad(val, list(leaves.values()), allow_unused=True) if val.requires_grad
                         else [None] * len(leaves))
            loss_sum.add_(val.detach().to(f32).reshape(1))
            for (l, leaf), gr in zip(leaves.items(), grads):
                leaf.requires_grad_(False)
                if gr is not None:
                    dmask_ext[l] = gr.detach().to(f32).reshape(B, n_img).contiguous()
        # ---- loss, head, final LayerNorm on the class row
        dlogits = ws.get("ft_dlogits", (B, pm.num_classes), f32)
        ops.softmax_xent(logits, labels, inv_count, loss_sum, dlogits)
        dfeat = ws.get("ft_dfeat", (B, D), f32)
        ops.head_bwd(dlogits, feat, pm.head_w, acc["head.weight"], acc["head.bias"], dfeat)
        g = ws.get("ft_g", (rows, D), f32)
        g.zero_()
        cls_rows = fwd._const(f"ft_cls_rows_{B}_{seq}_1", lambda: (torch.arange(B, device=g.device, dtype=torch.int32) * seq).contiguous())
        ops.layernorm_bwd(xl, dfeat, pm.ln_w, pm.ln_eps, g, B, row_index=cls_rows, dy_div=1, accumulate=False)
        dm = ws.get("ft_dm", (rows,), f32)
        # ---- blocks, last to first
        for l in range(L - 1, -1, -1):
            lw = pm.layers[l]
            F = lw.mlp[0].w_fc1.shape[0]
            qkv, att, x1, hpre, m, mi, gate = saved[l]
            xin = xs[l]
            gb = ws.get("ft_gb", (rows, D), bf)
            da = ws.get("ft_da", (rows, D), f32)
            w2t, w1t, wot, wqkvt = self._transposed(pm, l)
            # MLP branch: out = x1 + W2 gelu(W1 (m * LN2(x1)) + b1) + b2
            ops.cast_bf16(g, gb)
            dhid = ops.gemm(gb, w2t, None, ws.get("ft_dhid", (rows, F), bf), PK_EPI_BIAS_BF16)
            ops.gelu_bwd_bf16(hpre, dhid, dhid)
            ops.gemm(dhid, w1t, None, da, PK_EPI_BIAS_F32)
            if gate is None:
                ops.layernorm_bwd(x1, da, lw.ln2_w, lw.eps, g, rows)
                ops.cast_bf16(g, gb)
            else:
                dm.zero_()
                ops.layernorm_bwd_gated(x1, da, lw.ln2_w, lw.ln2_b, lw.eps, g, rows, m, dm)      # g = dL/dx1; dm += dy . LN2(x1)
                # x1 = mi + m * proj: dm += g . proj with proj = (x1 - mi) / m on the rows the relu lets through
                ops.rowdot(g, x1, dm, rows, c=mi, div=m)
                ops.cast_rows_bf16(g, gb, m, rows)
            datt = ops.gemm(gb, wot, None, ws.get("ft_datt", (rows, D), bf), PK_EPI_BIAS_BF16)
            dqkv = ops.attention_bwd(qkv, att, datt, ws.get("ft_dqkv", (rows, 3 * D), bf), B, H, dh, seq)
            ops.gemm(dqkv, wqkvt, None, da, PK_EPI_BIAS_F32)
            if gate is None:
                ops.layernorm_bwd(xin, da, lw.ln1_w, lw.eps, g, rows)
                continue
            ops.layernorm_bwd_gated(mi, da, lw.ln1_w, lw.ln1_b, lw.eps, g, rows, m, dm)           # g = dL/dmi; dm += da . LN1(mi)
            ops.rowdot(g, xin, dm, rows)                                                         # mi = m * x: dm += g . x
            ops.row_scale_add(g, g, m, rows)                                                     # g = dL/dx through the masking
            pre = f"encoder.layers.{l}"
            ops.residual_gate_train_bwd(xin, dm, dmask_ext.get(l), gate["mask"], gate["sig"], gate["thr"], B, seq, n_special, bpos,
                                        gate["w"], gate["temp"], gate["bt_w"], g, acc[pre + ".residual_gate.projection.weight"],
                                        acc[pre + ".residual_gate.projection.bias"], acc[pre + ".budget_token_gate.weight"],
                                        acc[pre + ".budget_token_gate.bias"])
        # ---- inputs of the encoder: the class token is row 0 of every sample, the budget token row 1 = budget_b * learnable token
        ops.sum_token_rows(g, B, seq, 0, 1, acc["class_tokens"])
        acc["learnable_budget_token_1"].add_((g.view(B, seq, D)[:, bpos] * budgets.view(B, 1)).sum(0))
        return logits
