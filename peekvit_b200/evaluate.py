"""Evaluation loop of the reference (``validate/test.py:97-156``) on the CUDA path: per budget (and noise
value) one pass over the data with the top-1 count kept on the device (``pk_argmax_count``: no per-batch ``argmax`` /
torchmetrics round trip), and the cost per image taken analytically from the token counts the forward itself reports instead
of the reference's second, hooked ptflops forward (``utils/flops_count.py``).  CUDA-graph replay per micro-batch comes with
``runner.run``.
"""
from __future__ import annotations

import time
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

from . import flops, ops, runner, sharding


def _layer_token_totals(model, aux: dict, batch: int) -> Optional[list]:
    """Per encoder layer, the number of token rows the micro-batches of one forward processed: a Python number, or a device
    tensor (ResidualViT / A-ViT keep their live-row counts on the device; nothing is read back here).  None: dense."""
    fam = getattr(model, "_family", "vit")
    n_layers = sum(1 for b in model.encoder.layers if hasattr(b, "mlp"))
    if fam == "rankvit" and "seq_lens" in aux:
        lens = aux["seq_lens"]
        lens = lens[0] if lens and isinstance(lens[0], (list, tuple)) else lens      # merged micro-batches: identical lists
        return [float(n) * batch for n in lens]
    if fam in ("residualvit", "eeresidualvit", "adavit") and "rows" in aux:
        rows = aux["rows"]
        full = float(getattr(model, "seq_length", 0) + (1 if getattr(model, "num_budget_tokens", 0) else 0)) * batch
        out = []
        for i in range(n_layers):
            if isinstance(rows, dict):
                parts = rows.get(i)
            else:                                   # A-ViT: one list per micro-batch, or a flat list for a single one
                per_mb = rows if rows and isinstance(rows[0], (list, tuple)) else [rows]
                parts = [mb[i] for mb in per_mb if i < len(mb)]
            if not parts:
                out.append(full)                    # an ungated layer: counted at full length
                continue
            parts = parts if isinstance(parts, (list, tuple)) else [parts]
            out.append(torch.stack([p.reshape(()) for p in parts]).sum())
        return out
    return None


def _residual_mask_totals(model, batch: int) -> list:
    """ResidualViT in the reference's accounting (its FLOP hooks count the tokens whose gate value is non-zero): per layer,
    special tokens + tokens kept by the mask the last forward published on the block (device tensors, no read-back).  The rows
    actually computed can be fewer: dropped tokens that a later gate re-admits are identical and travel as ONE row with a
    multiplicity."""
    n_special = 1 + (1 if getattr(model, "num_budget_tokens", 0) else 0)
    full = float(getattr(model, "seq_length", 0) + n_special - 1) * batch
    out = []
    for blk in model.encoder.layers:
        if not hasattr(blk, "mlp"):
            continue
        m = getattr(blk, "mask", None)
        gated = getattr(blk, "skip", None) in ("attention", "mlp", "attention+mlp") and m is not None
        out.append((m > 0).sum() + n_special * batch if gated else full)
    return out


def tokens_per_layer(model, aux: dict, batch: int) -> Optional[list]:
    """Average number of tokens that entered each encoder layer in the forward that filled ``aux`` (None: every layer saw
    the full sequence).  RankViT reports its per-layer sequence lengths, ResidualViT / A-ViT the number of live rows."""
    tot = _layer_token_totals(model, aux, batch)
    return None if tot is None else [float(t) / batch for t in tot]


@torch.no_grad()
def _stage(batch, dev, copy_stream):
    """(images, labels) on ``dev``; host tensors are copied on ``copy_stream`` (asynchronously when pinned) and come with the
    event the compute stream has to wait for."""
    images, labels = batch
    if images.device == dev and labels.device == dev:
        return images, labels, None
    with torch.cuda.stream(copy_stream):
        images = images.to(dev, non_blocking=True)
        labels = labels.to(dev, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(copy_stream)
    return images, labels, ready


def evaluate(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], budgets: Sequence = (None,),
             noise_module=None, noise_vals: Sequence = (None,), count_flops: bool = True) -> Dict:
    """``{budget: {noise: {accuracy, images_per_second, gmacs_per_image, tokens_per_layer, computed_rows_per_layer}}}`` (noise
    level omitted when there is no noise module).  ``batches`` yields ``(images, labels)``; tensors are moved to the model's device if needed.
    Mirrors validate/test.py: ``set_budget`` per budget, ``noise_module.set_value`` per noise value, accuracy over the whole
    set, throughput from the wall clock of the pass, cost in MACs per image like ``compute_flops(..., flops_units='Mac')``."""
    dev = next(model.parameters()).device
    batches = list(batches)
    copy_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
    results: Dict = {}
    for budget in budgets:
        if budget is not None and hasattr(model, "set_budget"):
            model.set_budget(budget)
        per_noise = {}
        for nv in noise_vals:
            if noise_module is not None and nv is not None:
                noise_module.set_value(nv)
            counts = torch.zeros(2, dtype=torch.int64, device=dev)
            tok_sum, row_sum, n_img = None, None, 0
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            # host-resident batches: the next batch's copy runs on a side stream while this one computes (the reference's loop
            # copies in front of every forward, validate/test.py:117-118); device-resident batches pass through
            cur = torch.cuda.current_stream(dev)
            staged = _stage(batches[0], dev, copy_stream) if batches else None
            for bi in range(len(batches)):
                images, labels, ready = staged
                staged = _stage(batches[bi + 1], dev, copy_stream) if bi + 1 < len(batches) else None
                if ready is not None:
                    cur.wait_event(ready)
                    images.record_stream(cur)
                    labels.record_stream(cur)
                aux = {} if count_flops else None
                logits = runner.run(model, images, aux)
                logits = logits[-1] if logits.dim() == 3 else logits            # EE-ResidualViT: final head
                ops.argmax_count(logits, labels, counts)
                if count_flops:
                    tot = _layer_token_totals(model, aux, images.shape[0])
                    if tot is not None:
                        row_sum = tot if row_sum is None else [a + t for a, t in zip(row_sum, tot)]
                        if getattr(model, "_family", "") in ("residualvit", "eeresidualvit"):
                            tot = _residual_mask_totals(model, images.shape[0])
                        tok_sum = tot if tok_sum is None else [a + t for a, t in zip(tok_sum, tot)]
                n_img += images.shape[0]
            counts = sharding.reduce_counts(counts[0], counts[1])                 # sample-sharded eval: one all-reduce per pass
            correct, total = (int(v) for v in counts.tolist())                    # the pass's only device -> host read
            dt = time.perf_counter() - t0
            with torch.cuda.device(dev):
                ops.raise_if_flagged(dev.index, sync=True)                        # the read above synchronised: the watchdog word is final
            dt = sharding.max_over_ranks(dt, dev)                                 # whole-job rate: all images / slowest rank
            # ``total`` is the global image count when sharded; token statistics below stay those of this rank's shard
            entry = {"accuracy": correct / max(total, 1), "images_per_second": total / dt, "images": total}
            if count_flops:
                tpl = [float(t) / n_img for t in tok_sum] if tok_sum is not None else None
                entry["gmacs_per_image"] = flops.model_gflops_per_image(model, tpl) / 2.0
                entry["tokens_per_layer"] = tpl
                # rows the kernels actually processed (<= tokens: merged identical rows, retired A-ViT samples)
                entry["computed_rows_per_layer"] = [float(t) / n_img for t in row_sum] if row_sum is not None else None
            per_noise[nv] = entry
        results[budget] = per_noise[None] if list(noise_vals) == [None] else per_noise
    return results
