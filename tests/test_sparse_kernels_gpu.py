"""GPU parity of the ragged-batch / sparsification kernels against plain torch expressions of the
reference lines they replace (index outputs bit-exact, fp32 outputs 1e-5)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from peekvit_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("n", [1, 5, 1024, 1025, 3000])
def test_exclusive_scan(ops, n):
    lens = torch.randint(0, 300, (n,), device=DEV, dtype=torch.int32)
    cu = torch.empty(n + 1, device=DEV, dtype=torch.int32)
    tot = torch.empty(1, device=DEV, dtype=torch.int32)
    ops.exclusive_scan(lens, cu, tot)
    exp = torch.cat([torch.zeros(1, device=DEV, dtype=torch.int64), lens.long().cumsum(0)])
    assert torch.equal(cu.long(), exp) and int(tot) == int(exp[-1])


def _ragged(B, max_len, D, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    lens = torch.randint(3, max_len + 1, (B,), device=DEV, generator=g)
    cu = torch.cat([torch.zeros(1, device=DEV, dtype=torch.int64), lens.cumsum(0)]).to(torch.int32)
    rows = int(cu[-1])
    x = torch.randn(rows, D, device=DEV, generator=g)
    return lens, cu, rows, x


@pytest.mark.parametrize("thr_mode", [0, 1, 2])
@pytest.mark.parametrize("gate_type", [0, 1])
def test_residual_gate_plan_and_compact(ops, thr_mode, gate_type):
    """mask = relu(sigmoid((w.x+b)/temp + bias) - thr) (residualvit.py:54-69), keep = mask>0 & mult>0,
    compaction = mask*row for kept rows + ghost slot; M_drop = multiplicity of the dropped rows."""
    B, D, max_len = 9, 128, 70
    lens, cu, rows, x = _ragged(B, max_len, D, seed=thr_mode * 2 + gate_type)
    mult = torch.randint(0, 4, (rows,), device=DEV).float()
    gw, btw = torch.randn(D, device=DEV) / math.sqrt(D), torch.randn(D, device=DEV) / math.sqrt(D)
    gb, temp, gbias, btb = 0.1, 0.7, -0.2, 0.05
    thr_dev = torch.tensor([0.45], device=DEV)
    mask = torch.empty(rows, device=DEV)
    dst = torch.empty(rows, device=DEV, dtype=torch.int32)
    sample_of = torch.empty(rows, device=DEV, dtype=torch.int32)
    new_len = torch.empty(B, device=DEV, dtype=torch.int32)
    mdrop = torch.empty(B, device=DEV)
    ops.residual_gate_plan(x, cu, mult, B, max_len, n_special=2, budget_pos=1, gated=True, gate_w=gw, gate_b=gb, gate_temp=temp,
                           gate_bias=gbias, gate_type=gate_type, thr_mode=thr_mode, bt_w=btw, bt_b=btb, thr_dev=thr_dev, thr_const=0.5,
                           mask=mask, dst_local=dst, sample_of=sample_of, new_len=new_len, mdrop=mdrop)
    # torch restatement
    logit = x @ gw + gb
    exp_mask = torch.ones(rows, device=DEV)
    exp_keep = torch.ones(rows, device=DEV, dtype=torch.bool)
    for b in range(B):
        s, e = int(cu[b]), int(cu[b + 1])
        thr = {0: torch.sigmoid(x[s + 1] @ btw + btb), 1: thr_dev[0], 2: torch.tensor(0.5, device=DEV)}[thr_mode]
        m = torch.relu(torch.sigmoid(logit[s + 2:e] / temp + gbias) - thr) if gate_type == 0 else torch.round(torch.sigmoid(logit[s + 2:e]))
        exp_mask[s + 2:e] = m
        exp_keep[s + 2:e] = (m > 0) & (mult[s + 2:e] > 0)
    near_tie = (exp_mask.abs() < 1e-6) & (mask.abs() < 1e-6)
    assert torch.allclose(mask, exp_mask, atol=1e-6)
    keep = dst >= 0
    assert torch.equal(keep | near_tie, exp_keep | near_tie)
    for b in range(B):
        s, e = int(cu[b]), int(cu[b + 1])
        k = keep[s:e]
        assert int(new_len[b]) == int(k.sum()) + 1
        assert torch.equal(dst[s:e][k].long(), torch.arange(int(k.sum()), device=DEV))
        assert abs(float(mdrop[b]) - float(mult[s:e][~k].sum())) < 1e-4
        assert (sample_of[s:e] == b).all()
    cu_out = torch.empty(B + 1, device=DEV, dtype=torch.int32)
    tot = torch.empty(1, device=DEV, dtype=torch.int32)
    ops.exclusive_scan(new_len, cu_out, tot)
    y = torch.full((int(tot), D), float("nan"), device=DEV)
    mult_out = torch.full((int(tot),), -1.0, device=DEV)
    scale_out = torch.full((int(tot),), -1.0, device=DEV)
    ops.compact_rows(x, y, cu, cu_out, B, rows, dst, sample_of, scale_in=mask, scale_out=scale_out, attrs=[(mult, mult_out)], ghost=True)
    for b in range(B):
        s, e, so, eo = int(cu[b]), int(cu[b + 1]), int(cu_out[b]), int(cu_out[b + 1])
        k = keep[s:e]
        assert torch.equal(y[so:eo - 1], mask[s:e][k][:, None] * x[s:e][k])
        assert torch.equal(mult_out[so:eo - 1], mult[s:e][k]) and torch.equal(scale_out[so:eo - 1], mask[s:e][k])
        assert (y[eo - 1] == 0).all() and mult_out[eo - 1] == 0 and scale_out[eo - 1] == 1       # ghost slot
    mlp0 = torch.randn(D, device=DEV)
    ops.residual_ghost(y, mult_out, cu_out, mdrop, mlp0, B)
    for b in range(B):
        eo = int(cu_out[b + 1])
        assert torch.equal(y[eo - 1], mlp0) and mult_out[eo - 1] == mdrop[b]


def test_residual_publish_tracks_tokens(ops):
    B, n_img = 2, 5
    # packed rows: sample 0 = [cls, bud, t0..t4] (7 rows), sample 1 likewise
    tok_row = (torch.arange(B, device=DEV, dtype=torch.int32)[:, None] * 7 + 2 + torch.arange(n_img, device=DEV, dtype=torch.int32)).contiguous()
    mask = torch.arange(14, device=DEV).float() / 10
    dst = torch.tensor([0, 1, 2, -1, 3, -1, -1, 0, 1, -1, -1, -1, -1, -1], device=DEV, dtype=torch.int32)
    cu_out = torch.tensor([0, 5, 8], device=DEV, dtype=torch.int32)      # 4 kept + ghost, 2 kept + ghost
    pub = torch.empty(B, n_img, 1, device=DEV)
    ops.residual_publish(mask, dst, cu_out, tok_row, pub, B, n_img)
    assert pub.view(-1).tolist() == pytest.approx([0.2, 0.3, 0.4, 0.5, 0.6, 0.9, 1.0, 1.1, 1.2, 1.3])
    assert tok_row.tolist() == [[2, 4, 3, 4, 4], [7, 7, 7, 7, 7]]        # dropped tokens -> ghost row (last of the sample)


def test_compact_rows_fused_publish_matches_separate_kernel(ops):
    """pk_compact_rows with pub_* set = pk_compact_rows + pk_residual_publish (block.mask, utils/utils.py:100-122)."""
    g = torch.Generator(device=DEV).manual_seed(11)
    B, n_special, n_img, D = 7, 2, 37, 64
    seq, cap = n_special + n_img, n_special + n_img + 1
    rows = B * seq
    x = torch.randn(B * cap, D, device=DEV, generator=g)
    cu = torch.arange(B + 1, device=DEV, dtype=torch.int32) * seq
    mask = torch.rand(B * cap, device=DEV, generator=g)
    keep = torch.rand(rows, device=DEV, generator=g) > 0.5
    keep.view(B, seq)[:, :n_special] = True
    mask[:rows][~keep] = 0.0
    pos = keep.view(B, seq).int().cumsum(1) - 1
    dst = torch.where(keep.view(B, seq), pos, torch.full_like(pos, -1)).reshape(-1).to(torch.int32).contiguous()
    new_len = (keep.view(B, seq).sum(1) + 1).to(torch.int32)             # + ghost slot
    cu_out = torch.zeros(B + 1, device=DEV, dtype=torch.int32)
    cu_out[1:] = new_len.cumsum(0)
    smp = torch.arange(B, device=DEV, dtype=torch.int32).repeat_interleave(seq).contiguous()
    mult = torch.ones(B * cap, device=DEV)
    tok0 = (torch.arange(B, device=DEV, dtype=torch.int32)[:, None] * seq + n_special
            + torch.arange(n_img, device=DEV, dtype=torch.int32)[None, :]).contiguous()
    outs = []
    for fused in (False, True):
        y = torch.full((B * cap, D), 7.0, device=DEV); rs = torch.zeros(B * cap, device=DEV); mo = torch.zeros(B * cap, device=DEV)
        tok, pub = tok0.clone(), torch.empty(B, n_img, 1, device=DEV)
        ops.compact_rows(x, y, cu, cu_out, B, B * cap, dst, smp, scale_in=mask, scale_out=rs, attrs=[(mult, mo)], ghost=True,
                         publish=(tok, pub, n_img) if fused else None)
        if not fused:
            ops.residual_publish(mask, dst, cu_out, tok, pub, B, n_img)
        outs.append((y, rs, mo, tok, pub))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert torch.equal(outs[1][4].view(B, n_img), mask[:rows].view(B, seq)[:, n_special:])


@pytest.mark.parametrize("last_layer,early_exit", [(False, True), (False, False), (True, True)])
def test_avit_halt_plan(ops, last_layer, early_exit):
    """adavit.py:186-210 on packed active rows."""
    B, D, seq = 6, 64, 20
    g = torch.Generator(device=DEV).manual_seed(7)
    lens = torch.tensor([20, 7, 1, 12, 0, 20], device=DEV)
    cu = torch.cat([torch.zeros(1, device=DEV, dtype=torch.int64), lens.cumsum(0)]).to(torch.int32)
    rows = int(cu[-1])
    x = torch.randn(rows, D, device=DEV, generator=g)
    c = torch.rand(rows, device=DEV, generator=g) * 0.9
    R = 1 - c
    tok = torch.cat([torch.arange(int(n), device=DEV) if b % 2 == 0 else torch.arange(1, int(n) + 1, device=DEV)
                     for b, n in enumerate(lens)]).float()          # odd samples have already lost their class token
    scale, center, eps = 3.0, -0.3, 0.01
    c0, R0 = c.clone(), R.clone()
    out_acc = torch.zeros(B, D, device=DEV)
    rho, counter = torch.zeros(B, seq + 1, device=DEV), torch.ones(B, seq + 1, device=DEV)
    dst = torch.empty(rows, device=DEV, dtype=torch.int32)
    sample_of = torch.empty(rows, device=DEV, dtype=torch.int32)
    new_len = torch.empty(B, device=DEV, dtype=torch.int32)
    n_halted = torch.empty(B, device=DEV)
    ops.avit_halt_plan(x, cu, B, seq + 1, c, R, tok, gate_scale=scale, gate_center=center, eps=eps, last_layer=last_layer,
                       early_exit=early_exit, out_acc=out_acc, rho=rho, counter=counter, dst_local=dst, sample_of=sample_of,
                       new_len=new_len, n_halted=n_halted)
    h = torch.ones(rows, device=DEV) if last_layer else torch.sigmoid(x[:, 0] * scale - center)
    c_new = c0 + h
    reached, not_reached = c_new > 1 - eps, c_new < 1 - eps
    w = torch.where(reached, R0, torch.where(not_reached, h, torch.zeros_like(h)))
    assert torch.allclose(c, c_new, atol=1e-6) and torch.allclose(R, torch.where(not_reached, R0 - h, R0), atol=1e-6)
    for b in range(B):
        s, e = int(cu[b]), int(cu[b + 1])
        has_cls = e > s and tok[s] == 0
        exp_acc = w[s] * x[s] if has_cls else torch.zeros(D, device=DEV)
        assert torch.allclose(out_acc[b], exp_acc, atol=1e-5)
        keep = not_reached[s:e].clone()
        if early_exit and not (has_cls and bool(not_reached[s])):
            keep[:] = False
        assert torch.equal(dst[s:e] >= 0, keep) and int(new_len[b]) == int(keep.sum())
        assert float(n_halted[b]) == seq + 1 - int(keep.sum())
        t = tok[s:e].long()
        assert torch.allclose(rho[b, t], 1 + torch.where(reached[s:e], R0[s:e], torch.zeros_like(h[s:e])), atol=1e-6)
        assert torch.equal(counter[b, t], 1 + not_reached[s:e].float())


@pytest.mark.parametrize("E,D", [(4, 384), (2, 128), (8, 768)])
def test_moe_route_and_grouped_mlp(ops, E, D):
    """Routing = argmax(Linear(LN2(x))) (moevit.py:23-32, blocks.py:23-25); grouped fc1/fc2 over the
    expert-sorted rows == every expert dense + one-hot select (moevit.py:54-59)."""
    from peekvit_b200._lib import PK_EPI_BIAS_GELU_BF16, PK_EPI_BIAS_RESID_F32
    rows, F = 2500, 2 * D
    g = torch.Generator(device=DEV).manual_seed(E)
    x = torch.randn(rows, D, device=DEV, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(D, device=DEV, generator=g), 0.1 * torch.randn(D, device=DEV, generator=g)
    gw, gb = torch.randn(E, D, device=DEV, generator=g) * 0.3, torch.randn(E, device=DEV, generator=g) * 0.1
    expert = torch.empty(rows, device=DEV, dtype=torch.int32)
    offsets = torch.empty(E + 1, device=DEV, dtype=torch.int32)
    counts = torch.empty(E, device=DEV, dtype=torch.int32)
    src_of = torch.empty(rows, device=DEV, dtype=torch.int32)
    ops.moe_route(x, gamma, beta, 1e-5, gw, gb, rows, expert, offsets, counts, src_of)
    ln = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5)
    scores = ln @ gw.t() + gb
    top2 = scores.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4                  # ignore fp32-summation-order near-ties
    assert torch.equal(expert.long()[clear], scores.argmax(-1)[clear])
    assert torch.equal(counts.long(), torch.bincount(expert.long(), minlength=E))
    assert torch.equal(offsets.long(), torch.cat([torch.zeros(1, device=DEV, dtype=torch.int64), counts.long().cumsum(0)]))
    exp_src = torch.argsort(expert.long(), stable=True)       # stable counting sort
    assert torch.equal(src_of.long(), exp_src)
    # grouped MLP
    w1 = [(torch.randn(F, D, device=DEV, generator=g) / math.sqrt(D)).to(torch.bfloat16) for _ in range(E)]
    w2 = [(torch.randn(D, F, device=DEV, generator=g) / math.sqrt(F)).to(torch.bfloat16) for _ in range(E)]
    b1 = [torch.randn(F, device=DEV, generator=g) * 0.1 for _ in range(E)]
    b2 = [torch.randn(D, device=DEV, generator=g) * 0.1 for _ in range(E)]
    a = ops.layernorm(x, gamma, beta, 1e-5, row_index=src_of)
    hid = torch.zeros(rows, F, device=DEV, dtype=torch.bfloat16)
    lnb = ln.to(torch.bfloat16).float()
    ref = x.clone()
    for e in range(E):
        sel = expert.long() == e
        h = torch.nn.functional.gelu(lnb[sel] @ w1[e].float().t() + b1[e]).to(torch.bfloat16).float()
        ref[sel] += h @ w2[e].float().t() + b2[e]
    # the un-permuting fc2 on the CTA-pair kernel (row-indexed reductions at L2; what the model runs) and on the single-CTA one
    for cta_pair in (2, 1):
        y = x.clone()
        for e in range(E):
            ops.gemm(a, w1[e], b1[e], hid, PK_EPI_BIAS_GELU_BF16, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1])
            ops.gemm(hid, w2[e], b2[e], y, PK_EPI_BIAS_RESID_F32, resid=y, m_dev=counts[e:e + 1], row_begin_dev=offsets[e:e + 1],
                     out_row_index=src_of, cta_pair=cta_pair)
        assert ops.device_flag() == 0
        assert ((y - ref).abs().max() / ref.abs().max()).item() < 1e-2
    # the same as ONE grouped launch per projection (stacked weights, the tile scheduler walks the expert segments)
    y = x.clone()
    hid.fill_(float("nan"))
    ops.gemm(a, torch.cat(w1), torch.cat(b1), hid, PK_EPI_BIAS_GELU_BF16, group_offsets=offsets, n_groups=E, cta_pair=2)
    ops.gemm(hid, torch.cat(w2), torch.cat(b2), y, PK_EPI_BIAS_RESID_F32, resid=y, out_row_index=src_of, group_offsets=offsets,
             n_groups=E, cta_pair=2)
    assert ops.device_flag() == 0
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 1e-2
