"""Achieved HBM bandwidth of the memory-bound kernels at the BASELINE shapes (config D: RankViT on ViT-B/16, 512 images;
config C: ResidualViT-S), against MEASURED_PEAKS.json hbm_gbs.  Algorithmic bytes per unit as in SURVEY.md §8(d)/DESIGN.md §4.

    python tools/membound_probe.py [--json gpurun_out/membound.json]

DRAM counters of the same launches: PK_PROBE_ITERS=1 PK_PROBE_WARM=1 under ``ncu --metrics dram__bytes_read.sum,...`` (the L2 flush
before every timed launch is a torch kernel and is filtered out by ``-k regex:``).
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peekvit_b200 import ops

DEV = "cuda:0"
peak = 6471.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
res = {}


ITERS, WARM = int(os.environ.get("PK_PROBE_ITERS", "20")), int(os.environ.get("PK_PROBE_WARM", "3"))   # 1 / 1 under ncu


def time_us(fn, iters=ITERS, warm=WARM, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                 # > L2: the next read comes from HBM
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


def report(name, us, nbytes, note=""):
    gbs = nbytes / us / 1e3
    res[name] = dict(us=round(us, 1), algorithmic_MB=round(nbytes / 1e6, 1), GBps=round(gbs), frac_of_measured_hbm_peak=round(gbs / peak, 3), note=note)
    print(name, json.dumps(res[name]), flush=True)


flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
B, seq, D = 512, 197, 768
rows = B * seq
x = torch.randn(rows, D, device=DEV)
g, b = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
y = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
report("layernorm_bf16 (ViT-B, 512 img)", time_us(lambda: ops.layernorm(x, g, b, 1e-5, y), flush=flush), rows * D * 6)
xb = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
st = torch.empty(rows, ops.gemm_row_stat_parts(D), 2, device=DEV)
report("row_stats_cast (ViT-B, 512 img)", time_us(lambda: ops.row_stats_cast(x, xb, st), flush=flush), rows * D * 6)
sc = torch.empty(B, seq - 1, device=DEV)
report("token_norm_score (K9)", time_us(lambda: ops.token_norm_score(x, B, seq, sc), flush=flush), rows * D * 4 + B * (seq - 1) * 4)
for k in (98, 49):
    kept = torch.empty(B, k, device=DEV, dtype=torch.int32)
    report(f"topk_select k={k} of 196 (K10)", time_us(lambda: ops.topk_select(sc, k, kept)), B * (196 * 4 + k * 4),
           note="latency-bound: 0.5 MB per launch, one CTA per sample")
    out = torch.empty(B * (k + 1), D, device=DEV)
    report(f"gather_rows k={k} (K11)", time_us(lambda: ops.gather_rows(x, kept, B, seq, out), flush=flush), 2 * B * (k + 1) * D * 4)
imgs = torch.randn(B, 3, 224, 224, device=DEV)
patches = torch.empty(B * 196, 768, device=DEV, dtype=torch.bfloat16)
report("patchify (K1 im2col)", time_us(lambda: ops.patchify(imgs, 16, patches), flush=flush), B * 3 * 224 * 224 * 6)
# ResidualViT-S gate plan + compaction at keep ~0.4
Ds, n = 384, 198
cap = n + 1
xs = torch.randn(B * cap, Ds, device=DEV)
cu = (torch.arange(B + 1, device=DEV, dtype=torch.int32) * n)
mult = torch.ones(B * cap, device=DEV)
mask = torch.empty(B * cap, device=DEV); dst = torch.empty(B * cap, device=DEV, dtype=torch.int32)
smp = torch.empty(B * cap, device=DEV, dtype=torch.int32); newlen = torch.empty(B, device=DEV, dtype=torch.int32)
mdrop = torch.empty(B, device=DEV)
gw = torch.randn(Ds, device=DEV) * 0.05
btw = torch.randn(Ds, device=DEV) * 0.05
def plan():
    ops.residual_gate_plan(xs, cu, mult, B, cap, n_special=2, budget_pos=1, gated=True, gate_w=gw, gate_b=0.0, gate_temp=1.0,
                           gate_bias=0.0, gate_type=0, thr_mode=0, bt_w=btw, bt_b=0.0, thr_dev=None, mask=mask, dst_local=dst,
                           sample_of=smp, new_len=newlen, mdrop=mdrop)
report("residual_gate_plan (K12, ViT-S, 512 img)", time_us(plan, flush=flush), B * n * (Ds * 4 + 16))
cu_out = torch.empty(B + 1, device=DEV, dtype=torch.int32); total = torch.empty(1, device=DEV, dtype=torch.int32)
ops.exclusive_scan(newlen, cu_out, total)
kept_rows = int(total.item())
ys = torch.empty(B * cap, Ds, device=DEV); rs = torch.empty(B * cap, device=DEV); mo = torch.empty(B * cap, device=DEV)
def compact():
    ops.compact_rows(xs, ys, cu, cu_out, B, B * cap, dst, smp, scale_in=mask, scale_out=rs, attrs=[(mult, mo)], ghost=True)
report(f"compact_rows (K13, {kept_rows} of {B * n} rows kept)", time_us(compact, flush=flush), kept_rows * Ds * 8 + B * n * 12)
# kernels added later in the round: fp32-mode split rows, NoiseBlock, MoE routing helpers
a6 = torch.empty(rows, 6 * D, device=DEV, dtype=torch.bfloat16)
report("split3 + LayerNorm (fp32 mode, ViT-B, 512 img)", time_us(lambda: ops.split3(x, a6, ops.SPLIT_LAYERNORM, g, b, 1e-5), flush=flush), rows * D * 16)
hid32 = torch.randn(B // 4 * seq, 3072, device=DEV)
h6 = torch.empty(B // 4 * seq, 6 * 3072, device=DEV, dtype=torch.bfloat16)
report("split3 + exact GELU (fp32 mode, 128 img hidden rows)", time_us(lambda: ops.split3(hid32, h6, ops.SPLIT_GELU), flush=flush), hid32.numel() * 16)
noise = torch.randn(rows, D, device=DEV)
report("noise_snr (NoiseBlock, ViT-B, 512 img)", time_us(lambda: ops.noise_snr(x, noise, 10.0), flush=flush), rows * D * 12)
xs2 = torch.randn(rows, 384, device=DEV); ysrt = torch.randn(rows, 384, device=DEV)
perm = torch.randperm(rows, device=DEV).to(torch.int32)
report("scatter_add_rows (MoE un-permute, ViT-S, 512 img)", time_us(lambda: ops.scatter_add_rows(xs2, ysrt, perm), flush=flush), rows * 384 * 12)
gwm, gbm = torch.randn(4, 384, device=DEV), torch.zeros(4, device=DEV)
gs, bs = torch.ones(384, device=DEV), torch.zeros(384, device=DEV)
ex = torch.empty(rows, device=DEV, dtype=torch.int32); off = torch.empty(5, device=DEV, dtype=torch.int32); cnt = torch.empty(4, device=DEV, dtype=torch.int32)
src = torch.empty(rows, device=DEV, dtype=torch.int32); scr = torch.empty(ops.MOE_SORT_SCRATCH_INTS, device=DEV, dtype=torch.int32)
report("moe_route (LN + gate + arg-max + counting sort, ViT-S, 4 experts)",
       time_us(lambda: ops.moe_route(xs2, gs, bs, 1e-5, gwm, gbm, rows, ex, off, cnt, src, scratch=scr), flush=flush), rows * (384 * 4 + 8),
       note="four kernels: route, histogram, scan, scatter")
if "--json" in sys.argv:
    p = sys.argv[sys.argv.index("--json") + 1]
    os.makedirs(os.path.dirname(p), exist_ok=True)
    json.dump(dict(hbm_peak_GBps=peak, kernels=res), open(p, "w"), indent=1)
