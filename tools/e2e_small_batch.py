"""ViT-B/16 at 256 images per step (one GPU's share of BASELINE's 2048-image batch on 8 GPUs): device-resident vs the
host-resident path, for the first-micro-batch split given in PEEKVIT_B200_HOST_FIRST_SPLIT.  python tools/e2e_small_batch.py [B]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import ops, runner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev)
x = torch.randn(B, 3, 224, 224).pin_memory()
xd = x.to(dev)
def t(fn, n=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
d = t(lambda: model(xd))
h = t(lambda: model.forward_host(x))
print(f"split {runner.HOST_FIRST_SPLIT} B={B}: device {d:.2f} ms ({B / d * 1e3:.0f} img/s), host-resident {h:.2f} ms ({B / h * 1e3:.0f} img/s), ratio {d / h:.3f}, flag {ops.device_flag()}")
