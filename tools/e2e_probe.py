"""Where the end-to-end time goes: device-resident forward vs forward_host vs raw H2D bandwidth."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
model, sd = bench.build_model(dev)
B = 2048
images = torch.randn(B, 3, 224, 224, device=dev)
host = torch.empty(B, 3, 224, 224, dtype=torch.float32, pin_memory=True); host.copy_(images)
out_host = torch.empty(B, 1000, dtype=torch.float32, pin_memory=True)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("device forward ms", timed(lambda: model(images)))
print("forward_host ms  ", timed(lambda: model.forward_host(host, out_host)))
stage = torch.empty(256, 3, 224, 224, device=dev)
def h2d_all():
    for s in range(0, B, 256): stage.copy_(host[s:s + 256], non_blocking=True)
print("H2D 1.23 GB ms   ", timed(h2d_all), "-> GB/s", 1.233 / (timed(h2d_all) * 1e-3))
for mb in (128, 256, 512):
    model.pk_micro_batch = mb
    print(f"mb={mb}: device {timed(lambda: model(images)):.2f} host {timed(lambda: model.forward_host(host, out_host)):.2f}")
