"""Why forward_host is slower than the device-resident forward: ramp schedule, exposed first copy, or DMA interference."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import runner
dev = torch.device("cuda", 0)
model, sd = bench.build_model(dev)
B = 2048
images = torch.randn(B, 3, 224, 224, device=dev)
host = torch.empty(B, 3, 224, 224, dtype=torch.float32, pin_memory=True); host.copy_(images)
out_host = torch.empty(B, 1000, dtype=torch.float32, pin_memory=True)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3
print("device forward           ev/wall ms", timed(lambda: model(images)))
print("forward_host             ev/wall ms", timed(lambda: model.forward_host(host, out_host)))
def ramp():
    for s, n in ((0, 128), (128, 384), (512, 512), (1024, 512), (1536, 512)):
        model(images[s:s + n])
print("device, ramp schedule    ev/wall ms", timed(ramp))
def synced():
    model(images); torch.cuda.synchronize()
print("device + sync per step   ev/wall ms", timed(synced))
side = torch.cuda.Stream()
stage = torch.empty(512, 3, 224, 224, device=dev)
def with_dma():
    with torch.cuda.stream(side):
        for s in range(0, B, 512): stage.copy_(host[s:s + 512], non_blocking=True)
    model(images)
print("device + concurrent H2D  ev/wall ms", timed(with_dma))
print("device forward again     ev/wall ms", timed(lambda: model(images)))
