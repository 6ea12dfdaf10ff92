import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for mb in [int(a) for a in sys.argv[1:]] or (256, 512, 683, 1024, 2048, 512, 256):
    model.pk_micro_batch = mb
    print(f"mb={mb}: device {timed(lambda: model(images)):.2f} host {timed(lambda: model.forward_host(host, out_host)):.2f}", flush=True)
