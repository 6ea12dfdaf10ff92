"""Device-resident ViT-B/16 throughput: python tools/model_time.py [batch] [micro_batch] [steps]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from peekvit_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
model, sd = bench.build_model(dev)
if mb: model.pk_micro_batch = mb
images = torch.randn(B, 3, 224, 224, device=dev)
for _ in range(3): out = model(images)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): out = model(images)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f"graphs={os.environ.get('PEEKVIT_B200_CUDA_GRAPHS','1')} B={B} mb={mb} ms/step={ms:.2f} img/s={B/ms*1e3:.0f} flag={ops.device_flag()} nan={bool(torch.isnan(out).any())}")
