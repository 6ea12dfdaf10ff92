"""A few fp32-mode ViT-B/16 forwards (for ncu launch lists): python tools/fp32_run.py [images]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, sd = bench.build_model(torch.device("cuda", 0))
model.pk_precision = "fp32"
x = torch.randn(N, 3, 224, 224, device="cuda")
for _ in range(3): model(x)
torch.cuda.synchronize()
