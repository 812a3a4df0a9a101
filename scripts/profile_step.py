#!/usr/bin/env python
"""torch.profiler breakdown of one training step (all CUDA kernels, ours and the library ones)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
tr = bench.build_trainer(dev)
x = torch.rand(16, 3, 256, 256, device=dev) * 2 - 1
for _ in range(4):
    tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
