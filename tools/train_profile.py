"""Kernel-level breakdown of one eager training step (torch.profiler; development aid for training.py)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
import gecco_b200 as G
from gecco_b200 import training as T

dev = torch.device("cuda:0")
cfg = bench.CONFIGS[5]
model = bench.build_model(dev, cfg)
B = int(os.environ.get("CLOUDS", "32"))
g = torch.Generator("cpu").manual_seed(1)
ctx = G.Context3d(image=torch.rand(B, 3, 137, 137, generator=g).to(dev), K=torch.tensor(cfg["K"]).expand(B, 3, 3).contiguous().to(dev))
data = model.reparam.diffusion_to_data(torch.randn(B, 2048, 3, generator=g).to(dev), ctx).float()
tr = T.Trainer(model, graph=False)
for _ in range(3):
    tr.step((data, ctx))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step((data, ctx))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
