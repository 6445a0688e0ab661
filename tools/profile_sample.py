"""A short sample_stochastic call of the bench workload (B = 64, N = 2048, num_steps = 3 -> 5 evaluations) for the ncu
launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_sample.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import gecco_b200 as G  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev)
B = bench.CLOUDS_PER_GPU
g = torch.Generator("cpu").manual_seed(123)
ctx = G.Context3d(image=torch.rand(B, 3, bench.IMAGE, bench.IMAGE, generator=g).to(dev),
                  K=torch.tensor(bench.K_CAM).expand(B, 3, 3).contiguous().to(dev))
out = model.sample_stochastic((B, bench.POINTS, 3), ctx, rng=torch.Generator(dev).manual_seed(42), num_steps=3)
torch.cuda.synchronize()
assert torch.isfinite(out).all()
