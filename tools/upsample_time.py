"""Timing of BASELINE.json config 4: conditional upsampling of a 2048-point seed to 16384 points through
Diffusion.upsample (inducer states cached once per step, conditioner once per call) with the bench model.
Per cloud: 64 x 33.27 GFLOP full evaluations @2048 + 635 x 192.37 GFLOP cached evaluations @16384 = 124.28 TFLOP
(SURVEY.md §8d).  Usage: python tools/upsample_time.py [clouds] [num_steps]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
import gecco_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
model = bench.build_model(dev)
g = torch.Generator("cpu").manual_seed(123)
ctx = G.Context3d(image=torch.rand(B, 3, bench.IMAGE, bench.IMAGE, generator=g).to(dev),
                  K=torch.tensor(bench.K_CAM).expand(B, 3, 3).contiguous().to(dev))
# an in-frustum seed: diffusion-space N(0, 1) mapped to data space
seed = model.reparam.diffusion_to_data(torch.randn(B, 2048, 3, generator=g).to(dev), ctx)
n_cached = steps * 5 + (steps - 1) * 5
flop = (steps * 33.266e9 + n_cached * 192.37e9) * B
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.upsample(seed, n_new=16384, context=ctx, seed=42, num_substeps=5, num_steps=steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
assert out.shape == (B, 16384, 3) and torch.isfinite(out).all()
print(json.dumps({"workload": "config 4: upsample 2048 -> 16384 points, 137^2 images, num_substeps 5", "clouds": B, "num_steps": steps,
                  "seconds": dt, "clouds_per_s": B / dt, "tflops": flop / dt / 1e12,
                  "evaluations": {"full@2048": steps, "cached@16384": n_cached}}))
