import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import synth
from tests.models_b200 import build
from tests.test_denoiser_gpu import load, rms
cuda = torch.device("cuda:0")
g, r = load("uncond.pt")
model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda)
x = torch.randn(r["B"], r["N"], 3, generator=synth.gen(r["x_seed"])) * r["x_scale"]
sig = r["noise_sigma"]
runs = []
for _ in range(6):
    D, hs = model(x.to(cuda), sig.to(cuda), None, do_cache=True)
    runs.append([h.clone() for h in hs] + [D.clone()])
torch.cuda.synchronize()
base = runs[-1]
for i, run in enumerate(runs[:-1]):
    msg = []
    for l, (a, b) in enumerate(zip(run, base)):
        d = (a - b).abs()
        if d.max() > 0:
            if l < 6:
                bad = (d > 0)
                heads = sorted(set((bad.nonzero()[:, 2] // 48).tolist()))
                clouds = sorted(set(bad.nonzero()[:, 0].tolist()))
                inds = bad.any(dim=2).sum().item()
                msg.append(f"h{l}: max {d.max().item():.2e} clouds {clouds} heads(ch//48) {heads} inducers {inds} elems {int(bad.sum())}")
            else:
                msg.append(f"D: max {d.max().item():.2e}")
            break
    print(f"run {i} vs run 5:", msg if msg else "identical")
