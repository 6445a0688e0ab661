import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build
import gecco_b200 as G
from gecco_b200.engine import engine_for

cuda = torch.device("cuda:0")
g = torch.load(ROOT / "tests/golden/cond_uvl.pt", weights_only=False); r = g["recipe"]
feats = synth.synth_features(r["B"], r["feat_sizes"], r["feat_seed"])
model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda, feats)
K = synth.camera(r["B"], r["K"])
ctx = G.Context3d(image=torch.zeros(r["B"], 3, 8, 8, device=cuda), K=K.to(cuda))
cfg = O.OracleConfig(kind="cond", reparam="uvl", sigma_max=r["sigma_max"])
sdf = synth.full_state_dict(r["kind"], r["reparam"], r["mean"], r["sigma"], r["weight_seed"])
rel = lambda a, b: ((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()
B = r["B"]
gen = synth.gen(77)
data_diff = torch.randn(B, 128, 3, generator=gen)
for sigma in (2.0, 0.002):
    sig = torch.full((B,), sigma)
    data_ctx = data_diff + torch.randn(B, 128, 3, generator=gen) * sigma
    D_o, hs_o = O.denoise(cfg, sdf, data_ctx, sig, feats, K, return_h=True)
    D_g, hs_g = model(data_ctx.to(cuda), sig.to(cuda), ctx, do_cache=True)
    print(f"sigma {sigma}: full pass D rel {rel(D_g.cpu(), D_o):.3e}; cache rel per layer", [f"{rel(a.cpu(), b):.2e}" for a, b in zip(hs_g, hs_o)])
    for s2 in (sigma * 1.25, 0.002):
        x2 = torch.randn(B, 200, 3, generator=gen) * max(s2, 1.0)
        sg2 = torch.full((B,), s2)
        Dc_o = O.denoise(cfg, sdf, x2, sg2, feats, K, hs=hs_o)
        Dc_g = model(x2.to(cuda), sg2.to(cuda), ctx, cache=hs_g)
        Dc_g2 = model(x2.to(cuda), sg2.to(cuda), ctx, cache=[h.to(cuda) for h in hs_o])
        print(f"   cached pass sigma {s2}: D rel {rel(Dc_g.cpu(), Dc_o):.3e} (own cache), {rel(Dc_g2.cpu(), Dc_o):.3e} (oracle cache)")
us = model.upsample(O.diffusion_to_data(cfg, synth.reparam_buffers("uvl", r["mean"], r["sigma"]), data_diff, K).to(cuda), n_new=200, context=ctx, num_substeps=2, num_steps=2, sigma_max=2.0, rng=synth.gen(9))
print("upsample finite", torch.isfinite(us).all().item())

# same host loop (the oracle's), two denoisers
sd = synth.reparam_buffers("uvl", r["mean"], r["sigma"])
seed_cloud = O.diffusion_to_data(cfg, sd, data_diff, K)
uo = O.upsample(cfg, sdf, seed_cloud, n_new=200, features=feats, K=K, seed=9, num_substeps=2, num_steps=2, sigma_max=2.0)
orig = O.denoise
def gpu_denoise(cfg_, sd_, x, sigma, features=None, K_=None, hs=None, return_h=False):
    if return_h:
        D, h = model(x.to(cuda), sigma.to(cuda), ctx, do_cache=True)
        return D.cpu(), [t.cpu() for t in h]
    if hs is not None:
        return model(x.to(cuda), sigma.to(cuda), ctx, cache=[t.to(cuda) for t in hs]).cpu()
    return model(x.to(cuda), sigma.to(cuda), ctx).cpu()
O.denoise = gpu_denoise
ug = O.upsample(cfg, sdf, seed_cloud, n_new=200, features=feats, K=K, seed=9, num_substeps=2, num_steps=2, sigma_max=2.0)
O.denoise = orig
to_diff = lambda d: O.data_to_diffusion(cfg, sd, d.cpu(), K.double())
print("oracle loop, GPU denoiser vs oracle denoiser (diffusion space):", rel(to_diff(ug), to_diff(uo)))
um = model.upsample(seed_cloud.to(cuda), n_new=200, context=ctx, num_substeps=2, num_steps=2, sigma_max=2.0, rng=synth.gen(9))
print("gecco_b200 upsample vs oracle:", rel(to_diff(um), to_diff(uo)), " vs oracle-loop+GPU denoiser:", rel(to_diff(um), to_diff(ug)))
print("data space:", rel(um.cpu(), uo), rel(ug, uo))
