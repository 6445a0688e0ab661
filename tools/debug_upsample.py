import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build
import gecco_b200 as G
from gecco_b200 import _abi

cuda = torch.device("cuda:0")
g = torch.load(ROOT / "tests/golden/cond_uvl.pt", weights_only=False); r = g["recipe"]
feats = synth.synth_features(r["B"], r["feat_sizes"], r["feat_seed"])
model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda, feats)
K = synth.camera(r["B"], r["K"])
ctx = G.Context3d(image=torch.zeros(r["B"], 3, 8, 8, device=cuda), K=K.to(cuda))
cfg = O.OracleConfig(kind="cond", reparam="uvl", sigma_max=r["sigma_max"])
sd = synth.reparam_buffers("uvl", r["mean"], r["sigma"])
seed_cloud = O.diffusion_to_data(cfg, sd, torch.randn(r["B"], r["ups_n_seed"], 3, generator=synth.gen(r["ups_seed_cloud_seed"])), K)
rms = lambda t: t.double().pow(2).mean().sqrt().item()
to_diff = lambda d: O.data_to_diffusion(cfg, sd, d.cpu(), K.double())
for pairs in (1, 0):
    _abi.check(_abi.load().gecco_set_option(ctypes.c_char_p(b"gemm_pairs"), pairs))
    u = model.upsample(seed_cloud.to(cuda), n_new=r["ups_n_new"], context=ctx, num_substeps=r["ups_substeps"],
                       num_steps=r["ups_steps"], rng=synth.gen(r["ups_seed"]))
    ud, gd = to_diff(u), to_diff(g["upsample"])
    print("pairs", pairs, "finite", torch.isfinite(u).all().item(), "nonfinite count", (~torch.isfinite(u)).sum().item(),
          "diff-space finite", torch.isfinite(ud).all().item(), (~torch.isfinite(ud)).sum().item(), "golden finite", torch.isfinite(gd).all().item())
    ok = torch.isfinite(ud).all(dim=-1) & torch.isfinite(gd).all(dim=-1)
    print("  rel rms on finite rows", rms(ud[ok] - gd[ok]) / rms(gd[ok]), "rows", ok.sum().item(), "of", ok.numel())
    # where do non-finite values appear
    bad = ~torch.isfinite(u).all(dim=-1)
    if bad.any():
        idx = bad.nonzero()[:5]
        print("  bad rows", idx.tolist(), u[bad][:3].tolist(), "golden there", g["upsample"][bad.cpu()][:3].tolist())
