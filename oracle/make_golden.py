"""Generates tests/golden/*.pt from the UNMODIFIED reference package.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference is not available on the GPU box):

    python oracle/make_golden.py

It imports /root/reference/gecco-torch/src/gecco_torch through the import stubs in oracle/stubs/
(lightning, kornia, h5py, imageio are not installed here; kornia's two projection functions are
restated in the stub — see oracle/stubs/README.md), builds the reference modules with the reference
constructors, overwrites their parameters with the seeded synthetic weights of tests/synth.py, runs the
reference forward / sample_stochastic / upsample on seeded synthetic inputs on CPU in fp32, and stores
the recipe + outputs.  Weights and inputs are NOT stored: tests regenerate them from the same seeds.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "stubs"))
sys.path.insert(0, "/root/reference/gecco-torch/src")
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from gecco_torch.diffusion import Diffusion, EDMLoss, EDMPrecond, IdleConditioner, LogUniformSchedule  # noqa: E402
from gecco_torch.models.activation import GaussianActivation  # noqa: E402
from gecco_torch.models.feature_pyramid import FeaturePyramidContext  # noqa: E402
from gecco_torch.models.linear_lift import LinearLift  # noqa: E402
from gecco_torch.models.ray import RayNetwork  # noqa: E402
from gecco_torch.models.set_transformer import SetTransformer  # noqa: E402
from gecco_torch.reparam import GaussianReparam, UVLReparam  # noqa: E402
from gecco_torch.structs import Context3d  # noqa: E402

from tests import synth  # noqa: E402

OUT = ROOT / "tests" / "golden"
WEIGHT_SEED = 1234


class FixedConditioner(torch.nn.Module):
    """Stands in for ConvNeXtExtractor (models/feature_pyramid.py:28-73): returns a fixed synthetic pyramid.
    The conditioner is upstream of the hot path; the path under test starts at its output."""

    def __init__(self, features):
        super().__init__()
        self.features = features

    def forward(self, raw_ctx):
        return FeaturePyramidContext(features=self.features, K=raw_ctx.K)


def build_reference(kind: str, reparam: str, mean, sigma, sigma_max: float, features=None):
    st = SetTransformer(n_layers=synth.N_LAYERS, num_inducers=synth.NUM_INDUCERS, feature_dim=synth.FEATURE_DIM,
                        t_embed_dim=1, num_heads=synth.NUM_HEADS, activation=GaussianActivation)
    m, s = torch.tensor(mean), torch.tensor(sigma)
    rp = GaussianReparam(m, s) if reparam == "gaussian" else UVLReparam(m, s)
    if kind == "uncond":
        net = LinearLift(inner=st, feature_dim=synth.FEATURE_DIM)
        cond = IdleConditioner()
    else:
        net = RayNetwork(backbone=st, reparam=rp, context_dims=synth.CONTEXT_DIMS)
        cond = FixedConditioner(features)
    model = Diffusion(backbone=EDMPrecond(model=net), conditioner=cond, reparam=rp,
                      loss=EDMLoss(schedule=LogUniformSchedule(max=sigma_max)))
    sd = synth.full_state_dict(kind, reparam, mean, sigma, WEIGHT_SEED)
    ref_sd = model.state_dict()
    # the synthetic schema must be exactly the reference's learnable schema
    assert set(ref_sd) == set(sd), (sorted(set(ref_sd) ^ set(sd)))
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    model.load_state_dict(sd)
    return model.eval()


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()


def bf16_drift(fn) -> torch.Tensor:
    """Runs `fn` under the reference's own mixed precision (torch CPU autocast, bfloat16): the drift of the UNMODIFIED
    reference at the precision the CUDA path computes in.  Stored next to the fp32 goldens to calibrate tolerances."""
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        return fn()


def sub(t: torch.Tensor) -> torch.Tensor:
    """strided sub-sample of an inducer state [B, 64, C] (keeps fixtures small)"""
    return t[:, ::8, ::8].contiguous()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    OUT.mkdir(parents=True, exist_ok=True)
    recipe_common = dict(weight_seed=WEIGHT_SEED, feature_dim=synth.FEATURE_DIM, n_layers=synth.N_LAYERS,
                         num_heads=synth.NUM_HEADS, num_inducers=synth.NUM_INDUCERS, torch=torch.__version__)

    # ---------------------------------------------------------------- unconditional (config 1)
    kind, rp = "uncond", "gaussian"
    cfg = dict(kind=kind, reparam=rp, **synth.UNCOND_REPARAM, sigma_max=165.0)
    model = build_reference(kind, rp, cfg["mean"], cfg["sigma"], cfg["sigma_max"])
    B, N = 2, 300
    x = torch.randn(B, N, 3, generator=synth.gen(11)) * 3
    sig = torch.tensor([0.05, 7.0])
    with torch.no_grad():
        D, hs = model(x, sig, None, do_cache=True)
        D_plain = model(x, sig, None)
        assert torch.equal(D, D_plain)
        x2 = torch.randn(B, 200, 3, generator=synth.gen(12)) * 3
        D_cached = model(x2, sig, None, cache=hs)
        samp = model.sample_stochastic((2, 256, 3), None, rng=synth.gen(42), num_steps=6)
        ts = model.t_steps(64, 165.0, 0.002, 7)
    drift = dict(D=rel_rms(bf16_drift(lambda: model(x, sig, None)).float(), D),
                 sample=rel_rms(bf16_drift(lambda: model.sample_stochastic((2, 256, 3), None, rng=synth.gen(42), num_steps=6)), samp))
    print("uncond bf16-autocast drift of the reference", drift)
    torch.save(dict(drift=drift, recipe={**recipe_common, **cfg, "x_seed": 11, "x_scale": 3.0, "B": B, "N": N, "noise_sigma": sig,
                            "x2_seed": 12, "N2": 200, "sample_shape": (2, 256, 3), "sample_seed": 42, "sample_steps": 6},
                    D=D, hs_sub=[sub(h) for h in hs], D_cached=D_cached, sample=samp, t_steps=ts),
               OUT / "uncond.pt")
    print("uncond", D.abs().mean().item(), samp.abs().mean().item())

    # ---------------------------------------------------------------- conditional, GaussianReparam (config 2)
    kind, rp = "cond", "gaussian"
    cfg = dict(kind=kind, reparam=rp, **synth.SHAPENET_VOL_REPARAM, sigma_max=165.0)
    B, N = 2, 333
    feats = synth.synth_features(B, (34, 17, 8), 21)
    K = synth.camera(B, synth.K_SHAPENET)
    ctx = Context3d(image=torch.zeros(B, 3, 137, 137), K=K)
    model = build_reference(kind, rp, cfg["mean"], cfg["sigma"], cfg["sigma_max"], feats)
    x = torch.randn(B, N, 3, generator=synth.gen(22)) * 2
    sig = torch.tensor([0.3, 40.0])
    with torch.no_grad():
        D = model(x, sig, ctx)
        c_in = 1 / (1 + sig**2).sqrt()
        look = model.backbone.model.extract_image_features(x * c_in[:, None, None], feats, ctx)
        samp = model.sample_stochastic((2, 200, 3), ctx, rng=synth.gen(43), num_steps=5)
        samp2 = model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2)  # short trajectory
    drift = dict(D=rel_rms(bf16_drift(lambda: model(x, sig, ctx)).float(), D),
                 sample=rel_rms(bf16_drift(lambda: model.sample_stochastic((2, 200, 3), ctx, rng=synth.gen(43), num_steps=5)), samp),
                 sample2=rel_rms(bf16_drift(lambda: model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2)), samp2))
    print("cond_gaussian bf16-autocast drift of the reference", drift)
    torch.save(dict(drift=drift, recipe={**recipe_common, **cfg, "K": synth.K_SHAPENET, "feat_sizes": (34, 17, 8), "feat_seed": 21,
                            "x_seed": 22, "x_scale": 2.0, "B": B, "N": N, "noise_sigma": sig,
                            "sample_shape": (2, 200, 3), "sample_seed": 43, "sample_steps": 5},
                    D=D, lookup_sub=look[:, ::3].contiguous(), sample=samp, sample2=samp2),
               OUT / "cond_gaussian.pt")
    print("cond_gaussian", D.abs().mean().item(), look.abs().mean().item(), samp.abs().mean().item())

    # ---------------------------------------------------------------- conditional, UVLReparam (configs 3, 4)
    kind, rp = "cond", "uvl"
    cfg = dict(kind=kind, reparam=rp, **synth.UVL_REPARAM, sigma_max=180.0)
    B, N = 2, 256
    feats = synth.synth_features(B, (64, 32, 16), 31)
    K = synth.camera(B, synth.K_TASKONOMY)
    ctx = Context3d(image=torch.zeros(B, 3, 256, 256), K=K)
    model = build_reference(kind, rp, cfg["mean"], cfg["sigma"], cfg["sigma_max"], feats)
    x = torch.randn(B, N, 3, generator=synth.gen(32)) * 1.5
    sig = torch.tensor([0.002, 180.0])
    with torch.no_grad():
        D, hs = model(x, sig, ctx, do_cache=True)
        c_in = 1 / (1 + sig**2).sqrt()
        look = model.backbone.model.extract_image_features(x * c_in[:, None, None], feats, ctx)
        x2 = torch.randn(B, 500, 3, generator=synth.gen(33)) * 1.5
        D_cached = model(x2, sig, ctx, cache=hs)
        samp = model.sample_stochastic((2, 192, 3), ctx, rng=synth.gen(44), num_steps=5)
        # short trajectory (sigma_max lowered so that exp() of the ray-length coordinate stays finite after one giant step)
        samp2 = model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2, sigma_max=2.0)
        # reparam round trip on in-frustum data
        diff = torch.randn(B, 64, 3, generator=synth.gen(34))
        data = model.reparam.diffusion_to_data(diff, ctx)
        back = model.reparam.data_to_diffusion(data, ctx)
        # upsample (diffusion.py:354-470): seed cloud in data space, in frustum
        seed_cloud = model.reparam.diffusion_to_data(torch.randn(B, 128, 3, generator=synth.gen(35)), ctx)
        ups = model.upsample(seed_cloud, n_new=384, context=ctx, seed=7, num_substeps=2, num_steps=3)
    to_diff = lambda d: model.reparam.data_to_diffusion(d, Context3d(image=ctx.image, K=K.double()))
    drift = dict(D=rel_rms(bf16_drift(lambda: model(x, sig, ctx)).float(), D),
                 sample=rel_rms(to_diff(bf16_drift(lambda: model.sample_stochastic((2, 192, 3), ctx, rng=synth.gen(44), num_steps=5))),
                                to_diff(samp)),
                 sample2=rel_rms(to_diff(bf16_drift(lambda: model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2,
                                                                                     sigma_max=2.0))), to_diff(samp2)),
                 upsample=rel_rms(to_diff(bf16_drift(lambda: model.upsample(seed_cloud, n_new=384, context=ctx, seed=7,
                                                                              num_substeps=2, num_steps=3))), to_diff(ups)))
    print("cond_uvl bf16-autocast drift of the reference (samples in diffusion space)", drift)
    torch.save(dict(drift=drift, recipe={**recipe_common, **cfg, "K": synth.K_TASKONOMY, "feat_sizes": (64, 32, 16), "feat_seed": 31,
                            "x_seed": 32, "x_scale": 1.5, "B": B, "N": N, "noise_sigma": sig, "x2_seed": 33, "N2": 500,
                            "sample_shape": (2, 192, 3), "sample_seed": 44, "sample_steps": 5,
                            "rt_seed": 34, "ups_seed_cloud_seed": 35, "ups_n_seed": 128, "ups_n_new": 384,
                            "ups_seed": 7, "ups_substeps": 2, "ups_steps": 3},
                    D=D, hs_sub=[sub(h) for h in hs], lookup_sub=look[:, ::3].contiguous(), D_cached=D_cached,
                    sample=samp, sample2=samp2, rt_data=data, rt_back=back, upsample=ups),
               OUT / "cond_uvl.pt")
    print("cond_uvl", D.abs().mean().item(), look.abs().mean().item(), samp.abs().mean().item(), ups.abs().mean().item(),
          (back - diff).abs().max().item())
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size)


# ------------------------------------------------------------------------------------------------
# Goldens at the BENCHMARKED shape (bench.py: 2048 points per cloud, 16 row tiles, pool key splits, CTA-pair GEMM,
# staged lookup) and on a "tame" weight recipe for which the full 64-step sampler is contractive.
TAME_OUT_SCALE = 0.15  # output-head weights of the tame recipe: Lipschitz constant of F around 1 => the probability-flow
#                        map contracts and the reference's own bf16 drift over 64 steps stays ~1e-3 (printed below)


def tame_state_dict(kind, reparam, mean, sigma, seed=WEIGHT_SEED):
    return synth.tame(synth.full_state_dict(kind, reparam, mean, sigma, seed), TAME_OUT_SCALE)


def build_reference_convnext(reparam: str, mean, sigma, sigma_max: float, convnext_seed: int):
    from gecco_torch.models.feature_pyramid import ConvNeXtExtractor

    st = SetTransformer(n_layers=synth.N_LAYERS, num_inducers=synth.NUM_INDUCERS, feature_dim=synth.FEATURE_DIM,
                        t_embed_dim=1, num_heads=synth.NUM_HEADS, activation=GaussianActivation)
    m, s = torch.tensor(mean), torch.tensor(sigma)
    rp = GaussianReparam(m, s) if reparam == "gaussian" else UVLReparam(m, s)
    net = RayNetwork(backbone=st, reparam=rp, context_dims=synth.CONTEXT_DIMS)
    torch.manual_seed(convnext_seed)  # torchvision's random init draws from the global CPU generator
    cond = ConvNeXtExtractor(n_stages=3, model="tiny", pretrained=False)
    model = Diffusion(backbone=EDMPrecond(model=net), conditioner=cond, reparam=rp,
                      loss=EDMLoss(schedule=LogUniformSchedule(max=sigma_max)))
    sd = synth.full_state_dict("cond", reparam, mean, sigma, WEIGHT_SEED)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("conditioner.") for k in missing), (missing, unexpected)
    return model.eval()


def bench_shape():
    """tests/golden/bench_*.pt: Diffusion.forward at B = 4, N = 2048 for configs 1-3 (config 2 / 3 through the reference's
    own ConvNeXtExtractor, random init), the EDM training loss value on the same models, and full 64-step sampler runs
    at N = 2048 on the tame recipe."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    OUT.mkdir(parents=True, exist_ok=True)
    common = dict(weight_seed=WEIGHT_SEED, feature_dim=synth.FEATURE_DIM, n_layers=synth.N_LAYERS, num_heads=synth.NUM_HEADS,
                  num_inducers=synth.NUM_INDUCERS, torch=torch.__version__, tame_out_scale=TAME_OUT_SCALE)
    B, N = 4, 2048
    specs = [
        dict(name="bench_uncond", kind="uncond", reparam="gaussian", **synth.UNCOND_REPARAM, sigma_max=165.0, x_scale=3.0,
             noise_sigma=[0.002, 0.5, 1.0, 165.0]),
        dict(name="bench_cond_gaussian", kind="cond", reparam="gaussian", **synth.SHAPENET_VOL_REPARAM, sigma_max=165.0, x_scale=2.0,
             noise_sigma=[0.002, 0.5, 1.0, 165.0], K=synth.K_SHAPENET, image=137, convnext_seed=0, image_seed=123),
        dict(name="bench_cond_uvl", kind="cond", reparam="uvl", **synth.UVL_REPARAM, sigma_max=180.0, x_scale=1.5,
             noise_sigma=[0.002, 0.5, 1.0, 180.0], K=synth.K_TASKONOMY, image=256, convnext_seed=0, image_seed=123),
    ]
    for sp in specs:
        kind, rp = sp["kind"], sp["reparam"]
        out = {}
        if kind == "uncond":
            model = build_reference(kind, rp, sp["mean"], sp["sigma"], sp["sigma_max"])
            ctx, feats, Kc = None, None, None
        else:
            model = build_reference_convnext(rp, sp["mean"], sp["sigma"], sp["sigma_max"], sp["convnext_seed"])
            img = torch.rand(B, 3, sp["image"], sp["image"], generator=synth.gen(sp["image_seed"]))
            Kc = synth.camera(B, sp["K"])
            ctx = Context3d(image=img, K=Kc)
            with torch.no_grad():
                feats = model.conditioner(ctx).features
            out["pyramid_sub"] = [f[:, ::8, ::3, ::3].contiguous() for f in feats]
            out["pyramid_rms"] = [f.pow(2).mean().sqrt().item() for f in feats]
        sig = torch.tensor(sp["noise_sigma"])
        x = synth.noisy_input(B, N, sig, 51, 54)  # the EDM input distribution: unit-variance data + sigma * noise
        with torch.no_grad():
            D, hs = model(x, sig, ctx, do_cache=True)
        out["D"], out["hs_sub"] = D, [sub(h) for h in hs]
        drift = dict(D=rel_rms(bf16_drift(lambda: model(x, sig, ctx)).float(), D))
        # EDM training loss value (diffusion.py:118-143) on seeded draws: u ~ rand(B), noise ~ randn_like(examples)
        ex = model.reparam.diffusion_to_data(torch.randn(B, N, 3, generator=synth.gen(52)), ctx)  # in-frustum data
        torch.manual_seed(53)
        with torch.no_grad():
            loss = model.loss(model, ex, ctx)
        torch.manual_seed(53)
        u = torch.rand(B)
        n = torch.randn_like(ex)
        out["loss"], out["loss_u"], out["loss_noise_sub"] = loss, u, n[:, ::64].contiguous()
        print(sp["name"], "loss", loss.item(), "D rms", D.pow(2).mean().sqrt().item(), "bf16 drift", drift)
        # full 64-step sampler at N = 2048 on the tame recipe (same model object, weights swapped)
        tame = tame_state_dict(kind, rp, sp["mean"], sp["sigma"])
        model.load_state_dict(tame, strict=False)
        Bs = 2
        ctx_s = None if ctx is None else Context3d(image=ctx.image[:Bs], K=ctx.K[:Bs])
        with torch.no_grad():
            samp = model.sample_stochastic((Bs, N, 3), ctx_s, rng=synth.gen(61), num_steps=64)
        sb = bf16_drift(lambda: model.sample_stochastic((Bs, N, 3), ctx_s, rng=synth.gen(61), num_steps=64))
        if rp == "uvl":
            to_diff = lambda d: model.reparam.data_to_diffusion(d, Context3d(image=ctx_s.image, K=ctx_s.K.double()))
        else:
            to_diff = lambda d: model.reparam.data_to_diffusion(d, ctx_s)
        drift["sample64"] = rel_rms(to_diff(sb), to_diff(samp))
        out["sample64"] = samp
        print(sp["name"], "tame 64-step sampler: bf16 drift of the reference (diffusion space)", drift["sample64"],
              "finite", torch.isfinite(to_diff(samp)).all().item())
        recipe = {**common, **{k: v for k, v in sp.items() if k != "name"}, "B": B, "N": N, "x_seed": 51, "x_noise_seed": 54, "ex_seed": 52,
                  "loss_seed": 53, "sample_B": Bs, "sample_seed": 61, "sample_steps": 64}
        recipe["noise_sigma"] = sig
        torch.save(dict(recipe=recipe, drift=drift, **out), OUT / (sp["name"] + ".pt"))
    for f in sorted(OUT.glob("bench_*.pt")):
        print(f.name, f.stat().st_size)


def module_goldens():
    """tests/golden/modules.pt: every module on the call surface run STAND-ALONE in the reference
    (set_transformer.py:47-216, mlp.py, activation.py, normalization.py) with randomised AdaGN weights and alphas."""
    from gecco_torch.models.mlp import MLP
    from gecco_torch.models.normalization import AdaGN  # noqa: F401

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    L = 2
    st = SetTransformer(n_layers=L, num_inducers=synth.NUM_INDUCERS, feature_dim=synth.FEATURE_DIM, t_embed_dim=1,
                        num_heads=synth.NUM_HEADS, activation=GaussianActivation).eval()
    pre = "backbone.model.inner."
    sd = {k[len(pre):]: v for k, v in synth.synth_state_dict(synth.network_shapes("uncond", n_layers=L), WEIGHT_SEED).items()
          if k.startswith(pre)}
    assert set(sd) == set(st.state_dict()), sorted(set(sd) ^ set(st.state_dict()))
    st.load_state_dict(sd)
    B, N, C = 2, 300, synth.FEATURE_DIM
    x = torch.randn(B, N, C, generator=synth.gen(71))
    t = torch.randn(B, 1, 1, generator=synth.gen(72)) * 0.8
    x2 = torch.randn(B, 450, C, generator=synth.gen(73))
    lay = st.layers[0]
    relu_mlp = MLP(C, 96, 256, depth=2).eval()  # reference default activation (nn.ReLU), depth 2
    rsd = synth.synth_state_dict({k: tuple(v.shape) for k, v in relu_mlp.state_dict().items()}, 77)
    relu_mlp.load_state_dict(rsd)
    out = {}
    sub = lambda t_: t_[:, ::5, ::8].contiguous()
    with torch.no_grad():
        out["act"] = sub(lay.mlp[1](x))
        out["act_raw"] = sub(GaussianActivation(normalized=False)(x))
        out["adagn"] = sub(lay.broadcast_norm(x, t))
        out["mlp"] = sub(lay.mlp(x))
        out["relu_mlp"] = sub(relu_mlp(x))
        out["pool"] = sub(lay.broadcast.pool(x))
        attn, h = lay.broadcast(x, t, return_h=True)
        out["broadcast"], out["broadcast_h"] = sub(attn), sub(h)
        attn2, none = lay.broadcast(x2, t, return_h=False, h=h)
        assert none is None
        out["broadcast_cached"] = sub(attn2)
        y, h1 = lay(x, t, return_h=True)
        out["layer"], out["layer_h"] = sub(y), sub(h1)
        f, hs = st(x, t, return_h=True)
        out["st"], out["st_hs"] = sub(f), [sub(h_) for h_ in hs]
        f2, none = st(x2, t, return_h=False, hs=hs)
        assert none is None
        out["st_cached"] = sub(f2)
    recipe = dict(weight_seed=WEIGHT_SEED, n_layers=L, B=B, N=N, N2=450, x_seed=71, t_seed=72, t_scale=0.8, x2_seed=73,
                  relu_mlp=dict(out=96, width=256, depth=2, seed=77), torch=torch.__version__)
    torch.save(dict(recipe=recipe, **out), OUT / "modules.pt")
    print("modules.pt", (OUT / "modules.pt").stat().st_size, {k: (v.pow(2).mean().sqrt().item() if torch.is_tensor(v) else None) for k, v in out.items()})


def grad_goldens():
    """tests/golden/grads_*.pt: the EDM training loss (diffusion.py:118-143, 210-222) and its GRADIENTS w.r.t. every
    parameter, from the unmodified reference (fp32, CPU) on seeded draws: per-parameter gradient norm and a strided
    sub-sample of every gradient.  Mode-dependent layers (torchvision's StochasticDepth inside ConvNeXt) are in eval mode."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    OUT.mkdir(parents=True, exist_ok=True)
    common = dict(weight_seed=WEIGHT_SEED, feature_dim=synth.FEATURE_DIM, n_layers=synth.N_LAYERS, num_heads=synth.NUM_HEADS,
                  num_inducers=synth.NUM_INDUCERS, torch=torch.__version__)
    specs = [
        dict(name="grads_uncond", kind="uncond", reparam="gaussian", **synth.UNCOND_REPARAM, sigma_max=165.0, B=2, N=1000),
        dict(name="grads_cond_gaussian", kind="cond", reparam="gaussian", **synth.SHAPENET_VOL_REPARAM, sigma_max=165.0, B=2, N=1024,
             K=synth.K_SHAPENET, image=137, convnext_seed=0, image_seed=123),
    ]
    for sp in specs:
        kind, rp, B, N = sp["kind"], sp["reparam"], sp["B"], sp["N"]
        if kind == "uncond":
            model = build_reference(kind, rp, sp["mean"], sp["sigma"], sp["sigma_max"])
            ctx = None
        else:
            model = build_reference_convnext(rp, sp["mean"], sp["sigma"], sp["sigma_max"], sp["convnext_seed"])
            img = torch.rand(B, 3, sp["image"], sp["image"], generator=synth.gen(sp["image_seed"]))
            ctx = Context3d(image=img, K=synth.camera(B, sp["K"]))
        model.eval()
        with torch.no_grad():
            ex = model.reparam.diffusion_to_data(torch.randn(B, N, 3, generator=synth.gen(52)), ctx)
        torch.manual_seed(53)
        loss = model.loss(model, ex, ctx)
        loss.backward()
        torch.manual_seed(53)
        u = torch.rand(B)
        n = torch.randn_like(ex)
        norms, subs = {}, {}
        for k, p_ in model.named_parameters():
            if p_.grad is None:
                continue
            g = p_.grad.detach().flatten()
            norms[k] = g.double().norm().item()
            stride = max(1, g.numel() // 256)
            subs[k] = g[::stride].clone()
        total = sum(v * v for v in norms.values()) ** 0.5
        print(sp["name"], "loss", loss.item(), "params with grad", len(norms), "total grad norm", total)
        recipe = {**common, **{k: v for k, v in sp.items() if k != "name"}, "ex_seed": 52, "loss_seed": 53}
        torch.save(dict(recipe=recipe, loss=loss.detach(), loss_u=u, loss_noise_sub=n[:, ::64].contiguous(), grad_norm=norms,
                        grad_sub=subs), OUT / (sp["name"] + ".pt"))
        print((OUT / (sp["name"] + ".pt")).stat().st_size)


if __name__ == "__main__":
    if "--grads" in sys.argv:
        grad_goldens()
    elif "--modules" in sys.argv:
        module_goldens()
    elif "--bench-shape" in sys.argv:
        bench_shape()
    else:
        main()
