"""End-to-end parity of the CUDA denoiser / sampler (through the gecco-torch call surface and the C ABI) against
the golden vectors of the unmodified reference and against the CPU oracle on the same seeded inputs.

Tolerances (bf16 tensor-core operands, fp32 accumulation, fp32 residual stream; SURVEY.md §8c measured the
reference's own bf16-autocast drift at rms 3.9e-3 * rms(F), max 1.7e-2):
  raw network output F:   rms error <= 1.5e-2 * rms(F), max error <= 8e-2 * max|F|
  inducer states h:       rms error <= 1.5e-2 * rms(h)
  sampled points:         rms error <= max(2e-2, the reference's own bf16-autocast drift on the same call) * rms(sample).
The multi-step sampler amplifies any perturbation (with the randomised synthetic weights the conditional sampler is
chaotic: the UNMODIFIED reference drifts by 9e-2 .. 3.6e-1 under its own bf16 autocast), so the golden files carry that
drift (`drift`, measured by oracle/make_golden.py).  2-step sampler runs are additionally checked against the oracle and
the reference golden with tolerance max(2e-2, half the reference's own bf16 drift on that call).
"""
from pathlib import Path

import pytest
import torch

from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def rms(t):
    return t.double().pow(2).mean().sqrt().item()


def check_F(D_gpu, D_ref, x, sigma, what):
    """compares in units of the raw network output F = (D - c_skip x) / c_out"""
    s = sigma.reshape(-1, 1, 1).double()
    c_skip, c_out = 1 / (s**2 + 1), s / (s**2 + 1).sqrt()
    F_gpu = (D_gpu.double().cpu() - c_skip * x.double()) / c_out
    F_ref = (D_ref.double() - c_skip * x.double()) / c_out
    err = F_gpu - F_ref
    # at tiny sigma, D is dominated by c_skip x and F is recovered with cancellation noise: weight by c_out
    for b in range(x.shape[0]):
        if c_out[b].item() < 0.01:
            continue
        r, mx = rms(err[b]) / rms(F_ref[b]), err[b].abs().max().item() / F_ref[b].abs().max().item()
        print(f"{what}[{b}] sigma={sigma[b].item():g}: rel rms {r:.2e}, rel max {mx:.2e}")
        assert r < 1.5e-2 and mx < 8e-2, (what, b, r, mx)
    assert torch.isfinite(D_gpu).all()
    d = (D_gpu.double().cpu() - D_ref.double())
    assert rms(d) < 1.5e-2 * max(rms(D_ref), 1e-3), (what, rms(d), rms(D_ref))


def load(name):
    g = torch.load(GOLD / name, weights_only=False)
    return g, g["recipe"]


def test_uncond(cuda):
    g, r = load("uncond.pt")
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda)
    x = torch.randn(r["B"], r["N"], 3, generator=synth.gen(r["x_seed"])) * r["x_scale"]
    sig = r["noise_sigma"]
    D = model(x.to(cuda), sig.to(cuda), None)
    check_F(D, g["D"], x, sig, "uncond D")
    D2, hs = model(x.to(cuda), sig.to(cuda), None, do_cache=True)
    assert torch.equal(D, D2) and len(hs) == r["n_layers"]
    for l, (h, hg) in enumerate(zip(hs, g["hs_sub"])):
        e = rms(h[:, ::8, ::8].cpu() - hg) / rms(hg)
        assert e < 1.5e-2, (l, e)
    x2 = torch.randn(r["B"], r["N2"], 3, generator=synth.gen(r["x2_seed"])) * r["x_scale"]
    Dc = model(x2.to(cuda), sig.to(cuda), None, cache=hs)
    check_F(Dc, g["D_cached"], x2, sig, "uncond cached")
    s = model.sample_stochastic(r["sample_shape"], None, rng=synth.gen(r["sample_seed"]), num_steps=r["sample_steps"])
    assert s.dtype == torch.float64 and s.shape == tuple(r["sample_shape"])
    e = rms(s.cpu() - g["sample"]) / rms(g["sample"])
    print("uncond sample rel rms", e, "reference bf16 drift", g["drift"]["sample"])
    assert e < max(2e-2, g["drift"]["sample"])
    assert torch.equal(model.t_steps(64, 165.0, 0.002, 7).cpu(), g["t_steps"])


def test_cond_gaussian(cuda):
    g, r = load("cond_gaussian.pt")
    feats = synth.synth_features(r["B"], r["feat_sizes"], r["feat_seed"])
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda, feats)
    import gecco_b200 as G

    ctx = G.Context3d(image=torch.zeros(r["B"], 3, 8, 8, device=cuda), K=synth.camera(r["B"], r["K"]).to(cuda))
    x = torch.randn(r["B"], r["N"], 3, generator=synth.gen(r["x_seed"])) * r["x_scale"]
    sig = r["noise_sigma"]
    D = model(x.to(cuda), sig.to(cuda), ctx)
    check_F(D, g["D"], x, sig, "cond_gaussian D")
    c_in = 1 / (1 + sig**2).sqrt()
    look = model.backbone.model.extract_image_features((x * c_in[:, None, None]).to(cuda), [f.to(cuda) for f in feats], ctx)
    e = rms(look[:, ::3].cpu() - g["lookup_sub"]) / rms(g["lookup_sub"])
    print("lookup rel rms", e)
    assert e < 5e-3  # bf16 feature maps, fp32 interpolation
    s = model.sample_stochastic(r["sample_shape"], ctx, rng=synth.gen(r["sample_seed"]), num_steps=r["sample_steps"])
    e = rms(s.cpu() - g["sample"]) / rms(g["sample"])
    print("cond_gaussian sample rel rms", e, "reference bf16 drift", g["drift"]["sample"])
    assert e < max(2e-2, g["drift"]["sample"])
    # short trajectory against the oracle: little room for chaotic amplification
    cfg = O.OracleConfig(kind="cond", reparam="gaussian", sigma_max=r["sigma_max"])
    sd = synth.full_state_dict(r["kind"], r["reparam"], r["mean"], r["sigma"], r["weight_seed"])
    s2 = model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2)
    o2 = O.sample_stochastic(cfg, sd, (2, 160, 3), feats, synth.camera(r["B"], r["K"]), rng=synth.gen(5), num_steps=2)
    e = rms(s2.cpu() - o2) / rms(o2)
    eg = rms(s2.cpu() - g["sample2"]) / rms(g["sample2"])
    tol = max(2e-2, 0.5 * g["drift"]["sample2"])  # at least twice as close as the reference's own bf16 autocast
    print("cond_gaussian 2-step sample rel rms vs oracle", e, "vs reference golden", eg, "tolerance", tol)
    assert e < tol and eg < tol


def test_cond_uvl(cuda):
    g, r = load("cond_uvl.pt")
    feats = synth.synth_features(r["B"], r["feat_sizes"], r["feat_seed"])
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda, feats)
    import gecco_b200 as G

    K = synth.camera(r["B"], r["K"])
    ctx = G.Context3d(image=torch.zeros(r["B"], 3, 8, 8, device=cuda), K=K.to(cuda))
    x = torch.randn(r["B"], r["N"], 3, generator=synth.gen(r["x_seed"])) * r["x_scale"]
    sig = r["noise_sigma"]
    D, hs = model(x.to(cuda), sig.to(cuda), ctx, do_cache=True)
    check_F(D, g["D"], x, sig, "cond_uvl D")
    for l, (h, hg) in enumerate(zip(hs, g["hs_sub"])):
        e = rms(h[:, ::8, ::8].cpu() - hg) / rms(hg)
        assert e < 1.5e-2, (l, e)
    x2 = torch.randn(r["B"], r["N2"], 3, generator=synth.gen(r["x2_seed"])) * r["x_scale"]
    Dc = model(x2.to(cuda), sig.to(cuda), ctx, cache=hs)
    check_F(Dc, g["D_cached"], x2, sig, "cond_uvl cached")
    # reparam round trip (float32 and float64)
    diff = torch.randn(r["B"], 64, 3, generator=synth.gen(r["rt_seed"]))
    data = model.reparam.diffusion_to_data(diff.to(cuda), ctx)
    assert (data.cpu() - g["rt_data"]).abs().max().item() < 1e-4 * g["rt_data"].abs().max().item()
    back = model.reparam.data_to_diffusion(data.double(), ctx)
    assert (back.cpu() - diff.double()).abs().max().item() < 1e-4
    # samples are compared in diffusion space (data space goes through exp() of the ray length)
    cfg = O.OracleConfig(kind="cond", reparam="uvl", sigma_max=r["sigma_max"])
    sd = synth.reparam_buffers("uvl", r["mean"], r["sigma"])
    to_diff = lambda d: O.data_to_diffusion(cfg, sd, d.cpu(), K.double())
    s = model.sample_stochastic(r["sample_shape"], ctx, rng=synth.gen(r["sample_seed"]), num_steps=r["sample_steps"])
    e = rms(to_diff(s) - to_diff(g["sample"])) / rms(to_diff(g["sample"]))
    print("cond_uvl sample rel rms", e, "reference bf16 drift", g["drift"]["sample"])
    assert e < max(2e-2, g["drift"]["sample"])
    sdf = synth.full_state_dict(r["kind"], r["reparam"], r["mean"], r["sigma"], r["weight_seed"])
    # (sigma_max lowered so that exp() of the ray-length coordinate stays finite after one giant step)
    s2 = model.sample_stochastic((2, 160, 3), ctx, rng=synth.gen(5), num_steps=2, sigma_max=2.0)
    o2 = O.sample_stochastic(cfg, sdf, (2, 160, 3), feats, K, rng=synth.gen(5), num_steps=2, sigma_max=2.0)
    assert torch.isfinite(o2).all() and torch.isfinite(s2).all()
    e = rms(to_diff(s2) - to_diff(o2)) / rms(to_diff(o2))
    eg = rms(to_diff(s2) - to_diff(g["sample2"])) / rms(to_diff(g["sample2"]))
    tol = max(2e-2, 0.5 * g["drift"]["sample2"])  # at least twice as close as the reference's own bf16 autocast
    print("cond_uvl 2-step sample rel rms vs oracle", e, "vs reference golden", eg, "tolerance", tol)
    assert e < tol and eg < tol
    seed_cloud = O.diffusion_to_data(cfg, sd, torch.randn(r["B"], r["ups_n_seed"], 3, generator=synth.gen(r["ups_seed_cloud_seed"])), K)
    # Upsampling against the oracle with the same draws from a CPU generator.  With the randomised synthetic weights the
    # upsampling dynamics are violently chaotic at the reference's noise levels (perturbing the ORACLE's own denoiser
    # output by 1e-3 moves its result by 11 %, by 5e-3 by 23 %; tools/debug_upsample.py), so the host loop (draw
    # order, churn / Euler / Heun / re-noise arithmetic, cached-inducer plumbing) is pinned at a small sigma_max, where the
    # map is contractive; the numerics of the cached evaluations themselves are pinned above (D_cached).
    us = model.upsample(seed_cloud.to(cuda), n_new=200, context=ctx, num_substeps=3, num_steps=3, sigma_max=0.05, rng=synth.gen(9))
    uo = O.upsample(cfg, sdf, seed_cloud, n_new=200, features=feats, K=K, seed=9, num_substeps=3, num_steps=3, sigma_max=0.05)
    assert torch.isfinite(us).all() and torch.isfinite(uo).all()
    e = rms(to_diff(us) - to_diff(uo)) / rms(to_diff(uo))
    print("cond_uvl small-sigma upsample rel rms vs oracle", e)
    assert e < 1e-2
    # the long trajectory of the golden file is chaotic (the reference drifts by 0.28 under its own bf16 autocast) and a
    # few points saturate tanh / exp on the way back to diffusion space: compare the rows that are finite in both
    u = model.upsample(seed_cloud.to(cuda), n_new=r["ups_n_new"], context=ctx, num_substeps=r["ups_substeps"],
                       num_steps=r["ups_steps"], rng=synth.gen(r["ups_seed"]))
    assert torch.isfinite(u).all() and u.dtype == torch.float64 and u.shape == g["upsample"].shape
    ud, gd = to_diff(u), to_diff(g["upsample"])
    ok = torch.isfinite(ud).all(dim=-1) & torch.isfinite(gd).all(dim=-1)
    assert ok.float().mean().item() > 0.7
    e = rms(ud[ok] - gd[ok]) / rms(gd[ok])
    print("cond_uvl upsample rel rms", e, "reference bf16 drift", g["drift"]["upsample"], "finite rows", ok.float().mean().item())
    assert e < max(3e-2, 1.25 * g["drift"]["upsample"])


def test_errors(cuda):
    g, r = load("uncond.pt")
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], None, cuda)
    with pytest.raises(Exception):
        model(torch.zeros(1, 8, 3), torch.ones(1), None)  # CPU tensors: no fallback
    with pytest.raises(ValueError):
        model.upsample(torch.zeros(1, 8, 3, device=cuda))  # diffusion.py:398-401
    with pytest.raises(ValueError):
        model(torch.zeros(1, 8, 2, device=cuda), torch.ones(1, device=cuda), None)
