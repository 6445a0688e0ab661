"""Parity at the BENCHMARKED shape: 2048 points per cloud (16 row tiles per cloud, pool key splits, CTA-pair GEMM, staged
lookup), B = 4, for the three model configurations of BASELINE.json (configs 1-3), against goldens minted from the
unmodified reference by `oracle/make_golden.py --bench-shape`:

  * Diffusion.forward at sigma in {0.002, 0.5, 1, sigma_max}: rms(F error) <= 1.5e-2 rms(F), per cloud;
  * configs 2 / 3 run through the reference's own ConvNeXtExtractor (random init, seeded) rather than a fixed pyramid;
    the pyramid itself is compared with the reference's (conditioner parity);
  * the EDM training-loss VALUE (diffusion.py:118-143) on the same seeded draws;
  * a full 64-step stochastic sampler run at N = 2048 on the "tame" weight recipe (tests/synth.py: tame), for which the
    probability-flow map is contractive: rms error of the samples (diffusion space) <= 2e-2 rms, with NO allowance for the
    reference's own drift.
"""
from pathlib import Path

import pytest
import torch

from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build
from tests.test_denoiser_gpu import check_F, rms

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
NAMES = ["bench_uncond", "bench_cond_gaussian", "bench_cond_uvl"]


def _setup(name, cuda, state_dict=None):
    import gecco_b200 as G

    g = torch.load(GOLD / (name + ".pt"), weights_only=False)
    r = g["recipe"]
    cond = r["kind"] == "cond"
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda,
                  convnext_seed=r["convnext_seed"] if cond else None, state_dict=state_dict)
    ctx = None
    if cond:
        img = torch.rand(r["B"], 3, r["image"], r["image"], generator=synth.gen(r["image_seed"]))
        ctx = G.Context3d(image=img.to(cuda), K=synth.camera(r["B"], r["K"]).to(cuda))
    return g, r, model, ctx


@pytest.fixture(autouse=True)
def _exact_convs():
    # the conditioner is compared with the reference's fp32 CPU run: no TF32 inside cuDNN / cuBLAS for these tests
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", NAMES)
def test_forward_and_loss_at_bench_shape(cuda, monkeypatch, name):
    g, r, model, ctx = _setup(name, cuda)
    B, N = r["B"], r["N"]
    if ctx is not None:
        feats = model.conditioner(ctx).features
        for l, (f, sub, frms) in enumerate(zip(feats, g["pyramid_sub"], g["pyramid_rms"])):
            e = rms(f[:, ::8, ::3, ::3].cpu() - sub) / frms
            print(f"{name}: ConvNeXt level {l} rel rms vs reference {e:.2e}")
            assert e < 1e-3, (l, e)
    sig = r["noise_sigma"]
    x = synth.noisy_input(B, N, sig, r["x_seed"], r["x_noise_seed"])
    D, hs = model(x.to(cuda), sig.to(cuda), ctx, do_cache=True)
    check_F(D, g["D"], x, sig, name + " D")
    for l, (h, hg) in enumerate(zip(hs, g["hs_sub"])):
        e = rms(h[:, ::8, ::8].cpu() - hg) / rms(hg)
        assert e < 1.5e-2, (l, e)

    # EDM loss value on the reference's draws: u ~ rand(B) and noise ~ randn_like(examples) after manual_seed
    cfg = O.OracleConfig(kind=r["kind"], reparam=r["reparam"], sigma_max=r["sigma_max"])
    bufs = synth.reparam_buffers(r["reparam"], r["mean"], r["sigma"])
    Kc = None if ctx is None else synth.camera(B, r["K"])
    ex = O.diffusion_to_data(cfg, bufs, torch.randn(B, N, 3, generator=synth.gen(r["ex_seed"])), Kc)
    torch.manual_seed(r["loss_seed"])
    u = torch.rand(B)
    noise = torch.randn_like(ex)
    assert torch.equal(u, g["loss_u"]) and torch.equal(noise[:, ::64], g["loss_noise_sub"])
    monkeypatch.setattr(torch, "rand", lambda n, device=None, **kw: u.to(device))
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: noise.to(t.device, t.dtype))
    loss = model.loss(model, ex.to(cuda), ctx)
    monkeypatch.undo()
    rel = abs(loss.item() - g["loss"].item()) / abs(g["loss"].item())
    print(f"{name}: EDM loss {loss.item():.6g} vs reference {g['loss'].item():.6g} (rel {rel:.2e})")
    assert rel < 1e-2
    vloss = model.validation_step((ex.to(cuda), ctx), 0)  # same entry, fresh draws: finite and of the same magnitude
    assert torch.isfinite(vloss) and 0.2 < vloss.item() / g["loss"].item() < 5


@pytest.mark.parametrize("name", NAMES)
def test_full_sampler_tame_weights(cuda, name):
    """64 steps (127 evaluations) at N = 2048, through the public sampler (CUDA-graph path), <= 2e-2 in diffusion space."""
    import gecco_b200 as G
    from gecco_b200.engine import engine_for

    g0 = torch.load(GOLD / (name + ".pt"), weights_only=False)
    r = g0["recipe"]
    sd = synth.tame(synth.full_state_dict(r["kind"], r["reparam"], r["mean"], r["sigma"], r["weight_seed"]), r["tame_out_scale"])
    g, r, model, ctx = _setup(name, cuda, state_dict=sd)
    Bs, N = r["sample_B"], r["N"]
    ctx_s = None if ctx is None else G.Context3d(image=ctx.image[:Bs], K=ctx.K[:Bs])
    s = model.sample_stochastic((Bs, N, 3), ctx_s, rng=synth.gen(r["sample_seed"]), num_steps=r["sample_steps"])
    assert s.dtype == torch.float64 and torch.isfinite(s).all()
    cfg = O.OracleConfig(kind=r["kind"], reparam=r["reparam"], sigma_max=r["sigma_max"])
    bufs = synth.reparam_buffers(r["reparam"], r["mean"], r["sigma"])
    Kd = None if ctx is None else synth.camera(Bs, r["K"]).double()
    to_diff = lambda d: O.data_to_diffusion(cfg, bufs, d.cpu().double(), Kd)
    e = rms(to_diff(s) - to_diff(g["sample64"])) / rms(to_diff(g["sample64"]))
    print(f"{name}: 64-step sampler at N=2048, rel rms (diffusion space) {e:.2e}; reference's own bf16 drift {g['drift']['sample64']:.2e}")
    assert e < 2e-2
    eng = engine_for(*model._network())
    assert eng.graph_status() in (1, 2), "the sampler loop did not run as a CUDA graph"
    # a second call replays the captured graph and reproduces the first bit for bit
    s2 = model.sample_stochastic((Bs, N, 3), ctx_s, rng=synth.gen(r["sample_seed"]), num_steps=r["sample_steps"])
    assert eng.graph_status() == 2 and torch.equal(s, s2)


def test_anorm_and_fold_paths_agree(cuda):
    """AdaGN inside the consuming GEMM (A-operand transform, default at this shape) against the per-cloud weight fold
    (`gecco_set_option("anorm", 0)`): two implementations of the same arithmetic, both within tolerance of the reference."""
    from gecco_b200 import ops

    g, r, model, ctx = _setup("bench_cond_gaussian", cuda)
    sig = r["noise_sigma"]
    x = synth.noisy_input(r["B"], r["N"], sig, r["x_seed"], r["x_noise_seed"])
    try:
        ops.set_option("anorm", 0)
        D_fold = model(x.to(cuda), sig.to(cuda), ctx).clone()
        ops.set_option("anorm", 1)
        D_an = model(x.to(cuda), sig.to(cuda), ctx).clone()
    finally:
        ops.set_option("anorm", 1)
    check_F(D_fold, g["D"], x, sig, "fold path")
    check_F(D_an, g["D"], x, sig, "anorm path")
    assert not torch.equal(D_fold, D_an)  # they really are two code paths
    e_fold, e_an = rms(D_fold.cpu() - g["D"]), rms(D_an.cpu() - g["D"])
    print(f"rms error vs reference: fold {e_fold:.3e}, anorm {e_an:.3e}")
