"""The samplers gecco-torch lists as missing and gecco-jax has: deterministic probability-flow ODE
sampling (Heun) and inpainting, both on the engine, against the CPU oracle on tame weights (contractive map) with the same
CPU-generator draws.  gecco-jax cannot be imported in the build container, so these are pinned to the oracle's restatement
(oracle/gecco_oracle.py: sample_stochastic with S_churn = 0, sample_inpaint), not to reference outputs."""
import pytest
import torch

from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build

pytestmark = pytest.mark.gpu


def rms(t):
    return t.double().pow(2).mean().sqrt().item()


def _model(cuda, kind, feats=None):
    rp = synth.SHAPENET_VOL_REPARAM if kind == "cond" else synth.UNCOND_REPARAM
    sd = synth.tame(synth.full_state_dict(kind, "gaussian", rp["mean"], rp["sigma"], 1234), 0.15)
    model = build(kind, "gaussian", rp["mean"], rp["sigma"], 165.0, None, cuda, feats, state_dict=sd)
    return model, sd, rp


def test_ode_sampler_matches_oracle(cuda):
    import gecco_b200 as G

    B, N, steps = 2, 384, 12
    feats = synth.synth_features(B, (34, 17, 8), 21)
    model, sd, rp = _model(cuda, "cond", feats)
    K = synth.camera(B, synth.K_SHAPENET)
    ctx = G.Context3d(image=torch.zeros(B, 3, 8, 8, device=cuda), K=K.to(cuda))
    cfg = O.OracleConfig(kind="cond", reparam="gaussian", sigma_max=165.0)
    got = model.sample_ode((B, N, 3), ctx, rng=synth.gen(3), num_steps=steps)
    ref = O.sample_stochastic(cfg, sd, (B, N, 3), feats, K, rng=synth.gen(3), num_steps=steps, S_churn=0.0)
    to_diff = lambda d: O.data_to_diffusion(cfg, sd, d.cpu().double(), None)
    e = rms(to_diff(got) - to_diff(ref)) / rms(to_diff(ref))
    print("ODE sampler rel rms (diffusion space)", e)
    assert got.dtype == torch.float64 and e < 1e-2
    # deterministic: the same latents give the same cloud, bit for bit, and no noise is consumed from the generator
    g = synth.gen(3)
    again = model.sample_ode((B, N, 3), ctx, rng=g, num_steps=steps)
    assert torch.equal(got, again)
    assert torch.equal(torch.randn(4, generator=g), torch.randn(4, generator=_after_latents(3, (B, N, 3))))


def _after_latents(seed, shape):
    g = synth.gen(seed)
    torch.randn(shape, generator=g)
    return g


def test_inpaint_matches_oracle(cuda):
    B, N, M, steps = 2, 200, 184, 6
    model, sd, rp = _model(cuda, "uncond")
    cfg = O.OracleConfig(kind="uncond", reparam="gaussian", sigma_max=165.0)
    known = O.diffusion_to_data(cfg, sd, torch.randn(B, N, 3, generator=synth.gen(5)), None)
    got = model.sample_inpaint(known.to(cuda), M, None, rng=synth.gen(6), num_substeps=2, num_steps=steps, S_churn=0.5)
    ref = O.sample_inpaint(cfg, sd, known, M, rng=synth.gen(6), num_substeps=2, num_steps=steps, S_churn=0.5)
    to_diff = lambda d: O.data_to_diffusion(cfg, sd, d.cpu().double(), None)
    e = rms(to_diff(got) - to_diff(ref)) / rms(to_diff(ref))
    print("inpaint rel rms (diffusion space)", e)
    assert got.shape == (B, M, 3) and got.dtype == torch.float64 and e < 1e-2


@pytest.mark.parametrize("kind", ["uncond", "cond"])
def test_log_likelihood_matches_oracle(cuda, kind):
    """`Diffusion.log_likelihood` (gecco-jax `evaluate_logp`, models/diffusion.py:444-541) on the tcgen05 path (forward and
    input-gradient products in bf16) against the fp32 CPU oracle with the same Rademacher probes."""
    import gecco_b200 as G

    B, N, steps = 2, 384, 6
    feats = synth.synth_features(B, (34, 17, 8), 21) if kind == "cond" else None
    model, sd, rp = _model(cuda, kind, feats)
    K = synth.camera(B, synth.K_SHAPENET) if kind == "cond" else None
    ctx = G.Context3d(image=torch.zeros(B, 3, 8, 8, device=cuda), K=K.to(cuda)) if kind == "cond" else None
    cfg = O.OracleConfig(kind=kind, reparam="gaussian", sigma_max=165.0)
    data = O.diffusion_to_data(cfg, sd, torch.randn(B, N, 3, generator=synth.gen(2)) * 0.8, K)
    noise = (torch.randint(0, 2, (2, B, N, 3), generator=synth.gen(3)) * 2 - 1).float()
    ref = O.log_likelihood(cfg, sd, data, noise, feats, K, num_steps=steps)
    got = model.log_likelihood(data.to(cuda), ctx, noise=noise.to(cuda), num_steps=steps, return_details=True)
    rel = lambda k: ((got[k].cpu() - ref[k]).abs() / ref[k].abs().clamp_min(1.0)).max().item()
    print("log-likelihood", kind, {k: rel(k) for k in ("prior_logp", "delta_jacobian", "delta_reparam", "logp")},
          "latent", rms(got["latent"].cpu() - ref["latent"]) / rms(ref["latent"]))
    assert got["logp"].shape == (B,) and got["logp"].dtype == torch.float64
    assert rel("delta_reparam") < 1e-5
    assert rms(got["latent"].cpu() - ref["latent"]) / rms(ref["latent"]) < 1e-2
    # measured on B200: prior 1.6e-5 / 1.6e-6, divergence integral 4e-5 / 9.4e-4, log p 6e-5 / 2.9e-3 (uncond / cond)
    assert rel("prior_logp") < 1e-3 and rel("delta_jacobian") < 5e-3 and rel("logp") < 1e-2


def test_log_likelihood_uvl_matches_oracle(cuda):
    """The same with the UVL reparametrisation (256^2-style pyramid): the lookup positions go through tanh / exp / the
    camera, and the log-det of data -> diffusion is a full 3 x 3 block per point."""
    import gecco_b200 as G

    B, N, steps = 2, 256, 5
    rp = synth.UVL_REPARAM
    feats = synth.synth_features(B, (64, 32, 16), 21)
    sd = synth.tame(synth.full_state_dict("cond", "uvl", rp["mean"], rp["sigma"], 1234), 0.15)
    model = build("cond", "uvl", rp["mean"], rp["sigma"], 180.0, None, cuda, feats, state_dict=sd)
    K = synth.camera(B, synth.K_TASKONOMY)
    ctx = G.Context3d(image=torch.zeros(B, 3, 8, 8, device=cuda), K=K.to(cuda))
    cfg = O.OracleConfig(kind="cond", reparam="uvl", sigma_max=180.0)
    data = O.diffusion_to_data(cfg, sd, torch.randn(B, N, 3, generator=synth.gen(2)) * 0.8, K)
    noise = (torch.randint(0, 2, (1, B, N, 3), generator=synth.gen(3)) * 2 - 1).float()
    ref = O.log_likelihood(cfg, sd, data, noise, feats, K, num_steps=steps)
    got = model.log_likelihood(data.to(cuda), ctx, noise=noise.to(cuda), num_steps=steps, return_details=True)
    rel = lambda k: ((got[k].cpu() - ref[k]).abs() / ref[k].abs().clamp_min(1.0)).max().item()
    print("log-likelihood uvl", {k: rel(k) for k in ("prior_logp", "delta_jacobian", "delta_reparam", "logp")},
          "latent", rms(got["latent"].cpu() - ref["latent"]) / rms(ref["latent"]))
    assert rel("delta_reparam") < 1e-4
    assert rms(got["latent"].cpu() - ref["latent"]) / rms(ref["latent"]) < 1e-2
    # log p is a sum of large terms of both signs here: its error is held against their magnitude, not against |log p|
    scale = ref["prior_logp"].abs() + ref["delta_jacobian"].abs() + ref["delta_reparam"].abs()
    e_logp = ((got["logp"].cpu() - ref["logp"]).abs() / scale).max().item()
    print("log-likelihood uvl: log p error relative to the sum of its terms", e_logp)
    # measured on B200: prior 7.7e-6, divergence integral 6.1e-3, log-det 6e-7, latent 2.9e-3
    assert rel("prior_logp") < 1e-3 and rel("delta_jacobian") < 2e-2 and e_logp < 1e-2
