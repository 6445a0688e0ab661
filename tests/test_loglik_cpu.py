"""Log-likelihood through the probability-flow ODE (gecco-jax `evaluate_logp`,
models/diffusion.py:444-541).  gecco-jax cannot be imported here, so (i) the oracle's restatement is checked against the
closed-form likelihood of the optimal denoiser of Gaussian data, and (ii) the host logic of `Diffusion.log_likelihood`
(ODE integration, Hutchinson estimator, reparametrisation log-det) is compared with the oracle on the CPU autograd
restatement of the network (2 layers); the GPU test (tests/test_samplers_gpu.py) covers the tcgen05 path."""
import math

import pytest
import torch

from oracle import gecco_oracle as O
from tests import synth
from tests.models_b200 import build


def test_oracle_log_likelihood_closed_form_gaussian():
    # data ~ N(0, s^2 I): D(x; sigma) = x s^2 / (s^2 + sigma^2), dx/dsigma = x sigma / (s^2 + sigma^2),
    # x(sigma) = x0 sqrt((s^2 + sigma^2) / (s^2 + smin^2)), divergence = d sigma / (s^2 + sigma^2)
    s, smin, smax, B, N = 0.7, 0.002, 80.0, 3, 16
    cfg = O.OracleConfig(kind="uncond", reparam="none", sigma_max=smax)
    x0 = torch.randn(B, N, 3, generator=synth.gen(0)) * s
    noise = (torch.randint(0, 2, (2, B, N, 3), generator=synth.gen(1)) * 2 - 1).float()
    den = lambda x, sig: x * s**2 / (s**2 + sig.reshape(-1, 1, 1) ** 2)
    out = O.log_likelihood(cfg, {}, x0, noise, denoise_fn=den, num_steps=256, sigma_min=smin)
    d = 3 * N
    ratio = (s**2 + smax**2) / (s**2 + smin**2)
    xT = x0.double() * math.sqrt(ratio)
    prior = (-0.5 * (xT / smax) ** 2 - math.log(smax) - 0.5 * math.log(2 * math.pi)).flatten(1).sum(1)
    want = prior + 0.5 * d * math.log(ratio)
    assert torch.allclose(out["latent"], xT, rtol=2e-3, atol=1e-3)
    assert torch.allclose(out["delta_jacobian"], torch.full((B,), 0.5 * d * math.log(ratio), dtype=torch.float64), rtol=2e-3)
    assert torch.allclose(out["logp"], want, rtol=3e-3)
    # and it is close to the exact density of the data smoothed at smin (the prior N(0, smax) vs N(0, s^2 + smax^2) differs by O(s^2 / smax^2))
    exact = (-0.5 * x0.double() ** 2 / (s**2 + smin**2) - 0.5 * math.log(2 * math.pi * (s**2 + smin**2))).flatten(1).sum(1)
    assert torch.allclose(out["logp"], exact, rtol=5e-3, atol=0.5)


@pytest.mark.parametrize("kind,reparam", [("uncond", "gaussian"), ("cond", "gaussian"), ("cond", "uvl")])
def test_log_likelihood_host_logic_matches_oracle(kind, reparam):
    import gecco_b200 as G

    torch.manual_seed(0)
    B, N, L, steps = 2, 32, 2, 3
    if reparam == "uvl":
        rp, Kc, smax = synth.UVL_REPARAM, synth.K_TASKONOMY, 180.0
    else:
        rp, Kc, smax = (synth.SHAPENET_VOL_REPARAM if kind == "cond" else synth.UNCOND_REPARAM), synth.K_SHAPENET, 165.0
    feats = synth.synth_features(B, (34, 17, 8), 21) if kind == "cond" else None
    sd = synth.tame(synth.full_state_dict(kind, reparam, rp["mean"], rp["sigma"], 77, n_layers=L), 0.15)
    model = build(kind, reparam, rp["mean"], rp["sigma"], smax, None, "cpu", feats, n_layers=L, state_dict=sd)
    cfg = O.OracleConfig(kind=kind, reparam=reparam, sigma_max=smax, n_layers=L)
    K = synth.camera(B, Kc) if kind == "cond" else None
    ctx = G.Context3d(image=torch.zeros(B, 3, 8, 8), K=K) if kind == "cond" else None
    data = O.diffusion_to_data(cfg, sd, torch.randn(B, N, 3, generator=synth.gen(2)) * 0.8, K)
    noise = (torch.randint(0, 2, (1, B, N, 3), generator=synth.gen(3)) * 2 - 1).float()
    ref = O.log_likelihood(cfg, sd, data, noise, feats, K, num_steps=steps)
    got = model.log_likelihood(data, ctx, noise=noise, num_steps=steps, return_details=True)
    for k in ("prior_logp", "delta_reparam", "delta_jacobian", "logp"):
        assert torch.allclose(got[k], ref[k], rtol=2e-4, atol=2e-3), (k, got[k], ref[k])
    assert torch.allclose(got["latent"], ref["latent"], rtol=1e-4, atol=1e-4)
    # drawn probes: deterministic for a seeded generator, [B] float64
    a = model.log_likelihood(data, ctx, rng=synth.gen(9), num_steps=2)
    b = model.log_likelihood(data, ctx, rng=synth.gen(9), num_steps=2)
    assert a.shape == (B,) and a.dtype == torch.float64 and torch.equal(a, b)
