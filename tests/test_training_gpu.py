"""Training step (BASELINE config 5) on the GPU:
  * TCLinear (tcgen05 forward + input gradient, cuBLAS weight gradient) against F.linear's autograd;
  * the EDM loss and its GRADIENTS w.r.t. every parameter against goldens from the unmodified reference
    (`oracle/make_golden.py --grads`, fp32 CPU): loss 1e-2, total gradient norm 3e-2, per-parameter norms and directions;
  * the differentiable path against the fused sampling engine on the same input (the two implementations of the network);
  * the fused Adam + EMA kernel against torch.optim.Adam and the reference's EMA formula (ema.py:187-194);
  * Trainer.step: the loss of a fixed batch goes down and the EMA weights trail the trained ones.
"""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from tests import synth
from tests.models_b200 import build

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def rms(t):
    return t.double().pow(2).mean().sqrt().item()


@pytest.mark.parametrize("M,K,N,bias", [(2048, 384, 1152, True), (1000, 384, 768, False), (130, 768, 384, True)])
def test_tc_linear_matches_autograd(cuda, M, K, N, bias):
    from gecco_b200.training import TCLinear

    g = synth.gen(5)
    x = torch.randn(M, K, generator=g).to(cuda).requires_grad_()
    w = (torch.randn(N, K, generator=g) / K**0.5).to(cuda).requires_grad_()
    b = torch.randn(N, generator=g).to(cuda).requires_grad_() if bias else None
    dy = torch.randn(M, N, generator=g).to(cuda)
    y = TCLinear.apply(x, w, b)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), None if b is None else b.grad.clone())
    x.grad = w.grad = None
    if b is not None:
        b.grad = None
    # fp32 reference on the same bf16-rounded operands
    xr, wr = x.detach().bfloat16().float().requires_grad_(), w.detach().bfloat16().float().requires_grad_()
    br = None if b is None else b.detach().clone().requires_grad_()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yr = F.linear(xr, wr, br)
        yr.backward(dy.bfloat16().float())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert rms(got[0] - yr.detach()) <= 2e-3 * rms(yr)
    assert rms(got[1] - xr.grad) <= 2e-3 * rms(xr.grad)
    assert rms(got[2] - wr.grad) <= 5e-3 * rms(wr.grad)
    if b is not None:
        assert rms(got[3] - dy.sum(0)) <= 1e-5 * rms(dy.sum(0))


def _setup(name, cuda):
    import gecco_b200 as G

    g = torch.load(GOLD / (name + ".pt"), weights_only=False)
    r = g["recipe"]
    cond = r["kind"] == "cond"
    model = build(r["kind"], r["reparam"], r["mean"], r["sigma"], r["sigma_max"], r["weight_seed"], cuda,
                  convnext_seed=r["convnext_seed"] if cond else None)
    ctx = None
    if cond:
        img = torch.rand(r["B"], 3, r["image"], r["image"], generator=synth.gen(r["image_seed"]))
        ctx = G.Context3d(image=img.to(cuda), K=synth.camera(r["B"], r["K"]).to(cuda))
    ex = model.reparam.diffusion_to_data(torch.randn(r["B"], r["N"], 3, generator=synth.gen(r["ex_seed"])).to(cuda), ctx)
    return g, r, model, ctx, ex


@pytest.fixture()
def _exact():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", ["grads_uncond", "grads_cond_gaussian"])
def test_loss_and_gradients_match_reference(cuda, monkeypatch, _exact, name):
    g, r, model, ctx, ex = _setup(name, cuda)
    torch.manual_seed(r["loss_seed"])
    u = torch.rand(r["B"])
    noise = torch.randn(r["B"], r["N"], 3)
    assert torch.equal(u, g["loss_u"]) and torch.equal(noise[:, ::64], g["loss_noise_sub"])
    monkeypatch.setattr(torch, "rand", lambda n, device=None, **kw: u.to(device))
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: noise.to(t.device, t.dtype))
    model.train()
    model.conditioner.eval()
    loss = model.training_step((ex, ctx), 0)
    loss.backward()
    monkeypatch.undo()
    rel = abs(loss.item() - g["loss"].item()) / abs(g["loss"].item())
    print(f"{name}: loss {loss.item():.6g} vs reference {g['loss'].item():.6g} (rel {rel:.2e})")
    assert rel < 1e-2
    norms, subs = g["grad_norm"], g["grad_sub"]
    total_ref = sum(v * v for v in norms.values()) ** 0.5
    got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(norms), sorted(set(got) ^ set(norms))
    total = sum(v.double().pow(2).sum().item() for v in got.values()) ** 0.5
    print(f"{name}: total gradient norm {total:.6g} vs reference {total_ref:.6g}")
    assert abs(total - total_ref) <= 3e-2 * total_ref
    worst_norm, worst_cos, worst_key = 0.0, 1.0, None
    for k, gr in got.items():
        n_ref = norms[k]
        if n_ref < 1e-3 * total_ref:  # negligible gradients are compared on the absolute scale of the step
            assert gr.double().norm().item() < 2e-3 * total_ref, k
            continue
        # small gradients (scalars such as the activation widths, biases) are judged on the scale of the step
        e_norm = abs(gr.double().norm().item() - n_ref) / max(n_ref, 1e-2 * total_ref)
        if e_norm > worst_norm:
            worst_norm, worst_key = e_norm, k
        flat = gr.detach().flatten()
        stride = max(1, flat.numel() // 256)
        a, b = flat[::stride].cpu().double(), subs[k].double()
        if b.numel() >= 16 and b.norm() > 0:
            worst_cos = min(worst_cos, (a @ b / (a.norm() * b.norm())).item())
    print(f"{name}: worst per-parameter norm error {worst_norm:.3e} ({worst_key}), worst direction cosine {worst_cos:.5f}")
    assert worst_norm < 5e-2 and worst_cos > 0.97


def test_training_path_matches_engine(cuda):
    """The differentiable network and the fused sampling engine are two implementations of the same function."""
    g, r, model, ctx, ex = _setup("grads_cond_gaussian", cuda)
    sig = torch.tensor([0.3, 5.0], device=cuda)
    x = synth.noisy_input(r["B"], r["N"], sig.cpu(), 51, 54).to(cuda)
    with torch.no_grad():
        D_engine = model(x, sig, ctx)
    model.train()
    model.conditioner.eval()
    D_train = model(x, sig, ctx)
    assert D_train.requires_grad
    e = rms(D_train.detach() - D_engine) / rms(D_engine)
    print(f"training path vs engine: rel rms {e:.2e}")
    assert e < 1e-2
    with pytest.raises(ValueError):
        model(x, sig, ctx, do_cache=True)


def test_adam_ema_kernel_matches_torch(cuda):
    from gecco_b200 import ops

    g = synth.gen(9)
    n = 100003  # not a multiple of 4: scalar tail
    p0 = torch.randn(n + 1, generator=g).to(cuda)[:n + 1]
    p = p0[:n].clone()
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=3e-3)
    m, v, ema = torch.zeros_like(p), torch.zeros_like(p), p.clone()
    ema_ref = p.clone()
    decay, world = 0.99, 4
    for step in range(1, 6):
        grad = torch.randn(n, generator=g).to(cuda) * (10.0 ** (step - 3))
        ref_p.grad = grad.clone()
        opt.step()
        ema_ref.mul_(decay).add_(ref_p.detach(), alpha=1 - decay)
        ops.adam_ema_step(p, grad * world, m, v, ema, step, lr=3e-3, grad_scale=1.0 / world, ema_decay=decay)
        assert torch.allclose(p, ref_p.detach(), rtol=2e-5, atol=2e-6), (step, (p - ref_p.detach()).abs().max())
        assert torch.allclose(ema, ema_ref, rtol=2e-5, atol=2e-6)
    st = opt.state[ref_p]
    assert torch.allclose(m, st["exp_avg"], rtol=1e-5, atol=1e-6 * m.abs().max().item())
    assert torch.allclose(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-6 * v.abs().max().item())


def test_trainer_reduces_loss(cuda):
    from gecco_b200.training import Trainer

    g, r, model, ctx, ex = _setup("grads_uncond", cuda)
    tr = Trainer(model, lr=5e-5, ema_decay=0.9)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    gen_state = torch.random.get_rng_state()
    losses = []
    for _ in range(8):
        torch.manual_seed(7)  # the same noise levels and noise every step: a fixed objective
        losses.append(tr.step((ex, ctx)).item())
    torch.random.set_rng_state(gen_state)
    print("losses", [round(l, 3) for l in losses])
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < 0.9 * losses[0]
    moved = sum((model.state_dict()[k] - before[k]).abs().sum().item() for k in before if before[k].is_floating_point())
    assert moved > 0
    ema_sd = tr.ema_state_dict()
    k = "backbone.model.inner.layers.0.mlp.0.weight"
    cur = model.state_dict()[k]
    d_cur, d_ema = (cur - before[k]).norm().item(), (ema_sd[k] - before[k]).norm().item()
    assert 0 < d_ema < d_cur  # the EMA trails the trained weights
    assert torch.equal(model.state_dict()[k], cur)  # the swap restored the trained weights
    # the sampling engine sees the trained weights (its snapshot is keyed on content)
    model.eval()
    with torch.no_grad():
        s = model.sample_stochastic((2, 256, 3), None, rng=torch.Generator(cuda).manual_seed(0), num_steps=4)
    assert torch.isfinite(s).all()


def test_graphed_step_matches_eager(cuda):
    """The CUDA-graph replay of forward + backward produces the eager step's loss and gradients (fixed draws), keeps
    working when the batch changes, and trains."""
    import gecco_b200 as G
    from gecco_b200.training import Trainer

    class FixedDrawLoss(G.EDMLoss):  # the reference loss with its two draws replaced by stored device tensors
        def forward(self, net, examples, context):
            ex_diff = net.reparam.data_to_diffusion(examples, context)
            sigma = self.fixed_sigma
            weight = (sigma**2 + self.sigma_data**2) / ((sigma * self.sigma_data) ** 2)
            D = net(ex_diff + self.fixed_noise * sigma, sigma, context)
            return (self.loss_scale * weight * (D - ex_diff) ** 2).mean()

    results = []
    for graph in (False, True):
        g, r, model, ctx, ex = _setup("grads_uncond", cuda)
        loss_mod = FixedDrawLoss(schedule=model.loss.schedule)
        loss_mod.fixed_sigma = torch.tensor([0.4, 6.0], device=cuda).reshape(-1, 1, 1)
        loss_mod.fixed_noise = torch.randn(r["B"], r["N"], 3, generator=synth.gen(4)).to(cuda)
        model.loss = loss_mod
        tr = Trainer(model, lr=5e-5, graph=graph)
        l0 = tr.step((ex, ctx))
        g0 = tr.state.g.clone()
        ex2 = ex.flip(0).contiguous()  # a different batch of the same shape goes through the static buffers
        l1 = tr.step((ex2, ctx))
        results.append((l0.item(), g0, l1.item(), tr.state.p.clone()))
    (a0, ga, a1, pa), (b0, gb, b1, pb) = results
    print(f"eager losses {a0:.6g} {a1:.6g}; graphed {b0:.6g} {b1:.6g}")
    assert abs(a0 - b0) <= 1e-5 * abs(a0) and abs(a1 - b1) <= 1e-3 * abs(a1)
    assert rms(ga - gb) <= 1e-4 * rms(ga)
    assert rms(pa - pb) <= 1e-4 * rms(pa)


@pytest.mark.parametrize("B,N,C,groups", [(3, 200, 384, 32), (2, 77, 672, 16), (1, 64, 384, 32)])
def test_group_affine_norm_kernels_match_torch(cuda, B, N, C, groups):
    """GroupAffineNorm (gecco_group_stats + gecco_train_affine / gecco_train_colsum2) against the torch expression of
    models/normalization.py:36-44 in float64: output and the gradients of x, gamma, beta."""
    from gecco_b200 import training as T

    g = torch.Generator("cpu").manual_seed(B * 1000 + N)
    x = (torch.randn(B, N, C, generator=g) * 1.7 + 0.4).to(cuda).requires_grad_(True)
    gamma = (torch.randn(B, C, generator=g) * 0.3 + 1.0).to(cuda).requires_grad_(True)
    beta = (torch.randn(B, C, generator=g) * 0.2).to(cuda).requires_grad_(True)
    dy = torch.randn(B, N, C, generator=g).to(cuda)
    y = T.GroupAffineNorm.apply(x, gamma, beta, groups, 1e-5)
    gx, gg, gb = torch.autograd.grad(y, (x, gamma, beta), dy)
    xd, gd, bd = (t.detach().double().requires_grad_(True) for t in (x, gamma, beta))
    ref = gd[:, None] * torch.nn.functional.group_norm(xd.transpose(1, 2), groups, eps=1e-5).transpose(1, 2) + bd[:, None]
    rx, rg, rb = torch.autograd.grad(ref, (xd, gd, bd), dy.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    errs = dict(y=rel(y, ref), dx=rel(gx, rx), dgamma=rel(gg, rg), dbeta=rel(gb, rb))
    print("GroupAffineNorm", (B, N, C, groups), errs)
    assert max(errs.values()) < 2e-5, errs


@pytest.mark.parametrize("normalized", [True, False])
def test_gauss_act_kernels_match_torch(cuda, normalized):
    """GaussAct (models/activation.py:17-24) forward / backward kernels against torch in float64, odd element count."""
    from gecco_b200 import training as T

    g = torch.Generator("cpu").manual_seed(5)
    x = (torch.randn(3, 37, 771, generator=g) * 1.5).to(cuda).requires_grad_(True)
    alpha = torch.tensor(1.3, device=cuda, requires_grad=True)
    dy = torch.randn(3, 37, 771, generator=g).to(cuda)
    y = T.GaussAct.apply(x, alpha, normalized)
    gx, ga = torch.autograd.grad(y, (x, alpha), dy)
    xd, ad = x.detach().double().requires_grad_(True), alpha.detach().double().requires_grad_(True)
    ref = (-(xd**2) / (2 * ad**2)).exp()
    if normalized:
        ref = (ref - 0.7) / 0.28
    rx, ra = torch.autograd.grad(ref, (xd, ad), dy.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    errs = dict(y=rel(y, ref), dx=rel(gx, rx), dalpha=abs(ga.item() - ra.item()) / abs(ra.item()))
    print("GaussAct", normalized, errs)
    assert errs["y"] < 5e-6 and errs["dx"] < 5e-6 and errs["dalpha"] < 1e-4, errs
