"""Fused point-side MLP (gecco_mlp: GEMM -> Gaussian activation -> GEMM -> + residual, hidden on chip) against a plain
fp32 torch reference on the same bf16-rounded operands (models/set_transformer.py:165-166, models/mlp.py:5-39)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w1, b1, alpha, w2, b2, res):
    h = a.float() @ w1.float().t() + b1
    h = ((-(h**2) / (2 * alpha**2)).exp() - 0.7) / 0.28
    h = h.bfloat16().float()  # the hidden activation is the bf16 A operand of the second projection
    return res + h @ w2.float().t() + b2


@pytest.mark.parametrize("clouds,Np,N,hidden", [(1, 256, 256, 768), (3, 256, 200, 768), (2, 512, 512, 256), (40, 512, 509, 768)])
def test_mlp_fused(cuda, clouds, Np, N, hidden):
    from gecco_b200 import ops

    C = 384
    g = torch.Generator(device="cpu").manual_seed(clouds * 1000 + N + hidden)
    a = torch.randn(clouds * Np, C, generator=g)
    a.view(clouds, Np, C)[:, N:] = 0
    a = a.to(cuda).bfloat16()
    w1 = (torch.randn(clouds * hidden, C, generator=g) / math.sqrt(C)).to(cuda).bfloat16()
    b1 = (0.3 * torch.randn(clouds, hidden, generator=g)).to(cuda)
    w2 = (torch.randn(C, hidden, generator=g) / math.sqrt(hidden)).to(cuda).bfloat16()
    b2 = torch.randn(C, generator=g).to(cuda)
    res = torch.randn(clouds * Np, C, generator=g).to(cuda)
    stats = torch.zeros(clouds, C // 12, 2, dtype=torch.float64, device=cuda)
    x = res.clone()
    xb = torch.full((clouds * Np, C), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.mlp(a, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=N,
            w1_rows_per_cloud=hidden, b1_stride=hidden)
    torch.cuda.synchronize()
    xv, xbv = x.view(clouds, Np, C), xb.view(clouds, Np, C)
    for b in range(clouds):
        sl = slice(b * Np, (b + 1) * Np)
        ref = _ref(a[sl], w1[b * hidden:(b + 1) * hidden], b1[b], 1.3, w2, b2, res[sl])
        # bf16 rounding of the hidden activation may flip by one ulp against the reference (ex2.approx): 768 terms of
        # magnitude <= 2.6 * 2^-8 * |w2| average out to ~1e-3
        assert (xv[b, :N] - ref[:N]).abs().max().item() < 8e-3, (b, (xv[b, :N] - ref[:N]).abs().max().item())
        assert (xbv[b, :N].float() - ref[:N]).abs().max().item() < 5e-2
        v = ref[:N].view(N, C // 12, 12).double()
        s1, s2 = v.sum(dim=(0, 2)), (v * v).sum(dim=(0, 2))
        assert torch.allclose(stats[b, :, 0], s1, rtol=1e-3, atol=0.5), (stats[b, :, 0] - s1).abs().max()
        assert torch.allclose(stats[b, :, 1], s2, rtol=1e-3, atol=0.5), (stats[b, :, 1] - s2).abs().max()
    if N < Np:
        assert xv[:, N:].abs().max().item() == 0.0 and xbv[:, N:].abs().max().item() == 0.0


def test_mlp_fused_shared_weights_no_stats(cuda):
    from gecco_b200 import ops

    C, hidden, M = 384, 768, 1024
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(M, C, generator=g).to(cuda).bfloat16()
    w1 = (torch.randn(hidden, C, generator=g) / math.sqrt(C)).to(cuda).bfloat16()
    b1 = (0.3 * torch.randn(hidden, generator=g)).to(cuda)
    w2 = (torch.randn(C, hidden, generator=g) / math.sqrt(hidden)).to(cuda).bfloat16()
    b2 = torch.randn(C, generator=g).to(cuda)
    res = torch.randn(M, C, generator=g).to(cuda)
    out, _ = ops.mlp(a, w1, b1, 0.9, w2, b2, res, rows_per_cloud=256)
    torch.cuda.synchronize()
    ref = _ref(a, w1, b1, 0.9, w2, b2, res)
    assert (out - ref).abs().max().item() < 8e-3


def test_mlp_fused_rejects_unsupported_shapes(cuda):
    from gecco_b200 import ops

    a = torch.zeros(128, 384, device=cuda, dtype=torch.bfloat16)
    w1 = torch.zeros(768, 384, device=cuda, dtype=torch.bfloat16)
    w2 = torch.zeros(384, 768, device=cuda, dtype=torch.bfloat16)
    b1 = torch.zeros(768, device=cuda)
    b2 = torch.zeros(384, device=cuda)
    res = torch.zeros(128, 384, device=cuda)
    with pytest.raises(ValueError):
        ops.mlp(a, w1, b1, 1.0, w2, b2, res, rows_per_cloud=128)


def _adagn_ref(x, stats_rows, t, sw, sb, bw, bb, clouds, Np, N):
    import torch.nn.functional as F

    C = x.shape[1]
    xv = x.view(clouds, Np, C)[:, :N].float()
    normed = F.group_norm(xv.transpose(1, 2), 32, eps=1e-5).transpose(1, 2)
    scale = t[:, None] * sw[None] + sb[None]
    bias = t[:, None] * bw[None] + bb[None]
    out = torch.zeros(clouds, Np, C, device=x.device)
    out[:, :N] = scale[:, None] * normed + bias[:, None]
    return out.view(clouds * Np, C)


@pytest.mark.parametrize("clouds,Np,N", [(1, 2048, 2048), (3, 2048, 2000), (8, 256, 256), (40, 512, 509), (64, 2048, 2048)])
def test_mlp_pair_anorm(cuda, clouds, Np, N):
    """CTA-pair MLP (mlp_pair.cu): AdaGN on the A operand, hidden tile through the L2 scratch, residual + statistics."""
    from gecco_b200 import ops

    C, hidden = 384, 768
    g = torch.Generator(device="cpu").manual_seed(clouds * 1000 + N)
    xf = torch.randn(clouds * Np, C, generator=g) * 1.5 + 0.3
    xf.view(clouds, Np, C)[:, N:] = 0
    xf = xf.to(cuda)
    xb = xf.bfloat16()
    w1 = (torch.randn(hidden, C, generator=g) / math.sqrt(C)).to(cuda).bfloat16()
    b1 = (0.3 * torch.randn(hidden, generator=g)).to(cuda)
    w2 = (torch.randn(C, hidden, generator=g) / math.sqrt(hidden)).to(cuda).bfloat16()
    b2 = torch.randn(C, generator=g).to(cuda)
    t = torch.randn(clouds, generator=g).to(cuda)
    sw, sb, bw, bb = [(torch.randn(C, generator=g) * k + o).to(cuda) for k, o in ((0.5, 0), (0.3, 1), (0.5, 0), (0.3, 0))]
    stats_in = ops.group_stats(xf, Np, N, 12)
    an = dict(stats=stats_in, t=t, scale_w=sw, scale_b=sb, bias_w=bw, bias_b=bb, groups=32)
    stats = torch.zeros(clouds, C // 12, 2, dtype=torch.float64, device=cuda)
    x = xf.clone()
    xo = torch.full((clouds * Np, C), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.mlp(xb, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xo, stats=stats, rows_per_cloud=Np, valid_rows=N, anorm=an)
    torch.cuda.synchronize()
    # reference: AdaGN of the bf16 copy with the fp32 tensor's statistics, rounded once to bf16 (what the kernel multiplies)
    a_n = _adagn_ref(xb.float(), None, t, sw, sb, bw, bb, clouds, Np, N)
    # (group statistics of the bf16 copy differ from those of the fp32 tensor by O(2^-9) relative: inside the tolerance)
    ref = _ref(a_n.bfloat16(), w1, b1, 1.3, w2, b2, xf)
    xv, rv = x.view(clouds, Np, C)[:, :N], ref.view(clouds, Np, C)[:, :N]
    err = (xv - rv).abs().max().item()
    assert err < 3e-2, err
    rms = (xv - rv).pow(2).mean().sqrt().item() / (rv - xf.view(clouds, Np, C)[:, :N]).pow(2).mean().sqrt().item()
    assert rms < 6e-3, rms
    assert torch.equal(xo.view(clouds, Np, C)[:, :N], xv.bfloat16())
    v = xv.double().view(clouds, N, C // 12, 12)
    assert torch.allclose(stats[:, :, 0], v.sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)
    assert torch.allclose(stats[:, :, 1], (v * v).sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)
    if N < Np:
        assert x.view(clouds, Np, C)[:, N:].abs().max().item() == 0.0 and xo.view(clouds, Np, C)[:, N:].abs().max().item() == 0.0


def test_mlp_pair_matches_two_gemms_in_engine(cuda):
    """One evaluation at a shape the pair kernels take (N = 2048) with the fused MLP against the two-GEMM path."""
    from gecco_b200 import ops
    from tests import synth
    from tests.models_b200 import build as build_model

    B, N = 2, 2048
    model = build_model("uncond", "gaussian", [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 80.0, 78, cuda, None)
    x = (torch.randn(B, N, 3, generator=synth.gen(6)) * 1.5).to(cuda)
    sig = torch.tensor([0.2, 5.0], device=cuda)
    try:
        ops.set_option("mlp_pair", 0)
        d0 = model(x, sig, None)
        ops.set_option("mlp_pair", 1)
        d1 = model(x, sig, None)
    finally:
        ops.set_option("mlp_pair", 1)
    torch.cuda.synchronize()
    rms = lambda v: v.float().pow(2).mean().sqrt().item()
    # same arithmetic (bf16 hidden, fp32 accumulation); only the statistics' atomic order differs
    assert rms(d1 - d0) < 2e-3 * rms(d0), rms(d1 - d0) / rms(d0)
