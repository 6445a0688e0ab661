"""tcgen05 GEMM (gecco_gemm) against a plain fp32 torch reference on the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None, alpha=None, res=None):
    out = a.float() @ w.float().t()
    if bias is not None:
        out = out + bias
    if alpha is not None:
        out = ((-(out**2) / (2 * alpha**2)).exp() - 0.7) / 0.28
    if res is not None:
        out = out + res
    return out


@pytest.mark.parametrize("m,n,k", [(128, 192, 64), (256, 384, 384), (1000, 768, 768), (4096, 1152, 384), (300, 200, 136)])
def test_gemm_plain(cuda, m, n, k):
    from gecco_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(cuda).bfloat16()
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda).bfloat16()
    o32, o16 = ops.gemm(a, w, out_f32=True, out_bf16=True)
    torch.cuda.synchronize()
    ref = _ref(a, w)
    err = (o32 - ref).abs().max().item()
    assert err < 2e-3, err
    assert (o16.float() - ref).abs().max().item() < 3e-2


def test_gemm_epilogue_full(cuda):
    from gecco_b200 import ops

    B, N, Np, C = 3, 200, 256, 384
    g = torch.Generator(device="cpu").manual_seed(7)
    a = torch.randn(B * Np, 768, generator=g).to(cuda).bfloat16()
    w = (torch.randn(C, 768, generator=g) / math.sqrt(768)).to(cuda).bfloat16()
    bias = torch.randn(C, generator=g).to(cuda)
    res = torch.randn(B * Np, C, generator=g).to(cuda)
    stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=cuda)
    x = res.clone()
    xb = torch.full((B * Np, C), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, res=x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=N)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias=bias, res=res)
    xv, rv = x.view(B, Np, C), ref.view(B, Np, C)
    assert (xv[:, :N] - rv[:, :N]).abs().max().item() < 2e-3
    assert (xb.view(B, Np, C)[:, :N].float() - rv[:, :N]).abs().max().item() < 4e-2
    # padding rows are written as exact zeros (they feed the next projection as keys of no weight)
    assert xv[:, N:].abs().max().item() == 0.0 and xb.view(B, Np, C)[:, N:].abs().max().item() == 0.0
    v = ref.view(B, Np, C // 12, 12)[:, :N].double()
    s1 = v.sum(dim=(1, 3))
    s2 = (v * v).sum(dim=(1, 3))
    assert torch.allclose(stats[..., 0], s1, rtol=1e-4, atol=1e-2), (stats[..., 0] - s1).abs().max()
    assert torch.allclose(stats[..., 1], s2, rtol=1e-4, atol=1e-2), (stats[..., 1] - s2).abs().max()


def test_gemm_act_and_percloud_weights(cuda):
    from gecco_b200 import ops

    B, Np, K, C = 4, 128, 704, 384
    g = torch.Generator(device="cpu").manual_seed(11)
    a = torch.randn(B * Np, K, generator=g).to(cuda).bfloat16()
    w = (torch.randn(B * C, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    bias = torch.randn(B, C, generator=g).to(cuda)
    o32, _ = ops.gemm(a, w, bias=bias, bias_stride=C, act_alpha=1.3, out_f32=True, rows_per_cloud=Np,
                      w_rows_per_cloud=C, n_out=C)
    torch.cuda.synchronize()
    for b in range(B):
        ref = _ref(a[b * Np:(b + 1) * Np], w[b * C:(b + 1) * C], bias=bias[b], alpha=1.3)
        assert (o32[b * Np:(b + 1) * Np] - ref).abs().max().item() < 5e-3


def test_gemm_xyz_embed(cuda):
    from gecco_b200 import ops

    B, N, Np, K, C = 2, 100, 128, 128, 192
    g = torch.Generator(device="cpu").manual_seed(13)
    a = torch.randn(B * Np, K, generator=g).to(cuda).bfloat16()
    w = (torch.randn(C, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    bias = torch.randn(C, generator=g).to(cuda)
    geom = torch.randn(B, N, 3, generator=g).to(cuda)
    wx = torch.randn(C, 3, generator=g).to(cuda)
    sigma = torch.tensor([0.5, 7.0], device=cuda)
    o32, _ = ops.gemm(a, w, bias=bias, out_f32=True, rows_per_cloud=Np, valid_rows=N, geom=geom, sigma=sigma,
                      sigma_stride=1, wx=wx)
    torch.cuda.synchronize()
    c_in = 1 / (1 + sigma**2).sqrt()
    gg = geom * c_in.view(B, 1, 1)
    ref = _ref(a, w, bias=bias).view(B, Np, C)
    ref[:, :N] += gg @ wx.t()
    assert (o32.view(B, Np, C)[:, :N] - ref[:, :N]).abs().max().item() < 2e-3
    assert o32.view(B, Np, C)[:, N:].abs().max().item() == 0.0


@pytest.mark.parametrize("m,n,k", [(8192, 384, 768), (148 * 128 * 2 + 128, 384, 384), (8192, 384, 384), (2048 * 40, 384, 320)])
def test_gemm_persistent_residual_inplace(cuda, m, n, k):
    """Several tiles per CTA with the in-place residual stream: exercises the residual prefetch ring, the double
    buffered staging and the TMEM slot hand-over across tiles."""
    from gecco_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(m + k)
    a = torch.randn(m, k, generator=g).to(cuda).bfloat16()
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda).bfloat16()
    bias = torch.randn(n, generator=g).to(cuda)
    res = torch.randn(m, n, generator=g).to(cuda)
    Np = 2048 if m % 2048 == 0 else 128
    stats = torch.zeros(m // Np, n // 12, 2, dtype=torch.float64, device=cuda)
    x = res.clone()
    xb = torch.empty(m, n, device=cuda, dtype=torch.bfloat16)
    for _ in range(2):  # twice: x is updated in place, so the second run checks the first one's stores too
        ref = _ref(a, w, bias=bias, res=x.clone())
        stats.zero_()
        ops.gemm(a, w, bias=bias, res=x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=Np)
        torch.cuda.synchronize()
        assert (x - ref).abs().max().item() < 3e-3
        assert (xb.float() - ref).abs().max().item() < 6e-2
        v = ref.view(m // Np, Np, n // 12, 12).double()
        assert torch.allclose(stats[..., 0], v.sum(dim=(1, 3)), rtol=1e-4, atol=5e-2)
        assert torch.allclose(stats[..., 1], (v * v).sum(dim=(1, 3)), rtol=1e-4, atol=5e-2)


@pytest.mark.parametrize("fast,B,O", [("1", 3, 1152), ("0", 3, 1152), ("1", 5, 776), ("0", 5, 776)])
def test_fold_adagn_matches_adagn_then_linear(cuda, monkeypatch, fast, B, O):
    """AdaGN -> Linear folded into per-cloud weights (gecco_fold_adagn: the register-prefetching C = 384 kernel and the
    generic one; cloud groups and row blocks with tails) + per-cloud GEMM on the raw stream == Linear(AdaGN(x))
    (models/normalization.py:36-44)."""
    from gecco_b200 import ops

    monkeypatch.setenv("GECCO_FOLD_FAST", fast)
    N, Np, C = 200, 256, 384
    g = torch.Generator(device="cpu").manual_seed(21)
    x = (torch.randn(B, Np, C, generator=g) * 1.7 + 0.3).to(cuda)
    x[:, N:] = 0
    t = torch.randn(B, generator=g).to(cuda)
    sw, sb, bw, bb = [torch.randn(C, generator=g).to(cuda) for _ in range(4)]
    W = (torch.randn(O, C, generator=g) / math.sqrt(C)).to(cuda)
    bias = torch.randn(O, generator=g).to(cuda)
    stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=cuda)
    ops.group_stats(x.view(B * Np, C), rows_per_cloud=Np, valid_rows=N, group_size=12, stats=stats)
    wf, bf = ops.fold_adagn(W, bias, stats, t, sw, sb, bw, bb, clouds=B, valid_rows=N)
    out, _ = ops.gemm(x.view(B * Np, C).bfloat16(), wf.view(B * O, C), bias=bf, bias_stride=O, out_f32=True,
                      rows_per_cloud=Np, valid_rows=N, w_rows_per_cloud=O, n_out=O)
    torch.cuda.synchronize()
    xv = x[:, :N].view(B, N, 32, 12)
    mean = xv.mean(dim=(1, 3), keepdim=True)
    var = xv.var(dim=(1, 3), unbiased=False, keepdim=True)
    xn = ((xv - mean) / (var + 1e-5).sqrt()).view(B, N, C)
    y = xn * (t[:, None, None] * sw + sb) + (t[:, None, None] * bw + bb)
    ref = y @ W.t() + bias
    got = out.view(B, Np, O)[:, :N]
    rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel < 6e-3, rel


@pytest.mark.parametrize("pairs", [1, 0])
def test_gemm_pair_percloud_act(cuda, pairs):
    """K <= 384, M % 256 == 0: the CTA-pair (cta_group::2) kernel with per-cloud weights, bias and activation; the same
    problem through the single-CTA kernel must agree."""
    import ctypes

    from gecco_b200 import _abi, ops

    B, Np, N, K, C = 5, 2048, 1900, 384, 768
    g = torch.Generator(device="cpu").manual_seed(17)
    a = torch.randn(B * Np, K, generator=g).to(cuda).bfloat16()
    w = (torch.randn(B * C, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    bias = torch.randn(B, C, generator=g).to(cuda)
    lib = _abi.load()
    _abi.check(lib.gecco_set_option(ctypes.c_char_p(b"gemm_pairs"), pairs))
    try:
        _, o16 = ops.gemm(a, w, bias=bias, bias_stride=C, act_alpha=1.3, out_bf16=True, rows_per_cloud=Np, valid_rows=N,
                          w_rows_per_cloud=C, n_out=C)
        torch.cuda.synchronize()
    finally:
        _abi.check(lib.gecco_set_option(ctypes.c_char_p(b"gemm_pairs"), 1))
    for b in range(B):
        ref = _ref(a[b * Np:b * Np + N], w[b * C:(b + 1) * C], bias=bias[b], alpha=1.3)
        got = o16[b * Np:b * Np + N].float()
        assert (got - ref).abs().max().item() < 4e-2, b
        assert o16[b * Np + N:(b + 1) * Np].abs().max().item() == 0.0


@pytest.mark.parametrize("src", ["bf16"])
@pytest.mark.parametrize("n_out,act", [(1152, None), (768, 1.3), (384, None)])
@pytest.mark.parametrize("clouds,rows_per_cloud,valid", [(3, 768, 700), (16, 2048, 2048)])
def test_gemm_anorm_operand(cuda, n_out, act, clouds, rows_per_cloud, valid, src):
    """AdaGN applied to the A operand inside the CTA-pair GEMM (gecco_anorm): bf16 x -> a * x + s in fp32 -> bf16, in place
    in shared memory -> tcgen05, against AdaGN in torch (models/normalization.py:36-44) of the same bf16 x followed by
    the same bf16 projection."""
    import torch.nn.functional as F

    from gecco_b200 import ops

    K, groups = 384, 32
    m = clouds * rows_per_cloud
    assert ops.gemm_anorm_supported(m, rows_per_cloud, n_out, K)
    g = torch.Generator(device="cpu").manual_seed(n_out + clouds)
    x = (torch.randn(clouds, rows_per_cloud, K, generator=g) * 0.7 + torch.randn(1, 1, K, generator=g) * (1.5 if src == "bf16" else 6.0)).to(cuda)
    x[:, valid:] = 0.0  # padding rows of the residual stream are zeros
    if src == "bf16":
        x = x.bfloat16().float()  # the operand the kernel reads is the bf16 copy of the residual stream
    x2 = x.view(m, K)
    w = (torch.randn(n_out, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    bias = torch.randn(n_out, generator=g).to(cuda)
    t = (torch.randn(clouds, generator=g) * 0.8).to(cuda)
    sw, sb = (torch.randn(K, 1, generator=g) * 0.3).to(cuda), (torch.randn(K, generator=g) * 0.1 + 1).to(cuda)
    bw, bb = (torch.randn(K, 1, generator=g) * 0.3).to(cuda), (torch.randn(K, generator=g) * 0.1).to(cuda)
    stats = ops.group_stats(x2, rows_per_cloud, valid, 12)
    extra = dict(x=x2) if src == "fp32" else {}
    _, o16 = ops.gemm(None if src == "fp32" else x2.bfloat16(), w, bias=bias, act_alpha=act, out_bf16=True, rows_per_cloud=rows_per_cloud,
                      valid_rows=valid, anorm=dict(stats=stats, **extra, t=t, scale_w=sw, scale_b=sb, bias_w=bw, bias_b=bb, groups=groups))
    torch.cuda.synchronize()
    xv = x[:, :valid]
    normed = F.group_norm(xv.transpose(1, 2), groups, eps=1e-5).transpose(1, 2)
    y = (t[:, None, None] * sw[:, 0] + sb) * normed + (t[:, None, None] * bw[:, 0] + bb)
    ref = _ref(y.bfloat16().reshape(-1, K), w, bias=bias, alpha=act).view(clouds, valid, n_out)
    got = o16.view(clouds, rows_per_cloud, n_out)[:, :valid].float()
    err = (got - ref).pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item()
    # bf16 output rounding (2^-9) plus the few operand elements that round to the neighbouring bf16 value
    assert err < 4e-3, err
    assert got.isfinite().all()
    if valid < rows_per_cloud:  # padding rows come out as exact zeros
        assert o16.view(clouds, rows_per_cloud, n_out)[:, valid:].abs().max().item() == 0.0
    # the same call is rejected, loudly, where the kernel cannot take it
    assert not ops.gemm_anorm_supported(1024, 512, n_out, K) and not ops.gemm_anorm_supported(m, rows_per_cloud, n_out, 136)
    with pytest.raises(ValueError):
        ops.gemm(x2[: 3 * 384].bfloat16(), w, out_bf16=True, rows_per_cloud=384, valid_rows=300,
                 anorm=dict(stats=stats, t=t, scale_w=sw, scale_b=sb, bias_w=bw, bias_b=bb, groups=groups))


@pytest.mark.parametrize("n_out,act", [(1152, None), (768, 1.3)])
@pytest.mark.parametrize("clouds,rows_per_cloud,valid", [(4, 768, 700), (8, 2048, 2048)])
def test_gemm_fast_epilogue_is_bit_identical(cuda, n_out, act, clouds, rows_per_cloud, valid):
    """The bf16-only fast epilogue of the CTA-pair kernel (three TMEM chunks in flight, early accumulator release, one
    bulk group per tile) against the generic epilogue: same arithmetic per element, so the outputs must agree bit for
    bit; both are also checked against torch."""
    from gecco_b200 import ops

    K = 384
    m = clouds * rows_per_cloud
    g = torch.Generator(device="cpu").manual_seed(n_out + rows_per_cloud)
    a = torch.randn(m, K, generator=g).to(cuda).bfloat16()
    w = (torch.randn(clouds * n_out, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    bias = torch.randn(clouds, n_out, generator=g).to(cuda)
    outs = []
    try:
        for fast in (0, 1):
            ops.set_option("fast_epilogue", fast)
            out = torch.full((m, n_out), 3.0, device=cuda, dtype=torch.bfloat16)
            ops.gemm(a, w, bias=bias, bias_stride=n_out, act_alpha=act, out_bf16=out, rows_per_cloud=rows_per_cloud, valid_rows=valid,
                     w_rows_per_cloud=n_out, n_out=n_out)
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        ops.set_option("fast_epilogue", 1)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))
    got = outs[1].view(clouds, rows_per_cloud, n_out)
    for c in range(clouds):
        ref = _ref(a.view(clouds, rows_per_cloud, K)[c, :valid], w.view(clouds, n_out, K)[c], bias=bias[c], alpha=act)
        assert (got[c, :valid].float() - ref).abs().max().item() < 6e-2
    if valid < rows_per_cloud:
        assert got[:, valid:].abs().max().item() == 0.0
