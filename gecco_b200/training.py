"""Training step of the diffusion model (BASELINE config 5): EDM denoising loss, forward + backward, gradient all-reduce,
Adam + weight EMA.  Reference: gecco_torch/diffusion.py:118-143 (EDMLoss), :207-222 (configure_optimizers /
training_step), gecco_torch/ema.py:187-194, 301-325 (EMA of the weights after every optimiser step).

The sampling engine (csrc/engine.cu) is forward-only and fused; training needs the activations, so the autograd path
below is a separate, differentiable restatement of the same network that runs when the model is in `train()` mode with
gradients enabled (`EDMPrecond.forward` dispatches here):

  * every nn.Linear of the SetTransformer / RayNetwork with K, N multiples of 8 goes through `TCLinear`: the FORWARD and
    the input gradient dX = dY W run on the hand-written tcgen05 GEMM of the sampling path (`gecco_gemm`, bf16 operands,
    fp32 accumulate and output); the weight gradient dW = dY^T X is a plain library GEMM (cuBLAS, bf16 operands, fp32
    output) -- the one K-major-in-M product the tcgen05 kernels have no layout for;
  * attention cores use torch's fused scaled-dot-product attention in bf16, normalisations, the Gaussian activation and
    the bilinear lookup (F.grid_sample, so that the conditioner receives gradients) are torch autograd ops in fp32;
  * parameters, gradients, both Adam moments and the EMA copy live in FLAT fp32 buffers (`FlatState`): gradients are
    all-reduced bucket by bucket while backward is still running (`GradReducer`, NCCL over NVLink; gloo in the CPU test)
    and the optimiser + EMA update is ONE kernel over the flat buffers (`gecco_adam_ema_step`, csrc/optimizer.cu).

Clouds are independent, so data parallelism shards the batch on dim 0 exactly like sampling; the only collective is the
gradient all-reduce.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Iterable, Sequence

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops


# ----------------------------------------------------------------------------------------------------------------------
# Linear layers on the tcgen05 GEMM
class TCLinear(torch.autograd.Function):
    """y = x W^T + b with x [M, K], W [N, K]: forward and dX on `gecco_gemm`, dW on cuBLAS, db a column sum (one GEMV-like
    pass).  x may already be bf16 (attention outputs) and y may be requested in bf16 (attention inputs): the tensors between
    a projection and an attention core then never exist in fp32, and their gradients arrive / leave in bf16 as well."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor | None, out_bf16: bool = False, residual: Tensor | None = None):
        xb = x.detach().to(torch.bfloat16).contiguous()
        wb = weight.detach().to(torch.bfloat16).contiguous()
        # `residual` [M, N] fp32 is added in the GEMM epilogue (the layer's `x + f(x)`): no separate element-wise pass
        y32, y16 = ops.gemm(xb, wb, bias=None if bias is None else bias.detach().float().contiguous(), out_f32=not out_bf16,
                            out_bf16=out_bf16, res=None if residual is None else residual.detach().contiguous())
        ctx.save_for_backward(xb, wb)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.x_bf16 = x.dtype == torch.bfloat16
        return y16 if out_bf16 else y32

    @staticmethod
    def backward(ctx, dy: Tensor):
        xb, wb = ctx.saved_tensors
        dyb = dy.to(torch.bfloat16).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dX [M, K] = dY [M, N] . W [N, K]: the same kernel with W^T as its (row-major [K, N]) weight operand
            d32, d16 = ops.gemm(dyb, wb.t().contiguous(), out_f32=not ctx.x_bf16, out_bf16=ctx.x_bf16)
            dx = d16 if ctx.x_bf16 else d32
        if ctx.needs_input_grad[1]:
            dw = _mm_f32(dyb.t(), xb)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy)
        return dx, dw, db, None, (dy if ctx.has_res and ctx.needs_input_grad[4] else None)


_ONES: dict = {}


def _colsum(dy: Tensor) -> Tensor:
    """Column sums of dY [M, N] in fp32 as ONE pass (ones^T . dY on the library GEMM / GEMV) instead of torch's strided
    reduction kernel."""
    key = (dy.device, dy.dtype, dy.shape[0])
    ones = _ONES.get(key)
    if ones is None:
        ones = _ONES[key] = torch.ones(dy.shape[0], device=dy.device, dtype=dy.dtype)
    if dy.dtype == torch.float32:
        return torch.mv(dy.t(), ones)
    return _mm_f32(ones.unsqueeze(0), dy).squeeze(0)


def _mm_f32(a: Tensor, b: Tensor) -> Tensor:
    """bf16 x bf16 -> fp32 library GEMM (fp32 accumulate and output when this torch exposes out_dtype)."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:
        return torch.mm(a, b).float()


# Input-gradient mode (Diffusion.log_likelihood): gradients flow to the network INPUT only -- weights are detached (no dW
# products) and the projective lookup becomes differentiable in the sample positions.
_INPUT_GRAD_ONLY = False


@contextlib.contextmanager
def input_gradients():
    global _INPUT_GRAD_ONLY
    prev, _INPUT_GRAD_ONLY = _INPUT_GRAD_ONLY, True
    try:
        yield
    finally:
        _INPUT_GRAD_ONLY = prev


def linear(x: Tensor, lin_w: Tensor, lin_b: Tensor | None, out_bf16: bool = False, residual: Tensor | None = None) -> Tensor:
    """F.linear on [..., K]; through the tcgen05 GEMM when the shape allows (K, N multiples of 8, CUDA).  `out_bf16`: the
    result feeds an attention core and is produced in bf16 directly (CUDA only; fp32 elsewhere)."""
    K, N = lin_w.shape[1], lin_w.shape[0]
    if _INPUT_GRAD_ONLY:
        lin_w, lin_b = lin_w.detach(), None if lin_b is None else lin_b.detach()
    if x.is_cuda and K % 8 == 0 and N % 8 == 0 and x.numel() > 0 and os.environ.get("GECCO_TRAIN_TC", "1") != "0":
        fuse = residual is not None and not out_bf16 and residual.dtype == torch.float32 and _fused()
        y = TCLinear.apply(x.reshape(-1, K), lin_w, lin_b, out_bf16, residual.reshape(-1, N) if fuse else None)
        y = y.view(*x.shape[:-1], N)
        return y if fuse or residual is None else residual + y
    if x.is_cuda:  # library arm of the A/B (GECCO_TRAIN_TC=0) and odd shapes: cuBLAS with the same bf16 operands
        y = F.linear(x.to(torch.bfloat16), lin_w.to(torch.bfloat16), None if lin_b is None else lin_b.to(torch.bfloat16))
        y = y if out_bf16 else y.float()
        return y if residual is None else residual + y
    y = F.linear(x, lin_w, lin_b)
    return y if residual is None else residual + y


# ----------------------------------------------------------------------------------------------------------------------
# The network, differentiable (same arithmetic as the engine; citations are the reference's)
def _fused() -> bool:
    """GECCO_TRAIN_FUSED=0: normalisation and activation through torch's own ops (A/B arm, and the CPU path of the tests)."""
    return os.environ.get("GECCO_TRAIN_FUSED", "1") != "0"


class GroupAffineNorm(torch.autograd.Function):
    """y[b,n,c] = gamma[b,c] * xhat[b,n,c] + beta[b,c], xhat = group normalisation of x over (rows x channels-in-group)
    per cloud (nn.GroupNorm statistics: biased variance, eps inside the root).  Forward: `gecco_group_stats` (fp64 sums)
    + one `gecco_train_affine` pass; backward: one `gecco_train_colsum2` reduction + one two-input affine pass:
        dbeta = sum_n dy,  dgamma = sum_n dy xhat,
        dx = rstd (gamma dy - mean_g(gamma dy) - xhat mean_g(gamma dy xhat)).
    x [B, N, C] fp32 contiguous, gamma / beta [B, C]."""

    @staticmethod
    def forward(ctx, x: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float):
        B, N, C = x.shape
        gs = C // groups
        x = x.contiguous()
        stats = ops.group_stats(x.view(B * N, C), N, N, gs)  # [B, G, 2] float64
        n = float(N * gs)
        mean = stats[..., 0] / n
        var = (stats[..., 1] / n - mean * mean).clamp_min(0.0)
        rstd_c = (var + eps).rsqrt().float().repeat_interleave(gs, dim=1)  # [B, C]
        mean_c = mean.float().repeat_interleave(gs, dim=1)
        p = (gamma * rstd_c).contiguous()
        y = ops.train_affine(x, None, p, None, (beta - p * mean_c).contiguous())
        ctx.save_for_backward(x, gamma, mean_c, rstd_c)
        ctx.gs, ctx.n = gs, n
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, gamma, mean_c, rstd_c = ctx.saved_tensors
        B, N, C = x.shape
        gs, n = ctx.gs, ctx.n
        dy = dy.contiguous()
        T = ops.train_colsum2(dy, x)
        t1, t2 = T[..., 0], T[..., 1]
        dgamma = rstd_c * (t2 - mean_c * t1)
        grp = lambda v: v.view(B, C // gs, gs).sum(-1, keepdim=True).expand(B, C // gs, gs).reshape(B, C)
        a_c, b_c = grp(gamma * t1), grp(gamma * dgamma)
        r2 = rstd_c * rstd_c
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.train_affine(dy, x, (rstd_c * gamma).contiguous(), (-r2 * b_c / n).contiguous(),
                                  ((r2 * mean_c * b_c - rstd_c * a_c) / n).contiguous())
        return dx, dgamma, t1, None, None


class GaussAct(torch.autograd.Function):
    """models/activation.py:17-24 in one pass forward and one backward (alpha stays on the device: graph-safe)."""

    @staticmethod
    def forward(ctx, x: Tensor, alpha: Tensor, normalized: bool):
        x = x.contiguous()
        a = alpha.detach().reshape(()).float()
        ctx.save_for_backward(x, a)
        ctx.normalized, ctx.alpha_shape = normalized, alpha.shape
        return ops.train_gauss_act_fwd(x, a, normalized)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, a = ctx.saved_tensors
        dx, dalpha = ops.train_gauss_act_bwd(x, dy.contiguous(), a, ctx.normalized)
        return dx, dalpha.reshape(ctx.alpha_shape), None


def _use_kernels(x: Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[-1] % 4 == 0 and x.shape[-1] <= 1024 and _fused()


def adagn(mod, x: Tensor, t: Tensor) -> Tensor:
    """models/normalization.py:36-44: GroupNorm over (points x channels-in-group), then the t-conditioned affine."""
    scale = F.linear(t, mod.scale.weight, mod.scale.bias)
    bias = F.linear(t, mod.bias.weight, mod.bias.bias)
    if _use_kernels(x) and x.shape[-1] % mod.gn.num_groups == 0:
        B, _, C = x.shape
        return GroupAffineNorm.apply(x, scale.reshape(B, C), bias.reshape(B, C), mod.gn.num_groups, mod.gn.eps)
    normed = F.group_norm(x.transpose(1, 2), mod.gn.num_groups, eps=mod.gn.eps).transpose(1, 2)
    return scale * normed + bias


def gaussian_activation(act, x: Tensor) -> Tensor:
    """models/activation.py:17-24."""
    if x.is_cuda and x.dtype == torch.float32 and _fused():
        return GaussAct.apply(x, act.alpha, bool(act.normalized))
    y = (-(x**2) / (2 * act.alpha**2)).exp()
    return (y - 0.7) / 0.28 if act.normalized else y


def mlp(mod, x: Tensor, residual: Tensor | None = None) -> Tensor:
    """models/mlp.py:5-39 (an nn.Sequential of Linear / activation modules); `residual` is added to the result (in the
    epilogue of the last projection)."""
    from .models.activation import GaussianActivation

    mods = list(mod)
    for i, m in enumerate(mods):
        if isinstance(m, nn.Linear):
            last = i == len(mods) - 1
            x = linear(x, m.weight, m.bias, residual=residual if last else None)
            if last:
                residual = None
        elif isinstance(m, GaussianActivation):
            x = gaussian_activation(m, x)
        else:
            x = m(x)
    return x if residual is None else residual + x


def _sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """Attention core; bf16 in and OUT on CUDA (the following projection takes bf16), fp32 on the CPU test path."""
    if q.is_cuda:
        return F.scaled_dot_product_attention(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16))
    return F.scaled_dot_product_attention(q, k, v)


def attention_pool(pool, y: Tensor) -> Tensor:
    """models/set_transformer.py:47-65."""
    B, N, C = y.shape
    H, d = pool.num_heads, pool.dims_per_head
    kv = linear(y, pool.kv_proj.weight, None, out_bf16=True)  # columns ordered (t h d)
    k = kv[..., :C].reshape(B, N, H, d).transpose(1, 2)
    v = kv[..., C:].reshape(B, N, H, d).transpose(1, 2)
    q = pool.inducers.expand(B, -1, -1, -1)
    attn = _sdpa(q, k, v).transpose(1, 2).reshape(B, -1, C)
    return linear(attn, pool.out_proj.weight, None)


def unpool(mha: nn.MultiheadAttention, y: Tensor, h: Tensor, residual: Tensor | None = None) -> Tensor:
    """nn.MultiheadAttention(batch_first)(query = y, key = value = h), models/set_transformer.py:90,112."""
    B, N, C = y.shape
    I, H = h.shape[1], mha.num_heads
    d = C // H
    w, b = mha.in_proj_weight, mha.in_proj_bias
    bq, bkv = (None, None) if b is None else (b[:C], b[C:])
    q = linear(y, w[:C], bq, out_bf16=True).reshape(B, N, H, d).transpose(1, 2)
    kv = linear(h, w[C:], bkv, out_bf16=True)
    k = kv[..., :C].reshape(B, I, H, d).transpose(1, 2)
    v = kv[..., C:].reshape(B, I, H, d).transpose(1, 2)
    a = _sdpa(q, k, v).transpose(1, 2).reshape(B, N, C)
    return linear(a, mha.out_proj.weight, mha.out_proj.bias, residual=residual)


def broadcasting_layer(layer, x: Tensor, t: Tensor) -> Tensor:
    """models/set_transformer.py:92-117 (Broadcast) and :155-168 (BroadcastingLayer)."""
    bc = layer.broadcast
    y = adagn(layer.broadcast_norm, x, t)
    h = attention_pool(bc.pool, y)
    h = adagn(bc.norm_1, h, t)
    h = mlp(bc.mlp, h)
    h = adagn(bc.norm_2, h, t)
    x = unpool(bc.unpool, y, h, residual=x)  # x + unpool(...): the addition runs in the out-projection's epilogue
    y = adagn(layer.mlp_norm, x, t)
    return mlp(layer.mlp, y, residual=x)


def set_transformer(st, x: Tensor, t: Tensor) -> Tensor:
    """models/set_transformer.py:198-216 (no inducer cache while training)."""
    for layer in st.layers:
        x = broadcasting_layer(layer, x, t)
    return x


def _project(points: Tensor, K: Tensor) -> Tensor:
    """kornia.geometry.camera.perspective.project_points as the reference calls it (models/ray.py:73): the guarded
    reciprocal 1 / (z + eps) for |z| > eps, else 1."""
    eps = 1e-8
    z = points[..., 2:]
    scale = torch.where(z.abs() > eps, 1.0 / (z + eps), torch.ones_like(z))
    xy = scale * points[..., :2]
    Kb = K.unsqueeze(1).to(points.dtype)
    return torch.stack([xy[..., 0] * Kb[..., 0, 0] + Kb[..., 0, 2], xy[..., 1] * Kb[..., 1, 1] + Kb[..., 1, 2]], dim=-1)


def reparam_to_data(reparam, diff: Tensor, ctx) -> Tensor:
    """`Reparam.diffusion_to_data` (reparam.py:31-40, 62-64, 191-201) as differentiable torch expressions (the product's
    `gecco_reparam` kernel has no backward)."""
    from .reparam import GaussianReparam, NoReparam, UVLReparam

    if isinstance(reparam, NoReparam):
        return diff
    if isinstance(reparam, GaussianReparam):
        return diff * reparam.sigma.to(diff) + reparam.mean.to(diff)
    if isinstance(reparam, UVLReparam):
        uvl = diff * reparam.uvl_std.to(diff) + reparam.uvl_mean.to(diff)
        K = ctx.K.unsqueeze(1).to(diff)
        hw = (torch.tanh(uvl[..., :2]) * reparam.logit_scale + 1.0) / 2
        x = (hw[..., 0] - K[..., 0, 2]) / K[..., 0, 0]
        y = (hw[..., 1] - K[..., 1, 2]) / K[..., 1, 1]
        ray = F.normalize(torch.stack([x, y, torch.ones_like(x)], dim=-1), dim=-1, p=2.0)
        return ray * torch.exp(uvl[..., 2:])
    raise TypeError(f"gecco_b200.training: unsupported reparametrisation {type(reparam).__name__}")


def reparam_to_diffusion(reparam, data: Tensor, ctx) -> Tensor:
    """`Reparam.data_to_diffusion` (reparam.py:58-60, 178-189), differentiable."""
    from .reparam import GaussianReparam, NoReparam, UVLReparam

    if isinstance(reparam, NoReparam):
        return data
    if isinstance(reparam, GaussianReparam):
        return (data - reparam.mean.to(data)) / reparam.sigma.to(data)
    if isinstance(reparam, UVLReparam):
        hw = _project(data, ctx.K.to(data))
        real = torch.arctanh((2 * hw - 1.0) / reparam.logit_scale)
        uvl = torch.cat([real, torch.log(torch.linalg.norm(data, dim=-1, keepdim=True))], dim=-1)
        return (uvl - reparam.uvl_mean.to(data)) / reparam.uvl_std.to(data)
    raise TypeError(f"gecco_b200.training: unsupported reparametrisation {type(reparam).__name__}")


def extract_image_features(net, geometry_diffusion: Tensor, features: Sequence[Tensor], raw_ctx) -> Tensor:
    """models/ray.py:64-87.  While training the sample positions carry no gradient (the network input is data + noise);
    the feature maps do (the conditioner trains), which is why this is F.grid_sample and not the sampling path's gather
    kernel.  In input-gradient mode (log-likelihood) the positions are differentiable too."""
    if _INPUT_GRAD_ONLY:
        data = reparam_to_data(net.reparam, geometry_diffusion.float(), raw_ctx)
        grid = (_project(data, raw_ctx.K.float()) * 2 - 1).unsqueeze(2)
        looks = [F.grid_sample(f.detach().float(), grid, align_corners=False)[..., 0].transpose(1, 2) for f in features]
        return torch.cat(looks, dim=-1)
    with torch.no_grad():
        data = net.reparam.diffusion_to_data(geometry_diffusion.detach().float().contiguous(), raw_ctx)
        grid = (_project(data, raw_ctx.K.float()) * 2 - 1).unsqueeze(2)  # [B, N, 1, 2]
    looks = [F.grid_sample(f.float(), grid, align_corners=False)[..., 0].transpose(1, 2) for f in features]
    return torch.cat(looks, dim=-1)


def _group_norm_bnc(mod, x: Tensor) -> Tensor:
    """models/ray.py:20-30."""
    if _use_kernels(x) and x.shape[-1] % mod.num_groups == 0:
        B, _, C = x.shape
        w = x.new_ones(B, C) if mod.weight is None else mod.weight.expand(B, C)
        b = x.new_zeros(B, C) if mod.bias is None else mod.bias.expand(B, C)
        return GroupAffineNorm.apply(x, w, b, mod.num_groups, mod.eps)
    return F.group_norm(x.transpose(1, 2), mod.num_groups, mod.weight, mod.bias, mod.eps).transpose(1, 2)


def network(net, geometry: Tensor, t: Tensor, raw_ctx, post_ctx) -> Tensor:
    """LinearLift.forward (models/linear_lift.py:33-46) / RayNetwork.forward (models/ray.py:89-120)."""
    from .models.linear_lift import LinearLift
    from .models.ray import RayNetwork

    if isinstance(net, LinearLift):
        x = F.linear(geometry, net.lift.weight, net.lift.bias)
        x = set_transformer(net.inner, x, t)
        if isinstance(net.lower, nn.Sequential):
            x = F.layer_norm(x, (x.shape[-1],), eps=net.lower[0].eps)
            return F.linear(x, net.lower[1].weight, net.lower[1].bias)
        return F.linear(x, net.lower.weight, net.lower.bias)
    if isinstance(net, RayNetwork):
        xyz = F.linear(geometry, net.xyz_embed.weight, net.xyz_embed.bias)
        img_raw = extract_image_features(net, geometry, post_ctx.features, raw_ctx)
        img = linear(_group_norm_bnc(net.img_feature_proj[0], img_raw), net.img_feature_proj[1].weight, net.img_feature_proj[1].bias)
        x = set_transformer(net.backbone, xyz + img, t)
        out = _group_norm_bnc(net.output_proj[0], x)
        return F.linear(out, net.output_proj[1].weight, net.output_proj[1].bias)
    raise TypeError(f"gecco_b200.training: unsupported network {type(net).__name__}")


def precond_forward(precond, x: Tensor, sigma: Tensor, raw_context, post_context) -> Tensor:
    """EDMPrecond.forward (diffusion.py:37-62), differentiable."""
    sigma = sigma.reshape(-1, *((1,) * (x.ndim - 1))).to(x.dtype)
    sd2 = precond.sigma_data**2
    c_skip = sd2 / (sigma**2 + sd2)
    c_out = sigma * precond.sigma_data / (sigma**2 + sd2).sqrt()
    c_in = 1 / (sd2 + sigma**2).sqrt()
    c_noise = sigma.log() / 4
    F_x = network(precond.model, c_in * x, c_noise, raw_context, post_context)
    return c_skip * x + c_out * F_x


# ----------------------------------------------------------------------------------------------------------------------
# Flat optimiser state, bucketed gradient all-reduce, the step
_ALIGN = 64  # elements: every parameter starts on a 256-byte boundary of the flat buffers


class FlatState:
    """Parameters, gradients, Adam moments and the EMA copy as five flat fp32 buffers.  `p.data` and `p.grad` of every
    trainable parameter become views into them (autograd then accumulates gradients in place)."""

    def __init__(self, params: Iterable[nn.Parameter], with_ema: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("training needs fp32 parameters on one device")
            self.offsets.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.g = torch.zeros_like(self.p)
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        for p, o in zip(self.params, self.offsets):
            self.p[o:o + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.p[o:o + p.numel()].view(p.shape)
            p.grad = self.g[o:o + p.numel()].view(p.shape)
        self.ema = self.p.clone() if with_ema else None

    def zero_grad(self) -> None:
        self.g.zero_()
        for p, o in zip(self.params, self.offsets):  # a `model.zero_grad(set_to_none=True)` from outside un-hooks the views
            if p.grad is None or p.grad.data_ptr() != self.g.data_ptr() + 4 * o:
                p.grad = self.g[o:o + p.numel()].view(p.shape)

    def ema_view(self, i: int) -> Tensor:
        o, p = self.offsets[i], self.params[i]
        return self.ema[o:o + p.numel()].view(p.shape)


class GradReducer:
    """Averages gradients over the ranks bucket by bucket, overlapped with backward: a bucket's all-reduce (SUM, async)
    starts from the post-accumulate hook of the last of its parameters to receive a gradient; `finish()` launches the
    buckets whose parameters took no part in the step, waits, and reports the 1 / world scale still to be applied (it is
    folded into the optimiser kernel).  Buckets are contiguous ranges of the flat gradient buffer, filled from the LAST
    parameter backwards since backward produces gradients roughly in reverse order."""

    def __init__(self, params: Sequence[nn.Parameter], offsets: Sequence[int], flat_grad: Tensor, bucket_bytes: int = 64 << 20,
                 group=None, hooks: bool = True):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: list[tuple[int, int]] = []   # [lo, hi) element ranges
        self.bucket_of: list[int] = [0] * len(params)
        cap = max(1, bucket_bytes // 4)
        hi = flat_grad.numel()
        members = 0
        for i in range(len(params) - 1, -1, -1):
            lo = offsets[i]
            if members and hi - lo > cap:
                self.buckets.append((offsets[i + 1], hi))
                hi, members = offsets[i + 1], 0
            self.bucket_of[i] = len(self.buckets)
            members += 1
        self.buckets.append((0, hi))
        self.count = [0] * len(self.buckets)
        for b in self.bucket_of:
            self.count[b] += 1
        self.pending = list(self.count)
        self.launched = [False] * len(self.buckets)
        self.works: list = []
        self.handles = []
        if self.world > 1 and hooks:  # hooks off: finish() reduces every bucket (after a CUDA-graph replay of backward)
            for i, p in enumerate(params):
                self.handles.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, i: int):
        def hook(_param):
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b: int) -> None:
        if self.launched[b]:
            return
        lo, hi = self.buckets[b]
        self.launched[b] = True
        self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> float:
        """Completes the step's all-reduces; returns the scale (1 / world size) that turns the sums into means."""
        if self.world > 1:
            for b in range(len(self.buckets)):
                self._launch(b)
            for w in self.works:
                w.wait()
        self.works.clear()
        self.pending = list(self.count)
        self.launched = [False] * len(self.buckets)
        return 1.0 / self.world

    def remove(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles.clear()


class Trainer:
    """forward + backward + gradient all-reduce + fused Adam / EMA for a `Diffusion` model (the reference runs the same
    sequence under Lightning with torch.optim.Adam(lr=1e-4) and its EMA callback).

    graph=False: eager autograd; bucket all-reduces start from gradient hooks while backward is still running.
    graph=True : forward + backward of a fixed batch shape are captured ONCE as a CUDA graph and replayed on static input
                 buffers (the eager step is host-bound: thousands of small autograd launches); the flat gradient buffer
                 is then all-reduced bucket by bucket and the optimiser kernel follows.  Noise levels and noise are drawn
                 inside the graph from torch's graph-safe CUDA generator, so every replay sees fresh draws."""

    def __init__(self, model: nn.Module, lr: float = 1e-4, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 ema_decay: float = 0.999, bucket_mb: int = 64, group=None, graph: bool = False):
        self.model = model
        self.lr, self.betas, self.eps, self.ema_decay = lr, betas, eps, ema_decay
        self.state = FlatState(model.parameters(), with_ema=ema_decay is not None)
        self.graph_mode = graph
        self.reducer = GradReducer(self.state.params, self.state.offsets, self.state.g, bucket_mb << 20, group, hooks=not graph)
        self.steps = 0
        self._graph = None
        self._static = None
        self.graph_library_launches = 0

    def step(self, batch) -> Tensor:
        """One optimisation step on `batch` = (data [B, N, 3], Context3d | None); returns the (detached) loss."""
        self.model.train()
        if self.graph_mode:
            loss = self._graphed_forward_backward(batch)
        else:
            self.state.zero_grad()
            loss = self.model.training_step(batch, self.steps)
            loss.backward()
        scale = self.reducer.finish()
        self.steps += 1
        ops.adam_ema_step(self.state.p, self.state.g, self.state.m, self.state.v, self.state.ema, self.steps, lr=self.lr,
                          betas=self.betas, eps=self.eps, grad_scale=scale,
                          ema_decay=self.ema_decay if self.ema_decay is not None else 0.0)
        return loss.detach()

    # ---- CUDA-graph path
    @staticmethod
    def _tensors(batch):
        data, ctx = batch
        return [data] + ([] if ctx is None else [ctx.image, ctx.K])

    def _graphed_forward_backward(self, batch) -> Tensor:
        ts = self._tensors(batch)
        key = tuple((tuple(t.shape), t.dtype) for t in ts)
        if self._graph is None or self._static["key"] != key:
            self._capture(batch, key)
        for dst, src in zip(self._static["tensors"], ts):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static["loss"].clone()

    def _capture(self, batch, key) -> None:
        data, ctx = batch
        s_data = data.clone()
        s_ctx = None if ctx is None else type(ctx)(image=ctx.image.clone(), K=ctx.K.clone())
        static_batch = (s_data, s_ctx)
        side = torch.cuda.Stream(device=data.device)
        side.wait_stream(torch.cuda.current_stream(data.device))
        with torch.cuda.stream(side):  # warm-up outside the capture (lazy initialisations, cuDNN plans, tensor-map cache)
            for _ in range(2):
                self.state.zero_grad()
                self.model.training_step(static_batch, 0).backward()
        torch.cuda.current_stream(data.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        from . import engine as _engine

        before = _engine.launch_count()
        with torch.cuda.graph(self._graph):
            self.state.g.zero_()
            loss = self.model.training_step(static_batch, 0)
            loss.backward()
        # launches of this library's kernels recorded in the graph (each replay issues them again without passing the C ABI)
        self.graph_library_launches = _engine.launch_count() - before
        self._static = dict(key=key, tensors=self._tensors(static_batch), loss=loss.detach())

    @contextlib.contextmanager
    def ema_weights(self):
        """Runs the body with the EMA weights in place of the trained ones (ema.py:327-349, swap_ema_weights)."""
        if self.state.ema is None:
            yield
            return
        self._swap()
        try:
            yield
        finally:
            self._swap()

    def _swap(self) -> None:
        tmp = self.state.p.clone()
        self.state.p.copy_(self.state.ema)
        self.state.ema.copy_(tmp)

    def ema_state_dict(self) -> dict:
        """state_dict of the model with the EMA weights (what the reference stores as `ema_state_dict`, ema.py:175-185)."""
        with self.ema_weights():
            return {k: v.detach().clone() for k, v in self.model.state_dict().items()}


def flops_per_cloud(n_points: int, n_layers: int = 6, c: int = 384, hidden: int = 768, inducers: int = 64) -> float:
    """Dense FLOPs of one forward pass per cloud (GEMMs + attention cores); a training step costs about three times that."""
    per_layer = 2 * n_points * c * (3 * c) + 2 * n_points * c * c + 4 * n_points * c * hidden + 8 * n_points * inducers * c
    per_layer += inducers * (2 * c * c * 3 + 4 * c * hidden)
    return float(n_layers * per_layer)
