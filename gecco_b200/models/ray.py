"""Image-conditional point network (reference: gecco_torch/models/ray.py:33-120): xyz embedding plus projective
lookup of the CNN feature pyramid, SetTransformer, GroupNorm + Linear head.

Parameter containers with the reference's names; the computation runs in the CUDA engine:
`extract_image_features` is the `gecco_lookup` gather kernel (reparam -> pinhole projection -> bilinear taps on
every level, csrc/lookup.cu) and `forward` is one `gecco_denoise` call in network mode.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from ..engine import engine_for
from ..reparam import Reparam
from ..structs import Context3d, FeaturePyramidContext
from .set_transformer import SetTransformer


class GroupNormBNC(nn.GroupNorm):
    """GroupNorm over a [batch, points, channels] tensor: statistics per (batch, group) over points x channels-in-group
    (ray.py:20-30).  Inside the engine it is folded into the following Linear (gecco_fold_group_norm)."""

    @torch.no_grad()
    def forward(self, tensor_bnc: Tensor) -> Tensor:
        assert tensor_bnc.ndim == 3
        B, N, C = tensor_bnc.shape
        x = tensor_bnc.to(torch.float32).contiguous().view(B * N, C)
        gs = C // self.num_groups
        stats = ops.group_stats(x, N, N, gs)
        one = torch.ones(C, device=x.device)
        zero_w = torch.zeros(C, 1, device=x.device)
        t = torch.zeros(B, 1, device=x.device)
        w = self.weight if self.affine else one
        b = self.bias if self.affine else torch.zeros(C, device=x.device)
        out, _ = ops.adagn(x, stats, gs, t, zero_w, w, zero_w, b, rows_per_cloud=N, valid_rows=N, groups=self.num_groups,
                           eps=self.eps, out_f32=True)
        return out.view(B, N, C).to(tensor_bnc.dtype)


class RayNetwork(nn.Module):
    def __init__(self, backbone: SetTransformer, reparam: Reparam, context_dims: list[int]):
        super().__init__()
        self.backbone = backbone
        self.reparam = reparam
        self.context_dims = context_dims
        self.xyz_embed = nn.Linear(reparam.dim, backbone.feature_dim)
        self.img_feature_proj = nn.Sequential(GroupNormBNC(16, sum(context_dims), affine=False),
                                              nn.Linear(sum(context_dims), backbone.feature_dim))
        self.output_proj = nn.Sequential(GroupNormBNC(16, backbone.feature_dim, affine=False),
                                         nn.Linear(backbone.feature_dim, reparam.dim))

    def extra_repr(self) -> str:
        return f"context_dims={self.context_dims}"

    @torch.no_grad()
    def extract_image_features(self, geometry_diffusion: Tensor, features: list[Tensor], ctx: Context3d) -> Tensor:
        """[B, N, sum(context_dims)] fp32 bilinear lookups of the pyramid at the projections of the points (ray.py:64-87)."""
        B, N, _ = geometry_diffusion.shape
        levels = [ops.pack_features(f) for f in features]
        mean, sigma, logit_scale = self.reparam._host_stats()
        out, _ = ops.lookup(geometry_diffusion.to(torch.float32).contiguous(), levels, ctx.K, reparam_kind=self.reparam._kind,
                            mean=mean, sigma_r=sigma, logit_scale=logit_scale, rows_per_cloud=N, out_f32=True)
        return out.view(B, N, -1)

    def forward(self, geometry: Tensor, t: Tensor, raw_ctx: Context3d, post_context: FeaturePyramidContext,
                do_cache: bool = False, cache: list[Tensor] | None = None):
        out, out_cache = engine_for(self).denoise(geometry, t_embed=t, post_context=post_context, K=raw_ctx.K, cache=cache,
                                                  do_cache=do_cache, mode=0)
        return out.to(geometry.dtype), out_cache
