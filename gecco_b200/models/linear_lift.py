"""Unconditional point network (reference: gecco_torch/models/linear_lift.py:7-46):
Linear(3 -> C) "lift", SetTransformer, LayerNorm(C, no affine) + Linear(C -> 3) "lower".
Parameter containers only; `forward` runs the whole stack in the CUDA engine."""
from typing import Any

from torch import Tensor, nn

from ..engine import engine_for
from .set_transformer import SetTransformer


class LinearLift(nn.Module):
    def __init__(self, inner: SetTransformer, feature_dim: int, geometry_dim: int = 3, do_norm: bool = True):
        super().__init__()
        if geometry_dim != 3:
            raise ValueError("gecco_b200 supports 3-dimensional geometry only")
        self.lift = nn.Linear(geometry_dim, feature_dim)
        self.inner = inner
        if do_norm:
            self.lower = nn.Sequential(nn.LayerNorm(feature_dim, elementwise_affine=False), nn.Linear(feature_dim, geometry_dim))
        else:
            self.lower = nn.Linear(feature_dim, geometry_dim)

    def forward(self, geometry: Tensor, embed: Tensor, raw_context: Any, post_context: Any, do_cache: bool = False,
                cache: list | None = None):
        """(features, cache) like the reference; `geometry` is the already scaled input, `embed` the noise embedding."""
        del raw_context, post_context
        out, out_cache = engine_for(self).denoise(geometry, t_embed=embed, cache=cache, do_cache=do_cache, mode=0)
        return out.to(geometry.dtype), out_cache
