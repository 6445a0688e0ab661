"""Adaptive group normalisation (reference: gecco_torch/models/normalization.py:14-44).

GroupNorm(num_groups, no affine) over (points x channels-in-group) of a [B, N, C] tensor followed by a
per-channel scale / bias that are linear functions of the noise-level embedding; zero-initialised to the
identity modulation like the reference (:30-34).  Standalone calls run the `gecco_group_stats` +
`gecco_adagn` kernels; inside the denoiser the statistics come from the producing GEMM's epilogue.
"""
import torch
from torch import Tensor, nn

from .. import ops


class AdaNorm(nn.Module):
    def forward(self, x: Tensor, ctx: Tensor) -> Tensor:
        raise NotImplementedError()


class AdaGN(nn.Module):
    def __init__(self, num_channels: int, ctx_dim: int, num_groups: int = 32):
        super().__init__()
        self.gn = nn.GroupNorm(num_groups=num_groups, num_channels=num_channels, affine=False)
        self.bias = nn.Linear(ctx_dim, num_channels)
        self.scale = nn.Linear(ctx_dim, num_channels)
        with torch.no_grad():
            self.bias.weight.zero_()
            self.bias.bias.zero_()
            self.scale.weight.zero_()
            self.scale.bias.fill_(1.0)

    @torch.no_grad()
    def forward(self, x: Tensor, ctx: Tensor) -> Tensor:
        assert x.ndim == 3, "AdaGN expects [batch, points, channels]"
        B, N, C = x.shape
        groups = self.gn.num_groups
        xf = x.to(torch.float32).contiguous().view(B * N, C)
        stats = ops.group_stats(xf, N, N, C // groups)
        t = ctx.to(torch.float32).reshape(B, -1).contiguous()
        out, _ = ops.adagn(xf, stats, C // groups, t, self.scale.weight, self.scale.bias, self.bias.weight,
                           self.bias.bias, rows_per_cloud=N, valid_rows=N, groups=groups, eps=self.gn.eps, out_f32=True)
        return out.view(B, N, C).to(x.dtype)
