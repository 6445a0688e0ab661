"""Gaussian activation (reference: gecco_torch/models/activation.py:5-24).

g(x) = exp(-x^2 / (2 alpha^2)), optionally normalised to (g - 0.7) / 0.28, with one learnable scalar alpha.
On the hot path it is applied inside the epilogue of the tcgen05 GEMM that produces its input
(csrc/epilogue.cuh, `act`); called on its own, `forward` runs the `gecco_gaussian_activation` kernel.
"""
import torch
import torch.nn as nn


class GaussianActivation(nn.Module):
    def __init__(self, normalized: bool = True):
        super().__init__()
        self.alpha = nn.Parameter(torch.tensor(1.0))
        self.normalized = normalized

    @torch.no_grad()
    def forward(self, x):
        from .. import ops

        return ops.gaussian_activation(x, float(self.alpha), self.normalized)
