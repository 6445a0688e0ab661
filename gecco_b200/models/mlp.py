"""Per-point MLP parameter container (reference: gecco_torch/models/mlp.py:5-39).

Linear -> activation -> [Linear -> activation] x (depth-1) -> Linear as an nn.Sequential, so the state_dict keys
are `0.weight, 0.bias, 1.alpha, 2.weight, 2.bias` for depth 1.  The CUDA path supports depth 1 with
GaussianActivation (the only configuration the reference uses, set_transformer.py:80-83,148-150).
"""
from typing import Callable

import torch.nn as nn


class MLP(nn.Sequential):
    def __init__(self, in_features: int, out_features: int, width_size: int, depth: int = 1,
                 activation: Callable = nn.ReLU):
        mods = [nn.Linear(in_features, width_size), activation()]
        for _ in range(depth - 1):
            mods += [nn.Linear(width_size, width_size), activation()]
        mods.append(nn.Linear(width_size, out_features))
        super().__init__(*mods)
        self.depth = depth

    def forward(self, x):
        raise NotImplementedError(
            "gecco_b200: MLP runs fused inside the denoiser engine; call Diffusion / EDMPrecond / the network module")
