"""Inducer-point SetTransformer: parameter containers with the reference's module tree
(gecco_torch/models/set_transformer.py:14-216), so `state_dict()` keys and random initialisation match.

The computation does not run module by module: `gecco_b200.engine.Engine` walks this tree once, packs the
weights and runs the whole stack in the CUDA engine (csrc/engine.cu).  `SetTransformer.forward` is routed
through the same engine (features in, features out).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .mlp import MLP
from .normalization import AdaGN


class AttentionPool(nn.Module):
    """points -> inducers cross-attention with learned queries (set_transformer.py:14-65)."""

    def __init__(self, feature_dim: int, num_heads: int, num_inducers: int):
        super().__init__()
        assert feature_dim % num_heads == 0, (feature_dim, num_heads)
        self.num_heads = num_heads
        self.feature_dim = feature_dim
        self.dims_per_head = feature_dim // num_heads
        self.inducers = nn.Parameter(torch.randn(1, num_heads, num_inducers, self.dims_per_head))
        self.kv_proj = nn.Linear(feature_dim, 2 * feature_dim, bias=False)
        self.out_proj = nn.Linear(feature_dim, feature_dim, bias=False)


class Broadcast(nn.Module):
    """pool -> AdaGN -> MLP -> AdaGN on the inducers, then inducers -> points attention (set_transformer.py:68-117).
    `unpool` is an nn.MultiheadAttention so that its parameter names and initialisation are the reference's."""

    def __init__(self, feature_dim: int, num_inducers: int, t_embed_dim: int, num_heads: int = 8, mlp_blowup: int = 2,
                 activation: nn.Module = nn.ReLU):
        super().__init__()
        self.pool = AttentionPool(feature_dim, num_heads, num_inducers)
        self.norm_1 = AdaGN(feature_dim, t_embed_dim)
        self.mlp = MLP(feature_dim, feature_dim, mlp_blowup * feature_dim, activation=activation)
        self.norm_2 = AdaGN(feature_dim, t_embed_dim)
        self.unpool = nn.MultiheadAttention(feature_dim, num_heads, batch_first=True)


class BroadcastingLayer(nn.Module):
    """Pre-norm residual block: x += Broadcast(AdaGN(x)); x += MLP(AdaGN(x)) (set_transformer.py:120-168)."""

    def __init__(self, feature_dim: int, num_inducers: int, embed_dim: int, num_heads: int = 8, mlp_blowup: int = 2,
                 activation: nn.Module = nn.ReLU):
        super().__init__()
        self.broadcast_norm = AdaGN(feature_dim, embed_dim)
        self.broadcast = Broadcast(feature_dim, num_inducers, embed_dim, num_heads, mlp_blowup=mlp_blowup,
                                   activation=activation)
        self.mlp_norm = AdaGN(feature_dim, embed_dim)
        self.mlp = MLP(feature_dim, feature_dim, mlp_blowup * feature_dim, activation=activation)
        with torch.no_grad():  # residual branches start small (:150-153)
            self.broadcast.unpool.out_proj.weight.mul_(0.1)
            self.mlp[-1].weight.mul_(0.1)


class SetTransformer(nn.Module):
    def __init__(self, n_layers: int, feature_dim: int, num_inducers: int, t_embed_dim: int, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList(
            [BroadcastingLayer(feature_dim=feature_dim, num_inducers=num_inducers, embed_dim=t_embed_dim, **kwargs)
             for _ in range(n_layers)])
        self.feature_dim = feature_dim

    def forward(self, features: Tensor, t_embed: Tensor, return_h: bool = False, hs: list | None = None):
        raise NotImplementedError(
            "gecco_b200: the SetTransformer stack runs inside the denoiser engine; call it through LinearLift / "
            "RayNetwork / EDMPrecond / Diffusion")
