"""Conditioner producing the feature pyramid the projective lookup gathers from
(reference: gecco_torch/models/feature_pyramid.py:28-73).

The ConvNeXt runs ONCE per sampling call and is upstream of the hot path: it stays a
torchvision module; its outputs are re-laid-out to bf16 channels-last by `gecco_pack_features` when they enter
the engine.
"""
from __future__ import annotations

from typing import Literal

import torch
from torch import nn

from ..structs import Context3d, FeaturePyramidContext  # noqa: F401  (re-exported like the reference)


class FeaturePyramidExtractor(nn.Module):
    def forward(self, ctx_raw: Context3d) -> FeaturePyramidContext:
        raise NotImplementedError()


class ConvNeXtExtractor(FeaturePyramidExtractor):
    def __init__(self, n_stages: int = 3, model: Literal["tiny", "small"] = "tiny", pretrained: bool = True):
        super().__init__()
        import torchvision.models as tvm

        if model == "tiny":
            net = tvm.convnext_tiny(weights=tvm.ConvNeXt_Tiny_Weights.DEFAULT if pretrained else None)
        elif model == "small":
            net = tvm.convnext_small(weights=tvm.ConvNeXt_Small_Weights.DEFAULT if pretrained else None)
        else:
            raise ValueError(f"Unknown model {model}")
        feats = list(net.features)
        # (downsampling, processing) pairs; keep the first n_stages
        self.stages = nn.ModuleList([nn.Sequential(feats[i], feats[i + 1]) for i in range(0, len(feats), 2)][:n_stages])
        for m in self.modules():  # stochastic depth off, like the reference (:56-60)
            if isinstance(m, tvm.convnext.CNBlock):
                m.stochastic_depth = torch.nn.Identity()

    def forward(self, raw_ctx: Context3d) -> FeaturePyramidContext:
        x = raw_ctx.image
        maps = []
        for stage in self.stages:
            x = stage(x)
            maps.append(x)
        return FeaturePyramidContext(features=maps, K=raw_ctx.K)
