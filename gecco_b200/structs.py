"""Containers on the call surface (reference: gecco_torch/structs.py:61-91, models/feature_pyramid.py:17-20)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional

from torch import Tensor


def _map_fields(obj, f: Callable[[Tensor], Tensor]):
    """Out-of-place map over the tensor fields of a NamedTuple, recursing into nested containers
    (same contract as structs.py:37-50)."""
    new = {}
    for name, value in obj._asdict().items():
        if hasattr(value, "apply_to_tensors"):
            new[name] = value.apply_to_tensors(f)
        elif isinstance(value, Tensor):
            new[name] = f(value)
        else:
            new[name] = value
    return type(obj)(**new)


def _describe(obj) -> str:
    parts = []
    for name, value in obj._asdict().items():
        parts.append(f"{name}={tuple(value.shape) if isinstance(value, Tensor) else value!r}")
    return f"{type(obj).__name__}({', '.join(parts)})"


class Context3d(NamedTuple):
    """Conditioning image [B,3,H,W] and 3x3 camera intrinsics [B,3,3] in normalised image units."""

    image: Tensor
    K: Tensor

    apply_to_tensors = _map_fields
    __repr__ = _describe


class Example(NamedTuple):
    """A point cloud [B,N,3] with its (optional) context."""

    data: Tensor
    ctx: Optional[Context3d]

    apply_to_tensors = _map_fields
    __repr__ = _describe


@dataclass
class FeaturePyramidContext:
    """What a conditioner hands to the denoiser: NCHW feature maps and the camera matrices."""

    features: list
    K: Tensor
