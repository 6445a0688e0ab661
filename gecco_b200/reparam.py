"""Reparametrisation schemes between "data" space (xyz) and "diffusion" space.

Same classes, constructor arguments, buffer names and method names as gecco_torch/reparam.py:14-201;
the arithmetic runs in the `gecco_reparam` CUDA kernel (csrc/elementwise.cu).  There is no CPU path.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops
from .structs import Context3d


class Reparam(torch.nn.Module):
    """Base class (reparam.py:14-28)."""

    _kind = 0

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self._host_cache = None

    # host copies of the (tiny) statistics buffers, refreshed when the buffers change
    def _host_stats(self):
        return None, None, 1.1

    def _run(self, x: Tensor, ctx, to_data: bool) -> Tensor:
        mean, sigma, logit_scale = self._host_stats()
        K = None
        if self._kind == 2:
            if not isinstance(ctx, Context3d):
                raise AssertionError("UVLReparam needs a Context3d")
            K = ctx.K
        in_dtype = x.dtype
        if x.dtype not in (torch.float32, torch.float64):
            x = x.float()
        out = ops.reparam(x, self._kind, to_data, mean, sigma, logit_scale, K)
        return out if out.dtype == in_dtype else out.to(in_dtype)

    def data_to_diffusion(self, data: Tensor, ctx: Context3d) -> Tensor:
        raise NotImplementedError()

    def diffusion_to_data(self, diff: Tensor, ctx: Context3d) -> Tensor:
        raise NotImplementedError()


class NoReparam(Reparam):
    """Identity (reparam.py:31-40)."""

    def data_to_diffusion(self, data: Tensor, ctx: Context3d) -> Tensor:
        return data

    def diffusion_to_data(self, diff: Tensor, ctx: Context3d) -> Tensor:
        return diff


class _StatReparam(Reparam):
    _mean_name = "mean"
    _sigma_name = "sigma"

    def _host_stats(self):
        m, s = getattr(self, self._mean_name), getattr(self, self._sigma_name)
        key = (m.data_ptr(), m._version, s.data_ptr(), s._version)
        if self._host_cache is None or self._host_cache[0] != key:
            self._host_cache = (key, [float(v) for v in m.detach().flatten().tolist()],
                                [float(v) for v in s.detach().flatten().tolist()])
        return self._host_cache[1], self._host_cache[2], float(getattr(self, "logit_scale", 1.1))


class GaussianReparam(_StatReparam):
    """(data - mean) / sigma and back (reparam.py:43-66)."""

    _kind = 1

    def __init__(self, mean: Tensor, sigma: Tensor):
        assert mean.ndim == 1
        assert mean.shape == sigma.shape
        super().__init__(mean.shape[0])
        if mean.shape[0] != 3:
            raise ValueError("gecco_b200 supports 3-dimensional geometry only")
        self.register_buffer("mean", mean)
        self.register_buffer("sigma", sigma)

    def data_to_diffusion(self, data: Tensor, ctx: Context3d) -> Tensor:
        return self._run(data, None, to_data=False)

    def diffusion_to_data(self, diff: Tensor, ctx: Context3d) -> Tensor:
        return self._run(diff, None, to_data=True)

    def extra_repr(self) -> str:
        return f"mean={self.mean.flatten().tolist()}, sigma={self.sigma.flatten().tolist()}"


class UVLReparam(_StatReparam):
    """Image-plane (u, v) through arctanh and ray length through log, then normalisation
    (reparam.py:69-201).  Buffers are named uvl_mean / uvl_std like the reference."""

    _kind = 2
    _mean_name = "uvl_mean"
    _sigma_name = "uvl_std"

    def __init__(self, mean: Tensor, sigma: Tensor, logit_scale: float = 1.1):
        assert mean.shape == (3,)
        assert sigma.shape == (3,)
        super().__init__(dim=3)
        self.register_buffer("uvl_mean", mean)
        self.register_buffer("uvl_std", sigma)
        self.logit_scale = logit_scale

    def data_to_diffusion(self, data: Tensor, ctx: Context3d) -> Tensor:
        assert isinstance(ctx, Context3d)
        return self._run(data, ctx, to_data=False)

    def diffusion_to_data(self, diff: Tensor, ctx: Context3d) -> Tensor:
        assert isinstance(ctx, Context3d)
        return self._run(diff, ctx, to_data=True)

    def extra_repr(self) -> str:
        return (f"uvl_mean={self.uvl_mean.flatten().tolist()}, uvl_std={self.uvl_std.flatten().tolist()}, "
                f"logit_scale={self.logit_scale}")
