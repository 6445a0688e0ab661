"""Op-level Python entry points over the C ABI (one function per exported kernel family).

Torch is used for memory, streams and dtype plumbing only; every computation below runs in
the hand-written sm_100a kernels of libgecco_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _abi


def _stream(t: Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _lib_for(t: Tensor):
    if not t.is_cuda:
        raise _abi.GeccoError("gecco_b200 ops need CUDA tensors (there is no CPU path)")
    return _abi.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def gemm(
    a: Tensor,
    w: Tensor,
    *,
    bias: Tensor | None = None,
    bias_stride: int = 0,
    act_alpha: float | None = None,
    res: Tensor | None = None,
    out_f32: Tensor | bool | None = None,
    out_bf16: Tensor | bool | None = None,
    stats: Tensor | None = None,
    rows_per_cloud: int | None = None,
    valid_rows: int | None = None,
    w_rows_per_cloud: int = 0,
    n_out: int | None = None,
    geom: Tensor | None = None,
    sigma: Tensor | None = None,
    sigma_stride: int = 0,
    wx: Tensor | None = None,
):
    """out = epilogue(a @ w.T) on the tcgen05 tensor cores; see gecco_gemm in include/gecco_b200.h."""
    lib = _lib_for(a)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    if n_out is None:
        n_out = w.shape[0]
    if rows_per_cloud is None:
        rows_per_cloud = _round_up(m, 32)
        valid_rows = m
    if valid_rows is None:
        valid_rows = rows_per_cloud
    if out_f32 is True:
        out_f32 = torch.empty((m, n_out), device=a.device, dtype=torch.float32)
    if out_bf16 is True:
        out_bf16 = torch.empty((m, n_out), device=a.device, dtype=torch.bfloat16)
    if out_f32 is False:
        out_f32 = None
    if out_bf16 is False:
        out_bf16 = None
    args = _abi.GemmArgs()
    args.a, args.lda = a.data_ptr(), a.stride(0)
    args.w, args.ldw = w.data_ptr(), w.stride(0)
    args.m, args.n_out, args.k = m, n_out, k
    args.rows_per_cloud, args.valid_rows, args.w_rows_per_cloud = rows_per_cloud, valid_rows, w_rows_per_cloud
    args.bias = 0 if bias is None else bias.data_ptr()
    args.bias_stride = bias_stride
    args.act = 0 if act_alpha is None else 1
    args.act_alpha = 1.0 if act_alpha is None else float(act_alpha)
    if res is not None:
        assert res.dtype == torch.float32 and res.stride(1) == 1
        args.res, args.ldr = res.data_ptr(), res.stride(0)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.stride(1) == 1
        args.out_f32, args.ldo32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.stride(1) == 1
        args.out_bf16, args.ldo16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if stats is not None:
        assert stats.dtype == torch.float64 and stats.is_contiguous()
        args.stats = stats.data_ptr()
    if geom is not None:
        assert geom.dtype == torch.float32 and geom.is_contiguous()
        assert sigma is not None and wx is not None and sigma.dtype == torch.float32 and wx.dtype == torch.float32
        args.geom, args.sigma, args.sigma_stride, args.wx = geom.data_ptr(), sigma.data_ptr(), sigma_stride, wx.data_ptr()
    _abi.check(lib.gecco_gemm(C.byref(args), _stream(a)))
    return out_f32, out_bf16
