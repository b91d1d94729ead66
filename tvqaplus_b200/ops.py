"""Thin Python wrappers over the C ABI (one function per entry point of ``include/stage_b200.h``).

All tensors must be contiguous fp32 CUDA tensors; outputs are allocated here with ``torch.empty`` (the library
itself never allocates).  There is no CPU path: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, current_stream, lib, ptr

HSZ = 128


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("tvqaplus_b200: expected float32, got %s" % t.dtype)
    return t.contiguous()


def kc_for(n: int) -> int:
    """K-chunk the GEMM kernels stream for an output width n."""
    return 64 if n <= 128 else 32


def pack_weight(w: torch.Tensor, kc: Optional[int] = None) -> torch.Tensor:
    """Split W[N,K] (or a 1x1 conv weight [N,K,1]) into the fp16 hi/lo operand image (stage_pack_weight)."""
    w = _f32(w.detach())
    if w.dim() == 3:
        w = w.reshape(w.shape[0], w.shape[1])
    n, k = w.shape
    kc = kc or kc_for(n)
    nbytes = lib().stage_packed_weight_bytes(n, k, kc)
    if nbytes == 0:
        raise RuntimeError("tvqaplus_b200: unsupported weight shape %s" % (tuple(w.shape),))
    out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib().stage_pack_weight(ptr(w), n, k, kc, ptr(out), current_stream()), "stage_pack_weight")
    out._stage_kc = kc  # type: ignore[attr-defined]
    return out


def linear_fwd(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, n: int, *, kc: Optional[int] = None,
               ln_in=None, ln_out=None, l2norm_in: bool = False, relu: bool = False) -> torch.Tensor:
    """y = LN_out(relu?(LN_in(l2norm?(x)) W^T + b)) over the last axis of x (stage_linear_fwd)."""
    x = _f32(x)
    k = x.shape[-1]
    m = x.numel() // k
    y = torch.empty(x.shape[:-1] + (n,), dtype=torch.float32, device=x.device)
    a = _lib.LinearArgs(m=m, k=k, n=n, kc=kc or kc_for(n), l2norm_in=int(l2norm_in), relu=int(relu),
                        x=ptr(x), w_packed=ptr(w_packed), bias=ptr(_f32(bias)),
                        ln_in_g=ptr(_f32(ln_in[0])) if ln_in else None, ln_in_b=ptr(_f32(ln_in[1])) if ln_in else None,
                        ln_out_g=ptr(_f32(ln_out[0])) if ln_out else None,
                        ln_out_b=ptr(_f32(ln_out[1])) if ln_out else None, y=ptr(y))
    check(lib().stage_linear_fwd(C.byref(a), current_stream()), "stage_linear_fwd")
    return y
