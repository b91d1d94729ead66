"""Thin wrappers over the training primitives of ``libstage_b200.so`` (``csrc/train.cuh`` / ``train_b200.cu``).

Every function launches one kernel on the current stream of the tensors' device and returns freshly allocated outputs
(``torch.empty`` - the caching allocator owns all memory).  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GemmArgs, current_stream, lib, on_device

LAUNCHES = 0  # kernels launched through this module (bench.py's gpu_launches claim for the training step)


def check(rc: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    _lib.check(rc, what)


def _p(t: Optional[torch.Tensor], off: int = 0) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("tvqaplus_b200: expected a CUDA tensor (no CPU path exists)")
    if not t.is_contiguous():
        raise RuntimeError("tvqaplus_b200: expected a contiguous tensor")
    return t.data_ptr() + off * t.element_size()


def gemm(a, b, c, m, n, k, a_st, b_st, ldc, *, batch=(1, 1), a_b=(0, 0), b_b=(0, 0), c_b=(0, 0), bias=None, relu=False,
         mode=0, split_k=1, alpha=1.0, a_off=0, b_off=0, c_off=0) -> None:
    """C (op)= alpha * A * B (+ bias)(ReLU) with element strides ``a_st = (a_sm, a_sk)``, ``b_st = (b_sk, b_sn)``."""
    g = GemmArgs(a=_p(a, a_off), b=_p(b, b_off), c=_p(c, c_off), bias=_p(bias), m=m, n=n, k=k, nb0=batch[0], nb1=batch[1],
                 a_sm=a_st[0], a_sk=a_st[1], b_sk=b_st[0], b_sn=b_st[1], ldc=ldc, a_b0=a_b[0], a_b1=a_b[1], b_b0=b_b[0],
                 b_b1=b_b[1], c_b0=c_b[0], c_b1=c_b[1], split_k=split_k, mode=mode, relu=int(relu), alpha=alpha)
    with on_device(a, b, c):
        check(lib().stage_gemm(C.byref(g), current_stream()), "stage_gemm")


def _split_for(m: int, n: int, k: int, target_ctas: int = 592) -> int:
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    return max(1, min(target_ctas // max(tiles, 1), (k + 511) // 512))


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False, out: Optional[torch.Tensor] = None,
           ldc: Optional[int] = None, c_off: int = 0) -> torch.Tensor:
    """y[M, N] = x[M, K] . w[N, K]^T + bias (F.linear)."""
    m, k = x.shape
    n = w.shape[0]
    w2 = w.reshape(n, k)
    y = torch.empty(m, n, device=x.device, dtype=torch.float32) if out is None else out
    gemm(x, w2, y, m, n, k, (k, 1), (1, k), n if ldc is None else ldc, bias=bias, relu=relu, c_off=c_off)
    return y


def linear_bwd_x(dy: torch.Tensor, w: torch.Tensor, dy_ld: Optional[int] = None, dy_off: int = 0, n: Optional[int] = None) -> torch.Tensor:
    """dx[M, K] = dy[M, N] . w[N, K]."""
    m = dy.shape[0]
    n = dy.shape[1] if n is None else n
    k = w.numel() // n
    dx = torch.empty(m, k, device=dy.device, dtype=torch.float32)
    gemm(dy, w.reshape(n, k), dx, m, k, n, (dy.shape[1] if dy_ld is None else dy_ld, 1), (k, 1), k, a_off=dy_off)
    return dx


def linear_bwd_w(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, db: Optional[torch.Tensor], dy_ld: Optional[int] = None,
                 dy_off: int = 0, n: Optional[int] = None) -> None:
    """dw[N, K] += dy[M, N]^T . x[M, K]; db[N] += column sums of dy (both accumulate: shared weights)."""
    m, k = x.shape
    n = dy.shape[1] if n is None else n
    ld = dy.shape[1] if dy_ld is None else dy_ld
    gemm(dy, x, dw, n, k, m, (1, ld), (k, 1), k, mode=2, split_k=_split_for(n, k, m), a_off=dy_off)
    if db is not None:
        with on_device(dy, db):
            check(lib().stage_colsum(_p(dy, dy_off), m, n, ld, _p(db), current_stream()), "stage_colsum")


def row_absmax(x: torch.Tensor) -> torch.Tensor:
    rows, d = x.shape
    out = torch.empty(rows, device=x.device, dtype=torch.float32)
    with on_device(x):
        check(lib().stage_row_absmax(_p(x), rows, d, _p(out), current_stream()), "stage_row_absmax")
    return out


def ln_fwd(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    rows, d = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    with on_device(x, g, b):
        check(lib().stage_ln_fwd(_p(x), _p(g), _p(b), _p(y), _p(mean), _p(rstd), rows, d, current_stream()), "stage_ln_fwd")
    return y, (mean, rstd)


def ln_bwd(dy, x, stats, g, dg, db, *, add=None, relu_x=False, want_dx=True, out=None) -> Optional[torch.Tensor]:
    rows, d = x.shape
    dx = None
    if want_dx:
        dx = torch.empty_like(x) if out is None else out
    with on_device(dy, x, g, dg, db):
        check(lib().stage_ln_bwd(_p(dy), _p(x), _p(stats[0]), _p(stats[1]), _p(g), _p(add), _p(dx), _p(dg), _p(db), rows, d,
                                 int(relu_x), current_stream()), "stage_ln_bwd")
    return dx


def relu_bwd(dy, y, out=None):
    dx = torch.empty_like(dy) if out is None else out
    with on_device(dy, y):
        check(lib().stage_relu_bwd(_p(dy), _p(y), _p(dx), dy.numel(), current_stream()), "stage_relu_bwd")
    return dx


def add(a, b, period: int = 0, out=None):
    y = torch.empty_like(a) if out is None else out
    with on_device(a, b):
        check(lib().stage_add(_p(a), _p(b), _p(y), a.numel(), period, current_stream()), "stage_add")
    return y


def scale_(x, scalar: torch.Tensor, mult: float = 1.0):
    with on_device(x, scalar):
        check(lib().stage_scale(_p(x), x.numel(), _p(scalar), mult, current_stream()), "stage_scale")
    return x


def dropout(x, p: float, seed_dev: torch.Tensor, site: int, out=None):
    """Inverted dropout keyed on (seed_dev, site, element index); ``out`` defaults to in place."""
    y = x if out is None else out
    with on_device(x, seed_dev):
        check(lib().stage_dropout(_p(x), _p(y), x.numel(), float(p), _p(seed_dev), site & 0xFFFFFFFF, current_stream()), "stage_dropout")
    return y


def permute_it(x, groups: int, i_dim: int, t_dim: int):
    d = x.shape[-1]
    y = torch.empty_like(x)
    with on_device(x):
        check(lib().stage_permute_it(_p(x), _p(y), groups, i_dim, t_dim, d, current_stream()), "stage_permute_it")
    return y


def dwconv_fwd(x, w, bias, n_seq: int, l: int, k: int):
    y = torch.empty_like(x)
    with on_device(x, w, bias):
        check(lib().stage_dwconv_fwd(_p(x), _p(w), _p(bias), _p(y), n_seq, l, x.shape[-1], k, current_stream()), "stage_dwconv_fwd")
    return y


def dwconv_bwd(dy, x, w, dw, db, n_seq: int, l: int, k: int):
    dx = torch.empty_like(dy)
    with on_device(dy, x, w, dw, db):
        check(lib().stage_dwconv_bwd(_p(dy), _p(x), _p(w), _p(dx), _p(dw), _p(db), n_seq, l, x.shape[-1], k, current_stream()),
              "stage_dwconv_bwd")
    return dx


def l2norm_fwd(x, want_inv: bool = True):
    rows, d = x.shape
    y = torch.empty_like(x)
    inv = torch.empty(rows, device=x.device, dtype=torch.float32) if want_inv else None
    with on_device(x):
        check(lib().stage_l2norm_fwd(_p(x), _p(y), _p(inv), rows, d, current_stream()), "stage_l2norm_fwd")
    return y, inv


def l2norm_bwd(dy, y, inv, dx, accumulate: bool):
    rows, d = y.shape
    with on_device(dy, y, inv, dx):
        check(lib().stage_l2norm_bwd(_p(dy), _p(y), _p(inv), _p(dx), rows, d, int(accumulate), current_stream()), "stage_l2norm_bwd")
    return dx


def attn_softmax_fwd(s, c_mask, q_mask, n: int, rows_per_n: int, t: int, lc: int, scale: float):
    p = torch.empty_like(s)
    with on_device(s, c_mask, q_mask):
        check(lib().stage_attn_softmax_fwd(_p(s), _p(c_mask), _p(q_mask), _p(p), n, rows_per_n, t, lc, scale, current_stream()),
              "stage_attn_softmax_fwd")
    return p


def attn_softmax_bwd_(dp, p, lc: int, scale: float):
    with on_device(dp, p):
        check(lib().stage_attn_softmax_bwd(_p(dp), _p(p), dp.numel() // lc, lc, scale, current_stream()), "stage_attn_softmax_bwd")
    return dp


def cat3_fwd(a, b, a_div: int):
    rows = b.shape[0]
    out = torch.empty(rows, 384, device=b.device, dtype=torch.float32)
    with on_device(a, b):
        check(lib().stage_cat3_fwd(_p(a), _p(b), _p(out), rows, a_div, current_stream()), "stage_cat3_fwd")
    return out


def cat3_bwd(dcat, a, b, da, a_div: int, accumulate_a: bool):
    db = torch.empty_like(b)
    with on_device(dcat, a, b, da):
        check(lib().stage_cat3_bwd(_p(dcat), _p(a), _p(b), _p(da), _p(db), a.shape[0], a_div, int(accumulate_a), current_stream()),
              "stage_cat3_bwd")
    return db


def maxpool_fwd(x, mask, n_seq: int, l: int):
    y = torch.empty(n_seq, 128, device=x.device, dtype=torch.float32)
    arg = torch.empty(n_seq, 128, device=x.device, dtype=torch.int32)
    with on_device(x, mask):
        check(lib().stage_maxpool_fwd(_p(x), _p(mask), _p(y), _p(arg), n_seq, l, current_stream()), "stage_maxpool_fwd")
    return y, arg


def maxpool_bwd(dy, mask, arg, n_seq: int, l: int):
    dx = torch.empty(n_seq * l, 128, device=dy.device, dtype=torch.float32)
    with on_device(dy, mask, arg):
        check(lib().stage_maxpool_bwd(_p(dy), _p(mask), _p(arg), _p(dx), n_seq, l, current_stream()), "stage_maxpool_bwd")
    return dx


def ts_loss_grad(t_scores, ts_mask, target, st, ed):
    """t_scores (N,5,T,2) masked -> (terms (N,2), probs (2,N,T), dscores (N,5,T,2) of 0.5*(CE_st+CE_ed))."""
    n, _, t, _ = t_scores.shape
    terms = torch.empty(n, 2, device=t_scores.device, dtype=torch.float32)
    probs = torch.empty(2, n, t, device=t_scores.device, dtype=torch.float32)
    ds = torch.zeros_like(t_scores)
    with on_device(t_scores, ts_mask, target, st, ed):
        check(lib().stage_ts_loss_grad(_p(t_scores), _p(ts_mask), _p(target), _p(st), _p(ed), n, t, _p(terms), _p(probs), _p(ds),
                                       current_stream()), "stage_ts_loss_grad")
    return terms, probs, ds


def proposal_pool_fwd(x, frame_mask, target, st, ed, spans, conf, n: int, t: int, extra: int, prob_thd: float):
    dev = x.device
    feat = torch.empty(2 * n * 5, 256, device=dev, dtype=torch.float32)
    arg = torch.empty(2 * n * 5, 256, device=dev, dtype=torch.int32)
    valid = torch.empty(2 * n, device=dev, dtype=torch.int32)
    row_target = torch.empty(2 * n, device=dev, dtype=torch.int64)
    with on_device(x, frame_mask, target, st, ed, spans, conf):
        check(lib().stage_proposal_pool_fwd(_p(x), _p(frame_mask), _p(target), _p(st), _p(ed), _p(spans), _p(conf), n, t, extra,
                                            prob_thd, _p(feat), _p(arg), _p(valid), _p(row_target), current_stream()),
              "stage_proposal_pool_fwd")
    return feat, arg, valid, row_target


def proposal_pool_bwd(dfeat, arg, valid, frame_mask, n: int, t: int, dx):
    with on_device(dfeat, arg, valid, frame_mask, dx):
        check(lib().stage_proposal_pool_bwd(_p(dfeat), _p(arg), _p(valid), _p(frame_mask), n, t, _p(dx), current_stream()),
              "stage_proposal_pool_bwd")
    return dx


def att_loss_bwd(scores, pos_off, neg_off, loss_type: int, alpha: float, margin: float, upstream, dscores):
    with on_device(scores, pos_off, neg_off, upstream, dscores):
        check(lib().stage_att_loss_bwd(_p(scores), _p(pos_off), _p(neg_off), pos_off.numel(), loss_type, alpha, margin,
                                       _p(upstream), _p(dscores), current_stream()), "stage_att_loss_bwd")
    return dscores


def answer_ce(rows, valid, row_target, n_examples: int):
    out = torch.empty(3, device=rows.device, dtype=torch.float32)
    drows = torch.empty_like(rows)
    with on_device(rows, valid, row_target):
        check(lib().stage_answer_ce(_p(rows), _p(valid), _p(row_target), rows.shape[0], n_examples, _p(out), _p(drows),
                                    current_stream()), "stage_answer_ce")
    return out, drows


def grad_sumsq(g, out):
    with on_device(g, out):
        check(lib().stage_grad_sumsq(_p(g), g.numel(), _p(out), current_stream()), "stage_grad_sumsq")
    return out


def adam_step(p, g, m, v, sumsq, *, max_norm: float, lr: float, beta1: float, beta2: float, eps: float, wd: float, step: int = 1,
              grad_scale: float = 1.0, step_dev: Optional[torch.Tensor] = None):
    """``step_dev``: int32 device tensor holding the step count (bias corrections computed on the device; graph-capturable)."""
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    with on_device(p, g, m, v):
        check(lib().stage_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(sumsq), max_norm, lr, beta1, beta2, eps, wd, bc1,
                                    bc2, grad_scale, _p(step_dev), current_stream()), "stage_adam_step")


def counter_inc(c: torch.Tensor):
    with on_device(c):
        check(lib().stage_counter_inc(_p(c), current_stream()), "stage_counter_inc")
