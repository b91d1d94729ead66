"""Cache of split-fp16 weight images, keyed on the parameter's storage and version counter so an optimizer step (or
``load_state_dict``) invalidates it; kernels otherwise read parameter storage in place.

The version counter does NOT move for in-place edits through ``.data`` (``p.data.copy_()``, EMA swaps, manual weight
surgery): after such an edit call ``invalidate_weight_cache()`` (or ``STAGE.invalidate_weight_cache()``), otherwise the
kernels keep using the images packed from the old values."""
from __future__ import annotations

import weakref
from typing import Dict, Tuple

import torch

from .. import ops

_cache: Dict[int, Tuple[tuple, torch.Tensor, "weakref.ref"]] = {}
_epoch = 0  # bumped by invalidate_weight_cache(); part of every tag (module-level tags such as STAGE._cqa_tag read it too)


def epoch() -> int:
    return _epoch


def invalidate_weight_cache() -> None:
    """Drop every packed weight image; the next forward re-packs from the current parameter values."""
    global _epoch
    _epoch += 1
    _cache.clear()
    _cache_ln.clear()


def packed(param: torch.Tensor, kc: int) -> torch.Tensor:
    key = id(param)
    tag = (param.data_ptr(), param._version, kc, str(param.device), tuple(param.shape))
    hit = _cache.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is param:
        return hit[1]
    img = ops.pack_weight(param.detach(), kc)
    _cache[key] = (tag, img, weakref.ref(param, lambda _r, k=key: _cache.pop(k, None)))
    return img


_cache_ln: Dict[tuple, Tuple[tuple, tuple, "weakref.ref"]] = {}


def packed_ln(lin, ln, kc: int):
    """(image of W*gamma, cs, cb) for ``lin`` (nn.Linear) applied after ``ln`` (nn.LayerNorm); cached on the versions of
    all four parameters."""
    params = (lin.weight, lin.bias, ln.weight, ln.bias)
    key = (id(lin.weight), id(ln.weight))
    tag = tuple((t.data_ptr(), t._version, str(t.device)) for t in params) + (kc,)
    hit = _cache_ln.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is lin.weight:
        return hit[1]
    out = ops.pack_weight_ln(lin.weight, lin.bias, ln.weight, ln.bias, kc)
    _cache_ln[key] = (tag, out, weakref.ref(lin.weight, lambda _r, k=key: _cache_ln.pop(k, None)))
    return out
