"""Host-side mirror of the reference's model/model_utils.py helpers that sit on the STAGE forward path."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .. import ops


def mask_logits(target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """``target * mask + (1 - mask) * (-1e10)`` (reference model/model_utils.py:14-15) as one CUDA kernel.

    ``mask`` must cover the leading axes of ``target`` (trailing singleton axes allowed), which is how every call site
    of the reference uses it (model/stage.py:425,429,456,458,503,521,533)."""
    return ops.mask_logits_fwd(target, _leading_mask(target, mask))


def _leading_mask(target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Bring ``mask`` to a contiguous tensor shaped like the leading axes of ``target``."""
    if mask.dim() > target.dim():
        raise RuntimeError("mask_logits: mask has more axes than target")
    m = mask
    while m.dim() < target.dim():
        m = m.unsqueeze(-1)
    # strip trailing broadcast axes (those where the mask is 1 and the target is not)
    keep = target.dim()
    while keep > 0 and m.shape[keep - 1] == 1 and (target.shape[keep - 1] != 1 or keep == target.dim()):
        keep -= 1
    m = m.reshape(m.shape[:keep]) if keep < m.dim() else m
    lead_shape = target.shape[:m.dim()]
    if tuple(m.shape) != tuple(lead_shape):
        m = m.expand(lead_shape)
    return m.contiguous().float()


def flat_list_of_lists(l: Sequence[Sequence]) -> list:
    """[[1, 2], [3, 4]] -> [1, 2, 3, 4] (model/model_utils.py:18-20)."""
    out: list = []
    for sub in l:
        out.extend(sub)
    return out


def compute_temporal_iou(pred, gt) -> float:
    """Temporal IoU whose denominator is the hull of the two spans, as in model/model_utils.py:37-51."""
    inter = min(pred[1], gt[1]) - max(pred[0], gt[0])
    hull = max(pred[1], gt[1]) - min(pred[0], gt[0])
    if hull == 0:
        return 0
    return 1.0 * max(0, inter) / hull


def get_high_iou_sapns(gt_ts_list, pred_ts_list, iou_thd: float = 0.5, add_gt: bool = True) -> List[list]:
    """Per example: [gt span] + predicted spans whose IoU with it reaches ``iou_thd`` (model/model_utils.py:54-75)."""
    out = []
    for gt, preds in zip(gt_ts_list, pred_ts_list):
        cur = [gt] if add_gt else []
        cur += [p[:2] for p in preds if compute_temporal_iou(p[:2], gt) >= iou_thd]
        out.append(cur)
    return out


def expand_span(span, expand_length: int = 2) -> list:
    """[max(0, st - e), ed + e]; the right end is left unclipped because it only slices (model/model_utils.py:78-89)."""
    return [max(0, span[0] - expand_length), span[1] + expand_length]


def find_max_triples(p1: torch.Tensor, p2: torch.Tensor, topN: int = 5, prob_thd: Optional[float] = None):
    """Per row the topN (st, ed, p1[st]*p2[ed]) with st <= ed, best first (model/model_utils.py:92-121).

    topN == 1 (the only value the forward path uses, model/stage.py:411-414,442-444) runs the O(T) prefix-max kernel;
    larger topN falls back to one device top-k over the masked outer product.  One device->host copy of the result."""
    if topN == 1:
        spans, conf = ops.find_max_span(p1, p2)
        spans, conf = spans.cpu().tolist(), conf.cpu().tolist()
        res = [[(s[0], s[1], c)] for s, c in zip(spans, conf)]
    else:
        T = p1.shape[1]
        prod = torch.triu(p1.unsqueeze(2) * p2.unsqueeze(1)).flatten(1)
        val, idx = prod.topk(min(topN, T * T), dim=1)
        val, idx = val.cpu().tolist(), idx.cpu().tolist()
        res = [[(i // T, i % T, v) for i, v in zip(ii, vv)] for ii, vv in zip(idx, val)]
    if prob_thd is not None:
        res = [[t for t in row if t[2] >= prob_thd] for row in res]
    return res
