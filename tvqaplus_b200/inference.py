"""Inference post-processing of the reference's ``inference.py`` that sits directly behind ``STAGE.forward``
(inference.py:13-35,63-72): the span of the predicted answer.  The Python loops over frames become one kernel over
the whole batch and one device->host copy."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ops


def find_max_pair(p1: Sequence[float], p2: Sequence[float], device="cuda") -> Tuple[Tuple[int, int], float]:
    """Reference signature (inference.py:13-35): lists of start / end probabilities -> ((st, ed), max_value).  Runs the
    batched device kernel on one row; use ``predicted_spans`` for a batch."""
    probs = torch.stack([torch.as_tensor(p1, dtype=torch.float32), torch.as_tensor(p2, dtype=torch.float32)], dim=1)
    spans, val = ops.find_max_pairs(probs.unsqueeze(0).to(device))
    return (int(spans[0, 0]), int(spans[0, 1])), float(val[0])


def predicted_spans(answer: torch.Tensor, t_scores: torch.Tensor):
    """The loop body of inference.py:58-72 for a whole batch: ``answer`` (N,5) scores and ``t_scores`` (N,5,T,2) span
    probabilities (the inference-mode outputs of STAGE) -> (pred_ids (N,), spans (N,2), values (N,)) on the device; the
    caller turns [st, ed] into seconds (``st*2 + offset, (ed+1)*2 + offset``)."""
    pred = answer.max(1)[1]
    rows = torch.arange(answer.shape[0], device=answer.device)
    spans, val = ops.find_max_pairs(t_scores[rows, pred].contiguous())
    return pred, spans, val
