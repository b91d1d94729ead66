"""Multi-head self-attention (reference model/self_attention.py:19-71): four ``Linear(d_model, d_model)`` kept under
``linears.0..3`` so the state_dict matches; the QUERY-row mask quirk of the reference (padded queries attend
uniformly, valid queries see padded keys) is reproduced by the kernel."""
import copy

import torch.nn as nn

from .. import ops
from ._packing import packed


class MultiHeadedAttention(nn.Module):
    def __init__(self, nh, d_model, dropout=0.1):
        super().__init__()
        assert d_model % nh == 0
        self.d_k = d_model // nh
        self.nh = nh
        proto = nn.Linear(d_model, d_model)  # the reference deep-copies ONE layer four times (self_attention.py:14-16,32)
        self.linears = nn.ModuleList([copy.deepcopy(proto) for _ in range(4)])
        self.dropout = nn.Dropout(p=dropout)

    def attn_args(self, ln=None) -> dict:
        return dict(ln_g=ln.weight if ln is not None else None, ln_b=ln.bias if ln is not None else None,
                    w_packed=[packed(l.weight, 64) for l in self.linears], b=[l.bias for l in self.linears])

    def forward(self, x, mask=None):
        """x (N, L, D), mask (N, L) -> (N, L, D)."""
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("tvqaplus_b200: attention dropout (train mode) is not built; call .eval()")
        return ops.mha_fwd(x, mask, self.attn_args(), self.nh)
