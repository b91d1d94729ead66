"""Cosine-similarity attention between QA tokens and per-frame context items (reference
model/context_query_attention.py:7-101).  ``StructuredAttention`` has no parameters.

Inside ``STAGE`` the attention is fused with the projection MLPs (``ops.cqa_fused_fwd``); this standalone module
runs the same kernel in its attention-only mode and returns the reference's 4-tuple ``(A, S, S_mask, S_)``.
"""
import torch
import torch.nn as nn

from .. import ops


class StructuredAttention(nn.Module):
    def __init__(self, dropout=0.1, scale=100, add_void=False):
        super().__init__()
        self.dropout = dropout
        self.scale = scale
        self.add_void = add_void

    def forward(self, C, Q, c_mask, q_mask, noun_mask=None, void_vector=None):
        """C (N,5,1|Li,Lqa,D), Q (N,1,Li,Lr,D), c_mask (N,5,1|Li,Lqa), q_mask (N,1,Li,Lr)
        -> A (N,5,Li,Lqa,D), S (N,5,Li,Lqa,Lr) masked cosine scores, S_mask, S_ (masked softmax weights)."""
        if noun_mask is not None or void_vector is not None:
            raise NotImplementedError("tvqaplus_b200: noun_mask / void_vector are never used by STAGE.forward "
                                      "(model/stage.py:247-249,268-270) and are not built")
        if self.training and self.dropout > 0:
            raise NotImplementedError("tvqaplus_b200: dropout on the normalised inputs (train mode) is not built")
        if C.shape[2] != 1:
            raise NotImplementedError("tvqaplus_b200: C must be frame-invariant (N,5,1,Lqa,D), as STAGE passes it")
        N, A, _, Lq, D = C.shape
        T, Lr = Q.shape[2], Q.shape[3]
        out = ops.attention_only_fwd(C.reshape(N, A, Lq, D), c_mask.reshape(N, A, Lq), Q.reshape(N, T, Lr, D),
                                     q_mask.reshape(N, T, Lr), float(self.scale))
        S_mask = torch.matmul(c_mask.unsqueeze(-1), q_mask.unsqueeze(-2))
        return out["att"], out["raw_s"], S_mask, out["norm_s"]
