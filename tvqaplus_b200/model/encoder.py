"""Conv + (optional) self-attention encoder blocks (reference model/encoder.py:9-74).

``EncoderBlock.forward`` is ONE kernel launch: position add, every [LayerNorm -> depthwise conv -> pointwise GEMM ->
ReLU -> residual] round and the closing LayerNorm run on a shared-memory-resident tile (csrc/encoder.cuh); with heads
the attention sub-layer runs between the conv rounds and the closing LayerNorm.
"""
import torch.nn as nn

from .. import ops
from . import _packing
from .cnn import DepthwiseSeparableConv
from .position_encoding import PositionEncoding
from .self_attention import MultiHeadedAttention


def _no_train_dropout(mod, p):
    if mod.training and p > 0:
        raise NotImplementedError(
            "tvqaplus_b200: the forward kernels implement the eval/inference path (dropout off, no autograd); "
            "call model.eval() or construct with dropout=0")


class EncoderBlock(nn.Module):
    def __init__(self, n_conv, kernel_size=7, n_filters=128, dropout=0.1, num_heads=4):
        super().__init__()
        if n_filters != ops.HSZ:
            raise NotImplementedError("tvqaplus_b200: kernels are specialised for hidden size 128")
        self.dropout = dropout
        self.n_conv = n_conv
        self.num_heads = num_heads
        self.kernel_size = kernel_size
        self.position_encoding = PositionEncoding(n_filters=n_filters)
        self.layer_norm = nn.ModuleList([nn.LayerNorm(n_filters) for _ in range(n_conv)])
        self.final_layer_norm = nn.LayerNorm(n_filters)
        self.conv = nn.ModuleList([DepthwiseSeparableConv(in_ch=n_filters, out_ch=n_filters, k=kernel_size, relu=True)
                                   for _ in range(n_conv)])
        if self.num_heads != 0:
            self.multi_head_attn = MultiHeadedAttention(nh=num_heads, d_model=n_filters)
            self.attn_layer_norm = nn.LayerNorm(n_filters)

    def _blocked_pe(self, L):
        pe = self.position_encoding.pe
        tag = (L, pe.data_ptr(), pe._version, _packing.epoch())
        if getattr(self, "_pe_blk_tag", None) != tag:
            self._pe_blk, self._pe_blk_tag = ops.blocked_pe(pe, L), tag
        return self._pe_blk

    def forward(self, x, mask, pool: bool = False, blocked_in=None, cqa_workspace=None):
        """x (N, L, D), mask (N, L) -> (N, L, D); ``pool=True`` returns max_l mask_logits(out, mask) (N, D) instead
        (the fused tail of STAGE.classfier_head_multi_proposal).  ``blocked_in=(N, L)``: x is the flat blocked tile buffer
        ``ops.cqa_fused_fwd(blocked=True)`` produced (internal fast path of STAGE.forward).  ``cqa_workspace``: write the
        output as the sub-title operand images of that ``ops.CqaWorkspace`` and return None (internal fast path, too)."""
        _no_train_dropout(self, self.dropout)
        convs = [c.conv_args(ln) for c, ln in zip(self.conv, self.layer_norm)]
        fin = (self.final_layer_norm.weight, self.final_layer_norm.bias)
        pe = self.position_encoding.pe if blocked_in is None else self._blocked_pe(blocked_in[1])
        if self.num_heads == 0:
            return ops.encoder_block_fwd(x, mask, pe, convs, fin, self.kernel_size, out_mode=1 if pool else 0,
                                         blocked_in=blocked_in, cqa_workspace=cqa_workspace)
        if cqa_workspace is not None:
            raise RuntimeError("tvqaplus_b200: the CQA-workspace output needs an attention-free encoder block")
        # convs (no closing LN) -> LN + MHA + residual -> closing LN (+ pool)
        y = ops.encoder_block_fwd(x, mask, pe, convs, None, self.kernel_size, skip_final_ln=True, blocked_in=blocked_in)
        return ops.mha_fwd(y, mask, self.multi_head_attn.attn_args(self.attn_layer_norm), self.num_heads, residual=True,
                           final_ln=fin, pool=pool)


class StackedEncoder(nn.Module):
    def __init__(self, n_blocks=7, n_conv=2, kernel_size=7, hidden_size=128, dropout=0.1, num_heads=4):
        super().__init__()
        self.n_blocks = n_blocks
        self.stacked_encoderBlocks = nn.ModuleList([
            EncoderBlock(n_conv=n_conv, kernel_size=kernel_size, n_filters=hidden_size, dropout=dropout,
                         num_heads=num_heads) for _ in range(n_blocks)])

    def forward(self, x, mask, pool: bool = False, blocked_in=None, cqa_workspace=None):
        for i, blk in enumerate(self.stacked_encoderBlocks):
            last = i == self.n_blocks - 1
            x = blk(x, mask, pool=pool and last, blocked_in=blocked_in if i == 0 else None,
                    cqa_workspace=cqa_workspace if last else None)
        return x

    def can_write_cqa_workspace(self, L: int) -> bool:
        """The last block is attention-free and L sits in the range ``ops.CqaWorkspace`` takes directly."""
        return self.n_blocks > 0 and self.stacked_encoderBlocks[-1].num_heads == 0 and ops.CqaWorkspace.takes_sub(L)
