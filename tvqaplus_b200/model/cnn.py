"""Depthwise-separable 1-D convolution (reference model/cnn.py:6-47): same parameter names/shapes
(``depthwise_conv.weight (C,1,k)``, ``pointwise_conv.weight (Cout,C,1)``), executed by the fused encoder kernel."""
import torch.nn as nn

from .. import ops
from ._packing import packed


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, in_ch, out_ch, k, dim=1, relu=True):
        super().__init__()
        if dim != 1:
            raise NotImplementedError("tvqaplus_b200: only the 1-D form used by STAGE is built (dim=%d)" % dim)
        self.relu = relu
        self.k = k
        self.depthwise_conv = nn.Conv1d(in_channels=in_ch, out_channels=in_ch, kernel_size=k, groups=in_ch,
                                        padding=k // 2)
        self.pointwise_conv = nn.Conv1d(in_channels=in_ch, out_channels=out_ch, kernel_size=1, padding=0)

    def conv_args(self, ln=None) -> dict:
        """Pointers the encoder kernel needs for this conv; ``ln`` = the LayerNorm applied in front (or None)."""
        return dict(ln_g=ln.weight if ln is not None else None, ln_b=ln.bias if ln is not None else None,
                    dw_w=self.depthwise_conv.weight, dw_b=self.depthwise_conv.bias,
                    pw_packed=packed(self.pointwise_conv.weight, 64), pw_b=self.pointwise_conv.bias)

    def forward(self, x):
        """(N, L, D) -> (N, L, D)."""
        if self.pointwise_conv.weight.shape[0] != ops.HSZ or self.pointwise_conv.weight.shape[1] != ops.HSZ:
            raise NotImplementedError("tvqaplus_b200: DepthwiseSeparableConv is built for 128 -> 128 channels")
        return ops.encoder_block_fwd(x, None, None, [self.conv_args()], None, self.k, conv_only=True, relu=self.relu)
