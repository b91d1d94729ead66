"""Sinusoidal position table (reference model/position_encoding.py:6-43); same ``pe`` buffer name and shape."""
import math

import torch
import torch.nn as nn


def build_pe_table(max_len: int, n_filters: int) -> torch.Tensor:
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, n_filters, 2, dtype=torch.float32) * -(math.log(10000.0) / n_filters))
    table = torch.zeros(max_len, n_filters)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table


class PositionEncoding(nn.Module):
    """Adds ``pe[:L]`` along axis -2.  Inside :class:`EncoderBlock` the add is fused into the encoder kernel; the
    standalone ``forward`` is a plain broadcast add on the tensor's own device."""

    def __init__(self, n_filters=128, max_len=500):
        super().__init__()
        self.register_buffer("pe", build_pe_table(max_len, n_filters))

    def forward(self, x):
        return x + self.pe[: x.size(-2), :]
