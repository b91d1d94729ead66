"""Drop-in replacements for the modules under the reference's ``model/`` directory.

Same class names, constructor arguments, ``forward`` signatures and ``state_dict`` keys as jayleicn/TVQAplus
(model/stage.py, encoder.py, cnn.py, self_attention.py, position_encoding.py, context_query_attention.py,
model_utils.py); every ``forward`` runs hand-written sm_100a kernels through ``libstage_b200.so``.
"""
from .cnn import DepthwiseSeparableConv
from .context_query_attention import StructuredAttention
from .encoder import EncoderBlock, StackedEncoder
from .model_utils import (compute_temporal_iou, expand_span, find_max_triples, flat_list_of_lists, get_high_iou_sapns,
                          mask_logits)
from .position_encoding import PositionEncoding
from .self_attention import MultiHeadedAttention
from .stage import STAGE, ConvLinear, LinearWrapper

__all__ = ["STAGE", "LinearWrapper", "ConvLinear", "StackedEncoder", "EncoderBlock", "DepthwiseSeparableConv",
           "MultiHeadedAttention", "StructuredAttention", "PositionEncoding", "mask_logits", "find_max_triples",
           "expand_span", "get_high_iou_sapns", "compute_temporal_iou", "flat_list_of_lists"]
