"""ctypes binding of ``libstage_b200.so`` (C ABI declared in ``include/stage_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstage_b200.so")

MAX_CONV = 4
ENC_NO_PE, ENC_NO_RESIDUAL, ENC_NO_FINAL_LN, ENC_NO_RELU, ENC_IN_BLOCKED = 1, 2, 4, 8, 16

_f = C.c_void_p  # device pointers travel as integers


class LinearArgs(C.Structure):
    _fields_ = [("m", C.c_int64), ("k", C.c_int), ("n", C.c_int), ("kc", C.c_int), ("l2norm_in", C.c_int),
                ("relu", C.c_int), ("x", _f), ("w_packed", _f), ("bias", _f), ("ln_in_cs", _f), ("ln_in_cb", _f),
                ("ln_out_g", _f), ("ln_out_b", _f), ("y", _f), ("timeline", _f), ("ldy", C.c_int64), ("x_rowmax", _f)]


class EncoderArgs(C.Structure):
    _fields_ = [("n_seq", C.c_int64), ("L", C.c_int), ("n_conv", C.c_int), ("ksize", C.c_int),
                ("out_mode", C.c_int), ("flags", C.c_int), ("x", _f), ("mask", _f), ("pe", _f),
                ("ln_g", _f * MAX_CONV), ("ln_b", _f * MAX_CONV), ("dw_w", _f * MAX_CONV), ("dw_b", _f * MAX_CONV),
                ("pw_packed", _f * MAX_CONV), ("pw_b", _f * MAX_CONV),
                ("final_ln_g", _f), ("final_ln_b", _f), ("y", _f), ("timeline", _f)]


class MhaArgs(C.Structure):
    _fields_ = [("n_seq", C.c_int64), ("L", C.c_int), ("n_heads", C.c_int), ("residual", C.c_int), ("out_mode", C.c_int),
                ("x", _f), ("mask", _f), ("ln_in_g", _f), ("ln_in_b", _f), ("w_packed", _f * 4), ("b", _f * 4),
                ("final_ln_g", _f), ("final_ln_b", _f), ("y", _f)]


class CqaArgs(C.Structure):
    _fields_ = [("n", C.c_int), ("t", C.c_int), ("lq", C.c_int), ("ls", C.c_int), ("r", C.c_int), ("scale", C.c_float),
                ("a_embed", _f), ("a_mask", _f), ("sub_embed", _f), ("sub_mask", _f), ("vid_embed", _f), ("vid_mask", _f),
                ("wblob", _f), ("cat_out_ln_g", _f), ("cat_out_ln_b", _f), ("workspace", _f),
                ("statement_blocked", C.c_int), ("statement", _f), ("statement_mask", _f), ("vid_raw_s", _f), ("sub_raw_s", _f),
                ("vid_norm_s", _f), ("sub_norm_s", _f), ("att_vid", _f), ("att_sub", _f),
                ("mixed_vid", _f), ("mixed_sub", _f), ("timeline", _f), ("sub_prepared", C.c_int)]


class HeadArgs(C.Structure):
    _fields_ = [("n", C.c_int), ("t", C.c_int), ("add_local", C.c_int), ("extra_span_length", C.c_int),
                ("max_statement", _f), ("statement_mask", _f), ("lq", C.c_int), ("ts_mask", _f),
                ("proj_ln_g", _f), ("proj_ln_b", _f), ("proj_w", _f), ("proj_b", _f),
                ("st_ln_g", _f), ("st_ln_b", _f), ("st_w", _f), ("st_b", _f),
                ("ed_ln_g", _f), ("ed_ln_b", _f), ("ed_w", _f), ("ed_b", _f),
                ("cls_ln_g", _f), ("cls_ln_b", _f), ("cls_w", _f), ("cls_b", _f),
                ("t_scores", _f), ("t_probs", _f), ("spans", _f), ("span_conf", _f), ("answer", _f), ("workspace", _f),
                ("target", _f), ("ts_st", _f), ("ts_ed", _f), ("ts_terms", _f), ("ts_loss", _f), ("frame_mask", _f), ("gfeat", _f)]


class ProposalArgs(C.Structure):
    _fields_ = [("n", C.c_int), ("t", C.c_int), ("extra_span_length", C.c_int), ("ce_prob_thd", C.c_float),
                ("encoded", _f), ("frame_mask", _f), ("gfeat", _f), ("spans", _f), ("span_conf", _f),
                ("target", _f), ("ts_st", _f), ("ts_ed", _f),
                ("cls_ln_g", _f), ("cls_ln_b", _f), ("cls_w", _f), ("cls_b", _f),
                ("answer_rows", _f), ("row_valid", _f), ("row_target", _f)]


class GemmArgs(C.Structure):
    _fields_ = [("a", _f), ("b", _f), ("c", _f), ("bias", _f),
                ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("nb0", C.c_int32), ("nb1", C.c_int32),
                ("a_sm", C.c_int64), ("a_sk", C.c_int64), ("b_sk", C.c_int64), ("b_sn", C.c_int64), ("ldc", C.c_int64),
                ("a_b0", C.c_int64), ("a_b1", C.c_int64), ("b_b0", C.c_int64), ("b_b1", C.c_int64), ("c_b0", C.c_int64),
                ("c_b1", C.c_int64), ("split_k", C.c_int32), ("mode", C.c_int32), ("relu", C.c_int32), ("alpha", C.c_float)]


class ProbeArgs(C.Structure):
    _fields_ = [("a_img", _f), ("b_img", _f), ("a_bytes", C.c_uint32), ("b_bytes", C.c_uint32),
                ("a_lbo", C.c_uint32), ("a_sbo", C.c_uint32), ("a_step", C.c_uint32),
                ("b_lbo", C.c_uint32), ("b_sbo", C.c_uint32), ("b_step", C.c_uint32),
                ("a_mn", C.c_uint32), ("b_mn", C.c_uint32),
                ("k_steps", C.c_uint32), ("n", C.c_uint32), ("use_bulk", C.c_uint32), ("d_out", _f),
                ("repeat", C.c_uint32), ("cycles_out", _f)]


# name -> (restype, argtypes); kept in one table so tests can check the .so exports everything the header declares
SYMBOLS = {
    "stage_abi_version": (C.c_int, []),
    "stage_last_error": (C.c_char_p, []),
    "stage_check_device_error": (C.c_int, [C.c_void_p]),
    "stage_post_device_error": (C.c_int, [C.c_void_p, C.c_void_p]),
    "stage_packed_weight_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "stage_pack_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "stage_pack_weight_ln": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7),
    "stage_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "stage_blocked_stride": (C.c_int, [C.c_int]),
    "stage_encoder_block_fwd": (C.c_int, [C.POINTER(EncoderArgs), C.c_void_p]),
    "stage_mha_fwd": (C.c_int, [C.POINTER(MhaArgs), C.c_void_p]),
    "stage_cqa_weight_blob_bytes": (C.c_size_t, []),
    "stage_cqa_prepare_weights": (C.c_int, [C.c_void_p] * 10),
    "stage_cqa_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "stage_cqa_ctx_blobs": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "stage_cqa_ctx_blob_bytes": (C.c_size_t, []),
    "stage_cqa_fused_fwd": (C.c_int, [C.POINTER(CqaArgs), C.c_void_p]),
    "stage_span_head_fwd": (C.c_int, [C.POINTER(HeadArgs), C.c_void_p]),
    "stage_train_proposals_fwd": (C.c_int, [C.POINTER(ProposalArgs), C.c_void_p]),
    "stage_find_max_span": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stage_find_max_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "stage_att_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "stage_pad_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "stage_mask_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "stage_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "stage_ln_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "stage_ln_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "stage_row_absmax": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "stage_colsum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "stage_relu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stage_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "stage_scale": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_void_p]),
    "stage_dropout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p]),
    "stage_permute_it": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "stage_dwconv_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "stage_dwconv_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "stage_l2norm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "stage_l2norm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "stage_attn_softmax_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "stage_attn_softmax_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p]),
    "stage_cat3_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "stage_cat3_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "stage_maxpool_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "stage_maxpool_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "stage_ts_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stage_proposal_pool_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stage_proposal_pool_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "stage_att_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stage_answer_ce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stage_grad_sumsq": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "stage_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "stage_counter_inc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "stage_umma_probe": (C.c_int, [C.POINTER(ProbeArgs), C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "tvqaplus_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the STAGE kernels)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.stage_abi_version() != 4:
            raise RuntimeError("tvqaplus_b200: ABI version mismatch in %s" % LIB_PATH)
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().stage_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, (msg or b"").decode()))


def ptr(t) -> Optional[int]:
    """Device pointer of a CUDA fp32/int tensor (None stays NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("tvqaplus_b200: expected a CUDA tensor (no CPU path exists)")
    if not t.is_contiguous():
        raise RuntimeError("tvqaplus_b200: expected a contiguous tensor")
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def check_device_error() -> None:
    check(lib().stage_check_device_error(current_stream()), "stage_check_device_error")


class on_device:
    """Context manager: make the device of ``t`` current for the launches inside (the library launches on the current
    device's stream; ``model.to('cuda:1')`` without ``torch.cuda.set_device`` must still work, as it does with PyTorch ops).
    Tensors on different devices raise."""

    def __init__(self, *tensors):
        import torch
        devs = {t.device for t in tensors if torch.is_tensor(t) and t.is_cuda}
        if len(devs) > 1:
            raise RuntimeError("tvqaplus_b200: tensors of one call live on different devices: %s" % sorted(map(str, devs)))
        cpu = [t for t in tensors if torch.is_tensor(t) and not t.is_cuda]
        if cpu:
            raise RuntimeError("tvqaplus_b200: expected a CUDA tensor (no CPU path exists)")
        self.guard = torch.cuda.device(next(iter(devs))) if devs else None

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


class DeviceErrorMonitor:
    """Non-blocking watch on the device error flag (a bounded mbarrier wait that timed out sets it and the kernel carries
    on with garbage).  ``post()`` enqueues a copy of the flag into pinned host memory behind the work already on the
    stream; ``poll()`` raises as soon as a posted copy has landed non-zero — without ever synchronising.  STAGE.forward
    posts at the end of every call and polls at the start of the next one and after every host read it does anyway, so a
    timed-out wait surfaces within one step instead of silently returning wrong answers."""

    def __init__(self, device):
        import torch
        self.device = torch.device(device)
        self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    def post(self) -> None:
        import torch
        with torch.cuda.device(self.device):
            check(lib().stage_post_device_error(self.flag.data_ptr(), current_stream()), "stage_post_device_error")
            self.event = torch.cuda.Event()
            self.event.record()

    def poll(self, wait: bool = False) -> None:
        if self.event is None:
            return
        if wait:
            self.event.synchronize()
        if self.event.query():
            v = int(self.flag.item()) & 0xFFFFFFFF
            if v != 0:
                raise RuntimeError("tvqaplus_b200: device error flag 0x%x — a bounded mbarrier wait timed out in an earlier "
                                   "kernel; the outputs of that forward are invalid" % v)
