"""Thin Python wrappers over the C ABI (one function per entry point of ``include/stage_b200.h``).

All tensors must be contiguous fp32 CUDA tensors; outputs are allocated here with ``torch.empty`` (the library
itself never allocates).  There is no CPU path: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, current_stream, lib, ptr

HSZ = 128
LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _tensors(objs):
    for o in objs:
        if torch.is_tensor(o):
            yield o
        elif isinstance(o, dict):
            yield from _tensors(o.values())
        elif isinstance(o, (list, tuple)):
            yield from _tensors(o)


def _guarded(fn):
    """Run ``fn`` with the device of its tensor arguments current (see ``_lib.on_device``)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        with _lib.on_device(*_tensors(args), *_tensors(kw.values())):
            return fn(*args, **kw)
    return wrapper


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("tvqaplus_b200: expected float32, got %s" % t.dtype)
    return t.contiguous()


def kc_for(n: int) -> int:
    """K-chunk the GEMM kernels stream for an output width n."""
    return 64 if n <= 128 else 32


@_guarded
def pack_weight(w: torch.Tensor, kc: Optional[int] = None) -> torch.Tensor:
    """Split W[N,K] (or a 1x1 conv weight [N,K,1]) into the fp16 hi/lo operand image (stage_pack_weight)."""
    w = _f32(w.detach())
    if w.dim() == 3:
        w = w.reshape(w.shape[0], w.shape[1])
    n, k = w.shape
    kc = kc or kc_for(n)
    nbytes = lib().stage_packed_weight_bytes(n, k, kc)
    if nbytes == 0:
        raise RuntimeError("tvqaplus_b200: unsupported weight shape %s" % (tuple(w.shape),))
    out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib().stage_pack_weight(ptr(w), n, k, kc, ptr(out), current_stream()), "stage_pack_weight")
    _count()
    out._stage_kc = kc  # type: ignore[attr-defined]
    return out


@_guarded
def pack_weight_ln(w: torch.Tensor, bias: torch.Tensor, ln_g: torch.Tensor, ln_b: torch.Tensor, kc: Optional[int] = None):
    """Weight of a Linear that follows a LayerNorm: the image of W*gamma plus the two epilogue vectors cs / cb
    (stage_pack_weight_ln).  Returns (image, cs, cb)."""
    w, bias, ln_g, ln_b = _f32(w.detach()), _f32(bias.detach()), _f32(ln_g.detach()), _f32(ln_b.detach())
    n, k = w.shape
    kc = kc or kc_for(n)
    nbytes = lib().stage_packed_weight_bytes(n, k, kc)
    if nbytes == 0:
        raise RuntimeError("tvqaplus_b200: unsupported weight shape %s" % (tuple(w.shape),))
    img = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    cs = torch.empty(n, dtype=torch.float32, device=w.device)
    cb = torch.empty(n, dtype=torch.float32, device=w.device)
    check(lib().stage_pack_weight_ln(ptr(w), n, k, kc, ptr(ln_g), ptr(ln_b), ptr(bias), ptr(img), ptr(cs), ptr(cb),
                                     current_stream()), "stage_pack_weight_ln")
    _count(2)
    return img, cs, cb


@_guarded
def linear_fwd(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], n: int, *, kc: Optional[int] = None,
               ln_in=None, ln_out=None, l2norm_in: bool = False, relu: bool = False,
               timeline: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, ldy: int = 0, y_off: int = 0,
               x_rowmax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = LN_out(relu?(LN_in(l2norm?(x)) W^T + b)) over the last axis of x (stage_linear_fwd).

    ``ln_in`` is the (cs, cb) pair ``pack_weight_ln`` returned together with ``w_packed`` (the LayerNorm is folded into
    the weight); ``ln_out`` is a plain (gamma, beta) pair."""
    x = _f32(x)
    k = x.shape[-1]
    m = x.numel() // k
    y = torch.empty(x.shape[:-1] + (n,), dtype=torch.float32, device=x.device) if out is None else out
    a = _lib.LinearArgs(m=m, k=k, n=n, kc=kc or kc_for(n), l2norm_in=int(l2norm_in), relu=int(relu),
                        x=ptr(x), w_packed=ptr(w_packed), bias=ptr(_f32(bias)) if bias is not None else None,
                        ln_in_cs=ptr(_f32(ln_in[0])) if ln_in else None, ln_in_cb=ptr(_f32(ln_in[1])) if ln_in else None,
                        ln_out_g=ptr(_f32(ln_out[0])) if ln_out else None,
                        ln_out_b=ptr(_f32(ln_out[1])) if ln_out else None, y=ptr(y) + 4 * y_off, timeline=ptr(timeline),
                        ldy=ldy, x_rowmax=ptr(x_rowmax))
    check(lib().stage_linear_fwd(C.byref(a), current_stream()), "stage_linear_fwd")
    _count()
    return y


def _arr(vals: Sequence[Optional[int]], n: int = _lib.MAX_CONV):
    a = (C.c_void_p * n)()
    for i, v in enumerate(vals):
        a[i] = v
    return a


@_guarded
def encoder_block_fwd(x: torch.Tensor, mask: Optional[torch.Tensor], pe: Optional[torch.Tensor], convs: Sequence[dict],
                      final_ln, ksize: int, *, out_mode: int = 0, timeline: Optional[torch.Tensor] = None,
                      blocked_in: Optional[tuple] = None,
                      conv_only: bool = False, pre_ln_only: bool = False, relu: bool = True,
                      skip_final_ln: bool = False, cqa_workspace: Optional["CqaWorkspace"] = None) -> torch.Tensor:
    """One EncoderBlock (stage_encoder_block_fwd).  x (n_seq, L, 128); convs[i] = dict(ln_g, ln_b, dw_w, dw_b,
    pw_packed, pw_b); final_ln = (g, b).  out_mode 1 returns the masked max-pool (n_seq, 128).

    ``conv_only``: just the depthwise-separable convs (no PE / LayerNorm / residual) — DepthwiseSeparableConv;
    ``pre_ln_only``: LayerNorm -> conv without PE / residual / final LN — ConvLinear."""
    x = _f32(x)
    if blocked_in is not None:  # x is the flat blocked tile buffer of cqa_fused_fwd(blocked=True); pe is blocked_pe(pe, L)
        n_seq, L = blocked_in
        D = HSZ
        if x.numel() != n_seq * 128 * blocked_stride(L):
            raise RuntimeError("tvqaplus_b200: blocked buffer size does not match (n_seq=%d, L=%d)" % (n_seq, L))
    else:
        n_seq, L, D = x.shape
    if D != HSZ:
        raise RuntimeError("tvqaplus_b200: hidden size %d != %d" % (D, HSZ))
    flags = _lib.ENC_IN_BLOCKED if blocked_in is not None else 0
    if conv_only or pre_ln_only:
        flags |= _lib.ENC_NO_PE | _lib.ENC_NO_RESIDUAL | _lib.ENC_NO_FINAL_LN
    if not relu:
        flags |= _lib.ENC_NO_RELU
    if skip_final_ln:
        flags |= _lib.ENC_NO_FINAL_LN
    if not (flags & _lib.ENC_NO_PE) and blocked_in is None and pe.shape[0] < L:
        raise RuntimeError("tvqaplus_b200: sequence length %d exceeds the position-encoding table (%d)" % (L, pe.shape[0]))
    if cqa_workspace is not None:
        # out_mode 2: the closing LayerNorm output is written as the sub-title operand image (+ inverse norms, masks) of the
        # CQA context blobs inside ``cqa_workspace``; nothing else is returned
        if n_seq != cqa_workspace.n * cqa_workspace.t or not cqa_workspace.takes_sub(L):
            raise RuntimeError("tvqaplus_b200: encoder output (n_seq=%d, L=%d) does not match the CQA workspace" % (n_seq, L))
        out_mode, y = 2, None
    else:
        y = torch.empty((n_seq, L, D) if out_mode == 0 else (n_seq, D), dtype=torch.float32, device=x.device)
    names = ("ln_g", "ln_b", "dw_w", "dw_b", "pw_b")
    keep = [[_f32(c[k]) if c.get(k) is not None else None for k in names] for c in convs]  # keep contiguous copies alive
    cols = {k: [ptr(row[i]) for row in keep] for i, k in enumerate(names)}
    mask_c = _f32(mask) if mask is not None else None
    pe_c = _f32(pe) if pe is not None else None
    fin = (_f32(final_ln[0]), _f32(final_ln[1])) if final_ln is not None else (None, None)
    a = _lib.EncoderArgs(n_seq=n_seq, L=L, n_conv=len(convs), ksize=ksize, out_mode=out_mode,
                         flags=flags, x=ptr(x), mask=ptr(mask_c), pe=ptr(pe_c),
                         ln_g=_arr(cols["ln_g"]), ln_b=_arr(cols["ln_b"]), dw_w=_arr(cols["dw_w"]),
                         dw_b=_arr(cols["dw_b"]), pw_packed=_arr([ptr(c["pw_packed"]) for c in convs]),
                         pw_b=_arr(cols["pw_b"]), final_ln_g=ptr(fin[0]), final_ln_b=ptr(fin[1]),
                         y=ptr(y) if cqa_workspace is None else cqa_workspace.ctx_blobs, timeline=ptr(timeline))
    check(lib().stage_encoder_block_fwd(C.byref(a), current_stream()), "stage_encoder_block_fwd")
    _count()
    if cqa_workspace is not None:
        cqa_workspace.sub_prepared = True
        cqa_workspace.sub_len = L
        return None
    return y


class CqaWorkspace:
    """Caller-owned scratch of one ``cqa_fused_fwd`` call (stage_cqa_workspace_bytes).  Allocated BEFORE the sub-title
    input encoder runs, that encoder can write its output straight into the context blobs (``encoder_block_fwd(...,
    cqa_workspace=ws)``) and the fused kernel's preparation pass then skips the sub-title stream."""

    def __init__(self, n: int, t: int, lq: int, device):
        self.n, self.t, self.lq = int(n), int(t), int(lq)
        nbytes = lib().stage_cqa_workspace_bytes(self.n, self.t, self.lq)
        self.buf = torch.empty(nbytes + 128, dtype=torch.uint8, device=device)
        self.ptr = (self.buf.data_ptr() + 127) // 128 * 128
        self.ctx_blobs = lib().stage_cqa_ctx_blobs(self.ptr, self.n, self.t, self.lq)
        self.sub_prepared = False
        self.sub_len = 0

    @staticmethod
    def takes_sub(L: int) -> bool:
        """Sequence lengths the encoder can write directly (one sequence per tile, within the CQA sub-title tile)."""
        return 64 < L <= 112


@_guarded
def cqa_prepare_weights(c2q_ln_g, c2q_ln_b, c2q_w, c2q_b, cat_ln_g, cat_ln_b, cat_w, cat_b) -> torch.Tensor:
    """Fold LayerNorm(384) into c2q_down_projection / concat_fc and build their split-fp16 images
    (stage_cqa_prepare_weights).  Re-run whenever one of the eight parameters changes."""
    ts = [_f32(t.detach()) for t in (c2q_ln_g, c2q_ln_b, c2q_w, c2q_b, cat_ln_g, cat_ln_b, cat_w, cat_b)]
    blob = torch.empty(lib().stage_cqa_weight_blob_bytes(), dtype=torch.uint8, device=ts[0].device)
    check(lib().stage_cqa_prepare_weights(*[ptr(t) for t in ts], ptr(blob), current_stream()),
          "stage_cqa_prepare_weights")
    _count(6)
    return blob


def blocked_stride(L: int) -> int:
    """Row stride S of the blocked tile layout (stage_blocked_stride)."""
    return int(lib().stage_blocked_stride(int(L)))


def blocked_fits(L: int) -> bool:
    """Can the encoder kernel take sequences of length L in the blocked layout?  (Any tile-sized sequence: the kernel reads
    the blocked rows straight from global memory, one row per lane.)"""
    return 0 < L <= 128


def blocked_pe(pe: torch.Tensor, L: int) -> torch.Tensor:
    """Position table rows [0, L) in the blocked form the encoder reads with STAGE_ENC_IN_BLOCKED: [32][L] float4."""
    return _f32(pe[:L]).view(L, 32, 4).permute(1, 0, 2).contiguous()


@_guarded
def cqa_fused_fwd(a_embed, a_mask, sub_embed, sub_mask, vid_embed, vid_mask, w: dict, scale: float, *,
                  want: Sequence[str] = (), timeline: Optional[torch.Tensor] = None, blocked: bool = False,
                  workspace: Optional[CqaWorkspace] = None) -> dict:
    """Fused StructuredAttention x2 + c2q_down_projection x2 + concat_fc (stage_cqa_fused_fwd).

    a_embed (N,5,Lq,128), a_mask (N,5,Lq), sub_embed (N,T,Ls,128), sub_mask (N,T,Ls), vid_embed (N,T,R,128),
    vid_mask (N,T,R).  ``w`` holds ``wblob`` (cqa_prepare_weights) and cat_out_ln_g/b.  ``want`` lists optional extra
    outputs: vid_raw_s, sub_raw_s, vid_norm_s, sub_norm_s, att_vid, att_sub, mixed_vid, mixed_sub.
    Returns dict with ``statement`` (N,5,T,Lq,128), ``statement_mask`` (N,5,T,Lq).
    """
    a_embed, a_mask = _f32(a_embed), _f32(a_mask)
    sub_mask, vid_embed, vid_mask = _f32(sub_mask), _f32(vid_embed), _f32(vid_mask)
    N, A, Lq, D = a_embed.shape
    prepared = workspace is not None and workspace.sub_prepared  # ``sub_embed`` may be None: its images are in the workspace
    sub_embed = None if prepared else _f32(sub_embed)
    T, Ls = sub_mask.shape[1:3]
    R = vid_embed.shape[2]
    if A != 5 or D != HSZ:
        raise RuntimeError("tvqaplus_b200: a_embed must be (N,5,Lq,128)")
    dev = a_embed.device
    # ``blocked``: the statement leaves in the blocked tile layout the cls encoder consumes directly (a flat buffer of
    # N*5*T tiles of 32 x S float4); the padding rows [Lq, S) of a tile are never read
    out = {"statement": torch.empty(N * 5 * T * 128 * blocked_stride(Lq), dtype=torch.float32, device=dev) if blocked
           else torch.empty(N, 5, T, Lq, D, dtype=torch.float32, device=dev),
           "statement_mask": torch.empty(N, 5, T, Lq, dtype=torch.float32, device=dev)}
    shapes = {"vid_raw_s": (N, 5, T, Lq, R), "sub_raw_s": (N, 5, T, Lq, Ls), "vid_norm_s": (N, 5, T, Lq, R),
              "sub_norm_s": (N, 5, T, Lq, Ls), "att_vid": (N, 5, T, Lq, D), "att_sub": (N, 5, T, Lq, D),
              "mixed_vid": (N, 5, T, Lq, D), "mixed_sub": (N, 5, T, Lq, D)}
    for k in want:
        out[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
    ws = workspace if workspace is not None else CqaWorkspace(N, T, Lq, dev)
    if (ws.n, ws.t, ws.lq) != (N, T, Lq) or (prepared and ws.sub_len != Ls):
        raise RuntimeError("tvqaplus_b200: CQA workspace was sized for another problem")
    ws_ptr = ws.ptr
    a = _lib.CqaArgs(n=N, t=T, lq=Lq, ls=Ls, r=R, scale=float(scale),
                     a_embed=ptr(a_embed), a_mask=ptr(a_mask), sub_embed=ptr(sub_embed), sub_mask=ptr(sub_mask),
                     vid_embed=ptr(vid_embed), vid_mask=ptr(vid_mask), wblob=ptr(w["wblob"]),
                     cat_out_ln_g=ptr(_f32(w["cat_out_ln_g"])), cat_out_ln_b=ptr(_f32(w["cat_out_ln_b"])),
                     workspace=ws_ptr, statement_blocked=int(blocked), statement=ptr(out["statement"]),
                     statement_mask=ptr(out["statement_mask"]),
                     timeline=ptr(timeline), sub_prepared=int(prepared), **{k: ptr(out.get(k)) for k in shapes})
    check(lib().stage_cqa_fused_fwd(C.byref(a), current_stream()), "stage_cqa_fused_fwd")
    _count(3)
    return out


@_guarded
def span_head_fwd(max_statement, statement_mask, ts_mask, w: dict, *, add_local: bool, extra_span_length: int,
                  labels=None, want_train: bool = False) -> dict:
    """Span + answer heads, eval branch (stage_span_head_fwd).

    max_statement (N*5, T, 128), statement_mask (N,5,T,Lq), ts_mask (N,T).  ``w``: proj_/st_/ed_/cls_ + ln_g, ln_b, w, b.
    Returns t_scores (N,5,T,2) masked raw logits, t_probs, spans (N*5,2) int32 inclusive, span_conf, answer (N,5),
    encoded (N,5,T,128).  ``labels = (target, ts_st, ts_ed)`` (int64 (N,) each) adds ``ts_loss`` (get_ts_loss fused into the
    launch); ``want_train`` adds ``frame_mask`` (N*5,T) and ``gfeat`` (N*5,128) for ``train_proposals_fwd``."""
    max_statement, statement_mask, ts_mask = _f32(max_statement), _f32(statement_mask), _f32(ts_mask)
    N, _, T, Lq = statement_mask.shape
    dev = max_statement.device
    o = dict(t_scores=torch.empty(N, 5, T, 2, dtype=torch.float32, device=dev),
             t_probs=torch.empty(N, 5, T, 2, dtype=torch.float32, device=dev),
             spans=torch.empty(N * 5, 2, dtype=torch.int32, device=dev),
             span_conf=torch.empty(N * 5, dtype=torch.float32, device=dev),
             answer=torch.empty(N, 5, dtype=torch.float32, device=dev),
             encoded=torch.empty(N, 5, T, HSZ, dtype=torch.float32, device=dev))
    extra = {}
    keep_lab = None
    if labels is not None:
        keep_lab = [t.to(device=dev, dtype=torch.int64).contiguous() for t in labels]
        o["ts_terms"] = torch.empty(N, 2, dtype=torch.float32, device=dev)
        o["ts_loss"] = torch.empty((), dtype=torch.float32, device=dev)
        extra.update(target=ptr(keep_lab[0]), ts_st=ptr(keep_lab[1]), ts_ed=ptr(keep_lab[2]), ts_terms=ptr(o["ts_terms"]),
                     ts_loss=ptr(o["ts_loss"]))
    if want_train:
        o["frame_mask"] = torch.empty(N * 5, T, dtype=torch.float32, device=dev)
        o["gfeat"] = torch.empty(N * 5, HSZ, dtype=torch.float32, device=dev)
        extra.update(frame_mask=ptr(o["frame_mask"]), gfeat=ptr(o["gfeat"]))
    keep = {k: _f32(v) for k, v in w.items()}
    a = _lib.HeadArgs(n=N, t=T, add_local=int(add_local), extra_span_length=int(extra_span_length),
                      max_statement=ptr(max_statement), statement_mask=ptr(statement_mask), lq=Lq, ts_mask=ptr(ts_mask),
                      **{"%s_%s" % (grp, f): ptr(keep["%s_%s" % (grp, f)])
                         for grp in ("proj", "st", "ed", "cls") for f in ("ln_g", "ln_b", "w", "b")},
                      t_scores=ptr(o["t_scores"]), t_probs=ptr(o["t_probs"]), spans=ptr(o["spans"]),
                      span_conf=ptr(o["span_conf"]), answer=ptr(o["answer"]), workspace=ptr(o["encoded"]), **extra)
    check(lib().stage_span_head_fwd(C.byref(a), current_stream()), "stage_span_head_fwd")
    _count(2 if labels is not None else 1)
    return o


@_guarded
def train_proposals_fwd(head: dict, labels, w: dict, *, extra_span_length: int, ce_prob_thd: float = 0.01) -> dict:
    """Train branch of get_proposals + classifier (stage_train_proposals_fwd) on the outputs of
    ``span_head_fwd(..., labels=labels, want_train=True)``.  Returns answer_rows (2N,5), row_valid (2N,) int32, row_target
    (2N,) int64: candidate row 2i = example i's ground-truth span, 2i+1 = its predicted span if it passed the IoU test."""
    enc = head["encoded"]
    N, _, T, _ = enc.shape
    dev = enc.device
    lab = [t.to(device=dev, dtype=torch.int64).contiguous() for t in labels]
    keep = {k: _f32(w[k]) for k in ("cls_ln_g", "cls_ln_b", "cls_w", "cls_b")}
    if keep["cls_w"].numel() != 2 * HSZ:
        raise RuntimeError("tvqaplus_b200: the train-mode proposal pass needs the add_local classifier (256 -> 1)")
    o = dict(answer_rows=torch.empty(2 * N, 5, dtype=torch.float32, device=dev),
             row_valid=torch.empty(2 * N, dtype=torch.int32, device=dev),
             row_target=torch.empty(2 * N, dtype=torch.int64, device=dev))
    a = _lib.ProposalArgs(n=N, t=T, extra_span_length=int(extra_span_length), ce_prob_thd=float(ce_prob_thd),
                          encoded=ptr(enc), frame_mask=ptr(head["frame_mask"]), gfeat=ptr(head["gfeat"]), spans=ptr(head["spans"]),
                          span_conf=ptr(head["span_conf"]), target=ptr(lab[0]), ts_st=ptr(lab[1]), ts_ed=ptr(lab[2]),
                          cls_ln_g=ptr(keep["cls_ln_g"]), cls_ln_b=ptr(keep["cls_ln_b"]), cls_w=ptr(keep["cls_w"]), cls_b=ptr(keep["cls_b"]),
                          answer_rows=ptr(o["answer_rows"]), row_valid=ptr(o["row_valid"]), row_target=ptr(o["row_target"]))
    check(lib().stage_train_proposals_fwd(C.byref(a), current_stream()), "stage_train_proposals_fwd")
    _count()
    return o


@_guarded
def find_max_span(p_st: torch.Tensor, p_ed: torch.Tensor):
    """Best (st <= ed) pair maximising p_st[st]*p_ed[ed] per row (stage_find_max_span)."""
    p_st, p_ed = _f32(p_st), _f32(p_ed)
    rows, T = p_st.shape
    spans = torch.empty(rows, 2, dtype=torch.int32, device=p_st.device)
    conf = torch.empty(rows, dtype=torch.float32, device=p_st.device)
    check(lib().stage_find_max_span(ptr(p_st), ptr(p_ed), rows, T, ptr(spans), ptr(conf), current_stream()),
          "stage_find_max_span")
    _count()
    return spans, conf


@_guarded
def find_max_pairs(probs: torch.Tensor):
    """Batched ``find_max_pair`` of inference.py:13-35 on the device: probs (rows, T, 2) = start / end probabilities
    -> spans (rows, 2) int32 [st, ed], values (rows,) float64 (stage_find_max_pair)."""
    probs = _f32(probs)
    rows, T, two = probs.shape
    if two != 2:
        raise RuntimeError("tvqaplus_b200: probs must be (rows, T, 2)")
    spans = torch.empty(rows, 2, dtype=torch.int32, device=probs.device)
    val = torch.empty(rows, dtype=torch.float64, device=probs.device)
    check(lib().stage_find_max_pair(ptr(probs), ptr(probs) + 4, rows, T, 2 * T, 2, ptr(spans), ptr(val), current_stream()),
          "stage_find_max_pair")
    _count()
    return spans, val


@_guarded
def att_loss_fwd(scores: torch.Tensor, pos_off: torch.Tensor, neg_off: torch.Tensor, *, loss_type: str, alpha: float,
                 margin: float) -> torch.Tensor:
    """Attention ranking loss over sampled (positive, negative) pairs (stage_att_loss_fwd): ``scores`` = vid_raw_s,
    ``pos_off`` / ``neg_off`` int64 flat offsets into it.  Returns a 0-d tensor."""
    scores = _f32(scores)
    if loss_type not in ("lse", "hinge"):
        raise NotImplementedError("Only support hinge and lse")
    m = pos_off.numel()
    dev = scores.device
    pos_off = pos_off.to(device=dev, dtype=torch.int64).contiguous()
    neg_off = neg_off.to(device=dev, dtype=torch.int64).contiguous()
    partials = torch.empty((m + 255) // 256, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    check(lib().stage_att_loss_fwd(ptr(scores), ptr(pos_off), ptr(neg_off), m, 0 if loss_type == "lse" else 1, float(alpha),
                                   float(margin), ptr(partials), ptr(loss), current_stream()), "stage_att_loss_fwd")
    _count(2)
    return loss


@_guarded
def pad_rows(rows: torch.Tensor, off: torch.Tensor, lens: torch.Tensor, l_pad: int):
    """``pad_sequences_2d`` on the device (stage_pad_rows): ``rows`` (R, D) = the valid rows of all sequences concatenated,
    ``off`` (n_seq,) int64 row offsets, ``lens`` (n_seq,) int32 -> (out (n_seq, l_pad, D) zero-padded, mask (n_seq, l_pad))."""
    rows = _f32(rows)
    if off.dtype != torch.int64 or lens.dtype != torch.int32:
        raise RuntimeError("tvqaplus_b200: pad_rows needs int64 offsets and int32 lengths")
    n_seq, D = lens.numel(), rows.shape[1]
    out = torch.empty(n_seq, l_pad, D, dtype=torch.float32, device=rows.device)
    mask = torch.empty(n_seq, l_pad, dtype=torch.float32, device=rows.device)
    check(lib().stage_pad_rows(ptr(rows) if rows.numel() else None, ptr(off.contiguous()), ptr(lens.contiguous()), n_seq, int(l_pad), D,
                               ptr(out), ptr(mask), current_stream()), "stage_pad_rows")
    _count()
    return out, mask


@_guarded
def mask_logits_fwd(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x*m + (1-m)*(-1e10) with m broadcast over the trailing axes of x it lacks (stage_mask_logits)."""
    x = _f32(x)
    mask = _f32(mask)
    while mask.dim() > 0 and mask.shape[-1] == 1 and mask.dim() == x.dim():
        mask = mask.squeeze(-1)
    if tuple(x.shape[:mask.dim()]) != tuple(mask.shape):
        mask = mask.expand(x.shape[:mask.dim()]).contiguous()
    outer = mask.numel()
    inner = x.numel() // max(1, outer)
    y = torch.empty_like(x)
    check(lib().stage_mask_logits(ptr(x), ptr(mask.contiguous()), outer, inner, ptr(y), current_stream()),
          "stage_mask_logits")
    _count()
    return y


_zero_cqa_blobs: dict = {}


@_guarded
def attention_only_fwd(c_embed, c_mask, q_embed, q_mask, scale: float) -> dict:
    """StructuredAttention alone for one stream: C (N,5,Lq,128), Q (N,T,Lr,128) -> att (N,5,T,Lq,128),
    raw_s / norm_s (N,5,T,Lq,Lr).  Runs the fused kernel with the query tensor in its subtitle slot (Lr <= 112), a
    one-item fully-masked dummy in the region slot and all-zero projection weights, and keeps only U = P.Q and the
    score tensors.  Not a hot path (STAGE.forward uses ``cqa_fused_fwd``)."""
    c_embed, q_embed = _f32(c_embed), _f32(q_embed)
    dev = c_embed.device
    N, T = q_embed.shape[0], q_embed.shape[1]
    key = str(dev)
    if key not in _zero_cqa_blobs:
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        _zero_cqa_blobs[key] = dict(wblob=cqa_prepare_weights(z(384), z(384), z(128, 384), z(128), z(384), z(384), z(128, 384), z(128)),
                                    cat_out_ln_g=z(128), cat_out_ln_b=z(128))
    dummy = torch.zeros(N, T, 1, HSZ, dtype=torch.float32, device=dev)
    dmask = torch.zeros(N, T, 1, dtype=torch.float32, device=dev)
    out = cqa_fused_fwd(c_embed, c_mask, q_embed, q_mask, dummy, dmask, _zero_cqa_blobs[key], scale,
                        want=("att_sub", "sub_raw_s", "sub_norm_s"))
    return dict(att=out["att_sub"], raw_s=out["sub_raw_s"], norm_s=out["sub_norm_s"])


@_guarded
def mha_fwd(x, mask, attn: dict, n_heads: int, *, residual: bool = False, final_ln=None, pool: bool = False):
    """Multi-head self-attention (stage_mha_fwd): y = LN_fin?((x +)? MHA(LN_in?(x))); ``attn`` = dict(ln_g, ln_b,
    w_packed[4], b[4]) as ``MultiHeadedAttention.attn_args`` builds it.  ``pool=True`` returns the masked max over L."""
    x = _f32(x)
    n_seq, L, D = x.shape
    if D != HSZ:
        raise RuntimeError("tvqaplus_b200: hidden size %d != %d" % (D, HSZ))
    if pool and mask is None:
        raise RuntimeError("tvqaplus_b200: the masked max-pool needs a mask")
    y = torch.empty((n_seq, D) if pool else (n_seq, L, D), dtype=torch.float32, device=x.device)
    mask_c = _f32(mask) if mask is not None else None
    ln = (_f32(attn["ln_g"]), _f32(attn["ln_b"])) if attn.get("ln_g") is not None else (None, None)
    fin = (_f32(final_ln[0]), _f32(final_ln[1])) if final_ln is not None else (None, None)
    bs = [_f32(b) for b in attn["b"]]
    a = _lib.MhaArgs(n_seq=n_seq, L=L, n_heads=n_heads, residual=int(residual), out_mode=int(pool), x=ptr(x),
                     mask=ptr(mask_c), ln_in_g=ptr(ln[0]), ln_in_b=ptr(ln[1]),
                     w_packed=_arr([ptr(w) for w in attn["w_packed"]], 4), b=_arr([ptr(b) for b in bs], 4),
                     final_ln_g=ptr(fin[0]), final_ln_b=ptr(fin[1]), y=ptr(y))
    check(lib().stage_mha_fwd(C.byref(a), current_stream()), "stage_mha_fwd")
    _count()
    return y
