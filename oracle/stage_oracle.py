"""CPU oracle for the STAGE forward hot path (TEST INFRASTRUCTURE — not product code).

This file is a functional, parameter-dict restatement (plain ``torch`` fp32 on the CPU) of the
forward path of ``STAGE`` in jayleicn/TVQAplus.  It exists only so the CUDA path can be checked:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package ``tvqaplus_b200`` never imports anything from ``oracle/``.

Pinning: the reference ships no golden vectors or tests for this path, so the
oracle is pinned by *execution against the live reference*: ``oracle/make_golden.py`` imports the
unmodified reference modules from ``/root/reference/model`` in the build container, runs them on
seeded synthetic batches and commits the outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those fixtures bit-for-bit (fp32, same torch build) and, when the
reference tree is present, against the live reference as well.

Every function cites the reference file:line it restates (paths relative to the reference root).
Weights come in as a ``state_dict``-style mapping that uses the reference's own key names
(e.g. ``input_encoder.stacked_encoderBlocks.0.conv.0.depthwise_conv.weight``).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NEG = -1e10  # model/model_utils.py:14-15 and model/context_query_attention.py:100

# Train-mode dropout (the reference's nn.Dropout / F.dropout sites): torch's random stream cannot be shared with the CUDA path,
# so the oracle takes the masks from the caller - ``DROP(tag, x)`` returns x with the site's mask applied (tests build the
# masks with the product's own generator and check loss + gradients against autograd through this restatement).  None = eval
# / dropout 0.  ``DROP_PREFIX`` names the stream ("q", "s", "v", "c") the shared modules are currently applied to.
DROP = None
DROP_PREFIX = ""


def _drop(tag, x):
    return x if DROP is None else DROP(tag, x)


# ----------------------------------------------------------------------------------------------
# options / synthetic data
# ----------------------------------------------------------------------------------------------
def default_opt(**over) -> SimpleNamespace:
    """The ``opt`` namespace ``run_main.sh --add_local`` produces (config.py:15-134, run_main.sh:34-46)."""
    o = dict(
        sub_flag=True, vfeat_flag=True, vfeat_size=300, embedding_size=768, hsz=128, dropout=0.1,
        t_iter=0, extra_span_length=3, add_local=True, use_sup_att=True, scale=10.0, alpha=20.0,
        margin=0.1, att_loss_type="lse", num_negatives=2, num_hard=2, negative_pool_size=0,
        drop_topk=0, add_non_visual=False,
        input_encoder_n_blocks=1, input_encoder_n_conv=2, input_encoder_kernel_size=7,
        input_encoder_n_heads=0,
        cls_encoder_n_blocks=1, cls_encoder_n_conv=2, cls_encoder_kernel_size=5,
        cls_encoder_n_heads=0,
    )
    o.update(over)
    return SimpleNamespace(**o)


class Batch(dict):
    """Attribute-access dict standing in for easydict.EasyDict (tvqa_dataset.py:631-688)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_synthetic_batch(N: int, T: int, R: int, Lq: int, Ls: int, *, seed: int = 2018,
                         dense: bool = False, wd: int = 768, vd: int = 300,
                         empty_frames: bool = True) -> Batch:
    """Seeded synthetic batch with the schema ``prepare_inputs`` emits (tvqa_dataset.py:592-688).

    Ragged variant: QA length ~U[0.5Lq, Lq]; subtitle length per frame ~U[0.3Ls, Ls]
    with ~5 % near-empty frames (1 token); regions ~U[0.5R, R]; frames per clip ~U[0.7T, T], tail frames
    fully masked in sub/vid/ts masks.  Padding slots are zero-filled like ``pad_sequences_2d``
    (tvqa_dataset.py:546-553).  ``dense=True`` sets every mask to 1 (the throughput shape).
    """
    g = torch.Generator().manual_seed(seed)

    def randint(lo, hi, shape):
        return torch.randint(lo, hi + 1, shape, generator=g)

    qas_bert = torch.randn(N, 5, Lq, wd, generator=g)
    sub_bert = torch.randn(N, T, Ls, wd, generator=g)
    vid = torch.randn(N, T, R, vd, generator=g)
    if dense:
        qas_mask = torch.ones(N, 5, Lq)
        sub_mask = torch.ones(N, T, Ls)
        vid_mask = torch.ones(N, T, R)
        ts_mask = torch.ones(N, T)
        n_frames = torch.full((N,), T, dtype=torch.long)
    else:
        q_len = randint(max(1, Lq // 2), Lq, (N, 5))
        qas_mask = (torch.arange(Lq)[None, None, :] < q_len[..., None]).float()
        n_frames = randint(max(1, int(0.7 * T)), T, (N,))
        frame_ok = (torch.arange(T)[None, :] < n_frames[:, None])
        s_len = randint(max(1, int(0.3 * Ls)), Ls, (N, T))
        if empty_frames:
            near_empty = torch.rand(N, T, generator=g) < 0.05
            s_len = torch.where(near_empty, torch.ones_like(s_len), s_len)
        sub_mask = (torch.arange(Ls)[None, None, :] < s_len[..., None]).float() * frame_ok[..., None].float()
        r_len = randint(max(1, R // 2), R, (N, T))
        vid_mask = (torch.arange(R)[None, None, :] < r_len[..., None]).float() * frame_ok[..., None].float()
        ts_mask = frame_ok.float()
        qas_bert = qas_bert * qas_mask[..., None]
        sub_bert = sub_bert * sub_mask[..., None]
        vid = vid * vid_mask[..., None]
    target = randint(0, 4, (N,))
    st = torch.stack([randint(0, int(n_frames[i]) - 1, (1,))[0] for i in range(N)])
    ed = torch.stack([randint(int(st[i]), int(n_frames[i]) - 1, (1,))[0] for i in range(N)])
    qas = randint(1, 15000, (N, 5, Lq)) * qas_mask.long()
    return Batch(
        qid=list(range(N)), vid_name=["syn_%d" % i for i in range(N)],
        qas_bert=qas_bert.contiguous(), qas_mask=qas_mask.contiguous(),
        sub_bert=sub_bert.contiguous(), sub_mask=sub_mask.contiguous(),
        vid=vid.contiguous(), vid_mask=vid_mask.contiguous(),
        target=target, ts_label={"st": st, "ed": ed}, ts_label_mask=ts_mask.contiguous(), qas=qas,
        att_labels=None, anno_st_idx=[0] * N, q_l=[int(Lq // 2)] * N, image_indices=None, boxes=None,
        use_hard_negatives=False, eval_object_word_ids=None,
    )


# ----------------------------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------------------------
def mask_logits(target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """model/model_utils.py:14-15."""
    return target * mask + (1 - mask) * NEG


def position_encoding_table(max_len: int = 500, n_filters: int = 128) -> torch.Tensor:
    """sin/cos table registered as buffer ``pe`` — model/position_encoding.py:24-31."""
    pe = torch.zeros(max_len, n_filters)
    pos = torch.arange(0, max_len).float().unsqueeze(1)
    div = torch.exp(torch.arange(0, n_filters, 2).float() * -(math.log(10000.0) / n_filters))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _ln(x, sd, key):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _lin(x, sd, key):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def depthwise_separable_conv(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, k: int,
                             relu: bool = True) -> torch.Tensor:
    """(N,L,D) → depthwise Conv1d(groups=D, pad=k//2) → pointwise 1x1 → ReLU — model/cnn.py:25-28,42-47."""
    y = x.transpose(1, 2)
    D = y.shape[1]
    y = F.conv1d(y, sd[prefix + ".depthwise_conv.weight"], sd[prefix + ".depthwise_conv.bias"],
                 padding=k // 2, groups=D)
    y = F.conv1d(y, sd[prefix + ".pointwise_conv.weight"], sd[prefix + ".pointwise_conv.bias"])
    if relu:
        y = F.relu(y)
    return y.transpose(1, 2)


def multi_headed_attention(x: torch.Tensor, mask: Optional[torch.Tensor], sd, prefix: str, nh: int) -> torch.Tensor:
    """4-Linear scaled-dot-product attention with the QUERY-row mask quirk — model/self_attention.py:37-71."""
    bsz, L, D = x.shape
    dk = D // nh
    q, k, v = [F.linear(x, sd["%s.linears.%d.weight" % (prefix, i)], sd["%s.linears.%d.bias" % (prefix, i)])
               .view(bsz, -1, nh, dk).transpose(1, 2) for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        scores = scores.masked_fill(mask.view(bsz, 1, -1, 1) == 0, -1e9)
    p = torch.softmax(scores, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).contiguous().view(bsz, -1, nh * dk)
    return F.linear(o, sd[prefix + ".linears.3.weight"], sd[prefix + ".linears.3.bias"])


def encoder_block(x: torch.Tensor, mask: torch.Tensor, sd, prefix: str, n_conv: int, k: int, nh: int) -> torch.Tensor:
    """PE → n_conv×[LN → DWSepConv → +res] → optional [LN → MHA → +res] → LN — model/encoder.py:29-52 (eval: no dropout)."""
    out = x + sd[prefix + ".position_encoding.pe"][: x.shape[-2], :]
    blk = prefix.rsplit(".", 1)[-1]
    for i in range(n_conv):
        res = out
        out = _ln(out, sd, "%s.layer_norm.%d" % (prefix, i))
        if i % 2 == 0:  # model/encoder.py:41-42
            out = _drop("%s.enc.%s.%d" % (DROP_PREFIX, blk, i), out)
        out = depthwise_separable_conv(out, sd, "%s.conv.%d" % (prefix, i), k)
        out = out + res
    if nh != 0:
        res = out
        out = _ln(out, sd, prefix + ".attn_layer_norm")
        out = multi_headed_attention(out, mask, sd, prefix + ".multi_head_attn", nh)
        out = out + res
    return _ln(out, sd, prefix + ".final_layer_norm")


def stacked_encoder(x, mask, sd, prefix: str, n_blocks: int, n_conv: int, k: int, nh: int):
    """model/encoder.py:66-74."""
    for b in range(n_blocks):
        x = encoder_block(x, mask, sd, "%s.stacked_encoderBlocks.%d" % (prefix, b), n_conv, k, nh)
    return x


def text_init_fc(x, sd):
    """bert_word_encoding_fc: LN768 → Lin768→300 → ReLU → LN300 — model/stage.py:85-91."""
    y = _drop(DROP_PREFIX + ".fc", _ln(x, sd, "bert_word_encoding_fc.0"))
    y = F.relu(_lin(y, sd, "bert_word_encoding_fc.2"))
    return _ln(y, sd, "bert_word_encoding_fc.4")


def vid_init_fc(x, sd):
    """vid_fc: LN300 → Lin300→300 → ReLU → LN300 — model/stage.py:98-104."""
    y = _drop(DROP_PREFIX + ".fc", _ln(x, sd, "vid_fc.0"))
    y = F.relu(_lin(y, sd, "vid_fc.2"))
    return _ln(y, sd, "vid_fc.4")


def input_embedding(x, sd):
    """Lin300→128 → ReLU → LN128 — model/stage.py:115-120."""
    return _ln(F.relu(_lin(_drop(DROP_PREFIX + ".emb", x), sd, "input_embedding.1")), sd, "input_embedding.3")


def base_encoder(x, mask, sd, opt, kind: str):
    """init_fc → input_embedding → input_encoder — model/stage.py:351-363."""
    y = text_init_fc(x, sd) if kind == "text" else vid_init_fc(x, sd)
    y = input_embedding(y, sd)
    return stacked_encoder(y, mask, sd, "input_encoder", opt.input_encoder_n_blocks, opt.input_encoder_n_conv,
                           opt.input_encoder_kernel_size, opt.input_encoder_n_heads)


def structured_attention(C, Q, c_mask, q_mask, scale: float):
    """Cosine-similarity attention — model/context_query_attention.py:35-101 (noun_mask/void_vector = None).

    C (N,5,1|Li,Lqa,D), Q (N,1,Li,Lr,D), c_mask (N,5,1|Li,Lqa), q_mask (N,1,Li,Lr)
    → A (N,5,Li,Lqa,D), S (masked raw cosine), S_mask, S_ (masked softmax weights).
    """
    Cn = _drop(DROP_PREFIX + ".cn", F.normalize(C, p=2, dim=-1))  # context_query_attention.py:95-96
    Qn = _drop(DROP_PREFIX + ".qn", F.normalize(Q, p=2, dim=-1))
    S_mask = torch.matmul(c_mask.unsqueeze(-1), q_mask.unsqueeze(-2))
    S = torch.matmul(Cn, Qn.transpose(-2, -1))
    S = S - 1e10 * (1 - S_mask)
    S_ = F.softmax(S * scale, dim=-1) * S_mask
    A = torch.matmul(S_, Q)
    return A, S, S_mask, S_


def qa_ctx_attention(qa_embed, ctx_embed, qa_mask, ctx_mask, sd, opt):
    """model/stage.py:365-387."""
    num_img = ctx_mask.shape[2]
    u_a, raw_s, s_mask, s_norm = structured_attention(qa_embed, ctx_embed, qa_mask, ctx_mask, opt.scale)
    qa_rep = qa_embed.repeat(1, 1, num_img, 1, 1)
    mixed = torch.cat([qa_rep, u_a, qa_rep * u_a], dim=-1)
    mixed = F.relu(_lin(_drop(DROP_PREFIX + ".c2q", _ln(mixed, sd, "c2q_down_projection.0")), sd, "c2q_down_projection.2"))
    mixed_mask = (s_mask.sum(-1) != 0).float()
    return mixed, mixed_mask, raw_s, s_norm


def concat_fc(sub_mix, vid_mix, sd):
    """cat[s, v, s⊙v] → LN384 → Lin384→128 → ReLU → LN128 — model/stage.py:107-113,276-279."""
    x = torch.cat([sub_mix, vid_mix, sub_mix * vid_mix], dim=-1)
    x = F.relu(_lin(_drop("cat", _ln(x, sd, "concat_fc.0")), sd, "concat_fc.2"))
    return _ln(x, sd, "concat_fc.4")


_WRAPPER_TAGS = {"cls_projection_layers.0": "proj", "temporal_scoring_st_layers.0": "st", "temporal_scoring_ed_layers.0": "ed",
                 "classifier": "cls"}


def linear_wrapper(x, sd, prefix: str, relu: bool):
    """LN → (dropout) → Linear [→ ReLU] — model/stage.py:15-32 (layer_norm=True)."""
    y = _lin(_drop(_WRAPPER_TAGS.get(prefix, prefix), _ln(x, sd, prefix + ".conv.0")), sd, prefix + ".conv.2")
    return F.relu(y) if relu else y


def conv_linear(x, sd, prefix: str, k: int = 3):
    """LN → (dropout) → DWSepConv(k) with ReLU — model/stage.py:35-52."""
    return depthwise_separable_conv(_ln(x, sd, prefix + ".conv.0"), sd, prefix + ".conv.2", k)


# ----------------------------------------------------------------------------------------------
# span search — model/model_utils.py:78-121
# ----------------------------------------------------------------------------------------------
def find_max_triples(p1: torch.Tensor, p2: torch.Tensor, topN: int = 5, prob_thd: Optional[float] = None):
    """outer product, keep upper triangle (ed ≥ st), numpy argsort descending — model_utils.py:92-121."""
    product = torch.bmm(p1.unsqueeze(2), p2.unsqueeze(1))
    upper = torch.stack([torch.triu(p) for p in product]).cpu().numpy()
    out = []
    for e in upper:
        rows, cols = np.unravel_index(np.argsort(e, axis=None), e.shape)
        rows, cols = rows[::-1][:topN], cols[::-1][:topN]
        triples = list(zip(rows.tolist(), cols.tolist(), e[rows, cols].tolist()))
        if prob_thd is not None:
            triples = [t for t in triples if t[2] >= prob_thd]
        out.append(triples)
    return out


def best_span_scan(p1: np.ndarray, p2: np.ndarray) -> Tuple[int, int, float]:
    """O(T) restatement of ``find_max_triples(topN=1)`` with the same fp32 products and the same tie rule.

    numpy's argsort is ascending and the reference takes the LAST flat index, so among equal fp32
    products the winner is the largest flat index st*T+ed; lower-triangle zeros also take part
    (model_utils.py:102-105,118-120), which matters only if every upper product is ≤ 0.
    """
    T = p1.shape[0]
    p1 = p1.astype(np.float32)
    p2 = p2.astype(np.float32)
    best_v, best_flat = np.float32(-np.inf), -1
    for i in range(T):
        for j in range(T):
            v = np.float32(p1[i] * p2[j]) if j >= i else np.float32(0.0)
            if v > best_v or (v == best_v and i * T + j > best_flat):
                best_v, best_flat = v, i * T + j
    return best_flat // T, best_flat % T, float(best_v)


def expand_span(span, expand_length: int = 2):
    """model/model_utils.py:78-89."""
    return [max(0, span[0] - expand_length), span[1] + expand_length]


def compute_temporal_iou(pred, gt) -> float:
    """model/model_utils.py:37-51 (the 'union' is the convex hull, as in the reference)."""
    inter = max(0, min(pred[1], gt[1]) - max(pred[0], gt[0]))
    union = max(pred[1], gt[1]) - min(pred[0], gt[0])
    return 0 if union == 0 else 1.0 * inter / union


# ----------------------------------------------------------------------------------------------
# heads — model/stage.py:389-555
# ----------------------------------------------------------------------------------------------
def residual_temporal_predictor(x, sd, layer_idx: int = 0):
    """x + LinearWrapper(x); st/ed = LinearWrapper(128→1) — model/stage.py:469-482 (layer 0 is a LinearWrapper)."""
    if layer_idx == 0:
        x = x + linear_wrapper(x, sd, "cls_projection_layers.0", relu=True)
    else:
        x = x + conv_linear(x, sd, "cls_projection_layers.%d" % layer_idx, 3)
    st = linear_wrapper(x, sd, "temporal_scoring_st_layers.%d" % layer_idx, relu=False)
    ed = linear_wrapper(x, sd, "temporal_scoring_ed_layers.%d" % layer_idx, relu=False)
    return x, torch.cat([st, ed], dim=2)


def get_proposals_eval(max_statement, max_statement_mask, t_scores, extra_span_length: int):
    """Eval branch of get_proposals — model/stage.py:439-467.  Returns (N,5,2D) features and (5N,2) [st, ed] spans."""
    bsz, num_a, num_img, _ = max_statement_mask.shape
    probs = F.softmax(t_scores, dim=2).view(bsz * num_a, -1, 2)
    triples = find_max_triples(probs[:, :, 0], probs[:, :, 1], topN=1, prob_thd=None)
    flat = torch.FloatTensor([t for sub in triples for t in sub])
    spans = flat[:, :2].long()
    ms = max_statement.view(bsz * num_a, num_img, -1)
    mm = max_statement_mask.view(bsz * num_a, num_img, -1)
    loc, glob = [], []
    for idx in range(bsz * num_a):
        sp = expand_span([int(spans[idx, 0]), int(spans[idx, 1]) + 1], extra_span_length)
        glob.append(torch.max(mask_logits(ms[idx], mm[idx]), 0)[0])
        loc.append(torch.max(mask_logits(ms[idx, sp[0]:sp[1]], mm[idx, sp[0]:sp[1]]), 0)[0])
    feat = torch.cat([torch.stack(loc), torch.stack(glob)], dim=-1)
    return feat.view(bsz, num_a, -1), spans


def get_proposals_train(max_statement, max_statement_mask, t_scores, targets, ts_labels,
                        extra_span_length: int, max_num_proposal=1, iou_thd=0.5, ce_prob_thd=0.01):
    """Train branch of get_proposals — model/stage.py:407-438: GT span + predicted span if IoU ≥ thd."""
    bsz = max_statement_mask.shape[0]
    ca = F.softmax(t_scores[torch.arange(bsz), targets].detach(), dim=1)
    pred = find_max_triples(ca[:, :, 0], ca[:, :, 1], topN=max_num_proposal, prob_thd=ce_prob_thd)
    pred = [[[e[0], e[1] + 1, e[2]] for e in sub] for sub in pred]
    gts = list(zip(ts_labels["st"].tolist(), (ts_labels["ed"] + 1).tolist()))
    spans = []
    for gt, sub in zip(gts, pred):
        cur = [gt]
        for p in sub:
            if compute_temporal_iou(p[:2], gt) >= iou_thd:
                cur.append(p[:2])
        spans.append(cur)
    loc, glob, tg = [], [], []
    for idx, (t, sub) in enumerate(zip(targets, spans)):
        tg.extend([t] * len(sub))
        g = torch.max(mask_logits(max_statement[idx], max_statement_mask[idx]), 1)[0]
        glob.extend([g] * len(sub))
        for sp in sub:
            sp = expand_span(list(sp), extra_span_length)
            loc.append(torch.max(mask_logits(max_statement[idx, :, sp[0]:sp[1]],
                                             max_statement_mask[idx, :, sp[0]:sp[1]]), 1)[0])
    feat = torch.cat([torch.stack(loc), torch.stack(glob)], dim=-1)
    return feat, targets.new_tensor(tg)


def classifier_head(statement, statement_mask, targets, ts_labels, ts_labels_mask, sd, opt, training: bool,
                    inter: Optional[dict] = None):
    """classfier_head_multi_proposal — model/stage.py:484-537 (t_iter layers ≥1 never reach the outputs, :515-518)."""
    bsz, num_a, num_img, num_words = statement_mask.shape
    st = statement.view(bsz * num_a * num_img, num_words, -1)
    sm = statement_mask.view(bsz * num_a * num_img, num_words)
    st = stacked_encoder(st, sm, sd, "cls_encoder", opt.cls_encoder_n_blocks, opt.cls_encoder_n_conv,
                         opt.cls_encoder_kernel_size, opt.cls_encoder_n_heads)
    max_statement = torch.max(mask_logits(st, sm.unsqueeze(2)), 1)[0]
    max_statement_mask = (sm.sum(1) != 0).float().view(bsz, num_a, num_img, 1)
    max_statement = max_statement.view(bsz * num_a, num_img, -1)
    enc0, t0 = residual_temporal_predictor(max_statement, sd, 0)
    t_scores = t0.view(bsz, num_a, num_img, 2)
    if opt.t_iter > 0:  # model/stage.py:515-516 averages layer 0 with itself
        t_scores = 0.5 * (t_scores + torch.stack([t_scores]).mean(0))
    t_scores = mask_logits(t_scores, ts_labels_mask.view(bsz, 1, num_img, 1))
    stacked = enc0.view(bsz, num_a, num_img, -1)
    spans = None
    if opt.add_local:
        if training:
            feat, targets = get_proposals_train(stacked, max_statement_mask, t_scores, targets, ts_labels,
                                                opt.extra_span_length)
        else:
            feat, spans = get_proposals_eval(stacked, max_statement_mask, t_scores, opt.extra_span_length)
    else:
        feat = torch.max(mask_logits(stacked, max_statement_mask), 2)[0]
    answer = linear_wrapper(feat, sd, "classifier", relu=False).squeeze(2)
    if inter is not None:
        inter.update(max_statement=max_statement.view(bsz, num_a, num_img, -1), max_statement_mask=max_statement_mask,
                     encoded_max_statement=stacked, spans=spans, cls_feat=feat)
    return answer, targets, t_scores


def get_ts_loss(t_scores, ts_labels, answer_indices):
    """model/stage.py:539-555 (CrossEntropyLoss(reduction='sum'), :180)."""
    bsz = len(answer_indices)
    ca = t_scores[torch.arange(bsz, dtype=torch.long), answer_indices]
    ls = F.cross_entropy(ca[:, :, 0], ts_labels["st"], reduction="sum")
    le = F.cross_entropy(ca[:, :, 1], ts_labels["ed"], reduction="sum")
    return (ls + le) / 2.0


# ----------------------------------------------------------------------------------------------
# the forward — model/stage.py:199-348
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def stage_forward(sd: Dict[str, torch.Tensor], opt, batch, training: bool = False, keep: bool = False) -> dict:
    """Eval/train-mode forward_main with dropout off.  Returns a dict:

    ``answer`` (N|N_new,5), ``target``, ``t_scores`` (masked raw logits, N,5,T,2), ``t_probs`` (softmax over T),
    ``temporal_loss``; with ``keep=True`` also the intermediates used by the kernel-level parity tests.
    """
    global DROP_PREFIX
    bsz, num_a, hsz = len(batch.qid), 5, opt.hsz
    inter: dict = {}
    DROP_PREFIX = "q"
    a_embed = base_encoder(batch.qas_bert.view(bsz * num_a, -1, opt.embedding_size),
                           batch.qas_mask.view(bsz * num_a, -1), sd, opt, "text")
    a_embed = a_embed.view(bsz, num_a, 1, -1, hsz)
    a_mask = batch.qas_mask.view(bsz, num_a, 1, -1)
    att_sub = att_vid = att_vid_mask = att_sub_mask = None
    if opt.sub_flag:
        T, Ls = batch.sub_bert.shape[1:3]
        DROP_PREFIX = "s"
        sub_embed = base_encoder(batch.sub_bert.view(bsz * T, Ls, -1), batch.sub_mask.view(bsz * T, Ls), sd, opt, "text")
        sub_embed = sub_embed.contiguous().view(bsz, 1, T, Ls, -1)
        att_sub, att_sub_mask, sub_raw_s, sub_norm_s = qa_ctx_attention(
            a_embed, sub_embed, a_mask, batch.sub_mask.view(bsz, 1, T, Ls), sd, opt)
        if keep:
            inter.update(sub_embed=sub_embed, sub_raw_s=sub_raw_s, sub_normalized_s=sub_norm_s, attended_sub=att_sub)
    if opt.vfeat_flag:
        T, R = batch.vid.shape[1:3]
        DROP_PREFIX = "v"
        v = F.normalize(batch.vid, p=2, dim=-1)  # model/stage.py:256
        vid_embed = base_encoder(v.view(bsz * T, R, -1), batch.vid_mask.view(bsz * T, R), sd, opt, "vid")
        vid_embed = vid_embed.contiguous().view(bsz, 1, T, R, -1)
        att_vid, att_vid_mask, vid_raw_s, vid_norm_s = qa_ctx_attention(
            a_embed, vid_embed, a_mask, batch.vid_mask.view(bsz, 1, T, R), sd, opt)
        inter.update(vid_raw_s=vid_raw_s)
        if keep:
            inter.update(vid_embed=vid_embed, vid_normalized_s=vid_norm_s, attended_vid=att_vid)
    if opt.sub_flag and opt.vfeat_flag:
        statement = concat_fc(att_sub, att_vid, sd)
        statement_mask = att_vid_mask  # model/stage.py:280-282: the vid mask only
    elif opt.sub_flag:
        statement, statement_mask = att_sub, att_sub_mask
    else:
        statement, statement_mask = att_vid, att_vid_mask
    head_inter = {} if keep else None
    DROP_PREFIX = "c"
    answer, target, t_scores = classifier_head(statement, statement_mask, batch.target, batch.ts_label,
                                               batch.ts_label_mask, sd, opt, training, head_inter)
    out = dict(answer=answer, target=target, t_scores=t_scores, t_probs=F.softmax(t_scores, dim=2),
               temporal_loss=get_ts_loss(t_scores, batch.ts_label, batch.target), vid_raw_s=inter.get("vid_raw_s"))
    if keep:
        out.update(a_embed=a_embed, statement=statement, statement_mask=statement_mask, **inter, **head_inter)
    return out


# ----------------------------------------------------------------------------------------------
# inference post-processing — model/stage.py:748-806, inference.py:13-35
# ----------------------------------------------------------------------------------------------
def get_att_prediction(scores, object_vocab, words, vid_names, qids, img_indices, boxes, start_indices,
                       score_thd: float = 0.2, num_a: int = 5):
    """Per (question, answer, annotated frame, object word): regions with raw cosine >= ``score_thd``, ascending by
    score — model/stage.py:775-805 (same loop order, same numpy calls, so ties sort the same way)."""
    out = []
    for b in range(len(scores)):
        start_idx = start_indices[b]
        per_q = {}
        for a in range(num_a):
            per_q[a] = []
            for local_i in range(len(boxes[b])):
                gi = local_i + start_idx
                cur = scores[b, a, gi]
                cur_words = words[b, a].tolist()
                cur_boxes = boxes[b][local_i]
                for wi, w in enumerate(cur_words):
                    if w in object_vocab:
                        row = cur[wi].detach().cpu().numpy()
                        ids = np.nonzero(row >= score_thd)[0].tolist()
                        sc = [float(row[i]) for i in ids]
                        bx = [cur_boxes[i] for i in ids]
                        order = np.argsort(sc)
                        per_q[a].append({"pred": [sc[i] for i in order], "bbox": [bx[i] for i in order],
                                         "word": int(words[b, a, wi]), "qid": int(qids[b]), "vid_name": vid_names[b],
                                         "img_idx": img_indices[b][gi]})
        out.append(per_q)
    return out


def find_max_pair(p1, p2):
    """inference.py:13-35: running first-argmax of p1, best p1[k1]*p2[k2] with k1 <= k2 under strict '>' (Python floats,
    i.e. double products of the fp32 probabilities); starts from span (0, 1), value 0."""
    max_val, best, k1 = 0, (0, 1), 0
    for i in range(len(p1)):
        v1 = p1[k1]
        if v1 < p1[i]:
            k1, v1 = i, p1[i]
        if v1 * p2[i] > max_val:
            best, max_val = (k1, i), v1 * p2[i]
    return best, float(max_val)


# ----------------------------------------------------------------------------------------------
# collate — tvqa_dataset.py:515-553 (pad_sequences_2d) and :631-688 (prepare_inputs' clipping)
# ----------------------------------------------------------------------------------------------
def pad_sequences_2d(sequences, dtype=torch.float):
    """Double-nested list of (len, D) tensors -> zero-padded (B, max_para, max_sen, D) tensor + (B, max_para, max_sen) mask."""
    bsz = len(sequences)
    max_para = max(len(seq) for seq in sequences)
    max_sen = max(len(w) for seq in sequences for w in seq)
    extra = tuple(sequences[0][0].shape[1:]) if torch.is_tensor(sequences[0][0]) else ()
    if not torch.is_tensor(sequences[0][0]):
        sequences = [[torch.LongTensor(w) for w in seq] for seq in sequences]
    padded = torch.zeros((bsz, max_para, max_sen) + extra, dtype=dtype)
    mask = torch.zeros(bsz, max_para, max_sen).float()
    for b in range(bsz):
        for s, w in enumerate(sequences[b]):
            padded[b, s, :len(w)] = w
            mask[b, s, :len(w)] = 1
    return padded, mask


def collate_and_prepare(items, max_len_dict):
    """pad_collate + prepare_inputs of the reference for the fields STAGE.forward reads (host tensors)."""
    qas_bert, qas_mask = pad_sequences_2d([d["qas_bert"] for d in items])
    sub_bert, sub_mask = pad_sequences_2d([d["sub_bert"] for d in items])
    vid, vid_mask = pad_sequences_2d([d["vfeat"] for d in items])
    qas, _ = pad_sequences_2d([[x.tolist() for x in d["qas"]] for d in items], dtype=torch.long)
    lens = [len(d["image_indices"]) for d in items]
    ts_mask = torch.zeros(len(items), max(lens))
    for i, l in enumerate(lens):
        ts_mask[i, :l] = 1
    ql = min(qas_bert.shape[2], max_len_dict["max_qa_l"])
    vl, sl = max_len_dict["max_vid_l"], max_len_dict["max_sub_l"]
    return Batch(qas=qas[:, :, :ql], qas_bert=qas_bert[:, :, :ql].contiguous(), qas_mask=qas_mask[:, :, :ql].contiguous(),
                 sub_bert=sub_bert[:, :vl, :sl].contiguous(), sub_mask=sub_mask[:, :vl, :sl].contiguous(),
                 vid=vid[:, :vl].contiguous(), vid_mask=vid_mask[:, :vl].contiguous(), ts_label_mask=ts_mask[:, :vl].contiguous(),
                 target=torch.tensor([d["target"] for d in items]),
                 ts_label=dict(st=torch.tensor([d["ts_label"][0] for d in items]), ed=torch.tensor([d["ts_label"][1] for d in items])),
                 qid=[d["qid"] for d in items], vid_name=[d["vid_name"] for d in items], att_labels=None, anno_st_idx=[0] * len(items),
                 q_l=[d["q_l"] for d in items], image_indices=[d["image_indices"] for d in items], boxes=None,
                 use_hard_negatives=False, eval_object_word_ids=None)


# ----------------------------------------------------------------------------------------------
# attention ranking loss — model/stage.py:556-746 (use_hard_negatives=False, the only mode the defaults ever reach)
# ----------------------------------------------------------------------------------------------
def att_loss_pairs(att_labels, target, start_indices, num_negatives: int = 2):
    """The (positive, negative) index lists of get_att_loss — model/stage.py:613-694 with sample_negatives' random branch
    (:603-606).  Rows are (batch, answer, image, word, region); the negatives are drawn with ``torch.randint`` on the global
    CPU generator, one call per (example, image, word) in loop order, exactly like the reference."""
    pos_c, neg_c = [], []
    for b in range(len(target)):
        ca = int(target[b])
        gt = att_labels[b]
        st = start_indices[b]
        bpos, bneg = [], []
        for i, lab in enumerate(gt):
            lab = lab.cpu()
            p = torch.nonzero(lab)
            if len(p) == 0:
                continue
            bpos.append(torch.cat([p.new_full([len(p), 1], st + i), p], dim=1))
            q = torch.nonzero(lab == 0)
            bneg.append(torch.cat([q.new_full([len(q), 1], st + i), q], dim=1))
        if not bpos:
            continue
        bpos, bneg = torch.cat(bpos, 0), torch.cat(bneg, 0)
        for img in bpos[:, 0].unique().tolist():
            ip, ineg = bpos[bpos[:, 0] == img], bneg[bneg[:, 0] == img]
            for w in ip[:, 1].unique().tolist():
                wp, wn = ip[ip[:, 1] == w], ineg[ineg[:, 1] == w]
                sp = torch.cat([wp] * num_negatives, dim=0)
                sn = wn[torch.randint(low=0, high=len(wn), size=(len(sp),), dtype=torch.long)]
                base = torch.LongTensor([[b, ca]] * len(sp))
                pos_c.append(torch.cat([base, sp], dim=1))
                neg_c.append(torch.cat([base, sn], dim=1))
    return torch.cat(pos_c, 0), torch.cat(neg_c, 0)


def get_att_loss(scores, att_labels, target, start_indices, num_negatives=2, att_loss_type="lse", alpha=20.0, margin=0.1):
    """model/stage.py:731-746 on the pairs above."""
    pos, neg = att_loss_pairs(att_labels, target, start_indices, num_negatives)
    ps = scores[pos[:, 0], pos[:, 1], pos[:, 2], pos[:, 3], pos[:, 4]]
    ns = scores[neg[:, 0], neg[:, 1], neg[:, 2], neg[:, 3], neg[:, 4]]
    if att_loss_type == "hinge":
        return torch.clamp(margin + ns - ps, min=0).sum()
    return torch.log1p(torch.exp(alpha * (ns - ps))).sum()


def make_att_labels(batch, n_img_max: int = 4, seed: int = 0):
    """Synthetic att_labels / anno_st_idx for a synthetic batch: per example a few annotated frames inside the clip, each a
    (Lq, R_valid) 0/1 matrix with a handful of positives (tvqa_dataset.py builds them from the bounding-box annotations)."""
    g = torch.Generator().manual_seed(seed)
    N, _, Lq = batch.qas_mask.shape
    labels, starts = [], []
    for b in range(N):
        n_frames = int(batch.ts_label_mask[b].sum())
        n_img = int(torch.randint(1, min(n_img_max, n_frames) + 1, (1,), generator=g))
        st = int(torch.randint(0, n_frames - n_img + 1, (1,), generator=g))
        q_len = int(batch.qas_mask[b, int(batch.target[b])].sum())
        cur = []
        for i in range(n_img):
            r_valid = max(2, int(batch.vid_mask[b, st + i].sum()))
            lab = (torch.rand(q_len, r_valid, generator=g) < 0.08).float()
            if i == 0:
                lab[0, 0] = 1  # at least one positive per example
            lab[:, -1] = 0     # and at least one negative per word
            cur.append(lab)
        labels.append(cur)
        starts.append(st)
    return labels, starts
