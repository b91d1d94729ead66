"""Collate and host->device staging for STAGE batches: the step directly in front of the hot path
(reference ``pad_collate`` / ``pad_sequences_2d`` / ``prepare_inputs``, tvqa_dataset.py:515-553,592-688).

The reference pads every feature tensor to the batch maximum on the HOST and uploads the padded fp32 tensors
(92 MB per example at configs[1]: as soon as the forward runs faster than ~500 examples/s per GPU the PCIe copy is the
bottleneck).  Here the collate keeps the features RAGGED — the valid rows of all sequences concatenated
in one pinned buffer per stream plus a length per sequence — ``prepare_inputs`` uploads exactly those rows and a kernel
(``stage_pad_rows``) scatters them into the zero-padded tensors and masks on the device.  The model sees the same
``(N, 5, Lq, 768)`` / ``(N, T, Ls, 768)`` / ``(N, T, R, 300)`` tensors and masks, bit for bit (pads are zeros, which the
reference's outputs depend on).

``pad_collate(data)`` and ``prepare_inputs(batch, max_len_dict, device)`` keep the reference's names and signatures; the
object travelling between them is a ``RaggedBatch`` instead of an edict of padded tensors.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .host import Batch


def _rows_and_lens(seqs: Sequence[Sequence[torch.Tensor]], d: int, pin: bool):
    """seqs[b][s] = tensor (len, d) -> (rows (R, d) concatenated in (b, s) order, lens [B][S_b] python ints)."""
    lens = [[int(t.shape[0]) for t in ex] for ex in seqs]
    total = sum(sum(l) for l in lens)
    rows = torch.empty(total, d, dtype=torch.float32, pin_memory=pin)
    at = 0
    for ex in seqs:
        for t in ex:
            n = int(t.shape[0])
            if n:
                rows[at:at + n] = t
                at += n
    return rows, lens


def _len_matrix(lens: List[List[int]], width: int) -> torch.Tensor:
    m = torch.zeros(len(lens), width, dtype=torch.int32)
    for b, row in enumerate(lens):
        if row:
            m[b, :len(row)] = torch.tensor(row[:width], dtype=torch.int32)
    return m


def _offsets(len_mat: torch.Tensor) -> torch.Tensor:
    flat = len_mat.reshape(-1).to(torch.int64)
    return torch.cumsum(flat, 0) - flat


class RaggedBatch(dict):
    """What ``pad_collate`` hands to ``prepare_inputs``: per stream the concatenated valid rows (pinned when CUDA is
    available), the per-sequence lengths as an (N, n_seq_max) int32 matrix, and the reference's pass-through fields."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def h2d_bytes(self) -> int:
        return sum(self[k].numel() * self[k].element_size() for k in ("qas_rows", "sub_rows", "vid_rows", "qas_len", "sub_len", "vid_len"))


def pad_collate(data: Sequence[Dict[str, Any]], pin: Optional[bool] = None) -> RaggedBatch:
    """List of dataset items (``TVQADataset.__getitem__``, tvqa_dataset.py:246-350) -> RaggedBatch.

    Item fields used: ``qas_bert`` (5 tensors (len, 768)), ``sub_bert`` (per frame (len, 768)), ``vfeat`` (per frame
    (regions, 300)), ``qas`` (5 word-id lists), ``target``, ``ts_label`` ([st, ed]) , ``image_indices``, ``vid_name``,
    ``qid``, ``att_labels``, ``anno_st_idx``, ``q_l``, ``boxes`` (the text-id streams ``sub`` / ``vcpt`` of the reference are
    not read by STAGE.forward and are not collated)."""
    pin = torch.cuda.is_available() if pin is None else pin
    b = RaggedBatch()
    b["qas_rows"], qas_lens = _rows_and_lens([d["qas_bert"] for d in data], int(data[0]["qas_bert"][0].shape[1]), pin)
    b["sub_rows"], sub_lens = _rows_and_lens([d["sub_bert"] for d in data], int(data[0]["sub_bert"][0].shape[1]), pin)
    b["vid_rows"], vid_lens = _rows_and_lens([d["vfeat"] for d in data], int(data[0]["vfeat"][0].shape[1]), pin)
    n_frames = [len(d["image_indices"]) for d in data]
    t_max = max(max(len(x) for x in sub_lens), max(len(x) for x in vid_lens), max(n_frames))
    b["qas_len"] = _len_matrix(qas_lens, 5)
    b["sub_len"] = _len_matrix(sub_lens, t_max)
    b["vid_len"] = _len_matrix(vid_lens, t_max)
    for k in ("qas", "sub", "vid"):  # row offset of every sequence in its stream's buffer
        b[k + "_off"] = _offsets(b[k + "_len"]).view_as(b[k + "_len"])
        if pin:
            b[k + "_len"], b[k + "_off"] = b[k + "_len"].pin_memory(), b[k + "_off"].pin_memory()
    b["n_frames"] = torch.tensor(n_frames, dtype=torch.int32)
    # word ids are tiny: padded on the host like the reference does (pad_sequences_2d, dtype long)
    lq = max(1, int(b["qas_len"].max()))
    qas = torch.zeros(len(data), 5, lq, dtype=torch.long)
    for i, d in enumerate(data):
        for a, ids in enumerate(d["qas"]):
            ids = torch.as_tensor(ids, dtype=torch.long)
            qas[i, a, :ids.numel()] = ids
    b["qas"] = qas
    b["target"] = torch.tensor([d["target"] for d in data], dtype=torch.long)
    if data[0].get("ts_label") is None:
        b["ts_label"] = None
    else:
        b["ts_label"] = dict(st=torch.tensor([d["ts_label"][0] for d in data], dtype=torch.long),
                             ed=torch.tensor([d["ts_label"][1] for d in data], dtype=torch.long))
    for k in ("vid_name", "qid", "att_labels", "anno_st_idx", "image_indices", "q_l", "boxes", "ts", "object_labels"):
        b[k] = [d.get(k) for d in data]
    return b


def prepare_inputs(batch: RaggedBatch, max_len_dict: Optional[Dict[str, int]] = None, device="cuda") -> Tuple[Batch, torch.Tensor, list]:
    """Upload the ragged rows and build the padded model inputs on the device (reference prepare_inputs,
    tvqa_dataset.py:631-688: clip to ``max_qa_l`` / ``max_vid_l`` / ``max_sub_l`` / ``max_vcpt_l``, move to the device).
    Returns ``(model_in_dict, targets, qids)`` like the reference."""
    mx = dict(max_qa_l=1 << 30, max_vid_l=1 << 30, max_sub_l=1 << 30, max_vcpt_l=1 << 30)
    mx.update(max_len_dict or {})
    dev = torch.device(device)
    n = batch["qas_len"].shape[0]
    out = Batch()

    def stream(rows_key, len_key, n_seq_clip, l_clip):
        lens, off = batch[len_key], batch[len_key[:3] + "_off"]  # offsets use the unclipped lengths (rows are stored unclipped)
        n_seq = min(lens.shape[1], n_seq_clip)
        l_pad = max(1, min(int(lens[:, :n_seq].max()), l_clip))
        lens_c = lens[:, :n_seq].contiguous().reshape(-1)     # (a no-op view, still pinned, unless max_vid_l clips frames away)
        off_c = off[:, :n_seq].contiguous().reshape(-1)
        rows = batch[rows_key].to(dev, non_blocking=True)
        x, m = ops.pad_rows(rows, off_c.to(dev, non_blocking=True), lens_c.to(dev, non_blocking=True), l_pad)
        return x.view(n, n_seq, l_pad, -1), m.view(n, n_seq, l_pad)

    out["qas_bert"], out["qas_mask"] = stream("qas_rows", "qas_len", 5, mx["max_qa_l"])
    out["sub_bert"], out["sub_mask"] = stream("sub_rows", "sub_len", mx["max_vid_l"], mx["max_sub_l"])
    out["vid"], out["vid_mask"] = stream("vid_rows", "vid_len", mx["max_vid_l"], 1 << 30)
    t = out["sub_bert"].shape[1]
    out["qas"] = batch["qas"][:, :, :out["qas_bert"].shape[2]].to(dev, non_blocking=True)
    out["ts_label_mask"] = (torch.arange(t)[None, :] < batch["n_frames"][:, None]).float().to(dev, non_blocking=True)
    out["ts_label"] = None if batch["ts_label"] is None else {k: v.to(dev, non_blocking=True) for k, v in batch["ts_label"].items()}
    out["target"] = batch["target"].to(dev, non_blocking=True)
    max_att_imgs = mx["max_vid_l"]
    out["att_labels"] = None if batch["att_labels"][0] is None else [
        [lab[:mx["max_qa_l"]].to(dev) for lab in d[:max_att_imgs]] for d in batch["att_labels"]]
    for k in ("anno_st_idx", "qid", "vid_name", "ts", "q_l", "image_indices", "boxes", "object_labels"):
        out[k] = batch[k]
    out["use_hard_negatives"] = False
    out["eval_object_word_ids"] = None
    return out, out["target"], out["qid"]


def synthetic_items(n: int, t: int, r: int, lq: int, ls: int, *, seed: int = 2018, dense: bool = False, wd: int = 768, vd: int = 300):
    """Synthetic dataset items with the length distributions of SURVEY.md §8d (QA length ~U[0.5 Lq, Lq], sub-title tokens per
    frame ~U[0.3 Ls, Ls], regions ~U[0.5 R, R], frames per clip ~U[0.7 T, T]); ``dense`` = every sequence at full length."""
    g = torch.Generator().manual_seed(seed)

    def ri(lo, hi):
        return int(torch.randint(lo, hi + 1, (1,), generator=g))

    items = []
    for i in range(n):
        nf = t if dense else ri(max(1, int(0.7 * t)), t)
        st = ri(0, nf - 1)
        items.append(dict(
            qas_bert=[torch.randn(lq if dense else ri(max(1, lq // 2), lq), wd, generator=g) for _ in range(5)],
            sub_bert=[torch.randn(ls if dense else ri(max(1, int(0.3 * ls)), ls), wd, generator=g) for _ in range(nf)],
            vfeat=[torch.randn(r if dense else ri(max(1, r // 2), r), vd, generator=g) for _ in range(nf)],
            qas=None, target=ri(0, 4), ts_label=[st, ri(st, nf - 1)], image_indices=list(range(nf)),
            vid_name="syn_%d" % i, qid=i, att_labels=None, anno_st_idx=0, q_l=lq // 2, boxes=None, ts=None, object_labels=None))
        items[-1]["qas"] = [torch.randint(1, 15000, (x.shape[0],), generator=g) for x in items[-1]["qas_bert"]]
    return items
