"""Ragged collate + on-device padding against the reference's pad_collate / prepare_inputs
(tvqa_dataset.py:515-553,592-688): host logic on CPU, the upload + stage_pad_rows path on the GPU."""
import ast
import os

import pytest
import torch

from oracle import stage_oracle as O
from tvqaplus_b200 import collate

REF = "/root/reference"
MAXL = dict(max_qa_l=20, max_vid_l=7, max_sub_l=11, max_vcpt_l=300)


def _rebuild(rows, len_mat, l_pad):
    """Host re-statement of stage_pad_rows for the CPU test."""
    n, s = len_mat.shape
    out = torch.zeros(n, s, l_pad, rows.shape[1])
    mask = torch.zeros(n, s, l_pad)
    at = 0
    for b in range(n):
        for j in range(s):
            l = int(len_mat[b, j])
            k = min(l, l_pad)
            out[b, j, :k] = rows[at:at + k]
            mask[b, j, :k] = 1
            at += l
    return out, mask


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present on this host")
def test_oracle_padding_matches_live_reference():
    """tvqa_dataset.py cannot be imported here (h5py is missing), so pad_sequences_2d / flat_list_of_lists are compiled from
    the reference's own source files at test time."""
    ns = {"torch": torch}
    for path, names in ((os.path.join(REF, "utils.py"), ("flat_list_of_lists",)), (os.path.join(REF, "tvqa_dataset.py"), ("pad_sequences_2d",))):
        tree = ast.parse(open(path).read())
        fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    items = collate.synthetic_items(3, 6, 5, 9, 8, seed=1, wd=16, vd=12)
    for key in ("qas_bert", "sub_bert", "vfeat"):
        want_x, want_m = ns["pad_sequences_2d"]([d[key] for d in items], dtype=torch.float)
        got_x, got_m = O.pad_sequences_2d([d[key] for d in items])
        assert torch.equal(got_x, want_x) and torch.equal(got_m, want_m)


def test_ragged_collate_carries_exactly_the_valid_rows():
    items = collate.synthetic_items(4, 9, 6, 12, 10, seed=3, wd=16, vd=8)
    rb = collate.pad_collate(items, pin=False)
    want = O.collate_and_prepare(items, dict(max_qa_l=1 << 30, max_vid_l=1 << 30, max_sub_l=1 << 30))
    for rows_key, len_key, x_key, m_key in (("qas_rows", "qas_len", "qas_bert", "qas_mask"), ("sub_rows", "sub_len", "sub_bert", "sub_mask"),
                                            ("vid_rows", "vid_len", "vid", "vid_mask")):
        assert int(rb[len_key].sum()) == rb[rows_key].shape[0] == int(want[m_key].sum())
        x, m = _rebuild(rb[rows_key], rb[len_key], want[x_key].shape[2])
        assert torch.equal(x, want[x_key]) and torch.equal(m, want[m_key])
    assert torch.equal(rb["qas"], want["qas"])
    dense = sum(want[k].numel() * 4 for k in ("qas_bert", "sub_bert", "vid"))
    assert rb.h2d_bytes() < dense  # the point of the exercise


@pytest.mark.gpu
@pytest.mark.parametrize("maxl", [MAXL, dict(max_qa_l=1 << 30, max_vid_l=1 << 30, max_sub_l=1 << 30)])
def test_prepare_inputs_on_device_equals_reference_padding(maxl):
    items = collate.synthetic_items(3, 9, 6, 24, 14, seed=5)
    want = O.collate_and_prepare(items, maxl)
    got, targets, qids = collate.prepare_inputs(collate.pad_collate(items), maxl, "cuda:0")
    torch.cuda.synchronize()
    for k in ("qas_bert", "qas_mask", "sub_bert", "sub_mask", "vid", "vid_mask", "ts_label_mask", "qas", "target"):
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        assert torch.equal(got[k].cpu(), want[k]), k
    assert torch.equal(got["ts_label"]["st"].cpu(), want["ts_label"]["st"]) and qids == want["qid"]
    assert torch.equal(targets.cpu(), want["target"])


@pytest.mark.gpu
def test_forward_from_ragged_upload_equals_forward_from_padded_batch():
    from tests import helpers as H
    from tvqaplus_b200.model import STAGE
    items = collate.synthetic_items(2, 14, 20, 30, 40, seed=9)
    maxl = dict(max_qa_l=100, max_vid_l=300, max_sub_l=100)
    opt = O.default_opt()
    torch.manual_seed(2018)
    model = STAGE(opt).eval()
    padded = O.collate_and_prepare(items, maxl)
    want = O.stage_forward({k: v.clone() for k, v in model.state_dict().items()}, opt, padded, keep=True)
    model = model.to("cuda:0")
    got_in, _, _ = collate.prepare_inputs(collate.pad_collate(items), maxl, "cuda:0")
    answer, _, _, t_loss, t_probs = model(got_in)
    ref_answer = model(H.batch_to(padded, torch.device("cuda:0")))[0]
    assert torch.equal(answer, ref_answer)  # identical inputs -> identical kernels -> identical bits
    assert H.max_rel(answer.cpu(), want["answer"]) < H.REL_TOL and H.max_rel(t_probs.cpu(), want["t_probs"]) < H.REL_TOL
    assert torch.equal(answer.argmax(1).cpu(), want["answer"].argmax(1))
    assert torch.equal(model._last_spans.cpu().long(), want["spans"])
