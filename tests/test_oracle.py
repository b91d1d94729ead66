"""The CPU oracle against the golden vectors produced by the live reference (oracle/make_golden.py), and — when the
reference tree is present (build container only) — against the reference itself."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import stage_oracle as O
from tests import helpers as H

REF = "/root/reference"


@pytest.mark.parametrize("name", H.CASES)
def test_oracle_matches_golden(name):
    gold = H.load_golden(name)
    opt, model, batch = H.seeded_model_and_batch(name, gold)
    sd = model.state_dict()
    if H.digest(sd.values()) != str(gold["weights_sha"]) or H.batch_digest(batch) != str(gold["batch_sha"]):
        pytest.skip("torch RNG stream on this host differs from the fixture's (weights/inputs digests do not match)")
    out = O.stage_forward(sd, opt, batch, training=False)
    # same torch build + same op sequence: bit-for-bit
    assert np.array_equal(out["answer"].numpy(), gold["answer"])
    assert np.array_equal(out["t_scores"].numpy(), gold["t_scores"])
    assert np.array_equal(out["t_probs"].numpy(), gold["t_probs"])
    assert np.float32(out["temporal_loss"].item()) == gold["temporal_loss"]
    assert np.array_equal(out["vid_raw_s"].numpy()[:, :, ::3, ::3], gold["vid_raw_s"])
    # span choice
    N, T = int(gold["shape"][0]), int(gold["shape"][1])
    probs = out["t_probs"].view(N * 5, T, 2)
    tri = O.find_max_triples(probs[:, :, 0], probs[:, :, 1], topN=1)
    assert np.array_equal(np.array([[t[0][0], t[0][1]] for t in tri]), gold["spans"])
    for r in range(N * 5):  # the O(T) scan picks the same pair with the same fp32 confidence
        st, ed, conf = O.best_span_scan(probs[r, :, 0].numpy(), probs[r, :, 1].numpy())
        assert (st, ed) == tuple(gold["spans"][r]) and np.float32(conf) == np.float32(tri[r][0][2])


@pytest.mark.parametrize("name", ["tiny_h0", "cfg1_h0", "wide_h0"])
def test_oracle_train_branch_matches_golden(name):
    # heads=4 is excluded: MultiHeadedAttention keeps its own Dropout(0.1) regardless of opt.dropout
    # (model/encoder.py:26, model/self_attention.py:20,33), so the reference's train forward is random there.
    gold = H.load_golden(name)
    N, T, R, Lq, Ls, heads = [int(v) for v in gold["shape"]]
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads, dropout=0.0, use_sup_att=False)
    torch.manual_seed(2018)
    sd = STAGE(opt).state_dict()
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=2018)
    if H.digest(sd.values()) != str(gold["weights_sha"]):
        pytest.skip("torch RNG stream on this host differs from the fixture's")
    out = O.stage_forward(sd, opt, batch, training=True)
    assert np.array_equal(out["answer"].numpy(), gold["train_answer"])
    assert np.array_equal(out["target"].numpy(), gold["train_target"])
    assert np.float32(out["temporal_loss"].item()) == gold["train_temporal_loss"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("heads", [0, 4])
def test_oracle_matches_live_reference(heads):
    sys.path.insert(0, os.path.join(REF, "model"))
    sys.path.insert(0, REF)
    import stage as ref_stage
    opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads, input_encoder_n_blocks=2, t_iter=1)
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_stage.STAGE(opt).eval()
    batch = O.make_synthetic_batch(3, 11, 7, 9, 13, seed=99)
    with torch.no_grad():
        out, _, _, t_loss, t_probs, other = ref.forward_main(batch)
    mine = O.stage_forward(ref.state_dict(), opt, batch)
    assert torch.equal(mine["answer"], out)
    assert torch.equal(mine["t_probs"], t_probs)
    assert torch.equal(mine["vid_raw_s"], other["vid_raw_s"])
    assert torch.equal(mine["temporal_loss"], t_loss)


def test_synthetic_batch_schema():
    b = O.make_synthetic_batch(3, 10, 6, 12, 9, seed=1)
    assert b.qas_bert.shape == (3, 5, 12, 768) and b.sub_bert.shape == (3, 10, 9, 768) and b.vid.shape == (3, 10, 6, 300)
    # padding slots are zero-filled and masks are prefix masks
    assert torch.all(b.sub_bert[b.sub_mask == 0] == 0) and torch.all(b.vid[b.vid_mask == 0] == 0)
    assert torch.all((b.ts_label["st"] <= b.ts_label["ed"]))
    d = O.make_synthetic_batch(2, 4, 3, 5, 6, dense=True)
    assert d.sub_mask.min() == 1 and d.vid_mask.min() == 1 and d.qas_mask.min() == 1


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not present on this host")
@pytest.mark.parametrize("loss_type", ["lse", "hinge"])
def test_att_loss_restatement_matches_live_reference(loss_type):
    """get_att_loss / sample_negatives (model/stage.py:556-746): the oracle's restatement against the unmodified reference
    method, same labels and the same host RNG seed -> bit-identical loss."""
    import contextlib
    import io
    from oracle.make_golden import load_reference
    ref_stage = load_reference()
    opt = O.default_opt(att_loss_type=loss_type)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_stage.STAGE(opt)
    ref.train()
    batch = O.make_synthetic_batch(3, 12, 9, 10, 8, seed=4)
    labels, starts = O.make_att_labels(batch, seed=1)
    scores = torch.rand(3, 5, 12, 10, 9, generator=torch.Generator().manual_seed(0)) * 2 - 1
    scores[:, :, :, :, -1] = -1e10
    torch.manual_seed(7)
    want, _ = ref.get_att_loss(scores, labels, batch.target, batch.qas, batch.vid_name, batch.qid, batch.q_l, None, None, starts,
                               num_negatives=2)
    torch.manual_seed(7)
    got = O.get_att_loss(scores, labels, batch.target, starts, 2, loss_type, opt.alpha, opt.margin)
    assert torch.equal(want, got)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not present on this host")
def test_oracle_gradients_match_live_reference_training_step():
    """The training parity tests compare the CUDA backward with autograd of the oracle; this pins that oracle: the unmodified
    reference in train mode (dropout 0, use_sup_att) with main.py:53-61's loss, ``loss.backward()`` - loss and every
    parameter gradient bit-identical to autograd through the restatement."""
    import contextlib
    import io
    import torch.nn.functional as F
    from oracle.make_golden import load_reference
    ref_stage = load_reference()
    opt = O.default_opt(dropout=0.0, use_sup_att=True)
    torch.manual_seed(2018)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_stage.STAGE(opt)
    ref.train()
    N = 3
    batch = O.make_synthetic_batch(N, 12, 9, 10, 8, seed=4)
    batch["att_labels"], batch["anno_st_idx"] = O.make_att_labels(batch, seed=1)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}

    torch.manual_seed(7)
    (out, target), att_loss, _, t_loss, _ = ref(batch)
    loss = F.cross_entropy(out, target, reduction="sum") * (1.0 * N / len(target)) + 0.1 * att_loss + 0.5 * t_loss
    loss.backward()

    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(".pe") else v.clone()) for k, v in sd.items()}
    with torch.enable_grad():
        mine = O.stage_forward.__wrapped__(sdg, opt, batch, training=True)
        torch.manual_seed(7)
        att = O.get_att_loss(mine["vid_raw_s"], batch.att_labels, batch.target, batch.anno_st_idx, opt.num_negatives,
                             opt.att_loss_type, opt.alpha, opt.margin)
        mine_loss = (F.cross_entropy(mine["answer"], mine["target"], reduction="sum") * (1.0 * N / len(mine["target"]))
                     + 0.1 * att + 0.5 * mine["temporal_loss"])
        mine_loss.backward()
    assert torch.equal(mine_loss.detach(), loss.detach())
    n = 0
    for k, p in ref.named_parameters():
        assert p.grad is not None and torch.allclose(sdg[k].grad, p.grad, rtol=0, atol=1e-6 * float(p.grad.abs().max()) + 1e-12), k
        n += 1
    assert n == 70


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not present on this host")
def test_oracle_dropout_sites_match_live_reference(monkeypatch):
    """Train mode with dropout 0.1: the oracle's dropout hooks sit exactly where the reference's nn.Dropout / F.dropout calls
    do.  Both sides get the SAME mask for the k-th dropout call (a function of k and the tensor's shape): identical placement
    and order -> bit-identical outputs and gradients."""
    import contextlib
    import io
    import torch.nn.functional as F
    from oracle.make_golden import load_reference
    ref_stage = load_reference()
    opt = O.default_opt(dropout=0.1, use_sup_att=False)
    torch.manual_seed(2018)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_stage.STAGE(opt)
    ref.train()
    N = 3
    batch = O.make_synthetic_batch(N, 12, 9, 10, 8, seed=4)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    calls = {"n": 0}

    def kth_mask(x, p):
        calls["n"] += 1
        g = torch.Generator().manual_seed(1000 + calls["n"])
        keep = (torch.rand(x.shape, generator=g) >= p).to(x.dtype)
        return x * keep / (1.0 - p)

    def fake_dropout(input, p=0.5, training=True, inplace=False):
        return kth_mask(input, p) if training and p > 0 else input

    monkeypatch.setattr(F, "dropout", fake_dropout)
    (out, target), _, _, t_loss, _ = ref(batch)
    loss = F.cross_entropy(out, target, reduction="sum") + 0.5 * t_loss
    loss.backward()
    n_ref_calls = calls["n"]

    calls["n"] = 0
    shapes = []
    monkeypatch.setattr(O, "DROP", lambda tag, x: (shapes.append(tag), kth_mask(x, 0.1))[1])
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(".pe") else v.clone()) for k, v in sd.items()}
    with torch.enable_grad():
        mine = O.stage_forward.__wrapped__(sdg, opt, batch, training=True)
        mine_loss = F.cross_entropy(mine["answer"], mine["target"], reduction="sum") + 0.5 * mine["temporal_loss"]
        mine_loss.backward()
    assert calls["n"] == n_ref_calls == len(shapes) == 21, (calls["n"], n_ref_calls, shapes)
    assert torch.equal(mine["answer"], out.detach()) and torch.equal(mine_loss.detach(), loss.detach())
    for k, p in ref.named_parameters():
        assert torch.allclose(sdg[k].grad, p.grad, rtol=0, atol=1e-6 * float(p.grad.abs().max()) + 1e-12), k
