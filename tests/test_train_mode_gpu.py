"""Train-mode forward (dropout = 0, no autograd): the train branch of get_proposals with its data-dependent N_new rows
(model/stage.py:407-438), the train return tuple (:345-348) and get_ts_loss fused into the head launch (:539-555),
against the golden vectors the live reference produced in train mode and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import stage_oracle as O
from tests import helpers as H
from tvqaplus_b200 import _lib


def _train_forward(opt, batch, seed=2018, perturb=False):
    from tvqaplus_b200.model import STAGE
    torch.manual_seed(seed)
    model = STAGE(opt)
    if perturb:
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for _, p_ in model.named_parameters():
                if p_.dim() == 1:
                    p_.add_(torch.randn(p_.shape, generator=g) * 0.1)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to("cuda:0").train()
    (out, target), att_loss, att_pred, t_loss, t_scores = model(H.batch_to(batch, torch.device("cuda:0")))
    torch.cuda.synchronize()
    _lib.check_device_error()
    return sd, out.cpu(), target.cpu(), t_loss, t_scores.cpu(), att_loss, att_pred


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_h0", "cfg1_h0", "wide_h0"])
def test_train_forward_matches_reference_golden(name):
    gold = H.load_golden(name)
    N, T, R, Lq, Ls, heads = [int(v) for v in gold["shape"]]
    opt = O.default_opt(dropout=0.0, use_sup_att=False)
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=2018)
    sd, out, target, t_loss, t_scores, att_loss, att_pred = _train_forward(opt, batch)
    if H.digest(sd.values()) != str(gold["weights_sha"]) or H.batch_digest(batch) != str(gold["batch_sha"]):
        pytest.skip("torch RNG stream on this host differs from the fixture's")
    assert out.shape == gold["train_answer"].shape, (out.shape, gold["train_answer"].shape)  # N_new rows
    assert np.array_equal(target.numpy(), gold["train_target"])
    assert H.max_rel(out, gold["train_answer"]) < H.REL_TOL
    assert abs(t_loss.item() - float(gold["train_temporal_loss"])) / abs(float(gold["train_temporal_loss"])) < H.REL_TOL
    assert H.masked_rel(t_scores, gold["t_scores"]) < H.REL_TOL  # train mode returns the raw masked logits
    assert att_loss == 0 and att_pred is None


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,R,Lq,Ls,seed", [(6, 30, 10, 12, 16, 3), (4, 9, 5, 7, 8, 4), (8, 60, 20, 30, 30, 6)])
def test_train_forward_matches_oracle(N, T, R, Lq, Ls, seed):
    """More examples than the fixtures hold, so that both kinds of rows occur (ground-truth span only / plus the predicted one)."""
    opt = O.default_opt(dropout=0.0, use_sup_att=False)
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=seed)
    sd, out, target, t_loss, t_scores, _, _ = _train_forward(opt, batch, seed=11, perturb=True)
    want = O.stage_forward(sd, opt, batch, training=True)
    assert out.shape == want["answer"].shape and torch.equal(target, want["target"])
    assert H.max_rel(out, want["answer"]) < H.REL_TOL
    assert abs(t_loss.item() - want["temporal_loss"].item()) / abs(want["temporal_loss"].item()) < H.REL_TOL
    print("train rows", out.shape[0], "of", N, "examples")


@pytest.mark.gpu
def test_fused_temporal_loss_equals_the_module_level_formula():
    """get_ts_loss rides in the head launch; STAGE.get_ts_loss (module-level, two F.cross_entropy calls) must agree."""
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt()
    torch.manual_seed(3)
    model = STAGE(opt).to("cuda:0").eval()
    batch = H.batch_to(O.make_synthetic_batch(5, 40, 10, 12, 16, seed=8), torch.device("cuda:0"))
    out, _, _, t_loss, t_probs, other = model.forward_main(batch)
    ref = model.get_ts_loss(other["temporal_scores"], batch.ts_label, batch.target)
    assert abs(t_loss.item() - ref.item()) <= 2e-6 * abs(ref.item())


@pytest.mark.gpu
@pytest.mark.parametrize("loss_type", ["lse", "hinge"])
def test_att_loss_on_device_matches_oracle(loss_type):
    """get_att_loss (model/stage.py:609-746): same labels, same host RNG seed -> the same sampled pairs; the device gather +
    loss against the oracle restatement (itself bit-identical to the live reference, tests/test_oracle.py)."""
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt(att_loss_type=loss_type, dropout=0.0)
    torch.manual_seed(2)
    model = STAGE(opt).to("cuda:0").train()
    batch = O.make_synthetic_batch(4, 14, 9, 12, 10, seed=4)
    labels, starts = O.make_att_labels(batch, seed=1)
    g = torch.Generator().manual_seed(3)
    scores = torch.rand(4, 5, 14, 12, 9, generator=g) * 2 - 1
    scores[:, :, :, :, -1] = -1e10  # a masked region: sampled as a negative it contributes exactly 0
    torch.manual_seed(7)
    want = O.get_att_loss(scores, labels, batch.target, starts, 2, loss_type, opt.alpha, opt.margin)
    torch.manual_seed(7)
    got, preds = model.get_att_loss(scores.cuda(), labels, batch.target, batch.qas, batch.vid_name, batch.qid, batch.q_l, None, None,
                                    starts, num_negatives=2)
    assert preds is None  # train mode
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item()), (got.item(), want.item())
    # eval mode additionally returns the detection lists (one gather instead of one copy per labelled word)
    model.eval()
    torch.manual_seed(7)
    _, preds = model.get_att_loss(scores.cuda(), labels, batch.target, batch.qas, batch.vid_name, batch.qid, batch.q_l,
                                  [list(range(14))] * 4, None, starts, num_negatives=2)
    torch.manual_seed(7)
    pos, _ = O.att_loss_pairs(labels, batch.target, starts, 2)
    n_unique = len(np.unique(pos.numpy(), axis=0))
    assert len(preds["det_q"]) + len(preds["det_ca"]) == n_unique
    d = (preds["det_q"] + preds["det_ca"])[0]
    assert d["pred"].dim() == 1 and d["vid_name"].startswith("syn_")


@pytest.mark.gpu
def test_train_forward_with_supervised_attention():
    """The whole train-mode forward with use_sup_att: att_loss from the kernel's own vid_raw_s against the oracle's."""
    opt = O.default_opt(dropout=0.0, use_sup_att=True)
    batch = O.make_synthetic_batch(3, 16, 9, 12, 10, seed=9)
    batch["att_labels"], batch["anno_st_idx"] = O.make_att_labels(batch, seed=2)
    from tvqaplus_b200.model import STAGE
    torch.manual_seed(2018)
    model = STAGE(opt)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    want = O.stage_forward(sd, opt, batch, training=True, keep=True)
    torch.manual_seed(5)
    want_att = O.get_att_loss(want["vid_raw_s"], batch.att_labels, batch.target, batch.anno_st_idx, opt.num_negatives,
                              opt.att_loss_type, opt.alpha, opt.margin)
    model = model.to("cuda:0").train()
    torch.manual_seed(5)
    (out, target), att_loss, att_pred, t_loss, _ = model(H.batch_to(batch, torch.device("cuda:0")))
    assert abs(att_loss.item() - want_att.item()) <= 1e-3 * abs(want_att.item()), (att_loss.item(), want_att.item())
    assert H.max_rel(out.cpu(), want["answer"]) < H.REL_TOL and torch.equal(target.cpu(), want["target"])
