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
