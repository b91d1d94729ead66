"""One training step against the golden fixture the LIVE reference produced (oracle/make_golden_train.py ->
tests/golden/train_step_h0.npz: loss terms and every parameter's gradient after main.py:53-61's loss.backward()).
CPU: the oracle's autograd reproduces the fixture (so the oracle-based gradient tests stand on the reference's numbers);
GPU: the CUDA backward reproduces it."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stage_oracle as O
from oracle.make_golden_train import CASE, STRIDE, WHOLE, make_inputs, perturb_
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step_h0.npz")


def _setup():
    from tvqaplus_b200.model import STAGE
    gold = np.load(GOLD, allow_pickle=False)
    opt = O.default_opt(dropout=0.0, use_sup_att=True)
    torch.manual_seed(CASE["weight_seed"])
    model = STAGE(opt)
    perturb_(model.named_parameters(), CASE["perturb_seed"])
    batch = make_inputs()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if H.digest(sd.values()) != str(gold["weights_sha"]) or H.batch_digest(batch) != str(gold["batch_sha"]):
        pytest.skip("torch RNG stream on this host differs from the fixture's")
    return gold, opt, model, sd, batch


def _sample(g):
    flat = g.detach().cpu().reshape(-1)
    return (flat if flat.numel() <= WHOLE else flat[::STRIDE]).numpy()


def _compare(gold, grads, tol):
    gscale = max(float(gold["s/" + str(n)][2]) for n in gold["names"])
    worst = (0.0, None)
    for n in gold["names"]:
        n = str(n)
        want, got = gold["g/" + n], _sample(grads[n])
        assert want.shape == got.shape, n
        den = max(float(gold["s/" + n][2]), 1e-3 * gscale)  # the floor: see tests/test_train_gpu.py (zero-gradient biases)
        err = float(np.abs(got - want).max()) / den
        nrm = float(grads[n].detach().double().norm())
        assert abs(nrm - float(gold["s/" + n][1])) <= tol * max(float(gold["s/" + n][1]), 1e-3 * gscale) * 10, n  # the whole tensor, not only the sample
        if err > worst[0]:
            worst = (err, n)
    assert worst[0] <= tol, "gradient of %s: max-norm relative error %.3e > %.1e" % (worst[1], worst[0], tol)
    return worst


def test_oracle_autograd_reproduces_the_reference_training_step():
    gold, opt, _, sd, batch = _setup()
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(".pe") else v.clone()) for k, v in sd.items()}
    with torch.enable_grad():
        out = O.stage_forward.__wrapped__(sdg, opt, batch, training=True)
        torch.manual_seed(CASE["att_seed"])
        att = O.get_att_loss(out["vid_raw_s"], batch.att_labels, batch.target, batch.anno_st_idx, opt.num_negatives,
                             opt.att_loss_type, opt.alpha, opt.margin)
        cls = F.cross_entropy(out["answer"], out["target"], reduction="sum") * (1.0 * CASE["N"] / len(out["target"]))
        loss = cls + 0.1 * att + 0.5 * out["temporal_loss"]
        loss.backward()
    assert np.float32(loss.item()) == gold["loss"] and np.float32(att.item()) == gold["att_loss"]
    assert np.array_equal(out["target"].numpy(), gold["target"])
    _compare(gold, {k: v.grad for k, v in sdg.items() if v.requires_grad}, 1e-6)


@pytest.mark.gpu
def test_cuda_backward_reproduces_the_reference_training_step():
    gold, opt, model, _, batch = _setup()
    dev = torch.device("cuda:0")
    model = model.to(dev).train()
    torch.manual_seed(CASE["att_seed"])
    (out, target), att_loss, _, t_loss, _ = model(H.batch_to(batch, dev))
    cls = F.cross_entropy(out, target, reduction="sum") * (1.0 * CASE["N"] / len(target))
    loss = cls + 0.1 * att_loss + 0.5 * t_loss
    loss.backward()
    assert np.array_equal(target.cpu().numpy(), gold["target"])
    for got, key, tol in ((loss, "loss", 1e-4), (cls, "cls_loss", 5e-4), (att_loss, "att_loss", 5e-4), (t_loss, "temporal_loss", 5e-4)):
        assert abs(got.item() - float(gold[key])) <= tol * abs(float(gold[key])), (key, got.item(), float(gold[key]))
    worst = _compare(gold, {k: p.grad for k, p in model.named_parameters()}, 2e-4)
    print("worst gradient error against the reference's fixture %.2e (%s)" % worst)
