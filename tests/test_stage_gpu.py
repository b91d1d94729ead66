"""End-to-end parity of the drop-in STAGE (CUDA kernels through the C ABI) with the reference.

Gate (BASELINE.json north_star / BASELINE.md §4.5): per output tensor max|x - ref| <= 1e-3 * max|ref|, masked (-1e10)
entries exact, answer argmax and span (st, ed) indices bit-exact.  References: the committed golden vectors produced
by the live reference, and the CPU oracle on the same seeded inputs.
"""
import numpy as np
import pytest
import torch

from oracle import stage_oracle as O
from tests import helpers as H
from tvqaplus_b200 import _lib


def run_product(model, batch, inference_mode=False):
    dev = torch.device("cuda:0")
    model = model.to(dev).eval()
    model.inference_mode = inference_mode
    out = model(H.batch_to(batch, dev))
    torch.cuda.synchronize()
    _lib.check_device_error()
    return model, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", H.CASES)
def test_stage_forward_matches_golden(name):
    gold = H.load_golden(name)
    opt, model, batch = H.seeded_model_and_batch(name, gold)
    if H.digest(model.state_dict().values()) != str(gold["weights_sha"]) or H.batch_digest(batch) != str(gold["batch_sha"]):
        pytest.skip("torch RNG stream on this host differs from the fixture's")
    model, (answer, att_loss, att_pred, t_loss, t_probs) = run_product(model, batch)
    errs = dict(answer=H.max_rel(answer.cpu(), gold["answer"]), t_probs=H.max_rel(t_probs.cpu(), gold["t_probs"]),
                t_loss=abs(t_loss.item() - float(gold["temporal_loss"])) / abs(float(gold["temporal_loss"])))
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < H.REL_TOL for v in errs.values()), errs
    assert np.array_equal(answer.argmax(1).cpu().numpy(), gold["answer_argmax"])
    assert np.array_equal(model._last_spans.cpu().numpy().astype(np.int64), gold["spans"])
    assert att_loss == 0 and att_pred is None
    # frames outside the clip get exactly zero probability (model/stage.py:521 + softmax)
    dead = torch.as_tensor(gold["t_scores"]) <= -1e9
    assert torch.equal(t_probs.cpu()[dead], torch.zeros(int(dead.sum())))


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,R,Lq,Ls,heads,dense", [(2, 40, 25, 40, 50, 0, False), (3, 17, 20, 100, 100, 0, False),
                                                     (2, 12, 40, 33, 100, 0, True), (1, 5, 1, 1, 1, 0, False)])
def test_stage_forward_matches_oracle(N, T, R, Lq, Ls, heads, dense):
    opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads)
    from tvqaplus_b200.model import STAGE
    torch.manual_seed(11)
    model = STAGE(opt).eval()
    # perturb LayerNorm / bias parameters so gamma != 1, beta != 0 paths are exercised
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if p_.dim() == 1:
                p_.add_(torch.randn(p_.shape, generator=g) * 0.1)
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=3, dense=dense)
    want = O.stage_forward({k: v.clone() for k, v in model.state_dict().items()}, opt, batch, keep=True)
    model, out = run_product(model, batch, inference_mode=True)
    # inference-mode contract: dict(answer, t_scores = softmax, att_predictions)
    assert set(out) == {"answer", "t_scores", "att_predictions"}
    errs = dict(answer=H.max_rel(out["answer"].cpu(), want["answer"]), t_probs=H.max_rel(out["t_scores"].cpu(), want["t_probs"]))
    print((N, T, R, Lq, Ls, heads), {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < H.REL_TOL for v in errs.values()), errs
    assert torch.equal(out["answer"].argmax(1).cpu(), want["answer"].argmax(1))
    assert torch.equal(model._last_spans.cpu().long(), want["spans"])


@pytest.mark.gpu
def test_module_level_drop_ins():
    """The sub-modules on their own keep the reference's forward signatures and numerics."""
    from tvqaplus_b200.model import (ConvLinear, DepthwiseSeparableConv, EncoderBlock, LinearWrapper, StackedEncoder,
                                     mask_logits)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(6, 30, 128)
    mask = (torch.arange(30)[None] < torch.tensor([30, 1, 17, 0, 29, 8])[:, None]).float()
    enc = StackedEncoder(n_blocks=2, n_conv=2, kernel_size=7, hidden_size=128, dropout=0.1, num_heads=0).eval()
    want = O.stacked_encoder(x, mask, {"e." + k: v for k, v in enc.state_dict().items()}, "e", 2, 2, 7, 0)
    got = enc.to(dev)(x.to(dev), mask.to(dev)).cpu()
    assert H.max_rel(got, want) < 1e-4
    lw = LinearWrapper(128, 128, layer_norm=True, dropout=0.1, relu=True).eval()
    want = O.linear_wrapper(x, {"l." + k: v for k, v in lw.state_dict().items()}, "l", relu=True)
    assert H.max_rel(lw.to(dev)(x.to(dev)).cpu(), want) < 1e-4
    cl = ConvLinear(128, 128, kernel_size=3).eval()
    want = O.conv_linear(x, {"c." + k: v for k, v in cl.state_dict().items()}, "c", 3)
    assert H.max_rel(cl.to(dev)(x.to(dev)).cpu(), want) < 1e-4
    dws = DepthwiseSeparableConv(128, 128, 5, relu=False).eval()
    want = O.depthwise_separable_conv(x, {"d." + k: v for k, v in dws.state_dict().items()}, "d", 5, relu=False)
    assert H.max_rel(dws.to(dev)(x.to(dev)).cpu(), want) < 1e-4
    m = (torch.rand(6, 30, 1) > 0.5).float()
    assert torch.equal(mask_logits(x.to(dev), m.to(dev)).cpu(), O.mask_logits(x, m))
    torch.cuda.synchronize()
    _lib.check_device_error()


@pytest.mark.gpu
@pytest.mark.parametrize("dense", [True, False])
def test_full_size_shape_against_oracle(dense):
    """BASELINE configs[1] shapes (T=300, R=20, Lq=Ls=100) for one example against the CPU oracle."""
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt()
    torch.manual_seed(2018)
    model = STAGE(opt).eval()
    batch = O.make_synthetic_batch(1, 300, 20, 100, 100, seed=7, dense=dense)
    want = O.stage_forward({k: v.clone() for k, v in model.state_dict().items()}, opt, batch, keep=True)
    model, (answer, _, _, t_loss, t_probs) = run_product(model, batch)
    errs = dict(answer=H.max_rel(answer.cpu(), want["answer"]), t_probs=H.max_rel(t_probs.cpu(), want["t_probs"]),
                t_loss=abs(t_loss.item() - want["temporal_loss"].item()) / abs(want["temporal_loss"].item()))
    print("full size dense=%s" % dense, {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < H.REL_TOL for v in errs.values()), errs
    assert torch.equal(answer.argmax(1).cpu(), want["answer"].argmax(1))
    assert torch.equal(model._last_spans.cpu().long(), want["spans"])


@pytest.mark.gpu
def test_full_size_batch_properties():
    """Size-independent properties at the BASELINE shapes (batch 6): examples are independent (a batch gives bit-identical
    results to its shards — the data-parallel contract), span probabilities are distributions over the valid frames, spans
    are ordered and inside the clip."""
    from tvqaplus_b200 import host
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt()
    torch.manual_seed(2018)
    dev = torch.device("cuda:0")
    model = STAGE(opt).to(dev).eval()
    batch = O.make_synthetic_batch(6, 300, 20, 100, 100, seed=11, dense=False)
    full = model(H.batch_to(batch, dev))
    spans_full = model._last_spans.clone()
    parts, spans = [], []
    for r in range(3):
        sh = host.shard_batch(batch, 3, r)
        parts.append(model(H.batch_to(sh, dev)))
        spans.append(model._last_spans.clone())
    torch.cuda.synchronize()
    _lib.check_device_error()
    assert torch.equal(torch.cat([p[0] for p in parts]), full[0]), "answers differ between a batch and its shards"
    assert torch.equal(torch.cat([p[4] for p in parts]), full[4]), "span probabilities differ between a batch and its shards"
    assert torch.equal(torch.cat(spans), spans_full)
    probs = full[4].cpu()  # (N, 5, T, 2)
    ts = batch.ts_label_mask
    assert torch.allclose(probs.sum(2), torch.ones(6, 5, 2), atol=1e-5)
    assert torch.equal(probs[(ts == 0)[:, None, :, None].expand_as(probs)], torch.zeros(int((ts == 0).sum()) * 10))
    sp = spans_full.cpu().view(6, 5, 2).long()
    n_frames = ts.sum(1).long()
    assert (sp[..., 0] <= sp[..., 1]).all() and (sp[..., 0] >= 0).all() and (sp[..., 1] < n_frames[:, None]).all()
    assert torch.isfinite(full[0]).all()


@pytest.mark.gpu
def test_long_context_shape_against_oracle():
    """BASELINE configs[4] shapes (T=600 frames, R=40 regions, Lq=Ls=100), one example, ragged masks."""
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt()
    torch.manual_seed(2018)
    model = STAGE(opt).eval()
    batch = O.make_synthetic_batch(1, 600, 40, 100, 100, seed=5, dense=False)
    want = O.stage_forward({k: v.clone() for k, v in model.state_dict().items()}, opt, batch, keep=True)
    model, (answer, _, _, t_loss, t_probs) = run_product(model, batch)
    errs = dict(answer=H.max_rel(answer.cpu(), want["answer"]), t_probs=H.max_rel(t_probs.cpu(), want["t_probs"]))
    print("long context", {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < H.REL_TOL for v in errs.values()), errs
    assert torch.equal(answer.argmax(1).cpu(), want["answer"].argmax(1))
    assert torch.equal(model._last_spans.cpu().long(), want["spans"])
