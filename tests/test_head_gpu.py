"""stage_span_head_fwd / stage_find_max_span / stage_mask_logits against the CPU oracle (model/stage.py:439-537)."""
import numpy as np
import pytest
import torch

from oracle import stage_oracle as O
from tvqaplus_b200 import _lib, ops

TOL = 1e-4


def head_weights(g, add_local):
    sd = {}
    for name, n_in, n_out in (("cls_projection_layers.0", 128, 128), ("temporal_scoring_st_layers.0", 128, 1),
                              ("temporal_scoring_ed_layers.0", 128, 1), ("classifier", 256 if add_local else 128, 1)):
        sd[name + ".conv.0.weight"] = torch.rand(n_in, generator=g) + 0.5
        sd[name + ".conv.0.bias"] = torch.randn(n_in, generator=g) * 0.1
        sd[name + ".conv.2.weight"] = torch.randn(n_out, n_in, generator=g) / n_in ** 0.5
        sd[name + ".conv.2.bias"] = torch.randn(n_out, generator=g) * 0.1
    return sd


def dev_weights(sd, dev):
    m = {"proj": "cls_projection_layers.0", "st": "temporal_scoring_st_layers.0", "ed": "temporal_scoring_ed_layers.0",
         "cls": "classifier"}
    w = {}
    for k, name in m.items():
        w[k + "_ln_g"] = sd[name + ".conv.0.weight"].to(dev)
        w[k + "_ln_b"] = sd[name + ".conv.0.bias"].to(dev)
        w[k + "_w"] = sd[name + ".conv.2.weight"].to(dev).contiguous()
        w[k + "_b"] = sd[name + ".conv.2.bias"].to(dev)
    return w


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,Lq,add_local", [(2, 40, 7, True), (3, 300, 5, True), (2, 17, 3, False), (1, 1, 2, True),
                                              (1, 600, 4, True)])
def test_span_head(N, T, Lq, add_local):
    g = torch.Generator().manual_seed(N * 1000 + T)
    sd = head_weights(g, add_local)
    n_frames = torch.randint(max(1, T // 2), T + 1, (N,), generator=g)
    frame_ok = (torch.arange(T)[None] < n_frames[:, None]).float()               # (N,T)
    smask = torch.ones(N, 5, T, Lq) * frame_ok[:, None, :, None]
    if T > 3:
        smask[0, 1, 2] = 0  # a valid frame whose statement is fully masked
    ms = torch.randn(N * 5, T, 128, generator=g)
    msm = (smask.sum(-1) != 0).float()                                          # (N,5,T)
    ms = O.mask_logits(ms, msm.view(N * 5, T, 1))                                # fully masked frames carry -1e10 rows
    # oracle
    enc, t0 = O.residual_temporal_predictor(ms, sd, 0)
    ts = O.mask_logits(t0.view(N, 5, T, 2), frame_ok.view(N, 1, T, 1))
    stacked = enc.view(N, 5, T, 128)
    if add_local:
        feat, spans = O.get_proposals_eval(stacked, msm.view(N, 5, T, 1), ts, 3)
    else:
        feat = torch.max(O.mask_logits(stacked, msm.view(N, 5, T, 1)), 2)[0]
        spans = None
    ans = O.linear_wrapper(feat, sd, "classifier", relu=False).squeeze(2)
    # device
    dev = torch.device("cuda:0")
    out = ops.span_head_fwd(ms.to(dev), smask.to(dev), frame_ok.to(dev), dev_weights(sd, dev), add_local=add_local,
                            extra_span_length=3)
    torch.cuda.synchronize()
    _lib.check_device_error()
    got_ts, got_p, got_ans = out["t_scores"].cpu(), out["t_probs"].cpu(), out["answer"].cpu()
    masked = ts <= -1e9
    assert torch.equal(got_ts <= -1e9, masked)
    assert torch.equal(got_ts[masked], ts[masked])
    e_ts = (got_ts.masked_fill(masked, 0) - ts.masked_fill(masked, 0)).abs().max().item() / ts.masked_fill(masked, 0).abs().max().item()
    e_p = (got_p - torch.softmax(ts, dim=2)).abs().max().item()
    e_a = (got_ans - ans).abs().max().item() / ans.abs().max().item()
    print("head", (N, T, Lq, add_local), "ts %.2e probs %.2e ans %.2e" % (e_ts, e_p, e_a))
    assert e_ts < TOL and e_p < 1e-6 and e_a < TOL
    assert torch.equal(got_p[masked], torch.zeros_like(got_p[masked]))
    if add_local:
        assert torch.equal(out["spans"].cpu().long(), spans), (out["spans"].cpu(), spans)
    assert torch.equal(got_ans.argmax(1), ans.argmax(1))


@pytest.mark.gpu
def test_find_max_span_matches_reference_argsort():
    g = torch.Generator().manual_seed(5)
    for T in (1, 2, 5, 37, 300):
        p1 = torch.softmax(torch.randn(64, T, generator=g) * 3, 1)
        p2 = torch.softmax(torch.randn(64, T, generator=g) * 3, 1)
        want = O.find_max_triples(p1, p2, topN=1)
        spans, conf = ops.find_max_span(p1.cuda(), p2.cuda())
        spans, conf = spans.cpu(), conf.cpu()
        for r in range(64):
            assert (int(spans[r, 0]), int(spans[r, 1])) == (want[r][0][0], want[r][0][1])
            assert np.float32(conf[r].item()) == np.float32(want[r][0][2])


@pytest.mark.gpu
def test_mask_logits():
    g = torch.Generator().manual_seed(6)
    x = torch.randn(7, 33, 128, generator=g)
    m = (torch.rand(7, 33, 1, generator=g) > 0.3).float()
    got = ops.mask_logits_fwd(x.cuda(), m.cuda()).cpu()
    assert torch.equal(got, O.mask_logits(x, m))
