"""Parity cases the reference's behaviour pins but random dense inputs never reach ("parity hardening"):

* quirk (i): values in PADDED token / region slots reach the outputs (the reference never masks before the convs,
  model/encoder.py:37-44) — garbage in the pads must move our outputs exactly as it moves the reference's;
* exact ties in the span search (model/model_utils.py:102-121);
* heads = 4 at the full configs[1] shape;
* softmax scales other than 10 in the structured attention (model/context_query_attention.py:59-61);
* the inference-mode post-processing (model/stage.py:748-806, inference.py:13-35).
"""
import numpy as np
import pytest
import torch

from oracle import stage_oracle as O
from tests import helpers as H
from tests.test_stage_gpu import run_product
from tvqaplus_b200 import _lib, ops


def _model(heads=0, seed=2018, **over):
    from tvqaplus_b200.model import STAGE
    opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads, **over)
    torch.manual_seed(seed)
    return opt, STAGE(opt).eval()


def _check(model, opt, batch, tag):
    want = O.stage_forward({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, opt, batch, keep=True)
    model, (answer, _, _, t_loss, t_probs) = run_product(model, batch)
    errs = dict(answer=H.max_rel(answer.cpu(), want["answer"]), t_probs=H.max_rel(t_probs.cpu(), want["t_probs"]),
                t_loss=abs(t_loss.item() - want["temporal_loss"].item()) / abs(want["temporal_loss"].item()))
    print(tag, {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < H.REL_TOL for v in errs.values()), errs
    assert torch.equal(answer.argmax(1).cpu(), want["answer"].argmax(1))
    assert torch.equal(model._last_spans.cpu().long(), want["spans"])
    return want, answer.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,R,Lq,Ls", [(2, 12, 20, 24, 30), (1, 40, 20, 100, 100)])
def test_garbage_in_padded_slots_moves_outputs_like_the_reference(N, T, R, Lq, Ls):
    """SURVEY.md §7 quirk (i).  Same masks, pads zero-filled vs pads filled with N(0,1) garbage: the oracle's answers move
    (the effect is real), and the CUDA path follows the oracle on both."""
    opt, model = _model()
    clean = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=21)
    dirty = O.Batch(clean)
    g = torch.Generator().manual_seed(99)
    for key, mkey in (("qas_bert", "qas_mask"), ("sub_bert", "sub_mask"), ("vid", "vid_mask")):
        x, m = clean[key], clean[mkey]
        assert float((1 - m).sum()) > 0, "case has no padding in %s" % key
        dirty[key] = (x + (1 - m)[..., None] * torch.randn(x.shape, generator=g)).contiguous()
    w_clean, a_clean = _check(model, opt, clean, "pads zero")
    w_dirty, a_dirty = _check(model, opt, dirty, "pads garbage")
    moved = float((w_dirty["answer"] - w_clean["answer"]).abs().max())
    assert moved > 1e-3, "the reference is expected to depend on pad values (got %.2e)" % moved
    ours = float((a_dirty - a_clean).abs().max())
    assert abs(ours - moved) < 1e-3 * max(1.0, moved)


def _is_maximiser(p1, p2, st, ed, conf):
    """(st, ed) is a valid answer of find_max_triples(topN=1): st <= ed (or every upper product is <= 0, where the
    reference's argsort also ranks the zeros of the lower triangle) and its fp32 product is the maximum."""
    T = p1.shape[0]
    prod = np.triu(np.outer(p1.astype(np.float32), p2.astype(np.float32)).astype(np.float32))
    return prod[st, ed] == prod.max() and np.float32(conf) == prod.max() and (st <= ed or prod.max() <= 0)


@pytest.mark.gpu
def test_span_search_exact_ties():
    """Ties between fp32 products.  The reference ranks with numpy's default (unstable) argsort and takes the last
    entry, so WHICH of several equal maxima it returns is implementation-defined (checked here: numpy 2.3 returns (0, 0)
    for a uniform T=17 row and (4, 4) for T=5).  The kernels document their rule — the largest flat index st*T + ed among
    the maxima, ``oracle.best_span_scan`` — and are checked against that rule, plus: the returned pair is always one of the
    reference's maximisers with the bit-identical fp32 confidence."""
    cases = []
    for T in (1, 2, 5, 17, 64, 300):
        u = np.full(T, 1.0 / T, dtype=np.float32)
        cases.append((u, u))                                        # every upper product equal
        a = np.zeros(T, dtype=np.float32); a[T - 1] = 1
        b = np.zeros(T, dtype=np.float32); b[0] = 1
        cases.append((a, b))                                        # every upper product is 0 (unless T == 1)
    g = np.random.RandomState(3)
    for T in (8, 33, 300):
        p1 = g.dirichlet(np.ones(T)).astype(np.float32)
        p2 = g.dirichlet(np.ones(T)).astype(np.float32)
        i, k = int(p1.argmax()), int(p2.argmax())
        p1[(i + 3) % T] = p1[i]                                     # duplicated maxima in both factors
        p2[(k + 2) % T] = p2[k]
        cases.append((p1, p2))
    for p1, p2 in cases:
        T = p1.shape[0]
        spans, conf = ops.find_max_span(torch.from_numpy(p1)[None].cuda(), torch.from_numpy(p2)[None].cuda())
        st, ed, c = int(spans[0, 0]), int(spans[0, 1]), float(conf[0])
        w_st, w_ed, w_c = O.best_span_scan(p1, p2)
        assert (st, ed) == (w_st, w_ed) and np.float32(c) == np.float32(w_c), (T, (st, ed, c), (w_st, w_ed, w_c))
        assert _is_maximiser(p1, p2, st, ed, c)
        ref = O.find_max_triples(torch.from_numpy(p1)[None], torch.from_numpy(p2)[None], topN=1)[0][0]
        assert np.float32(ref[2]) == np.float32(c)                  # same confidence as the reference's own pick


@pytest.mark.gpu
def test_head_kernel_span_ties():
    """The same rule inside head_kernel: identical frames give identical logits, hence exactly uniform probabilities."""
    from tests.test_head_gpu import dev_weights, head_weights
    g = torch.Generator().manual_seed(4)
    sd = head_weights(g, True)
    N, T, Lq = 2, 23, 4
    row = torch.randn(N * 5, 1, 128, generator=g)
    ms = row.expand(N * 5, T, 128).contiguous()
    smask = torch.ones(N, 5, T, Lq)
    ts_mask = torch.ones(N, T)
    ts_mask[1, 15:] = 0
    dev = torch.device("cuda:0")
    out = ops.span_head_fwd(ms.to(dev), smask.to(dev), ts_mask.to(dev), dev_weights(sd, dev), add_local=True, extra_span_length=3)
    torch.cuda.synchronize()
    _lib.check_device_error()
    probs = out["t_probs"].cpu().view(N * 5, T, 2).numpy()
    spans = out["spans"].cpu().numpy()
    for r in range(N * 5):
        n_valid = 15 if r >= 5 else T
        assert np.all(probs[r, :n_valid, 0] == probs[r, 0, 0]) and np.all(probs[r, n_valid:] == 0)
        w_st, w_ed, _ = O.best_span_scan(probs[r, :, 0], probs[r, :, 1])
        assert (int(spans[r, 0]), int(spans[r, 1])) == (w_st, w_ed) == (n_valid - 1, n_valid - 1)


@pytest.mark.gpu
def test_heads4_full_size_against_oracle():
    """heads = 4 (self-attention in both encoders) at the configs[1] shape: T=300, R=20, Lq=Ls=100, one example."""
    opt, model = _model(heads=4)
    batch = O.make_synthetic_batch(1, 300, 20, 100, 100, seed=13, dense=False)
    _check(model, opt, batch, "heads=4 full size")


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1.0, 20.0, 30.0, 100.0])
def test_cqa_other_softmax_scales(scale):
    """scale*log2(e) > 15 takes the exact-row-maximum path of the fused kernel; padded answer tokens, empty frames and
    frames without regions (fully masked rows) must stay finite and match (NaN at scale 100, fp16 underflow at
    15 < scale < 35)."""
    from tests.test_cqa_gpu import ALL, make_case, oracle, run_gpu
    case = make_case(2, 6, 40, 50, 25, seed=int(scale) + 7)
    want = oracle(*case, scale=scale)
    got = run_gpu(*case, scale=scale, want=ALL)
    for k in ("statement", "att_sub", "att_vid", "sub_norm_s", "vid_norm_s", "mixed_sub", "mixed_vid"):
        assert torch.isfinite(got[k]).all(), k
        err = H.max_rel(got[k], want[k])
        assert err < 1e-4, (k, scale, err)
    assert torch.equal(got["statement_mask"], want["statement_mask"])


@pytest.mark.gpu
def test_stage_forward_scale100_with_empty_frames():
    opt, model = _model(scale=100.0)
    batch = O.make_synthetic_batch(2, 20, 20, 30, 40, seed=17)
    batch["sub_mask"][0, 3] = 0  # a frame inside the clip without any sub-title token
    batch["vid_mask"][1, 5] = 0  # and one without any region
    _check(model, opt, batch, "scale=100")


# ------------------------------------------------------------------------------------------------ inference post-processing
def _att_inputs(N, T, R, Lq, seed):
    g = np.random.RandomState(seed)
    n_img = [int(g.randint(1, max(2, T // 2))) for _ in range(N)]
    start = [int(g.randint(0, T - n + 1)) for n in n_img]
    boxes = [[[[int(v) for v in g.randint(0, 200, 4)] for _ in range(R)] for _ in range(n)] for n in n_img]
    image_indices = [[int(3 * t + 1) for t in range(T)] for _ in range(N)]
    vocab = [int(v) for v in g.choice(np.arange(1, 15000), 400, replace=False)]
    return start, boxes, image_indices, vocab


@pytest.mark.gpu
def test_inference_mode_att_predictions_and_spans():
    """inference_mode forward with boxes: the dict contract of model/stage.py:297-312, att_predictions against the oracle's
    restatement of get_att_prediction on the oracle's own vid_raw_s, predicted spans against inference.py's find_max_pair."""
    from tvqaplus_b200 import inference
    N, T, R, Lq, Ls = 2, 14, 12, 20, 24
    opt, model = _model()
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=31)
    start, boxes, image_indices, vocab = _att_inputs(N, T, R, Lq, 5)
    words = batch.qas.clone()
    words[:, :, 1::3] = torch.tensor(vocab[:1])[0]  # make sure object words occur
    words = words * batch.qas_mask.long()
    batch.update(qas=words, anno_st_idx=start, boxes=boxes, image_indices=image_indices, eval_object_word_ids=vocab)
    want = O.stage_forward({k: v.clone() for k, v in model.state_dict().items()}, opt, batch, keep=True)
    want_att = O.get_att_prediction(want["vid_raw_s"], vocab, words, batch.vid_name, batch.qid, image_indices, boxes, start)
    model, out = run_product(model, batch, inference_mode=True)
    got_att = out["att_predictions"]
    assert len(got_att) == N and all(set(q) == set(range(5)) for q in got_att)
    n_items = 0
    for gq, wq in zip(got_att, want_att):
        for a in range(5):
            assert len(gq[a]) == len(wq[a])
            for gi, wi in zip(gq[a], wq[a]):
                assert {k: gi[k] for k in ("word", "qid", "vid_name", "img_idx")} == {k: wi[k] for k in ("word", "qid", "vid_name", "img_idx")}
                near = [s for s in wi["pred"] if abs(s - 0.2) < 1e-4]
                if near:
                    continue  # a score within rounding of the threshold may legitimately fall on either side
                assert gi["bbox"] == wi["bbox"]
                assert np.allclose(gi["pred"], wi["pred"], atol=2e-5)
                n_items += 1
    assert n_items > 0
    # span of the predicted answer (inference.py:58-72)
    pred, spans, vals = inference.predicted_spans(out["answer"], out["t_scores"])
    probs = out["t_scores"].cpu()
    for i in range(N):
        a = int(pred[i])
        (st, ed), v = O.find_max_pair(probs[i, a, :, 0].numpy().tolist(), probs[i, a, :, 1].numpy().tolist())
        assert (int(spans[i, 0]), int(spans[i, 1])) == (st, ed) and float(vals[i]) == v


@pytest.mark.gpu
def test_find_max_pair_matches_inference_py():
    from tvqaplus_b200 import inference
    g = torch.Generator().manual_seed(8)
    rows = []
    for T in (1, 2, 7, 300):
        p = torch.softmax(torch.randn(5, T, 2, generator=g) * 3, dim=1)
        p[1] = 0                      # every product 0 -> the reference keeps its initial (0, 1), 0.0
        p[2, :, 0] = 1.0 / T          # ties: the first maximum wins
        spans, vals = ops.find_max_pairs(p.cuda())
        for r in range(5):
            (st, ed), v = O.find_max_pair(p[r, :, 0].tolist(), p[r, :, 1].tolist())
            assert (int(spans[r, 0]), int(spans[r, 1])) == (st, ed) and float(vals[r]) == v, (T, r)
        rows.append(p)
    (st, ed), v = inference.find_max_pair(rows[2][0, :, 0].tolist(), rows[2][0, :, 1].tolist())
    assert ((st, ed), v) == O.find_max_pair(rows[2][0, :, 0].tolist(), rows[2][0, :, 1].tolist())


@pytest.mark.gpu
def test_weight_cache_invalidation_after_data_edit():
    """In-place edits through ``.data`` do not move the version counter the packed-weight caches are keyed on; after
    ``invalidate_weight_cache()`` the forward must see the new values."""
    opt, model = _model()
    batch = O.make_synthetic_batch(1, 6, 5, 8, 7, seed=2)
    model, (a0, *_r) = run_product(model, batch)
    with torch.no_grad():
        model.concat_fc[2].weight.data.mul_(1.5)
        model.cls_encoder.stacked_encoderBlocks[0].conv[0].pointwise_conv.weight.data.mul_(0.5)
    model.invalidate_weight_cache()
    want = O.stage_forward({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, opt, batch)
    model, (a1, *_r) = run_product(model, batch)
    assert H.max_rel(a1.cpu(), want["answer"]) < H.REL_TOL
    assert float((a1 - a0).abs().max()) > 1e-4


@pytest.mark.gpu
def test_device_error_monitor_is_silent_on_clean_runs():
    opt, model = _model()
    batch = O.make_synthetic_batch(1, 6, 5, 8, 7, seed=2)
    model, _ = run_product(model, batch)
    model(H.batch_to(batch, torch.device("cuda:0")))  # polls the flag posted by the previous call
    model.check_device_errors()
    assert model._err_monitor.event is not None and int(model._err_monitor.flag.item()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("sub_flag,vfeat_flag", [(True, False), (False, True)])
def test_single_stream_models(sub_flag, vfeat_flag):
    """sub-only / vfeat-only models (model/stage.py:283-290): statement = that stream's mixed tensor, mask = its own."""
    opt, model = _model(sub_flag=sub_flag, vfeat_flag=vfeat_flag)
    assert ("vid_fc.0.weight" in model.state_dict()) == vfeat_flag and "concat_fc.0.weight" not in model.state_dict()
    batch = O.make_synthetic_batch(2, 14, 12, 20, 24, seed=23)
    _check(model, opt, batch, "single stream sub=%s vfeat=%s" % (sub_flag, vfeat_flag))
