"""stage_cqa_fused_fwd against the CPU oracle (StructuredAttention + c2q_down_projection + concat_fc)."""
import pytest
import torch

from oracle import stage_oracle as O
from tvqaplus_b200 import _lib, ops

TOL = 1e-4


def make_case(N, T, Lq, Ls, R, seed, dense=False):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(N, 5, Lq, 128, generator=g)
    s = torch.randn(N, T, Ls, 128, generator=g)
    v = torch.randn(N, T, R, 128, generator=g)
    if dense:
        am, sm, vm = torch.ones(N, 5, Lq), torch.ones(N, T, Ls), torch.ones(N, T, R)
    else:
        am = (torch.arange(Lq)[None, None] < torch.randint(1, Lq + 1, (N, 5, 1), generator=g)).float()
        sm = (torch.arange(Ls)[None, None] < torch.randint(0, Ls + 1, (N, T, 1), generator=g)).float()
        vm = (torch.arange(R)[None, None] < torch.randint(0, R + 1, (N, T, 1), generator=g)).float()
        if T > 1:
            sm[:, -1] = 0  # a fully padded frame
            vm[:, -1] = 0
    sd = {}
    for name in ("c2q_down_projection", "concat_fc"):
        sd[name + ".0.weight"] = torch.rand(384, generator=g) + 0.5
        sd[name + ".0.bias"] = torch.randn(384, generator=g) * 0.1
        sd[name + ".2.weight"] = torch.randn(128, 384, generator=g) / 384 ** 0.5
        sd[name + ".2.bias"] = torch.randn(128, generator=g) * 0.1
    sd["concat_fc.4.weight"] = torch.rand(128, generator=g) + 0.5
    sd["concat_fc.4.bias"] = torch.randn(128, generator=g) * 0.1
    return a, am, s, sm, v, vm, sd


def oracle(a, am, s, sm, v, vm, sd, scale):
    N, _, Lq, _ = a.shape
    T, Ls = s.shape[1:3]
    R = v.shape[2]
    opt = O.default_opt(scale=scale)
    ae, amk = a.view(N, 5, 1, Lq, 128), am.view(N, 5, 1, Lq)
    ms, msm, sraw, snorm = O.qa_ctx_attention(ae, s.view(N, 1, T, Ls, 128), amk, sm.view(N, 1, T, Ls), sd, opt)
    mv, mvm, vraw, vnorm = O.qa_ctx_attention(ae, v.view(N, 1, T, R, 128), amk, vm.view(N, 1, T, R), sd, opt)
    att_s = O.structured_attention(ae, s.view(N, 1, T, Ls, 128), amk, sm.view(N, 1, T, Ls), scale)[0]
    att_v = O.structured_attention(ae, v.view(N, 1, T, R, 128), amk, vm.view(N, 1, T, R), scale)[0]
    st = O.concat_fc(ms, mv, sd)
    return dict(statement=st, statement_mask=mvm, sub_raw_s=sraw, vid_raw_s=vraw, sub_norm_s=snorm, vid_norm_s=vnorm,
                att_sub=att_s, att_vid=att_v, mixed_sub=ms, mixed_vid=mv)


def run_gpu(a, am, s, sm, v, vm, sd, scale, want):
    dev = torch.device("cuda:0")
    d = {k: v.to(dev) for k, v in sd.items()}
    w = dict(wblob=ops.cqa_prepare_weights(d["c2q_down_projection.0.weight"], d["c2q_down_projection.0.bias"],
                                           d["c2q_down_projection.2.weight"], d["c2q_down_projection.2.bias"],
                                           d["concat_fc.0.weight"], d["concat_fc.0.bias"], d["concat_fc.2.weight"],
                                           d["concat_fc.2.bias"]),
             cat_out_ln_g=d["concat_fc.4.weight"], cat_out_ln_b=d["concat_fc.4.bias"])
    out = ops.cqa_fused_fwd(a.to(dev), am.to(dev), s.to(dev), sm.to(dev), v.to(dev), vm.to(dev), w, scale, want=want)
    torch.cuda.synchronize()
    _lib.check_device_error()
    return {k: t.cpu() for k, t in out.items()}


ALL = ("sub_raw_s", "vid_raw_s", "sub_norm_s", "vid_norm_s", "att_sub", "att_vid", "mixed_sub", "mixed_vid")


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,Lq,Ls,R,dense", [(1, 2, 16, 16, 16, True), (2, 5, 40, 50, 25, False),
                                               (1, 200, 100, 100, 20, False), (1, 3, 128, 112, 48, False),
                                               (2, 4, 7, 3, 1, False), (1, 160, 100, 100, 40, True)])
def test_cqa_fused(N, T, Lq, Ls, R, dense):
    case = make_case(N, T, Lq, Ls, R, seed=N * 100 + T, dense=dense)
    want = oracle(*case, scale=10.0)
    got = run_gpu(*case, scale=10.0, want=ALL)
    errs = {}
    for k in ("sub_raw_s", "vid_raw_s", "sub_norm_s", "vid_norm_s", "att_sub", "att_vid", "mixed_sub", "mixed_vid",
              "statement", "statement_mask"):
        w_, g_ = want[k], got[k]
        assert g_.shape == w_.shape, (k, g_.shape, w_.shape)
        if k.endswith("raw_s"):
            masked = w_ <= -1e9
            assert torch.equal(g_ <= -1e9, masked), k
            assert torch.equal(g_[masked], w_[masked]), k  # exactly -1e10
            w_, g_ = w_.masked_fill(masked, 0), g_.masked_fill(masked, 0)
        errs[k] = (g_ - w_).abs().max().item() / max(1e-9, w_.abs().max().item())
    print("cqa", (N, T, Lq, Ls, R), {k: "%.2e" % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    # statement written without the optional outputs must be identical
    got2 = run_gpu(*case, scale=10.0, want=())
    assert torch.equal(got2["statement"], got["statement"])


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,Lq,Ls,R,k", [(2, 5, 40, 100, 20, 7), (1, 3, 100, 70, 9, 7), (2, 4, 24, 112, 30, 5)])
def test_encoder_writes_cqa_context_blobs(N, T, Lq, Ls, R, k):
    """Internal fast path of STAGE.forward: the sub-title input encoder writes its closing-LayerNorm output straight into
    the context blobs of the CQA workspace (stage_encoder_block_fwd out_mode 2 + stage_cqa_fused_fwd sub_prepared).  The
    result must match the two-step path (encoder -> fp32 -> preparation pass) to rounding of the inverse norms."""
    from tests.test_encoder_gpu import random_block
    dev = torch.device("cuda:0")
    a, am, s_raw, sm, v, vm, sd = make_case(N, T, Lq, Ls, R, seed=N * 100 + Ls)
    g = torch.Generator().manual_seed(5)
    esd = {kk: vv.to(dev) for kk, vv in random_block("enc", 2, k, g).items()}
    convs = [dict(ln_g=esd["enc.layer_norm.%d.weight" % i], ln_b=esd["enc.layer_norm.%d.bias" % i],
                  dw_w=esd["enc.conv.%d.depthwise_conv.weight" % i], dw_b=esd["enc.conv.%d.depthwise_conv.bias" % i],
                  pw_packed=ops.pack_weight(esd["enc.conv.%d.pointwise_conv.weight" % i], 64),
                  pw_b=esd["enc.conv.%d.pointwise_conv.bias" % i]) for i in range(2)]
    fin = (esd["enc.final_layer_norm.weight"], esd["enc.final_layer_norm.bias"])
    pe = esd["enc.position_encoding.pe"]
    d = {kk: vv.to(dev) for kk, vv in sd.items()}
    w = dict(wblob=ops.cqa_prepare_weights(d["c2q_down_projection.0.weight"], d["c2q_down_projection.0.bias"],
                                           d["c2q_down_projection.2.weight"], d["c2q_down_projection.2.bias"],
                                           d["concat_fc.0.weight"], d["concat_fc.0.bias"], d["concat_fc.2.weight"],
                                           d["concat_fc.2.bias"]),
             cat_out_ln_g=d["concat_fc.4.weight"], cat_out_ln_b=d["concat_fc.4.bias"])
    x = (s_raw * sm[..., None]).view(N * T, Ls, 128).to(dev)
    smd = sm.to(dev)
    want_keys = ("sub_raw_s", "att_sub")
    # two-step path
    sub = ops.encoder_block_fwd(x, smd.view(N * T, Ls), pe, convs, fin, k).view(N, T, Ls, 128)
    ref = ops.cqa_fused_fwd(a.to(dev), am.to(dev), sub, smd, v.to(dev), vm.to(dev), w, 10.0, want=want_keys)
    # fused path
    ws = ops.CqaWorkspace(N, T, Lq, dev)
    assert ops.encoder_block_fwd(x, smd.view(N * T, Ls), pe, convs, fin, k, cqa_workspace=ws) is None
    assert ws.sub_prepared and ws.sub_len == Ls
    got = ops.cqa_fused_fwd(a.to(dev), am.to(dev), None, smd, v.to(dev), vm.to(dev), w, 10.0, want=want_keys, workspace=ws)
    torch.cuda.synchronize()
    _lib.check_device_error()
    assert torch.equal(got["statement_mask"], ref["statement_mask"])
    for key in ("statement",) + want_keys:
        r_, g_ = ref[key], got[key]
        live = r_ > -1e9  # raw scores carry exact -1e10 entries at masked pairs
        assert torch.equal(live, g_ > -1e9)
        err = (g_[live] - r_[live]).abs().max().item() / max(r_[live].abs().max().item(), 1e-9)
        assert err < 1e-5, (key, err)
