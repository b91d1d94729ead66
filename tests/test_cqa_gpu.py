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
