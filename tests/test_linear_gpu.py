"""stage_linear_fwd (tcgen05 split-fp16 GEMM + fused LN / ReLU) against a plain torch fp32 CPU reference."""
import pytest
import torch
import torch.nn.functional as F

from tvqaplus_b200 import _lib, ops

TOL = 1e-4  # max-norm relative; the split-fp16 GEMM is good to ~1e-6, the gate for the whole path is 1e-3


def ref_linear(x, w, b, ln_in, ln_out, l2, relu):
    if l2:
        x = F.normalize(x, p=2, dim=-1)
    if ln_in is not None:
        x = F.layer_norm(x, (x.shape[-1],), ln_in[0], ln_in[1], 1e-5)
    y = F.linear(x, w, b)
    if relu:
        y = F.relu(y)
    if ln_out is not None:
        y = F.layer_norm(y, (y.shape[-1],), ln_out[0], ln_out[1], 1e-5)
    return y


CASES = [
    # m, k, n, ln_in, ln_out, l2, relu
    (300, 768, 300, True, True, False, True),      # bert_word_encoding_fc
    (77, 300, 300, True, True, True, True),        # vid_fc (+ F.normalize)
    (1000, 300, 128, False, True, False, True),    # input_embedding
    (128, 128, 128, False, False, False, False),   # plain
    (5, 128, 1, True, False, False, False),        # LinearWrapper(128 -> 1)
    (40000, 768, 300, True, True, False, True),    # many tiles per CTA (persistent loop, phase wrap)
    (129, 384, 128, True, False, False, True),     # c2q-like
    (257, 300, 300, False, False, True, False),    # F.normalize alone
    (9000, 300, 128, False, True, False, True),    # input_embedding, many tiles (5 chunks, prefetch wraps tiles)
]


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,n,ln_in,ln_out,l2,relu", CASES)
def test_linear_matches_reference(m, k, n, ln_in, ln_out, l2, relu):
    g = torch.Generator().manual_seed(m * 31 + k)
    x = torch.randn(m, k, generator=g)
    if m > 10:
        x[3].zero_()  # zero padded row, as pad_sequences_2d produces
        x[5] += 40.0  # a row with a large common offset (the one-pass variance is shifted by its first element)
        x[7] = 3.0    # a constant row: LayerNorm must give exactly beta
    w = torch.randn(n, k, generator=g) / k ** 0.5
    b = torch.randn(n, generator=g) * 0.1
    li = (torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g) * 0.1) if ln_in else None
    lo = (torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1) if ln_out else None
    want = ref_linear(x, w, b, li, lo, l2, relu)
    dev = torch.device("cuda:0")
    if li:
        wp, cs, cb = ops.pack_weight_ln(w.to(dev), b.to(dev), li[0].to(dev), li[1].to(dev))
    else:
        wp, cs, cb = ops.pack_weight(w.to(dev)), None, None
    got = ops.linear_fwd(x.to(dev), wp, None if li else b.to(dev), n, ln_in=(cs, cb) if li else None,
                         ln_out=tuple(t.to(dev) for t in lo) if lo else None, l2norm_in=l2, relu=relu)
    torch.cuda.synchronize()
    _lib.check_device_error()
    got = got.cpu()
    err = (got - want).abs().max().item() / max(1e-9, want.abs().max().item())
    print("linear m=%d k=%d n=%d err=%.3e" % (m, k, n, err))
    assert torch.isfinite(got).all()
    assert err < TOL, err
