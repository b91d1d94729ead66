"""stage_encoder_block_fwd against the CPU oracle's encoder_block (model/encoder.py:29-52)."""
import pytest
import torch

from oracle import stage_oracle as O
from tvqaplus_b200 import _lib, ops

TOL = 1e-4


def random_block(prefix, n_conv, k, g):
    sd = {prefix + ".position_encoding.pe": O.position_encoding_table()}
    for i in range(n_conv):
        sd["%s.layer_norm.%d.weight" % (prefix, i)] = torch.rand(128, generator=g) + 0.5
        sd["%s.layer_norm.%d.bias" % (prefix, i)] = torch.randn(128, generator=g) * 0.1
        sd["%s.conv.%d.depthwise_conv.weight" % (prefix, i)] = torch.randn(128, 1, k, generator=g) / k ** 0.5
        sd["%s.conv.%d.depthwise_conv.bias" % (prefix, i)] = torch.randn(128, generator=g) * 0.1
        sd["%s.conv.%d.pointwise_conv.weight" % (prefix, i)] = torch.randn(128, 128, 1, generator=g) / 128 ** 0.5
        sd["%s.conv.%d.pointwise_conv.bias" % (prefix, i)] = torch.randn(128, generator=g) * 0.1
    sd[prefix + ".final_layer_norm.weight"] = torch.rand(128, generator=g) + 0.5
    sd[prefix + ".final_layer_norm.bias"] = torch.randn(128, generator=g) * 0.1
    return sd


def run_gpu(x, mask, sd, prefix, n_conv, k, out_mode):
    dev = torch.device("cuda:0")
    d = {kk: v.to(dev) for kk, v in sd.items()}
    convs = [dict(ln_g=d["%s.layer_norm.%d.weight" % (prefix, i)], ln_b=d["%s.layer_norm.%d.bias" % (prefix, i)],
                  dw_w=d["%s.conv.%d.depthwise_conv.weight" % (prefix, i)],
                  dw_b=d["%s.conv.%d.depthwise_conv.bias" % (prefix, i)],
                  pw_packed=ops.pack_weight(d["%s.conv.%d.pointwise_conv.weight" % (prefix, i)], 64),
                  pw_b=d["%s.conv.%d.pointwise_conv.bias" % (prefix, i)]) for i in range(n_conv)]
    y = ops.encoder_block_fwd(x.to(dev), mask.to(dev), d[prefix + ".position_encoding.pe"], convs,
                              (d[prefix + ".final_layer_norm.weight"], d[prefix + ".final_layer_norm.bias"]), k,
                              out_mode=out_mode)
    torch.cuda.synchronize()
    _lib.check_device_error()
    return y.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("n_seq,L,n_conv,k", [(3, 100, 2, 7), (700, 20, 2, 7), (150, 100, 2, 5), (5, 128, 1, 3),
                                               (11, 25, 2, 7), (9, 40, 3, 5), (1, 1, 2, 7), (400, 50, 2, 7)])
def test_encoder_block(n_seq, L, n_conv, k):
    g = torch.Generator().manual_seed(n_seq * 7 + L)
    sd = random_block("enc", n_conv, k, g)
    x = torch.randn(n_seq, L, 128, generator=g)
    lens = torch.randint(0, L + 1, (n_seq,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).float()
    x = x * mask[..., None]  # zero padded rows like the real batches
    want = O.encoder_block(x, mask, sd, "enc", n_conv, k, 0)
    got = run_gpu(x, mask, sd, "enc", n_conv, k, 0)
    err = (got - want).abs().max().item() / want.abs().max().item()
    print("encoder n_seq=%d L=%d err=%.3e" % (n_seq, L, err))
    assert err < TOL
    # masked max-pool epilogue (model/stage.py:503): exact -1e10 for fully masked sequences
    want_p = torch.max(O.mask_logits(want, mask.unsqueeze(2)), 1)[0]
    got_p = run_gpu(x, mask, sd, "enc", n_conv, k, 1)
    dead = lens == 0
    assert torch.equal(got_p[dead], torch.full_like(got_p[dead], -1e10))
    if (~dead).any():
        errp = (got_p[~dead] - want_p[~dead]).abs().max().item() / want_p[~dead].abs().max().item()
        assert errp < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("n_seq,L,k", [(7, 100, 5), (10, 40, 5), (130, 1, 5), (9, 33, 7), (4, 128, 5)])
def test_encoder_block_blocked_input(n_seq, L, k):
    """The blocked tile layout (the internal hand-over format from the fused CQA kernel) gives the row-major result."""
    if not ops.blocked_fits(L):
        pytest.skip("L=%d does not fit the blocked tile buffer" % L)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n_seq + L)
    sd = random_block("enc", 2, k, g)
    x = torch.randn(n_seq, L, 128, generator=g)
    lens = torch.randint(0, L + 1, (n_seq,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).float()
    want = run_gpu(x, mask, sd, "enc", 2, k, 1)
    S = ops.blocked_stride(L)
    assert S >= L and S % 8 == 1
    blk = torch.full((n_seq, 32, S, 4), float("nan"))          # padding rows are never read
    blk[:, :, :L, :] = x.view(n_seq, L, 32, 4).permute(0, 2, 1, 3)
    d = {kk: v.to(dev) for kk, v in sd.items()}
    convs = [dict(ln_g=d["enc.layer_norm.%d.weight" % i], ln_b=d["enc.layer_norm.%d.bias" % i],
                  dw_w=d["enc.conv.%d.depthwise_conv.weight" % i], dw_b=d["enc.conv.%d.depthwise_conv.bias" % i],
                  pw_packed=ops.pack_weight(d["enc.conv.%d.pointwise_conv.weight" % i], 64),
                  pw_b=d["enc.conv.%d.pointwise_conv.bias" % i]) for i in range(2)]
    got = ops.encoder_block_fwd(blk.reshape(-1).to(dev), mask.to(dev), ops.blocked_pe(d["enc.position_encoding.pe"], L), convs,
                                (d["enc.final_layer_norm.weight"], d["enc.final_layer_norm.bias"]), k, out_mode=1,
                                blocked_in=(n_seq, L)).cpu()
    torch.cuda.synchronize()
    _lib.check_device_error()
    dead = lens == 0
    assert torch.equal(got[dead], want[dead])
    if (~dead).any():
        assert (got[~dead] - want[~dead]).abs().max().item() / want[~dead].abs().max().item() < 1e-5
