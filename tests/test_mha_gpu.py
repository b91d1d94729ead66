"""stage_mha_fwd against the CPU oracle: MultiHeadedAttention (model/self_attention.py:19-71) on its own, the
attention-carrying EncoderBlock (model/encoder.py:29-52), and the stand-alone StructuredAttention module."""
import pytest
import torch

from oracle import stage_oracle as O
from tests import helpers as H
from tvqaplus_b200 import _lib

TOL = 1e-4


def ragged_mask(n_seq, L, g, allow_empty=True):
    lens = torch.randint(0 if allow_empty else 1, L + 1, (n_seq,), generator=g)
    return (torch.arange(L)[None] < lens[:, None]).float(), lens


@pytest.mark.gpu
@pytest.mark.parametrize("n_seq,L,nh", [(3, 100, 4), (40, 20, 4), (7, 128, 4), (5, 33, 2), (4, 50, 1), (1, 1, 4),
                                         (301, 25, 4)])
def test_multi_headed_attention(n_seq, L, nh):
    from tvqaplus_b200.model import MultiHeadedAttention
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(100 * n_seq + L + nh)
    torch.manual_seed(n_seq + L)
    mha = MultiHeadedAttention(nh, 128).eval()
    with torch.no_grad():  # the reference deep-copies one Linear four times; make the four differ
        for l in mha.linears:
            l.weight.copy_(torch.randn(128, 128, generator=g) / 128 ** 0.5)
            l.bias.copy_(torch.randn(128, generator=g) * 0.1)
    x = torch.randn(n_seq, L, 128, generator=g)
    mask, _ = ragged_mask(n_seq, L, g)
    sd = {"m." + k: v for k, v in mha.state_dict().items()}
    mha = mha.to(dev)
    for m in (mask, None):
        want = O.multi_headed_attention(x, m, sd, "m", nh)
        got = mha(x.to(dev), m.to(dev) if m is not None else None).cpu()
        torch.cuda.synchronize()
        _lib.check_device_error()
        err = H.max_rel(got, want)
        print("mha n_seq=%d L=%d nh=%d mask=%s err=%.3e" % (n_seq, L, nh, m is not None, err))
        assert err < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("n_seq,L,k,nh", [(6, 100, 5, 4), (50, 20, 7, 4), (9, 40, 7, 2)])
def test_encoder_block_with_attention(n_seq, L, k, nh):
    from tvqaplus_b200.model import EncoderBlock
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n_seq * 3 + L)
    torch.manual_seed(L)
    blk = EncoderBlock(n_conv=2, kernel_size=k, n_filters=128, dropout=0.1, num_heads=nh).eval()
    with torch.no_grad():
        for n_, p_ in blk.named_parameters():
            if p_.dim() == 1:
                p_.add_(torch.randn(p_.shape, generator=g) * 0.1)
        for l in blk.multi_head_attn.linears:
            l.weight.copy_(torch.randn(128, 128, generator=g) / 128 ** 0.5)
    x = torch.randn(n_seq, L, 128, generator=g)
    mask, lens = ragged_mask(n_seq, L, g)
    x = x * mask[..., None]
    sd = {"e." + kk: v for kk, v in blk.state_dict().items()}
    want = O.encoder_block(x, mask, sd, "e", 2, k, nh)
    blk = blk.to(dev)
    got = blk(x.to(dev), mask.to(dev)).cpu()
    assert H.max_rel(got, want) < TOL
    # fused masked max-pool tail (model/stage.py:503)
    want_p = torch.max(O.mask_logits(want, mask.unsqueeze(2)), 1)[0]
    got_p = blk(x.to(dev), mask.to(dev), pool=True).cpu()
    torch.cuda.synchronize()
    _lib.check_device_error()
    dead = lens == 0
    assert torch.equal(got_p[dead], torch.full_like(got_p[dead], -1e10))
    if (~dead).any():
        assert H.max_rel(got_p[~dead], want_p[~dead]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,Lq,Lr", [(2, 7, 40, 25), (1, 3, 100, 100), (2, 2, 5, 1)])
def test_structured_attention_module(N, T, Lq, Lr):
    """The parameter-free attention module on its own returns the reference's (A, S, S_mask, S_) tuple."""
    from tvqaplus_b200.model import StructuredAttention
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 31 + T + Lq + Lr)
    C = torch.randn(N, 5, 1, Lq, 128, generator=g)
    Q = torch.randn(N, 1, T, Lr, 128, generator=g)
    cm = (torch.arange(Lq)[None, None, None] < torch.randint(1, Lq + 1, (N, 5, 1, 1), generator=g)).float()
    qm = (torch.arange(Lr)[None, None, None] < torch.randint(0, Lr + 1, (N, 1, T, 1), generator=g)).float()
    want = O.structured_attention(C, Q, cm, qm, 10.0)
    att = StructuredAttention(dropout=0.1, scale=10.0).eval()
    got = att(C.to(dev), Q.to(dev), cm.to(dev), qm.to(dev))
    torch.cuda.synchronize()
    _lib.check_device_error()
    assert H.max_rel(got[0].cpu(), want[0]) < TOL
    assert H.masked_rel(got[1].cpu(), want[1]) < TOL
    assert torch.equal(got[2].cpu(), want[2])
    assert H.max_rel(got[3].cpu(), want[3]) < TOL
