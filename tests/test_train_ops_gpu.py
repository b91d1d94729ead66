"""Training primitives (csrc/train.cuh) against plain PyTorch fp32 autograd on the same device.  Tolerances are relative to
the largest reference magnitude (summation order differs; everything is fp32 FMA, allow_tf32 off)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from tvqaplus_b200 import train_ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return train_ops


def _close(got, want, tol=2e-5):
    scale = float(want.abs().max()) + 1e-30
    err = float((got - want).abs().max()) / scale
    assert err <= tol, "max-norm relative error %.3e > %.1e" % (err, tol)


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (100, 25, 50), (513, 300, 768), (37, 1, 128), (260, 2, 33), (1000, 384, 128)])
def test_linear_fwd_bwd(m, n, k):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    x = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g)
    b = torch.randn(n, device="cuda", generator=g)
    dy = torch.randn(m, n, device="cuda", generator=g)
    _close(T.linear(x, w, b), F.linear(x, w, b))
    _close(T.linear(x, w, b, relu=True), F.relu(F.linear(x, w, b)))
    _close(T.linear_bwd_x(dy, w), dy @ w)
    dw = torch.ones(n, k, device="cuda")
    db = torch.ones(n, device="cuda")
    T.linear_bwd_w(dy, x, dw, db)
    _close(dw, 1 + dy.t() @ x)
    _close(db, 1 + dy.sum(0))


def test_linear_into_strided_column():
    T = _ops()
    x = torch.randn(333, 128, device="cuda")
    w = torch.randn(1, 128, device="cuda")
    b = torch.randn(1, device="cuda")
    out = torch.zeros(333, 2, device="cuda")
    T.linear(x, w, b, out=out, ldc=2, c_off=1)
    _close(out[:, 1], F.linear(x, w, b)[:, 0])
    assert float(out[:, 0].abs().max()) == 0.0
    # and the matching backward reads: dy taken from column 1 of a (M, 2) tensor
    dy = torch.randn(333, 2, device="cuda")
    _close(T.linear_bwd_x(dy, w, dy_ld=2, dy_off=1, n=1), dy[:, 1:2] @ w)
    dw = torch.zeros(1, 128, device="cuda")
    db = torch.zeros(1, device="cuda")
    T.linear_bwd_w(dy, x, dw, db, dy_ld=2, dy_off=1, n=1)
    _close(dw, dy[:, 1:2].t() @ x)
    _close(db, dy[:, 1].sum(0, keepdim=True))


def test_batched_attention_contractions():
    """The five strided batched GEMMs of the context-query attention in the (n, a*i, t, j) layout."""
    T = _ops()
    n, ai, t, lc, d = 2, 45, 7, 25, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    cn = torch.randn(n, ai, d, device="cuda", generator=g)
    q = torch.randn(n, t, lc, d, device="cuda", generator=g)
    s = torch.empty(n, ai, t, lc, device="cuda")
    T.gemm(cn, q, s, ai, t * lc, d, (d, 1), (1, d), t * lc, batch=(n, 1), a_b=(ai * d, 0), b_b=(t * lc * d, 0), c_b=(ai * t * lc, 0))
    _close(s, torch.einsum("nrd,ntjd->nrtj", cn, q))
    p = torch.softmax(s, -1)
    u = torch.empty(n, ai, t, d, device="cuda")
    T.gemm(p, q, u, ai, d, lc, (t * lc, 1), (d, 1), t * d, batch=(n, t), a_b=(ai * t * lc, lc), b_b=(t * lc * d, lc * d), c_b=(ai * t * d, d))
    _close(u, torch.einsum("nrtj,ntjd->nrtd", p, q))
    du = torch.randn_like(u)
    dp = torch.empty_like(p)
    T.gemm(du, q, dp, ai, lc, d, (t * d, 1), (1, d), t * lc, batch=(n, t), a_b=(ai * t * d, d), b_b=(t * lc * d, lc * d), c_b=(ai * t * lc, lc))
    _close(dp, torch.einsum("nrtd,ntjd->nrtj", du, q))
    dq = torch.empty_like(q)
    T.gemm(p, du, dq, lc, d, ai, (1, t * lc), (t * d, 1), d, batch=(n, t), a_b=(ai * t * lc, lc), b_b=(ai * t * d, d), c_b=(t * lc * d, lc * d))
    _close(dq, torch.einsum("nrtj,nrtd->ntjd", p, du))
    ds = torch.randn_like(s)
    dcn = torch.zeros_like(cn)
    T.gemm(ds, q, dcn, ai, d, t * lc, (t * lc, 1), (d, 1), d, batch=(n, 1), a_b=(ai * t * lc, 0), b_b=(t * lc * d, 0), c_b=(ai * d, 0), mode=2, split_k=3)
    _close(dcn, torch.einsum("nrtj,ntjd->nrd", ds, q))
    dqn = torch.empty_like(q)
    T.gemm(ds, cn, dqn, t * lc, d, ai, (1, t * lc), (d, 1), d, batch=(n, 1), a_b=(ai * t * lc, 0), b_b=(ai * d, 0), c_b=(t * lc * d, 0))
    _close(dqn, torch.einsum("nrtj,nrd->ntjd", ds, cn))


@pytest.mark.parametrize("rows,d", [(1000, 128), (77, 300), (513, 768), (64, 384), (10, 256)])
def test_layer_norm(rows, d):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(d)
    x = (torch.randn(rows, d, device="cuda", generator=g) * 2 + 0.5).requires_grad_()
    w = torch.randn(d, device="cuda", generator=g).requires_grad_()
    b = torch.randn(d, device="cuda", generator=g).requires_grad_()
    dy = torch.randn(rows, d, device="cuda", generator=g)
    want = F.layer_norm(x, (d,), w, b, 1e-5)
    want.backward(dy)
    y, st = T.ln_fwd(x.detach(), w.detach(), b.detach())
    _close(y, want.detach())
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    res = torch.randn(rows, d, device="cuda", generator=g)
    dx = T.ln_bwd(dy, x.detach(), st, w.detach(), dg, db, add=res)
    _close(dx, x.grad + res)
    _close(dg, w.grad)
    _close(db, b.grad)
    # ReLU-input gradient: x = relu(z) feeds the LayerNorm
    xr = F.relu(x.detach() - 0.5)
    z = (x.detach() - 0.5).requires_grad_()
    F.layer_norm(F.relu(z), (d,), w.detach(), b.detach(), 1e-5).backward(dy)
    y2, st2 = T.ln_fwd(xr, w.detach(), b.detach())
    dz = T.ln_bwd(dy, xr, st2, w.detach(), torch.zeros_like(dg), torch.zeros_like(db), relu_x=True)
    _close(dz, z.grad)


@pytest.mark.parametrize("k,L", [(7, 100), (5, 40), (7, 20), (3, 9)])
def test_depthwise_conv(k, L):
    T = _ops()
    n_seq, d = 13, 128
    g = torch.Generator(device="cuda").manual_seed(k)
    x = torch.randn(n_seq, L, d, device="cuda", generator=g).requires_grad_()
    w = torch.randn(d, 1, k, device="cuda", generator=g).requires_grad_()
    b = torch.randn(d, device="cuda", generator=g).requires_grad_()
    dy = torch.randn(n_seq, L, d, device="cuda", generator=g)
    want = F.conv1d(x.transpose(1, 2), w, b, padding=k // 2, groups=d).transpose(1, 2)
    want.backward(dy)
    y = T.dwconv_fwd(x.detach().reshape(-1, d), w.detach(), b.detach(), n_seq, L, k)
    _close(y.view(n_seq, L, d), want.detach())
    dw = torch.zeros(d, 1, k, device="cuda")
    db = torch.zeros(d, device="cuda")
    dx = T.dwconv_bwd(dy.reshape(-1, d), x.detach().reshape(-1, d), w.detach(), dw, db, n_seq, L, k)
    _close(dx.view(n_seq, L, d), x.grad)
    _close(dw, w.grad)
    _close(db, b.grad)


@pytest.mark.parametrize("d", [128, 300])
def test_l2norm(d):
    T = _ops()
    x = torch.randn(501, d, device="cuda").requires_grad_()
    dy = torch.randn(501, d, device="cuda")
    want = F.normalize(x, p=2, dim=-1)
    want.backward(dy)
    y, inv = T.l2norm_fwd(x.detach())
    _close(y, want.detach())
    base = torch.randn(501, d, device="cuda")
    dx = T.l2norm_bwd(dy, y, inv, base.clone(), True)
    _close(dx, base + x.grad)


def test_attention_softmax_masks_and_gradient():
    T = _ops()
    n, lq5, t, lc, scale = 2, 30, 6, 25, 10.0
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = (torch.rand(n, lq5, t, lc, device="cuda", generator=g) * 2 - 1)
    cm = (torch.rand(n, lq5, device="cuda", generator=g) > 0.2).float()
    qm = (torch.rand(n, t, lc, device="cuda", generator=g) > 0.3).float()
    qm[0, 2] = 0  # an empty frame
    m = cm[:, :, None, None] * qm[:, None]
    x = raw.clone().requires_grad_()
    s_want = x - 1e10 * (1 - m)
    p_want = torch.softmax(s_want * scale, -1) * m
    dp = torch.randn_like(raw)
    p_want.backward(dp)
    s = raw.clone()
    p = T.attn_softmax_fwd(s, cm, qm, n, lq5, t, lc, scale)
    assert torch.equal(s, s_want.detach())
    _close(p, p_want.detach())
    ds = T.attn_softmax_bwd_(dp.clone(), p, lc, scale)
    _close(ds, x.grad, 1e-4)


def test_cat3_and_its_backward():
    T = _ops()
    a_rows, t = 40, 7
    a = torch.randn(a_rows, 128, device="cuda").requires_grad_()
    b = torch.randn(a_rows * t, 128, device="cuda").requires_grad_()
    ar = a.repeat_interleave(t, 0)
    want = torch.cat([ar, b, ar * b], -1)
    d = torch.randn_like(want)
    want.backward(d)
    out = T.cat3_fwd(a.detach(), b.detach(), t)
    assert torch.equal(out, want.detach())
    da = torch.ones_like(a)
    db = T.cat3_bwd(d, a.detach(), b.detach(), da, t, True)
    _close(da, 1 + a.grad)
    _close(db, b.grad)
    a2 = torch.randn(a_rows, 128, device="cuda")
    b2 = torch.randn(a_rows, 128, device="cuda")
    assert torch.equal(T.cat3_fwd(a2, b2, 1), torch.cat([a2, b2, a2 * b2], -1))


def test_maxpool_permute_and_elementwise():
    T = _ops()
    n_seq, L = 37, 23
    x = torch.randn(n_seq, L, 128, device="cuda").requires_grad_()
    mask = (torch.rand(n_seq, L, device="cuda") > 0.3).float()
    mask[3] = 0
    want = torch.max(x * mask[..., None] + (1 - mask[..., None]) * -1e10, 1)[0]
    dy = torch.randn(n_seq, 128, device="cuda")
    want.backward(dy)
    y, arg = T.maxpool_fwd(x.detach().reshape(-1, 128), mask, n_seq, L)
    assert torch.equal(y, want.detach())
    dx = T.maxpool_bwd(dy, mask, arg, n_seq, L)
    live = mask.sum(1) > 0  # fully masked sequences: every row ties at -1e10 and the gradient is zero either way
    _close(dx.view(n_seq, L, 128)[live], x.grad[live])
    assert float(dx.view(n_seq, L, 128)[~live].abs().max()) == 0.0
    z = torch.randn(3, 5, 7, 128, device="cuda")
    assert torch.equal(T.permute_it(z, 3, 5, 7).view(3, 7, 5, 128), z.transpose(1, 2).contiguous())
    pe = torch.randn(L, 128, device="cuda")
    assert torch.equal(T.add(x.detach().reshape(-1, 128), pe, period=L * 128).view(n_seq, L, 128), x.detach() + pe)
    yv = torch.randn(1000, device="cuda")
    assert torch.equal(T.relu_bwd(dy.reshape(-1)[:1000].contiguous(), yv), dy.reshape(-1)[:1000] * (yv > 0))
    sc = torch.tensor([0.25], device="cuda")
    assert torch.equal(T.scale_(yv.clone(), sc, 2.0), yv * 0.5)


def test_adam_step_matches_torch_optimizer():
    T = _ops()
    torch.manual_seed(0)
    p = torch.randn(10007, device="cuda")
    ref = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=3e-7)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p) * 3
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 10.0)
        opt.step()
        ss = torch.zeros(1, device="cuda")
        T.grad_sumsq(g, ss)
        T.adam_step(p, g, m, v, ss, max_norm=10.0, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=3e-7, step=step)
        _close(p, ref.detach(), 1e-6)


@pytest.mark.parametrize("m,n,k,gscale", [(5000, 128, 384, 1e-6), (4100, 300, 768, 3e-9), (4500, 128, 128, 1.0), (4099, 128, 300, 1e-4)])
def test_tensor_core_linear_pair_for_training(m, n, k, gscale):
    """lin_fwd / lin_bwd_x of the training engine on the tcgen05 kernel: forward at fp32 grade, and the input gradient of a
    tensor whose entries are far below fp16's normal range (the power-of-two scaling through x_absmax)."""
    from tvqaplus_b200.train import lin_bwd_x, lin_fwd
    _ops()
    g = torch.Generator(device="cuda").manual_seed(n + k)
    x = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.1
    b = torch.randn(n, device="cuda", generator=g)
    _close(lin_fwd(x, w, b, relu=True), F.relu(F.linear(x, w, b)), 3e-5)
    dy = torch.randn(m, n, device="cuda", generator=g) * gscale
    dy[::7] *= 1e-3  # rows three orders of magnitude below the tensor's maximum
    _close(lin_bwd_x(dy, w), dy @ w, 3e-5)
