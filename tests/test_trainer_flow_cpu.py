"""Control flow of Trainer.step on CPU: the CUDA primitives are replaced by torch stand-ins so that the host logic - flat
parameter buffer, part A / all-reduce / part B order, step counting, the summed gradient reaching Adam on every rank of a
world_size-2 gloo group - runs without a GPU.  The kernels themselves are checked in the -m gpu tests."""
import os
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_parallel_cpu import free_port


def _fake_ops(monkeypatch_target):
    """torch stand-ins for the four primitives Trainer calls directly."""
    T = monkeypatch_target

    def answer_ce(rows, valid, row_target, n):
        return torch.tensor([float(rows.sum()), float(valid.sum()), 0.0]), torch.zeros_like(rows)

    def grad_sumsq(g, out):
        out += (g * g).sum()
        return out

    def counter_inc(c):
        c += 1

    def adam_step(p, g, m, v, ss, *, max_norm, lr, beta1, beta2, eps, wd, step=1, grad_scale=1.0, step_dev=None):
        t = int(step_dev.item())
        coef = min(1.0, max_norm / (float(ss.sqrt()) + 1e-6))
        g.mul_(coef)
        gg = g + wd * p
        m.mul_(beta1).add_(gg, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
        p.sub_(lr / (1 - beta1 ** t) * m / (v.sqrt() / (1 - beta2 ** t) ** 0.5 + eps))

    T.answer_ce, T.grad_sumsq, T.counter_inc, T.adam_step = answer_ce, grad_sumsq, counter_inc, adam_step


def _make_trainer(rank):
    from tvqaplus_b200 import host, train
    from tvqaplus_b200.model import STAGE
    _fake_ops(train.T)
    torch.manual_seed(0)
    model = STAGE(host.default_opt(dropout=0.0))
    tr = train.Trainer(model)
    n = tr.engine.n_params

    def forward(batch, att_offsets=None):
        return dict(answer_rows=torch.ones(4, 5), row_valid=torch.ones(4, dtype=torch.int32), row_target=torch.zeros(4, dtype=torch.int64),
                    temporal_loss=torch.tensor(2.0), att_loss=None, ctx={})

    def backward(ctx, d_rows, g_ts, g_att=None):
        g = torch.Generator().manual_seed(10 + rank)
        return torch.randn(n, generator=g)

    tr.engine.forward, tr.engine.backward = forward, backward
    return tr, model


def test_step_flow_single_process(monkeypatch):
    from tvqaplus_b200 import train_ops
    for name in ("answer_ce", "grad_sumsq", "counter_inc", "adam_step"):  # restored after the test: the stand-ins must not leak
        monkeypatch.setattr(train_ops, name, getattr(train_ops, name))
    tr, model = _make_trainer(0)
    before = tr.flat.clone()
    batch = types.SimpleNamespace(qas_bert=torch.zeros(2, 5, 3, 4))
    out = tr.step(batch)
    assert set(out) >= {"loss", "cls_loss", "n_rows", "n_correct", "temporal_loss", "att_loss", "grad", "grad_sumsq"}
    assert float(out["loss"]) == 20.0 + 0.5 * 2.0 and tr.steps == 1 and int(tr.step_dev) == 1
    assert not torch.equal(tr.flat, before)
    # the module's parameters are views of the flat buffer: the update is visible through state_dict()
    name, p = next(iter(model.named_parameters()))
    o = tr.engine.offsets[name]
    assert torch.equal(p.detach().reshape(-1), tr.flat[o:o + p.numel()])
    tr.step(batch)
    assert tr.steps == 2 and int(tr.step_dev) == 2


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr, _ = _make_trainer(rank)
        n = tr.engine.n_params
        out = tr.step(types.SimpleNamespace(qas_bert=torch.zeros(2, 5, 3, 4)))
        want = sum(torch.randn(n, generator=torch.Generator().manual_seed(10 + r)) for r in range(world))
        coef = min(1.0, 10.0 / (float(want.norm()) + 1e-6))
        q.put((rank, float((out["grad"] - want * coef).abs().max()), float(tr.flat.double().sum())))
    finally:
        dist.destroy_process_group()


def test_ranks_apply_the_summed_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] < 1e-5 for r in res), res            # every rank's Adam saw (clipped) sum of both ranks' gradients
    assert abs(res[0][2] - res[1][2]) < 1e-6, res        # and the replicas stay identical


def test_att_offsets_address_the_engine_score_layout():
    """The pair rows are (batch, answer, image, word, region) indices into the reference's (N, 5, T, Lq, R) score tensor; the
    training engine keeps its scores as (N, 5, Lq, T, R).  The offsets must pick the same elements."""
    from tvqaplus_b200 import host, train
    from tvqaplus_b200.model import STAGE
    eng = train.TrainEngine(STAGE(host.default_opt(dropout=0.0)))
    N, T, Lq, R = 3, 7, 5, 4
    ref = torch.arange(N * 5 * T * Lq * R, dtype=torch.float32).view(N, 5, T, Lq, R)
    ours = ref.permute(0, 1, 3, 2, 4).contiguous().reshape(-1)
    g = torch.Generator().manual_seed(0)
    rows = torch.stack([torch.randint(0, hi, (50,), generator=g) for hi in (N, 5, T, Lq, R)], 1)
    pos, neg = eng.att_offsets_from_pairs(rows, rows.flip(0), N, T, Lq, R)
    assert torch.equal(ours[pos], ref[rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 4]])
    r2 = rows.flip(0)
    assert torch.equal(ours[neg], ref[r2[:, 0], r2[:, 1], r2[:, 2], r2[:, 3], r2[:, 4]])
    # flat parameter / gradient buffer: state_dict order, every tensor on a 256-byte boundary, nothing overlaps
    prev_end = 0
    for n, p in eng.model.named_parameters():
        o = eng.offsets[n]
        assert o % 64 == 0 and o >= prev_end
        prev_end = o + p.numel()
    assert prev_end <= eng.n_params
