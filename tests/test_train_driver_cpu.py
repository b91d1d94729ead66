"""Host logic of the data-parallel training driver on CPU (tvqaplus_b200/train_driver.py): the per-epoch sharding rule, the
checkpoint round trip, and - with a world_size-2 gloo group - that summing the per-rank gradients of per-example-sum losses
reproduces the large-batch gradient (why the training collective is one SUM all-reduce)."""
import os
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_parallel_cpu import free_port
from tvqaplus_b200 import train_driver as D


def test_shard_indices_partition_every_epoch():
    n, world = 103, 4
    for epoch in (0, 1):
        shares = [D.shard_indices(n, world, r, epoch, seed=5) for r in range(world)]
        assert len({len(s) for s in shares}) == 1 and len(shares[0]) == n // world
        flat = [i for s in shares for i in s]
        assert len(set(flat)) == len(flat) and set(flat) <= set(range(n))
    assert D.shard_indices(n, world, 0, 0, seed=5) != D.shard_indices(n, world, 0, 1, seed=5)   # reshuffled per epoch
    assert D.shard_indices(n, world, 1, 3, seed=5) == D.shard_indices(n, world, 1, 3, seed=5)   # and deterministic
    assert D.shard_indices(8, 2, 1, 0, shuffle=False) == [1, 3, 5, 7]
    assert list(D.batches_of(list(range(7)), 3)) == [[0, 1, 2], [3, 4, 5], [6]]
    assert list(D.batches_of(list(range(7)), 3, drop_last=True)) == [[0, 1, 2], [3, 4, 5]]


class _StubModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(5, 3)
        self.invalidated = 0

    def invalidate_weight_cache(self):
        self.invalidated += 1


def test_checkpoint_round_trip(tmp_path):
    torch.manual_seed(0)
    tr = types.SimpleNamespace(model=_StubModel(), m=torch.randn(18), v=torch.rand(18), steps=41, step_dev=torch.zeros(1, dtype=torch.int32))
    path = str(tmp_path / "ck.pth")
    D.save_checkpoint(path, tr, epoch=3, best_acc=0.625)
    want = {k: v.clone() for k, v in tr.model.state_dict().items()}
    tr2 = types.SimpleNamespace(model=_StubModel(), m=torch.zeros(18), v=torch.zeros(18), steps=0, step_dev=torch.zeros(1, dtype=torch.int32))
    st = D.load_checkpoint(path, tr2)
    assert st == {"epoch": 3, "best_acc": 0.625} and tr2.steps == 41 and int(tr2.step_dev) == 41 and tr2.model.invalidated == 1
    assert torch.equal(tr2.m, tr.m) and torch.equal(tr2.v, tr.v)
    assert all(torch.equal(v, want[k]) for k, v in tr2.model.state_dict().items())
    assert not os.path.exists(path + ".tmp")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(3)
        w = torch.randn(6, 4, requires_grad=True)
        x = torch.randn(8, 4)
        y = torch.randint(0, 6, (8,))

        def grad_of(idx):
            loss = torch.nn.functional.cross_entropy(x[idx] @ w.t(), y[idx], reduction="sum")  # a sum over examples, like all three STAGE losses
            return torch.autograd.grad(loss, w)[0].reshape(-1)

        mine = D.shard_indices(8, world, rank, epoch=0, seed=1)
        flat = grad_of(mine)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # what Trainer.step does with its flat gradient buffer
        full = grad_of(list(range(8)))
        tot = D.sum_over_ranks([len(mine), 1.0], "cpu")
        q.put((rank, float((flat - full).abs().max()), tot))
    finally:
        dist.destroy_process_group()


def test_summed_rank_gradients_equal_the_large_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, err, tot in res:
        assert err < 1e-5 and tot == [8.0, 2.0]
