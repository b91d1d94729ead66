"""Host logic of the N > 1 path on CPU: example-wise sharding (one process per GPU, no data-path collective) and the
edge all-gather, exercised with world_size-2 (and 3) gloo process groups."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tvqaplus_b200 import host


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def per_example_fn(b):
    """Stand-in for the forward: any function that is independent across examples (what makes the path shard)."""
    a = b["qas_bert"].flatten(1).sum(1, keepdim=True) + b["sub_bert"].flatten(1).mean(1, keepdim=True)
    return torch.cat([a, b["vid"].flatten(1).amax(1, keepdim=True), b["target"].float().unsqueeze(1)], 1)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = host.synthetic_batch(n, 4, 3, 5, 6, seed=7, wd=16, vd=8)
        mine = host.shard_batch(full, world, rank)
        lo, hi = host.shard_bounds(n, world, rank)
        assert len(mine.qid) == hi - lo and mine.qas_bert.shape[0] == hi - lo
        assert mine.ts_label["st"].shape[0] == hi - lo and mine.use_hard_negatives is False
        got = host.gather_examples(per_example_fn(mine), n)
        want = per_example_fn(full)
        # bench.py's timing reduction: MAX over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(torch.equal(got, want)), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (2, 1), (3, 7)])
def test_sharded_forward_gathers_to_the_unsharded_result(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), "gathered shards differ from the unsharded result"
    assert all(r[2] == float(world) for r in res)


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [host.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        host.shard_bounds(4, 2, 2)


def test_merge_predictions_restores_batch_order():
    full = host.synthetic_batch(5, 2, 2, 3, 3, seed=1, wd=8, vd=4)
    parts = []
    for r in range(2):
        s = host.shard_batch(full, 2, r)
        parts.append({"answer": s.qas_bert.flatten(1)[:, :5], "t_scores": s.vid.flatten(1)[:, :4],
                      "att_predictions": [{"qid": q} for q in s.qid]})
    m = host.merge_predictions(parts)
    assert torch.equal(m["answer"], full.qas_bert.flatten(1)[:, :5])
    assert [d["qid"] for d in m["att_predictions"]] == full.qid
    parts[1]["att_predictions"] = None
    assert host.merge_predictions(parts)["att_predictions"] is None
