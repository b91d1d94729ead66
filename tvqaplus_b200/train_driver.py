"""Training driver, one process per GPU (reference main.py:18-148,186-251 with ``nn.DataParallel`` replaced by data-parallel
replicas and one NCCL SUM all-reduce of the flat gradient buffer per step).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m tvqaplus_b200.train_driver --epochs 2

* every rank builds the same model (same seed), takes its share of each epoch's permutation (``shard_indices``: the
  DistributedSampler rule - seeded by (seed, epoch), equal counts, no overlap) and runs ``Trainer.step``; the summed gradient
  is the large-batch gradient because all three losses are sums over examples
* validation every ``val_every`` steps and at the end of an epoch, each rank on its share, correct counts summed over ranks
  (main.py:99-110 validates inside the training loop); it runs the fused eval kernels (``model.eval()``)
* ``--resume`` restores parameters, Adam moments, step count, epoch and best accuracy (the reference saves only the
  state_dict and its ``--resume`` flag is never read, config.py:55)

The dataset (h5 features + annotations, tvqa_dataset.py) is an external download that is not available here, so the driver
takes any sequence of dataset items with ``TVQADataset.__getitem__``'s fields; ``--synthetic`` generates them.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, List, Optional, Sequence

import torch


def shard_indices(n: int, world: int, rank: int, epoch: int, seed: int = 2018, shuffle: bool = True) -> List[int]:
    """Indices rank ``rank`` visits in ``epoch``: a permutation seeded by (seed, epoch), truncated to a multiple of ``world`` and
    dealt round-robin, so every rank gets the same count and the ranks' shares are disjoint (torch DistributedSampler with
    ``drop_last=True``)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("shard_indices: rank %d of world %d" % (rank, world))
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    usable = n - n % world
    return order[rank:usable:world]


def batches_of(indices: Sequence[int], bsz: int, drop_last: bool = False):
    for i in range(0, len(indices), bsz):
        chunk = list(indices[i:i + bsz])
        if len(chunk) == bsz or (chunk and not drop_last):
            yield chunk


def save_checkpoint(path: str, trainer, epoch: int, best_acc: float) -> None:
    """Parameters (reference key names, loadable by the reference's ``load_state_dict``) + optimiser state; written
    atomically (rank 0 only calls this)."""
    state = {"model": {k: v.detach().cpu() for k, v in trainer.model.state_dict().items()},
             "adam_m": trainer.m.detach().cpu(), "adam_v": trainer.v.detach().cpu(), "steps": int(trainer.steps),
             "epoch": int(epoch), "best_acc": float(best_acc)}
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, trainer) -> Dict[str, float]:
    state = torch.load(path, map_location="cpu")
    trainer.model.load_state_dict(state["model"], strict=True)  # in place: the parameters stay views of the flat buffer
    trainer.m.copy_(state["adam_m"])
    trainer.v.copy_(state["adam_v"])
    trainer.steps = int(state["steps"])
    trainer.step_dev.fill_(trainer.steps)
    trainer.model.invalidate_weight_cache()
    return {"epoch": int(state["epoch"]), "best_acc": float(state["best_acc"])}


def sum_over_ranks(values: Sequence[float], device) -> List[float]:
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def validate(model, items, indices: Sequence[int], bsz: int, max_len: Optional[dict], device) -> Dict[str, float]:
    """Answer accuracy over this rank's share, counts summed over ranks (reference validate(), main.py:151-183)."""
    from . import collate
    was_training = model.training
    model.eval()
    correct = total = 0
    with torch.no_grad():
        for chunk in batches_of(indices, bsz):
            inp, target, _ = collate.prepare_inputs(collate.pad_collate([items[i] for i in chunk]), max_len, device)
            out = model(inp)[0]
            correct += int((out.argmax(1) == target).sum())
            total += len(chunk)
    model.train(was_training)
    c, t = sum_over_ranks([correct, total], device)
    return {"acc": c / max(t, 1.0), "n": t}


def train_epoch(trainer, items, indices: Sequence[int], bsz: int, max_len: Optional[dict], device, log_every: int = 0,
                graph: bool = False, on_step=None) -> Dict[str, float]:
    from . import collate
    model = trainer.model
    model.train()
    loss_sum = corr = rows = n_ex = 0.0
    t0 = time.time()
    for it, chunk in enumerate(batches_of(indices, bsz, drop_last=True)):
        inp, _, _ = collate.prepare_inputs(collate.pad_collate([items[i] for i in chunk]), max_len, device)
        offsets = None
        if model.use_sup_att and inp.get("att_labels") is not None:
            pos, neg = model._att_pairs(inp["att_labels"], inp["target"].cpu(), inp["anno_st_idx"], model.num_negatives, False, model.drop_topk)
            offsets = trainer.engine.att_offsets_from_pairs(pos, neg, len(chunk), inp["vid"].shape[1], inp["qas_bert"].shape[2], inp["vid"].shape[2])
        res = trainer.step(inp, offsets, graph=graph)
        if log_every and (it + 1) % log_every == 0:  # the only host reads of the loop
            loss_sum += float(res["loss"]); corr += float(res["n_correct"]); rows += float(res["n_rows"]); n_ex += len(chunk)
        if on_step is not None:
            on_step(trainer.steps)
    torch.cuda.synchronize(device)
    return {"seconds": time.time() - t0, "logged_loss_per_example": loss_sum / max(n_ex, 1.0), "logged_train_acc": corr / max(rows, 1.0)}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--bsz", type=int, default=16, help="examples per GPU per step (config.py:27; the global batch is bsz x world)")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--wd", type=float, default=3e-7)
    ap.add_argument("--clip", type=float, default=10.0)
    ap.add_argument("--val-every", type=int, default=800, help="validate every this many steps (main.py:99: 800)")
    ap.add_argument("--ckpt", default="stage_b200_ckpt.pth")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--seed", type=int, default=2018)
    ap.add_argument("--synthetic", type=int, nargs=5, metavar=("N_TRAIN", "T", "R", "LQ", "LS"), default=[64, 30, 10, 16, 20])
    ap.add_argument("--n-valid", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.1, help="config.py:105")
    args = ap.parse_args(argv)

    import torch.distributed as dist
    from . import collate, host
    from .model import STAGE
    from .train import Trainer
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(args.seed)  # same initial weights on every rank
    model = STAGE(host.default_opt(dropout=args.dropout, use_sup_att=False)).to(device)  # synthetic items carry no attention labels
    trainer = Trainer(model, lr=args.lr, wd=args.wd, clip=args.clip)
    n_train, T, R, Lq, Ls = args.synthetic
    train_items = collate.synthetic_items(n_train, T, R, Lq, Ls, seed=args.seed)
    valid_items = collate.synthetic_items(args.n_valid, T, R, Lq, Ls, seed=args.seed + 1)
    max_len = dict(max_qa_l=Lq, max_vid_l=T, max_sub_l=Ls)
    start_epoch, best = 0, 0.0
    if args.resume and os.path.exists(args.ckpt):
        st = load_checkpoint(args.ckpt, trainer)
        start_epoch, best = st["epoch"] + 1, st["best_acc"]
    val_idx = shard_indices(len(valid_items), world, rank, 0, shuffle=False) if len(valid_items) >= world else (list(range(len(valid_items))) if rank == 0 else [])
    for epoch in range(start_epoch, args.epochs):
        def on_step(step):
            if args.val_every and step % args.val_every == 0:
                v = validate(model, valid_items, val_idx, args.bsz, max_len, device)
                if rank == 0:
                    print("step %d valid acc %.4f (%d examples)" % (step, v["acc"], v["n"]), flush=True)
        stats = train_epoch(trainer, train_items, shard_indices(len(train_items), world, rank, epoch, args.seed), args.bsz, max_len, device,
                            log_every=10, on_step=on_step)
        v = validate(model, valid_items, val_idx, args.bsz, max_len, device)
        if rank == 0:
            print("epoch %d: %.1f s, logged loss/example %.4f, train acc %.4f, valid acc %.4f" % (
                epoch, stats["seconds"], stats["logged_loss_per_example"], stats["logged_train_acc"], v["acc"]), flush=True)
            best = max(best, v["acc"])
            save_checkpoint(args.ckpt, trainer, epoch, best)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
