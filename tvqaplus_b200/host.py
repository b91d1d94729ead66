"""Host-side helpers around the kernels: the option namespace and batch container the reference passes to
``STAGE`` (config.py:15-134, tvqa_dataset.py:631-688), a synthetic batch of that schema, and the example-wise
sharding used when one process per GPU runs the forward (the path shards over QA examples; no data-path collective).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List, Sequence, Tuple

import torch

# batch fields by kind (tvqa_dataset.py:592-688)
TENSOR_FIELDS = ("qas_bert", "qas_mask", "sub_bert", "sub_mask", "vid", "vid_mask", "ts_label_mask", "target", "qas")
LIST_FIELDS = ("qid", "vid_name", "att_labels", "anno_st_idx", "q_l", "image_indices", "boxes")


def default_opt(**over) -> SimpleNamespace:
    """The ``opt`` fields ``STAGE.__init__`` reads, with the values ``run_main.sh --add_local`` produces
    (config.py:20-134, run_main.sh:34-46)."""
    o = dict(sub_flag=True, vfeat_flag=True, vfeat_size=300, embedding_size=768, hsz=128, dropout=0.1, t_iter=0,
             extra_span_length=3, add_local=True, use_sup_att=True, scale=10.0, alpha=20.0, margin=0.1,
             att_loss_type="lse", num_negatives=2, num_hard=2, negative_pool_size=0, drop_topk=0, add_non_visual=False,
             input_encoder_n_blocks=1, input_encoder_n_conv=2, input_encoder_kernel_size=7, input_encoder_n_heads=0,
             cls_encoder_n_blocks=1, cls_encoder_n_conv=2, cls_encoder_kernel_size=5, cls_encoder_n_heads=0)
    o.update(over)
    return SimpleNamespace(**o)


class Batch(dict):
    """Attribute-access dict with the interface of the ``easydict.EasyDict`` the reference's ``prepare_inputs``
    returns (tvqa_dataset.py:631-688)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def synthetic_batch(N: int, T: int, R: int, Lq: int, Ls: int, *, seed: int = 2018, pin: bool = False,
                    wd: int = 768, vd: int = 300) -> Batch:
    """Dense (all masks 1) synthetic batch with the ``prepare_inputs`` schema, on the host."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        t = torch.empty(*shape, pin_memory=pin)
        t.normal_(generator=g)
        return t

    def ones(*shape):
        t = torch.ones(*shape)
        return t.pin_memory() if pin else t

    return Batch(qid=list(range(N)), vid_name=["syn_%d" % i for i in range(N)], qas_bert=rnd(N, 5, Lq, wd),
                 sub_bert=rnd(N, T, Ls, wd), vid=rnd(N, T, R, vd), qas_mask=ones(N, 5, Lq), sub_mask=ones(N, T, Ls),
                 vid_mask=ones(N, T, R), ts_label_mask=ones(N, T), target=torch.randint(0, 5, (N,), generator=g),
                 ts_label={"st": torch.randint(0, max(1, T // 2), (N,), generator=g),
                           "ed": torch.randint(T // 2, T, (N,), generator=g)},
                 qas=torch.ones(N, 5, Lq, dtype=torch.long), att_labels=None, anno_st_idx=[0] * N, q_l=[Lq // 2] * N,
                 image_indices=None, boxes=None, use_hard_negatives=False, eval_object_word_ids=None)


def synthetic_att_labels(batch: Dict[str, Any], n_img_max: int = 4, seed: int = 0, density: float = 0.08):
    """Synthetic ``att_labels`` / ``anno_st_idx`` for a synthetic batch, in the form tvqa_dataset.py:362-420 builds them from the
    bounding-box annotations: per example a few annotated frames inside the clip, each a (Lq_valid, R_valid) 0/1 matrix."""
    g = torch.Generator().manual_seed(seed)
    n = batch["qas_mask"].shape[0]
    labels, starts = [], []
    for b in range(n):
        n_frames = int(batch["ts_label_mask"][b].sum())
        n_img = int(torch.randint(1, min(n_img_max, n_frames) + 1, (1,), generator=g))
        st = int(torch.randint(0, n_frames - n_img + 1, (1,), generator=g))
        q_len = int(batch["qas_mask"][b, int(batch["target"][b])].sum())
        cur = []
        for i in range(n_img):
            r_valid = max(2, int(batch["vid_mask"][b, st + i].sum()))
            lab = (torch.rand(q_len, r_valid, generator=g) < density).float()
            lab[0, 0], lab[0, 1] = 1.0, 0.0  # at least one positive and one negative on one word
            cur.append(lab)
        labels.append(cur)
        starts.append(st)
    return labels, starts


def batch_to(batch: Dict[str, Any], device, non_blocking: bool = False) -> Batch:
    """``prepare_inputs``' final ``.to(device)`` step (tvqa_dataset.py:660-688) for an already padded batch."""
    out = Batch(batch)
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
    if batch.get("ts_label") is not None:
        out["ts_label"] = {k: v.to(device, non_blocking=non_blocking) for k, v in batch["ts_label"].items()}
    return out


# ----------------------------------------------------------------------------------------------------------------
# data parallelism: contiguous example ranges per rank (replaces nn.DataParallel's scatter, main.py:204-206)
# ----------------------------------------------------------------------------------------------------------------
def shard_bounds(n_examples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the examples rank ``rank`` owns: contiguous, sizes differ by at most one, earlier ranks larger."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("shard_bounds: rank %d of world %d" % (rank, world_size))
    base, extra = divmod(n_examples, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, Any], world_size: int, rank: int) -> Batch:
    """The slice of ``batch`` (any mix of tensors / per-example lists / scalars) that ``rank`` processes."""
    n = len(batch["qid"])
    lo, hi = shard_bounds(n, world_size, rank)
    out = Batch()
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == n:
            out[k] = list(v[lo:hi])
        elif isinstance(v, dict):
            out[k] = {kk: (vv[lo:hi] if torch.is_tensor(vv) and vv.dim() > 0 and vv.shape[0] == n else vv)
                      for kk, vv in v.items()}
        else:
            out[k] = v
    return out


def gather_examples(local: torch.Tensor, n_examples: int, group=None) -> torch.Tensor:
    """All-gather a per-example result (leading axis = this rank's examples) back into batch order on every rank.
    Used only at the edge (metrics / predictions); the forward itself never communicates."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_examples, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def merge_predictions(per_rank: Sequence[dict]) -> dict:
    """Concatenate the inference-mode dicts (model/stage.py:297-312) of consecutive shards back into one batch."""
    out = {"answer": torch.cat([d["answer"] for d in per_rank], 0), "t_scores": torch.cat([d["t_scores"] for d in per_rank], 0)}
    att = [d.get("att_predictions") for d in per_rank]
    out["att_predictions"] = None if any(a is None for a in att) else [q for a in att for q in a]
    return out
