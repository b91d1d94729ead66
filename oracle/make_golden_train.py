"""Generate tests/golden/train_step_h0.npz from the LIVE reference (build container only; needs /root/reference):
one training step of the unmodified reference - train mode, dropout 0, use_sup_att, main.py:53-61's loss, loss.backward() -
with the loss terms and every parameter's gradient (small tensors whole, large ones as a strided sample plus their sum and
L2 norm, to keep the fixture small).

    python oracle/make_golden_train.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stage_oracle as O  # noqa: E402
from oracle.make_golden import batch_digest, digest, load_reference  # noqa: E402

CASE = dict(N=4, T=14, R=9, Lq=12, Ls=10, batch_seed=9, label_seed=2, att_seed=7, weight_seed=2018, perturb_seed=5)
WHOLE = 4096   # gradients up to this many elements are stored whole
STRIDE = 37    # larger ones: every STRIDE-th element of the flattened tensor


def perturb_(named_parameters, seed):
    """LayerNorm weights / biases (all 1-D parameters) away from 1 / 0, so that their gradient formulas are exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in named_parameters:
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)


def make_inputs():
    c = CASE
    batch = O.make_synthetic_batch(c["N"], c["T"], c["R"], c["Lq"], c["Ls"], seed=c["batch_seed"])
    batch["att_labels"], batch["anno_st_idx"] = O.make_att_labels(batch, seed=c["label_seed"])
    return batch


def sample(g: torch.Tensor) -> np.ndarray:
    flat = g.detach().reshape(-1)
    return (flat if flat.numel() <= WHOLE else flat[::STRIDE]).numpy().copy()


def main():
    ref_stage = load_reference()
    c = CASE
    opt = O.default_opt(dropout=0.0, use_sup_att=True)
    torch.manual_seed(c["weight_seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_stage.STAGE(opt)
    perturb_(ref.named_parameters(), c["perturb_seed"])
    ref.train()
    batch = make_inputs()
    torch.manual_seed(c["att_seed"])
    (out, target), att_loss, _, t_loss, _ = ref(batch)
    cls = F.cross_entropy(out, target, reduction="sum") * (1.0 * c["N"] / len(target))
    loss = cls + 0.1 * att_loss + 0.5 * t_loss
    loss.backward()
    rec = {"loss": np.float32(loss.item()), "cls_loss": np.float32(cls.item()), "att_loss": np.float32(att_loss.item()),
           "temporal_loss": np.float32(t_loss.item()), "n_rows": np.int64(len(target)), "target": target.numpy(),
           "weights_sha": digest(ref.state_dict().values()), "batch_sha": batch_digest(batch),
           "names": np.array([n for n, _ in ref.named_parameters()])}
    for n, p in ref.named_parameters():
        rec["g/" + n] = sample(p.grad)
        rec["s/" + n] = np.array([float(p.grad.double().sum()), float(p.grad.double().norm()), float(p.grad.abs().max())])
    path = os.path.join(ROOT, "tests", "golden", "train_step_h0.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), "rows", len(target))


if __name__ == "__main__":
    main()
