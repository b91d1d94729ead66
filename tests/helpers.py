"""Shared test helpers: golden fixtures, seeded weights/inputs, the parity norm."""
import hashlib
import os

import numpy as np
import torch

from oracle import stage_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny_h0", "tiny_h4", "cfg1_h0", "cfg1_h4", "wide_h0"]
REL_TOL = 1e-3  # BASELINE.json north_star: 1e-3 relative fp32; norm = max|x - ref| / max|ref| per output tensor


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def digest(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def batch_digest(b) -> str:
    return digest([b.qas_bert, b.qas_mask, b.sub_bert, b.sub_mask, b.vid, b.vid_mask, b.ts_label_mask,
                   b.target.float(), b.ts_label["st"].float(), b.ts_label["ed"].float()])


def seeded_model_and_batch(name, gold=None):
    """Product STAGE (CPU-constructed, default init under seed 2018) + the synthetic batch of a golden case."""
    from tvqaplus_b200.model import STAGE
    gold = gold or load_golden(name)
    N, T, R, Lq, Ls, heads = [int(v) for v in gold["shape"]]
    opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads)
    torch.manual_seed(2018)
    model = STAGE(opt).eval()
    batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=2018)
    return opt, model, batch


def max_rel(got, want) -> float:
    got = torch.as_tensor(got).double()
    want = torch.as_tensor(want).double()
    return float((got - want).abs().max() / max(1e-12, float(want.abs().max())))


def masked_rel(got, want, neg=-1e9):
    """Masked (-1e10) entries must match exactly; the rest within the max-norm relative tolerance."""
    got, want = torch.as_tensor(got), torch.as_tensor(want)
    m = want <= neg
    assert torch.equal(got <= neg, m), "mask pattern differs"
    assert torch.equal(got[m], want[m]), "masked entries are not exactly -1e10"
    return max_rel(got.masked_fill(m, 0), want.masked_fill(m, 0))


def batch_to(batch, device):
    out = O.Batch(batch)
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.to(device)
    out["ts_label"] = {k: v.to(device) for k, v in batch["ts_label"].items()}
    return out
