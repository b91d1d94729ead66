"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only; needs /root/reference).

    python oracle/make_golden.py

The unmodified reference modules are imported from /root/reference/model (py2-style sibling imports, so that directory
goes on sys.path), weights are the reference's default init under torch.manual_seed(2018) (config.py:20), inputs come
from oracle.stage_oracle.make_synthetic_batch.  Stored per case: the reference outputs, plus SHA-256 digests of the
weights and inputs so a test can tell "RNG stream differs on this host" from "numerics differ".
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stage_oracle as O  # noqa: E402

REF = "/root/reference"
CASES = {
    # name: (N, T, R, Lq, Ls, heads)
    "tiny_h0": (2, 6, 5, 8, 7, 0),
    "tiny_h4": (2, 6, 5, 8, 7, 4),
    "cfg1_h0": (2, 40, 25, 40, 50, 0),   # BASELINE.json configs[0]: debug-mode shapes, batch 2
    "cfg1_h4": (2, 40, 25, 40, 50, 4),
    "wide_h0": (1, 9, 40, 100, 100, 0),  # BASELINE configs[1]/[4] token/region widths on a short clip
}


def digest(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def batch_digest(b) -> str:
    return digest([b.qas_bert, b.qas_mask, b.sub_bert, b.sub_mask, b.vid, b.vid_mask, b.ts_label_mask,
                   b.target.float(), b.ts_label["st"].float(), b.ts_label["ed"].float()])


def load_reference():
    sys.path.insert(0, os.path.join(REF, "model"))
    sys.path.insert(0, REF)
    import stage as ref_stage  # noqa
    return ref_stage


def main():
    ref_stage = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_flush_denormal(False)
    for name, (N, T, R, Lq, Ls, heads) in CASES.items():
        opt = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads)
        torch.manual_seed(2018)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref_stage.STAGE(opt)
        model.eval()
        batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=2018)
        rec = dict(shape=np.array([N, T, R, Lq, Ls, heads]), weights_sha=digest(model.state_dict().values()),
                   batch_sha=batch_digest(batch))
        with torch.no_grad():
            out, att_loss, _, t_loss, t_probs, other = model.forward_main(batch)
        rec.update(answer=out.numpy(), t_probs=t_probs.numpy(), t_scores=other["temporal_scores"].numpy(),
                   temporal_loss=np.float32(t_loss.item()), vid_raw_s=other["vid_raw_s"].numpy()[:, :, ::3, ::3],
                   sub_raw_s=other["sub_raw_s"].numpy()[:, :, ::3, ::3, ::3],
                   vid_norm_s=other["vid_normalized_s"].numpy()[:, :, ::3, ::3],
                   answer_argmax=out.argmax(1).numpy())
        # the reference's own span choice (find_max_triples on softmaxed scores), as get_proposals' eval branch does
        probs = torch.softmax(other["temporal_scores"], dim=2).view(N * 5, T, 2)
        sys_utils = sys.modules["model_utils"]
        tri = sys_utils.find_max_triples(probs[:, :, 0], probs[:, :, 1], topN=1)
        tri = [list(t) for t in tri]  # py3: the reference hands back zip iterators (model_utils.py:121)
        rec["spans"] = np.array([[t[0][0], t[0][1]] for t in tri], dtype=np.int64)
        # train-mode forward with dropout off (N_new >= N rows)
        opt_t = O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads, dropout=0.0, use_sup_att=False)
        torch.manual_seed(2018)
        with contextlib.redirect_stdout(io.StringIO()):
            model_t = ref_stage.STAGE(opt_t)
        model_t.train()
        with torch.no_grad():
            (out_t, tgt_t), _, _, t_loss_t, t_scores_t, _ = model_t.forward_main(batch)
        rec.update(train_answer=out_t.numpy(), train_target=tgt_t.numpy(), train_temporal_loss=np.float32(t_loss_t.item()))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "answer", out.numpy().round(4).tolist(), "rows(train)", len(out_t))


if __name__ == "__main__":
    main()
