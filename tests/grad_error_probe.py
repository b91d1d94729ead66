"""(test infrastructure, run by hand: python tests/grad_error_probe.py [N T R Lq Ls seed])
Per-parameter gradient error of the CUDA backward and of the fp32 CPU oracle, both against the fp64 CPU oracle
(how much of the CUDA-vs-oracle difference is fp32 rounding amplified by softmax(10 s) and exp(20 d))."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import stage_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_train_gpu import _model_and_state, oracle_grads  # noqa: E402

N, T, R, Lq, Ls, seed = [int(v) for v in sys.argv[1:7]] if len(sys.argv) > 6 else (2, 40, 25, 40, 50, 2018)
opt = O.default_opt(dropout=0.0, use_sup_att=True)
batch = O.make_synthetic_batch(N, T, R, Lq, Ls, seed=seed)
batch["att_labels"], batch["anno_st_idx"] = O.make_att_labels(batch, seed=2)
model, sd = _model_and_state(opt)
_, g32 = oracle_grads(sd, opt, batch, att_seed=7)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
b64 = O.Batch(batch)
for k in ("qas_bert", "sub_bert", "vid", "qas_mask", "sub_mask", "vid_mask", "ts_label_mask"):
    b64[k] = batch[k].double()
_, g64 = oracle_grads(sd64, opt, b64, att_seed=7)
model = model.to("cuda:0").train()
torch.manual_seed(7)
(out, target), att_loss, _, t_loss, _ = model(H.batch_to(batch, torch.device("cuda:0")))
loss = F.cross_entropy(out, target, reduction="sum") * (1.0 * N / len(target)) + 0.5 * t_loss + 0.1 * att_loss
loss.backward()
gs = max(float(v.abs().max()) for v in g64.values())
rows = []
for k, p in model.named_parameters():
    w = g64[k]
    den = max(float(w.abs().max()), 1e-3 * gs)
    rows.append((float((p.grad.cpu().double() - w).abs().max()) / den, float((g32[k].double() - w).abs().max()) / den, k))
rows.sort(reverse=True)
print("%-60s %10s %10s" % ("parameter", "cuda", "cpu fp32"))
for e1, e2, k in rows[:12]:
    print("%-60s %10.2e %10.2e" % (k, e1, e2))
