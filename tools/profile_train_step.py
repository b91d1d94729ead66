"""Two training steps at the cfg3 shape (batch 16 by default) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_train_step.py [batch]"""
import sys
import torch

sys.path.insert(0, ".")
from tvqaplus_b200 import host  # noqa: E402
from tvqaplus_b200.model import STAGE  # noqa: E402
from tvqaplus_b200.train import Trainer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T, R, Lq, Ls = 300, 20, 100, 100
dev = torch.device("cuda:0")
torch.manual_seed(2018)
model = STAGE(host.default_opt(dropout=0.0)).to(dev).train()
trainer = Trainer(model)
hb = host.synthetic_batch(N, T, R, Lq, Ls, seed=2018)
hb["att_labels"], hb["anno_st_idx"] = host.synthetic_att_labels(hb, seed=1)
db = host.batch_to(hb, dev)
for _ in range(2):
    pos, neg = model._att_pairs(hb["att_labels"], hb["target"], hb["anno_st_idx"], model.num_negatives, False, model.drop_topk)
    res = trainer.step(db, trainer.engine.att_offsets_from_pairs(pos, neg, N, T, Lq, R))
torch.cuda.synchronize()
print("loss", float(res["loss"]))
