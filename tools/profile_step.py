"""Runs a few eval forwards of the drop-in STAGE on cuda:0 (for ncu captures; not a benchmark)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from tvqaplus_b200 import host  # noqa: E402
from tvqaplus_b200.model import STAGE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(2018)
model = STAGE(host.default_opt()).to(dev).eval()
shape = {k: bench.CONFIGS["cfg2"][k] for k in ("T", "R", "Lq", "Ls")}
batch = bench.to_device(bench.make_host_batch(a.batch, pin=False, shape=shape), dev)
for _ in range(a.iters):
    model(batch)
torch.cuda.synchronize()
print("done")
