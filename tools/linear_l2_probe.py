"""Is the 768->300 linear kernel limited by how its A rows come out of HBM?  Per-tile time with x resident in L2 (one tile
per CTA, launched repeatedly) against x streamed from HBM (20 tiles per CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
k, n = 768, 300
w = torch.randn(n, k, device=dev) / k ** 0.5
b = torch.zeros(n, device=dev)
img, cs, cb = ops.pack_weight_ln(w, b, torch.ones(k, device=dev), torch.zeros(k, device=dev))


def run(tiles_per_cta, reps):
    m = 148 * 128 * tiles_per_cta
    x = torch.randn(m, k, device=dev)
    f = lambda: ops.linear_fwd(x, img, None, n, ln_in=(cs, cb), relu=True)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("tiles/CTA %2d: %.3f ms per launch, %.1f us per tile" % (tiles_per_cta, ms, 1e3 * ms / tiles_per_cta))


run(1, 20)
run(2, 20)
run(20, 5)
run(50, 3)
