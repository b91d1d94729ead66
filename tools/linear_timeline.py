"""Per-tile timeline of the linear kernel (CTA 0): producer first/last chunk, MMA start/last commit, epilogue start/end."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for (m, k, n, ln) in ((148 * 128 * 20, 768, 300, True), (148 * 128 * 20, 300, 128, False)):
    x = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.zeros(n, device=dev)
    g = torch.ones(k, device=dev)
    if ln:
        img, cs, cb = ops.pack_weight_ln(w, b, g, b.new_zeros(k))
    else:
        img, cs, cb = ops.pack_weight(w), None, None
    lo = (torch.ones(n, device=dev), torch.zeros(n, device=dev))
    tl = torch.zeros(32, 8, dtype=torch.int64, device=dev)
    for _ in range(2):
        # like the model: the 768->300 layer has its LayerNorm folded in and none behind it, the 300->128 layer ends in LN128
        ops.linear_fwd(x, img, None if ln else b, n, ln_in=(cs, cb) if ln else None, ln_out=None if ln else lo, relu=True, timeline=tl)
    torch.cuda.synchronize()
    t = tl.cpu()
    t0 = int(t[2][0])
    print("k=%d n=%d  (cycles relative to tile 2 producer start)" % (k, n))
    for i in range(2, 8):
        print("  tile %d: prod %d..%d  mma %d..%d  epi %d..%d (setup done %d, blocks done %d)" % ((i,) + tuple(int(t[i][j]) - t0 for j in range(8))))
    print("  mean tile period %d cycles" % int((t[12][0] - t[4][0]) / 8))
