"""Phase timeline of the encoder kernel (thread 0 of CTA 0): cls-encoder shape (L=100, k=5, masked max-pool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n_seq, L, k = 148 * 2 * 26, 100, 5
x = torch.randn(n_seq, L, 128, device=dev)
mask = torch.ones(n_seq, L, device=dev)
pe = torch.randn(500, 128, device=dev)
convs = [dict(ln_g=torch.ones(128, device=dev), ln_b=torch.zeros(128, device=dev), dw_w=torch.randn(128, 1, k, device=dev),
              dw_b=torch.zeros(128, device=dev), pw_packed=ops.pack_weight(torch.randn(128, 128, device=dev) * 0.05, 64),
              pw_b=torch.zeros(128, device=dev)) for _ in range(2)]
fin = (torch.ones(128, device=dev), torch.zeros(128, device=dev))
tl = torch.zeros(32, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.encoder_block_fwd(x, mask, pe, convs, fin, k, out_mode=1, timeline=tl)
torch.cuda.synchronize()
t = tl.cpu()
names = ["start", "phase0", "conv0", "mma0", "epi0", "-", "conv1", "mma1", "epi1", "-", "-", "-", "lnorm", "pool", "end", "-"]
order = [0, 1, 2, 3, 4, 6, 7, 8, 12, 13, 14]
for i in range(3, 6):
    print("tile", i, " ".join("%s %d" % (names[j], int(t[i][j]) - int(t[i][0])) for j in order))
d = {names[b]: int((t[4:20, b] - t[4:20, a]).float().mean()) for a, b in zip(order[:-1], order[1:])}
print("mean phase cycles:", d, "period", int((t[20][0] - t[4][0]) / 16))
