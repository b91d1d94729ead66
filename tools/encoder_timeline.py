"""Phase timeline of the encoder kernel (thread 0 of CTA 0): cls-encoder shape (L=100, k=5, masked max-pool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n_seq, L, k = 148 * 24, 100, 5
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
names = ["start", "rows landed", "phase0", "conv0", "sync", "mma0", "epi0", "stat0", "conv1", "sync", "mma1", "epi1", "stat1",
         "-", "tail", "end"]
for i in range(3, 6):
    print("tile", i, " ".join("%s %d" % (names[j], int(t[i][j]) - int(t[i][0])) for j in range(16) if j != 13))
d = (t[4:20, 1:] - t[4:20, :-1]).float().mean(0)
print("mean phase cycles:", {names[j + 1]: int(d[j]) for j in range(15) if j not in (12, 13)}, "period", int((t[20][0] - t[4][0]) / 16))
