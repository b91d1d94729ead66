"""One launch of the encoder kernel at the cls-encoder shape of configs[1] (for ncu; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n_seq, L, k = 48000, 100, 5
S = ops.blocked_stride(L)
x = torch.randn(n_seq * 32 * S * 4, device=dev)
mask = torch.ones(n_seq, L, device=dev)
pe = ops.blocked_pe(torch.randn(500, 128, device=dev), L)
convs = [dict(ln_g=torch.ones(128, device=dev), ln_b=torch.zeros(128, device=dev), dw_w=torch.randn(128, 1, k, device=dev),
              dw_b=torch.zeros(128, device=dev), pw_packed=ops.pack_weight(torch.randn(128, 128, device=dev) * 0.05, 64),
              pw_b=torch.zeros(128, device=dev)) for _ in range(2)]
fin = (torch.ones(128, device=dev), torch.zeros(128, device=dev))
for _ in range(2):
    ops.encoder_block_fwd(x, mask, pe, convs, fin, k, out_mode=1, blocked_in=(n_seq, L))
torch.cuda.synchronize()
