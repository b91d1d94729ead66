"""Stand-alone timing of the encoder kernel at the cls-encoder shape of configs[1] (48 000 x 100 x 128, k=5, masked max-pool,
blocked input) and the sub-title input-encoder shape (9 600 x 100 x 128, k=7, row-major); CUDA events, not a benchmark line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def run(n_seq, L, k, out_mode, blocked):
    S = ops.blocked_stride(L)
    x = torch.randn(n_seq * 32 * S * 4 if blocked else n_seq * L * 128, device=dev)
    if not blocked:
        x = x.view(n_seq, L, 128)
    mask = torch.ones(n_seq, L, device=dev)
    pe = torch.randn(500, 128, device=dev)
    pe_in = ops.blocked_pe(pe, L) if blocked else pe
    convs = [dict(ln_g=torch.ones(128, device=dev), ln_b=torch.zeros(128, device=dev), dw_w=torch.randn(128, 1, k, device=dev),
                  dw_b=torch.zeros(128, device=dev), pw_packed=ops.pack_weight(torch.randn(128, 128, device=dev) * 0.05, 64),
                  pw_b=torch.zeros(128, device=dev)) for _ in range(2)]
    fin = (torch.ones(128, device=dev), torch.zeros(128, device=dev))
    f = lambda: ops.encoder_block_fwd(x, mask, pe_in, convs, fin, k, out_mode=out_mode, blocked_in=(n_seq, L) if blocked else None)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    print("n_seq=%d L=%d k=%d out_mode=%d blocked=%d: %.3f ms" % (n_seq, L, k, out_mode, blocked, e0.elapsed_time(e1) / 5))


run(48000, 100, 5, 1, True)
run(9600, 100, 7, 0, False)
run(9600, 20, 7, 0, False)
