"""tcgen05.mma issue-to-completion rate of one CTA (stage_umma_probe with repeat): SS (A and B from shared memory) against
TS (A from tensor memory) for the tile shapes the STAGE kernels use.  Cycles per K=16 MMA; M = 128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tvqaplus_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")


def rate(n, ksteps, a_tmem, repeat=400, a_pad=0, b_pad=0):
    K = 16 * ksteps
    a_lbo, b_lbo = 128 * 16 + a_pad, n * 16 + b_pad
    a = torch.zeros(128 * K * 2 if a_tmem else (K // 8) * a_lbo, dtype=torch.uint8, device=dev)
    b = torch.zeros((K // 8) * b_lbo, dtype=torch.uint8, device=dev)
    d = torch.zeros(128, n, device=dev)
    cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    args = _lib.ProbeArgs(a_img=a.data_ptr(), b_img=b.data_ptr(), a_bytes=a.numel(), b_bytes=b.numel(), a_lbo=a_lbo, a_sbo=128,
                          a_step=2 * a_lbo, b_lbo=b_lbo, b_sbo=128, b_step=2 * b_lbo, a_mn=2 if a_tmem else 0, b_mn=0,
                          k_steps=ksteps, n=n, use_bulk=0, d_out=d.data_ptr(), repeat=repeat, cycles_out=cyc.data_ptr())
    for _ in range(2):
        _lib.check(_lib.lib().stage_umma_probe(args, _lib.current_stream()), "stage_umma_probe")
    torch.cuda.synchronize()
    _lib.check_device_error()
    return float(cyc.item()) / (repeat * ksteps)


for n in (64, 128, 160, 256):
    print("N=%3d  SS %.1f cycles/MMA   TS %.1f cycles/MMA   (tensor floor %.0f)" % (n, rate(n, 8, False), rate(n, 8, True), 128 * n / 256))

# padded LBO (the +16 / +32 bytes that spread the producers' stores over the banks): do the MMA operand reads pay for it?
for a_pad, b_pad in ((16, 0), (32, 0), (0, 16), (16, 16), (128, 0)):
    print("N=128  A LBO +%3d  B LBO +%3d:  SS %.1f cycles/MMA" % (a_pad, b_pad, rate(128, 8, False, a_pad=a_pad, b_pad=b_pad)))
