"""Pins the tcgen05 operand-layout conventions every STAGE kernel relies on (see csrc/umma.cuh header comment).

The probe kernel takes raw shared-memory images + descriptor fields, so each case states a hypothesis about the
hardware convention; the asserted cases are the ones the production kernels use, the others are recorded in
``gpurun_out/umma_probe.json`` for diagnosis.
"""
import json
import os

import numpy as np
import pytest
import torch

from tvqaplus_b200 import _lib

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def kmajor_image(x: np.ndarray, lbo: int) -> np.ndarray:
    """x [R,K] fp16 -> bytes; element (r,k) at (k//8)*lbo + r*16 + (k%8)*2."""
    R, K = x.shape
    img = np.zeros((K // 8) * lbo, dtype=np.uint8)
    v = img.view(np.float16)
    for kc in range(K // 8):
        base = kc * lbo // 2
        for r in range(R):
            v[base + r * 8: base + r * 8 + 8] = x[r, kc * 8: kc * 8 + 8]
    return img


def run_probe(a_img, b_img, n, k_steps, a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, b_mn=0, use_bulk=0, a_mn=0):
    dev = torch.device("cuda:0")
    ta = torch.from_numpy(a_img).to(dev)
    tb = torch.from_numpy(b_img).to(dev)
    d = torch.full((128, n), float("nan"), device=dev)
    args = _lib.ProbeArgs(a_img=ta.data_ptr(), b_img=tb.data_ptr(), a_bytes=ta.numel(), b_bytes=tb.numel(),
                          a_lbo=a_lbo, a_sbo=a_sbo, a_step=a_step, b_lbo=b_lbo, b_sbo=b_sbo, b_step=b_step,
                          a_mn=a_mn, b_mn=b_mn, k_steps=k_steps, n=n, use_bulk=use_bulk, d_out=d.data_ptr())
    _lib.check(_lib.lib().stage_umma_probe(args, _lib.current_stream()), "stage_umma_probe")
    torch.cuda.synchronize()
    _lib.check_device_error()
    return d.cpu().numpy()


def relerr(got, want):
    return float(np.abs(got - want).max() / max(1e-9, np.abs(want).max()))


@pytest.mark.gpu
def test_umma_conventions():
    rng = np.random.default_rng(0)
    report = {}
    K = 32
    A = rng.standard_normal((128, K)).astype(np.float16)

    def pad16(nbytes):
        return (nbytes + 15) // 16 * 16

    # ---- K-major B, several N, padded and unpadded LBO, bulk and generic paths
    for n, pad, bulk in [(64, 0, 0), (64, 32, 0), (256, 16, 0), (48, 16, 1), (112, 32, 1), (160, 16, 1), (144, 0, 0)]:
        B = rng.standard_normal((n, K)).astype(np.float16)
        want = A.astype(np.float32) @ B.astype(np.float32).T
        a_lbo, b_lbo = 128 * 16 + pad, n * 16 + pad
        ai, bi = kmajor_image(A, a_lbo), kmajor_image(B, b_lbo)
        got = run_probe(ai, bi, n, K // 16, a_lbo, 128, 2 * a_lbo, b_lbo, 128, 2 * b_lbo, use_bulk=bulk)
        report["kmajor_n%d_pad%d_bulk%d" % (n, pad, bulk)] = relerr(got, want)
        print("kmajor", n, pad, bulk, report["kmajor_n%d_pad%d_bulk%d" % (n, pad, bulk)], flush=True)

    # ---- MN-major B read out of a K-major image of ctx[j, h]  (U = P . ctx)
    for J, pad in [(32, 16), (112, 32)]:
        P = rng.standard_normal((128, J)).astype(np.float16)
        ctx = rng.standard_normal((J, 128)).astype(np.float16)
        want = P.astype(np.float32) @ ctx.astype(np.float32)
        p_lbo, c_lbo = 128 * 16 + 16, J * 16 + pad
        pi, ci = kmajor_image(P, p_lbo), kmajor_image(ctx, c_lbo)
        got = run_probe(pi, ci, 128, J // 16, p_lbo, 128, 2 * p_lbo, 128, c_lbo, 256, b_mn=1)
        report["mnmajor_j%d" % J] = relerr(got, want)
        print("mnmajor", J, report["mnmajor_j%d" % J], flush=True)

    # ---- A operand from tensor memory (row r in lane r, two fp16 per 32-bit column), B K-major in shared memory
    for n, ksteps in [(128, 2), (128, 8), (64, 4)]:
        Kt = 16 * ksteps
        At = rng.standard_normal((128, Kt)).astype(np.float16)
        B = rng.standard_normal((n, Kt)).astype(np.float16)
        want = At.astype(np.float32) @ B.astype(np.float32).T
        b_lbo = n * 16
        bi = kmajor_image(B, b_lbo)
        got = run_probe(np.ascontiguousarray(At).view(np.uint8).reshape(-1), bi, n, ksteps, 0, 0, 0, b_lbo, 128, 2 * b_lbo, a_mn=2)
        report["a_tmem_n%d_k%d" % (n, Kt)] = relerr(got, want)
        print("a_tmem", n, Kt, report["a_tmem_n%d_k%d" % (n, Kt)], flush=True)

    # ---- B sub-tile addressed by start-address offset (N split 304 = 160 + 144)
    n_full = 304
    B = rng.standard_normal((n_full, K)).astype(np.float16)
    b_lbo = n_full * 16 + 16
    bi = kmajor_image(B, b_lbo)
    a_lbo = 128 * 16 + 32
    ai = kmajor_image(A, a_lbo)
    want = A.astype(np.float32) @ B.astype(np.float32).T
    # second half: rows 160..303 -> image offset 160*16 bytes
    dev_img = np.concatenate([bi[160 * 16:], np.zeros(160 * 16, np.uint8)])
    got = run_probe(ai, dev_img, 144, K // 16, a_lbo, 128, 2 * a_lbo, b_lbo, 128, 2 * b_lbo)
    report["nsplit_second_half"] = relerr(got, want[:, 160:])

    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "umma_probe.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))
    bad = {k: v for k, v in report.items() if not k.endswith("_swapped") and not (v < 1e-3)}
    assert not bad, bad
