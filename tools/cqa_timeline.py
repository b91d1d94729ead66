"""Phase timeline of the fused CQA kernel (CTA 0, first tiles): SM-clock stamps taken by compute threads 0 / 160 / 480,
the MMA warp and epilogue thread 640 (see the CQA_STAMP calls in csrc/cqa.cuh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tvqaplus_b200 import ops  # noqa: E402

N, T, Lq, Ls, R = 8, 300, 100, 100, 20  # CTA 0 sees 17 frames x 4 packed tiles >= the 64 stamped tiles
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = torch.randn(N, 5, Lq, 128, generator=g).to(dev)
s = torch.randn(N, T, Ls, 128, generator=g).to(dev)
v = torch.randn(N, T, R, 128, generator=g).to(dev)
ones = lambda *sh: torch.ones(*sh, device=dev)
w = dict(wblob=ops.cqa_prepare_weights(torch.ones(384, device=dev), torch.zeros(384, device=dev), torch.randn(128, 384, device=dev) * 0.05,
                                       torch.zeros(128, device=dev), torch.ones(384, device=dev), torch.zeros(384, device=dev),
                                       torch.randn(128, 384, device=dev) * 0.05, torch.zeros(128, device=dev)),
         cat_out_ln_g=torch.ones(128, device=dev), cat_out_ln_b=torch.zeros(128, device=dev))
tl = torch.zeros(8, 64, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.cqa_fused_fwd(a, ones(N, 5, Lq), s, ones(N, T, Ls), v, ones(N, T, R), w, 10.0, timeline=tl, blocked=True)  # the statement layout STAGE.forward uses
torch.cuda.synchronize()
tl = tl.cpu()
c, m, ep = tl[0], tl[1], tl[5]
names_c = ["start", "pre-D1", "D1", "softmax done", "E2 sent", "D2", "c2q prod done", "stats done", "D3a", "ms done", "D3b", "mv done", "product done", "D4", "end"]
names_m = ["start", "E2 seen", "U issued", "c2q-sub issued", "c2q-vid issued", "concat issued", "next S issued"]
names_e = ["handed over", "-", "T2 released", "stored"]
t0 = int(c[2][0])
for tile in range(2, 7):
    print("tile", tile)
    print("  compute :", "  ".join("%s %d" % (names_c[i], int(c[tile][i]) - t0) for i in range(15)))
    print("  mma     :", "  ".join("%s %d" % (names_m[i], int(m[tile][i]) - t0) for i in range(7)))
    print("  epilogue:", "  ".join("%s %d" % (names_e[i], int(ep[tile][i]) - t0) for i in range(4)))
d = (c[3:60, 1:15] - c[3:60, 0:14]).float().mean(0)
print("compute mean phase cycles:", {names_c[i + 1]: int(d[i]) for i in range(14)}, "tile total", int((c[4:60, 0] - c[3:59, 0]).float().mean()))
dm = (m[3:60, 1:7] - m[3:60, 0:6]).float().mean(0)
print("mma mean phase cycles:", {names_m[i + 1]: int(dm[i]) for i in range(6)})
print("epilogue warps, mean cycles after the hand-over: accumulator released %d, stored %d" % (
    int((ep[3:60, 2] - ep[3:60, 0]).float().mean()), int((ep[3:60, 3] - ep[3:60, 0]).float().mean())))
print("mma warp, mean cycles per tile spent waiting: for weights %d, for operands %d" % (int(m[3:60, 14].float().mean()), int(m[3:60, 15].float().mean())))
# a*U stages (thread 0: ready-to-store / arrived) and the MMA warp's view of the sub stream (stage seen full / issued, U operand seen)
ps, ms = tl[2], tl[3]
for tile in (4, 5):
    base = int(c[tile][5])  # D2 seen by compute thread 0
    print("tile", tile, "a*U stages, cycles after D2 (thread 0 ready/arrived): " + "  ".join(
        "s%d %d/%d" % (k, int(ps[tile][2 * k]) - base, int(ps[tile][2 * k + 1]) - base) for k in range(4)))
    print("tile", tile, "mma: sub stage0 full/issued %d/%d  stage1 %d/%d  U operand seen %d  cat segments seen %d %d %d" % tuple(
        int(ms[tile][i]) - base for i in range(8)))
c15, c5 = tl[4], tl[7]
for tile in (4, 5):
    print("tile", tile, "warp 15 minus warp 0 per phase:", "  ".join("%s %+d" % (names_c[i], int(c15[tile][i]) - int(c[tile][i])) for i in range(15)))
    print("tile", tile, "warp 5 minus warp 0 per phase: ", "  ".join("%s %+d" % (names_c[i], int(c5[tile][i]) - int(c[tile][i])) for i in range(15)))

