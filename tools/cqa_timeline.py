"""Phase timeline of the fused CQA kernel (CTA 0, first tiles): SM-clock stamps taken by compute thread 0 and the MMA thread."""
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
c, m = tl[0], tl[1]
names_c = ["start", "pre-D1", "D1", "softmax done", "E2 sent", "D2", "c2q prod done", "stats done", "D3", "cat prod done", "D4", "end"]
names_m = ["start", "AQ landed", "S issued", "E2 seen", "U done", "c2q-sub issued", "D3 commit", "D4 commit", "D4 done"]
t0 = int(c[2][0])
for tile in range(2, 8):
    print("tile", tile)
    print("  compute:", "  ".join("%s %d" % (names_c[i], int(c[tile][i]) - t0) for i in range(12)))
    print("  mma    :", "  ".join("%s %d" % (names_m[i], int(m[tile][i]) - t0) for i in range(9)))
d = (c[3:60, 1:12] - c[3:60, 0:11]).float().mean(0)
print("compute mean phase cycles:", {names_c[i + 1]: int(d[i]) for i in range(11)}, "tile total", int((c[4:60, 0] - c[3:59, 0]).float().mean()))
dm = (m[3:60, 1:9] - m[3:60, 0:8]).float().mean(0)
print("mma mean phase cycles:", {names_m[i + 1]: int(dm[i]) for i in range(8)})

# per-stage hand-off of the first 8 A stages (c2q): producer (thread 0) ready-to-store / arrived, MMA warp saw full / issued
ps, ms = tl[2], tl[3]
for tile in (4, 5):
    base = int(c[tile][5])  # D2 seen by the compute thread
    print("tile", tile, "stages (cycles after D2): " + "  ".join(
        "s%d prod %d/%d mma %d/%d" % (k, int(ps[tile][2 * k]) - base, int(ps[tile][2 * k + 1]) - base, int(ms[tile][2 * k]) - base,
                                      int(ms[tile][2 * k + 1]) - base) for k in range(8)))

# skew between compute warps: the same phase stamps taken by warp 15 (row 4) and warp 5 (row 7), relative to warp 0's
c15, c5 = tl[4], tl[7]
for tile in (4, 5):
    print("tile", tile, "warp 15 minus warp 0 per phase:", "  ".join("%s %+d" % (names_c[i], int(c15[tile][i]) - int(c[tile][i])) for i in range(12)))
    print("tile", tile, "warp 5 minus warp 0 per phase: ", "  ".join("%s %+d" % (names_c[i], int(c5[tile][i]) - int(c[tile][i])) for i in range(12)))
    base = int(c[tile][5])
    p15 = tl[6]
    print("tile", tile, "warp 15 stages (cycles after warp 0's D2): " + "  ".join(
        "s%d %d/%d" % (k, int(p15[tile][2 * k]) - base, int(p15[tile][2 * k + 1]) - base) for k in range(8)))
# epilogue detail of thread 0 (row 5): accumulator read, statistics done, barrier passed, stores issued — cycles after D4
ep = tl[5]
d4 = c[3:60, 10]
print("epilogue (mean cycles after D4): acc read %d  stats %d  barrier %d  stored %d  end %d" % tuple(
    int((x - d4).float().mean()) for x in (ep[3:60, 0], ep[3:60, 1], ep[3:60, 2], ep[3:60, 3], c[3:60, 11])))
