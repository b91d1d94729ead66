"""Per CUDA-source-line aggregation of an ncu report (needs -lineinfo): samples, executed warp instructions.
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        si, ei = r.index("# Samples"), r.index("Instructions Executed")
    elif hdr and len(r) > ei and r[0].isdigit() and r[2] == "-":  # a CUDA source line row (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += int(r[si] or 0)
        a[1] += int(r[ei] or 0)
tot_s = sum(a[0] for a in agg.values()) or 1
tot_e = sum(a[1] for a in agg.values()) or 1
print("total samples %d, executed warp instructions %d" % (tot_s, tot_e))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% exe  %s:%d  %s" % (100 * a[0] / tot_s, 100 * a[1] / tot_e, f, ln, a[2][:110]))
