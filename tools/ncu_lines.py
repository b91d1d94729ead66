"""Executed instructions / stall samples of one kernel per CUDA source line: joins the SASS page of an ncu report with the
line table of the built library (nvdisasm -g).   usage: python tools/ncu_lines.py report.ncu-rep <mangled-substring> <demangled-substring> [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, pat, dpat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tvqaplus_b200", "libstage_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# line table of the kernel: instruction offset -> (file, line) of the innermost //## marker in front of it
line_of, cur, inside = {}, None, False
for l in dis:
    if l.startswith("//---------------------"):
        inside = (".text." in l) and (pat in l)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if "inlined at" not in l:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(starts):
    if dpat not in rows[s][1]:
        continue
    hdr = rows[s + 1]
    data = rows[s + 2:(starts[si + 1] if si + 1 < len(starts) else len(rows))]
    ix = {n: i for i, n in enumerate(hdr)}
    base = int(data[0][ix["Address"]], 16)
    ex, sa = collections.Counter(), collections.Counter()
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        ex[key] += int(r[ix["Instructions Executed"]] or 0)
        sa[key] += int(r[ix["# Samples"]] or 0)
    te, ts = sum(ex.values()), sum(sa.values())
    print(rows[s][1], "executed", te, "samples", ts)
    cache = {}
    for key, e in ex.most_common(top):
        f, ln = key
        text = ""
        path = os.path.join(os.path.dirname(so), "csrc", f)
        if os.path.exists(path):
            if path not in cache:
                cache[path] = open(path).read().split("\n")
            text = cache[path][ln - 1].strip()[:90] if 0 < ln <= len(cache[path]) else ""
        print("%5.1f%% exec %5.1f%% samp  %s:%d  %s" % (100.0 * e / te, 100.0 * sa[key] / max(1, ts), f, ln, text))
    break
