"""Summarise an ncu report: headline metrics, stall reasons, opcode histogram, hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size"]
for i, n in enumerate(h):
    if n in keys or any(n == k for k in keys):
        print("%-70s %-10s %s" % (n, u[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "SASS instrs", len(data))
reasons = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
agg = {x: sum(int(r[ix[x]] or 0) for r in data) for x in reasons}
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
op = collections.Counter()
ex = collections.Counter()
for r in data:
    s = r[ix["Source"]].strip().split()
    o = s[1] if s and s[0].startswith("@") else (s[0] if s else "")
    op[o.split(".")[0]] += int(r[ix["# Samples"]])
    ex[o.split(".")[0]] += int(r[ix["Instructions Executed"]])
print("samples by opcode:", op.most_common(18))
print("executed by opcode:", ex.most_common(18))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(r[ix["# Samples"]], r[ix["Source"]].strip()[:80], "| exec", r[ix["Instructions Executed"]])
