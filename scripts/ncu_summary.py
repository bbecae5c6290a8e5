"""Summarise an ncu report of the step kernel: headline metrics + stall mix + hottest SASS lines.
   python scripts/ncu_summary.py gpurun_out/<tag>_step_full.ncu-rep [out.json]"""
import csv, collections, json, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; vals = rows[2] if len(rows) > 2 else rows[1]; units = rows[1]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
def num(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = {"kernel": d.get("Kernel Name"), "report": rep}
for k in keys:
    if k in d: out[k] = {"value": num(k), "unit": u.get(k)}
stalls = {}
for h in hdr:
    if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        v = num(h)
        if v and v > 0.02: stalls[h.split("issue_stalled_")[1].split("_per_issue")[0]] = round(v, 3)
out["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}
tot = collections.Counter(); lines = []
for r in rows[2:]:
    try: ex = int(r[ix["Instructions Executed"]])
    except Exception: continue
    smp = int(r[ix["# Samples"]] or 0)
    for h in h2:
        if h.startswith("stall_") and "Not Issued" not in h: tot[h] += int(r[ix[h]] or 0)
    lines.append((smp, ex, r[ix["Source"]].strip()))
out["static_instructions"] = len(lines)
out["never_executed"] = sum(1 for l in lines if l[1] == 0)
out["stall_samples"] = dict(tot.most_common(8))
out["hottest_sass"] = [{"samples": a, "executed": b, "sass": c} for a, b, c in sorted(lines, reverse=True)[:25]]
dr, dw = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
def to_bytes(v, unit):
    if v is None: return None
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1)
out["dram_bytes_per_launch"] = (to_bytes(dr, u.get("dram__bytes_read.sum")) or 0) + (to_bytes(dw, u.get("dram__bytes_write.sum")) or 0)
js = json.dumps(out, indent=1)
if len(sys.argv) > 2: open(sys.argv[2], "w").write(js + "\n")
print(js[:6000])
