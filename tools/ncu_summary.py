#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: key raw metrics per profiled launch, the stall-reason
totals and the top stalled SASS instructions.  usage: ncu_summary.py <rep> <kernel-regex> [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {rep}  (ncu --set full --clock-control none; per-launch values, cold cache, serialised)")
for r in rows[2:]:
    print("\nkernel:", r[hdr.index("Kernel Name")])
    for w in WANT:
        if w in hdr and r[hdr.index(w)] not in ("", "n/a"):
            print(f"  {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
except StopIteration:
    sys.exit(0)
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
S, IE = col["# Samples"], col["Instructions Executed"]
tot = sum(int(r[S] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
print(f"\nsource page of the first launch: {len(body)} SASS instructions, {tot} stall samples, "
      f"{sum(int(r[IE] or 0) for r in body)} warp-instructions executed")
print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%"
                                    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
ops = {}
for r in body:
    op = (r[1].split()[1] if r[1].startswith("@") else r[1].split()[0]).split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[IE] or 0)
tie = sum(ops.values())
print("instruction mix: " + ", ".join(f"{k} {100 * v / tie:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10]))
print(f"top {top} instructions by stall samples:")
for r in sorted(body, key=lambda r: -int(r[S] or 0))[:top]:
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"  {100 * int(r[S]) / max(tot, 1):5.1f}%  {r[1][:70]:70s} {why[0][1]}/{why[1][1]}")
