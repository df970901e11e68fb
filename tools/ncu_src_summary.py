#!/usr/bin/env python
"""Summarise the source page of an ncu report: top SASS instructions by stall samples and the
stall-reason totals.  usage: ncu_src_summary.py <rep> <kernel-regex> [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[hdr_i - 1][1] if hdr_i else "")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
S = col["# Samples"]
tot = sum(int(r[S] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
print("total samples", tot, "instructions", len(body))
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(tot,1):.1f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
ie = col["Instructions Executed"]
print("executed warp-instr total", sum(int(r[ie] or 0) for r in body))
print(f"top {top} by samples:")
for r in sorted(body, key=lambda r: -int(r[S] or 0))[:top]:
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{int(r[S]):7d} {100*int(r[S])/max(tot,1):5.1f}%  exec={r[ie]:>9}  {r[1][:90]:90s} {why}")
