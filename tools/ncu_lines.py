#!/usr/bin/env python
"""Per SOURCE LINE totals of an ncu report (built with -lineinfo, captured with --import-source on):
warp-instructions executed, average active threads and stall samples, top N lines.
usage: ncu_lines.py <rep> <kernel-regex> [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
lines = []
n_fun, cur_file = 0, ""
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("File Path", "Function Name", "Line No"):
        if r and r[0] == "File Path" and len(r) > 1:
            cur_file = r[1].split("/")[-1]
        if r and r[0] == "Function Name":
            n_fun += 1
            if n_fun > 1 and "--all-files" not in sys.argv:
                pass  # every source file of the kernel has its own section: all of them are summed
        continue
    if r[0] == "":
        continue
    try:
        lines.append((int(r[0]), (cur_file + ": " if cur_file else "") + r[1].strip(), int(r[ie] or 0),
                      int(r[te] or 0), int(r[sm] or 0)))
    except ValueError:
        pass
tot_i = sum(x[2] for x in lines)
tot_s = sum(x[4] for x in lines)
print(f"{len(lines)} source lines, {tot_i} warp-instructions, {tot_s} samples")
for ln, src, i, t, s in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{ln:5d} {100 * i / max(tot_i, 1):5.1f}% inst  {t / max(i, 1):5.1f} thr  {100 * s / max(tot_s, 1):5.1f}% smp  {src[:110]}")
