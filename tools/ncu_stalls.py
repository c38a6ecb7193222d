#!/usr/bin/env python
"""Warp-stall samples per CUDA source line for each kernel launch of an ncu report
(needs -lineinfo and `ncu --import-source on`).  usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = csv.reader(io.StringIO(out))
launches, cur, seen, hdr, path, func = [], [], set(), None, "", ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]
        if path in seen:                      # a file repeats: next kernel launch starts here
            launches.append((func, cur))
            cur, seen = [], set()
        seen.add(path)
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        stall_cols = [(h, i) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr is not None and r[0].isdigit() and len(r) > 10:
        try:
            n = int(r[hdr["Warp Stall Sampling (All Samples)"]])
        except ValueError:
            continue
        if n > 0:
            st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for h, i in stall_cols), reverse=True)[:2]
            cur.append((n, f"{path.split('/')[-1]}:{r[0]}", r[1].strip()[:95], ", ".join(f"{h[6:]}={v}" for v, h in st if v)))
launches.append((func, cur))
for li, (func, data) in enumerate(launches):
    tot = sum(d[0] for d in data) or 1
    print(f"=== launch {li}: {func[:100]}  total samples {tot}")
    for n, loc, src, st in sorted(data, reverse=True)[:top]:
        print(f"{n:7d} {100.0 * n / tot:5.1f}%  {loc:18s} {src}   [{st}]")
