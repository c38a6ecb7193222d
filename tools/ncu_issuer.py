#!/usr/bin/env python
"""SASS-level warp-stall samples of the MMA-issuing thread's code region (the instructions around the UTCHMMA block) of the
first kernel launch in an ncu report captured with --set full --import-source on.
usage: python tools/ncu_issuer.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, per, launch = None, {}, -1
for r in rows:
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        launch += 1
        per[launch] = []
    elif hdr and len(r) > 10:
        try:
            int(r[hdr["Warp Stall Sampling (All Samples)"]])
        except ValueError:
            continue
        per[launch].append(r)
rr = list(csv.reader(io.StringIO(raw)))
if rr:
    h0 = rr[0]
    for key in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__cycles_elapsed.max.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
        if key in h0:
            print(f"# {key}: {[r[h0.index(key)][:60] for r in rr[2:]]}")
L = per[0]
S = hdr["Warp Stall Sampling (All Samples)"]
E = hdr["Instructions Executed"]
idx = [i for i, x in enumerate(L) if "UTCHMMA" in x[1]]
lo, hi = max(0, idx[0] - 260), idx[-1] + 80
reg = L[lo:hi]
tot = sum(int(r[S]) for r in L)
print(f"# launch 0: {tot} samples in the kernel, {sum(int(r[S]) for r in reg)} in the issuer region "
      f"({sum(int(r[E] or 0) for r in reg)} warp instructions executed there)")
stalls, ops = Counter(), Counter()
for r in reg:
    s = r[1].strip()
    ops[s.split()[1] if s.startswith("@") else s.split()[0]] += int(r[S])
    for h, i in hdr.items():
        if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit():
            stalls[h[6:]] += int(r[i])
print("# stall reasons:", stalls.most_common(8))
print("# samples by opcode:", ops.most_common(10))
for r in sorted(reg, key=lambda r: -int(r[S]))[:top_n]:
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:2]
    print(f"{r[0][-5:]} {r[1].strip()[:84]:84s} samples {r[S]:>5s} executed {r[E]:>8s} {st}")
