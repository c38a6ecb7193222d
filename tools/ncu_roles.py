#!/usr/bin/env python
"""Warp-stall samples of the marching conv kernel split by ROLE (code region): MMA issuer, input transform, the live epilogue
variant, producers.  Needs a report captured with --set full --import-source on.
usage: python tools/ncu_roles.py report.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if rr:
    h0 = rr[0]
    for key in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__cycles_elapsed.max.per_second"):
        if key in h0:
            print(f"# {key}: {[r[h0.index(key)][:60] for r in rr[2:]]}")
rows = list(csv.reader(io.StringIO(out)))
hdr, L = None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr:
            break
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr and len(r) > 10:
        try:
            int(r[hdr["Warp Stall Sampling (All Samples)"]])
        except ValueError:
            continue
        L.append(r)
S = hdr["Warp Stall Sampling (All Samples)"]
stall_cols = [(h, i) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in L)
tanh = [i for i, r in enumerate(L) if "MUFU.TANH" in r[1]]
mma = [i for i, r in enumerate(L) if "UTCHMMA" in r[1]]
ldtm = [i for i, r in enumerate(L) if "LDTM" in r[1]]


def summarize(lo, hi, name, warps):
    c, n = collections.Counter(), 0
    for r in L[lo:hi]:
        n += int(r[S])
        for h, i in stall_cols:
            try:
                c[h[6:]] += int(r[i])
            except ValueError:
                pass
    if n:
        print(f"{name}: {hi - lo} instr ({(hi - lo) * 16} B), {n} samples = {100 * n / tot:.1f}% of the kernel's "
              f"({100 * n / tot * 16 / warps:.0f}% of the time of its {warps} warp(s)); "
              + ", ".join(f"{k} {100 * v / n:.0f}%" for k, v in c.most_common(6)))


print(f"# {tot} samples, 16 warps per CTA")
summarize(0, mma[0] - 300, "producers + prologue", 2)
summarize(mma[0] - 300, mma[-1] + 80, "MMA issuer", 1)
if tanh:
    summarize(mma[-1] + 80, tanh[-1] + 90, "input transform", 4)
nxt = (tanh[-1] + 90) if tanh else (mma[-1] + 80)
for k in range(0, len(ldtm), 2):
    lo = max(nxt, ldtm[k] - 60) if k else nxt
    hi = ldtm[k + 2] - 60 if k + 2 < len(ldtm) else len(L)
    summarize(lo, hi, f"epilogue variant {k // 2}", 8 if tanh else 12)
