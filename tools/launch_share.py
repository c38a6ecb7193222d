#!/usr/bin/env python
"""Kernel share of the profiled window from an `ncu --metrics gpu__time_duration.sum,... --csv` launch list.
usage: python tools/launch_share.py gpurun_out/<tag>_launches.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
t, rd, wr, cnt = {}, {}, {}, {}
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")[:70]
    metric, val = r[12], float(r[14].replace(",", ""))
    if metric == "gpu__time_duration.sum":
        unit = r[13]
        t[name] = t.get(name, 0.0) + val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(unit, 1e-6)
        cnt[name] = cnt.get(name, 0) + 1
    elif metric == "dram__bytes_read.sum":
        rd[name] = rd.get(name, 0.0) + val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1)
    elif metric == "dram__bytes_write.sum":
        wr[name] = wr.get(name, 0.0) + val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1)
tot = sum(t.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {tot:.3f} ms of kernel time (ncu-serialised, cold cache)")
print(f"{'kernel':72s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
for k in sorted(t, key=lambda k: -t[k]):
    print(f"{k:72s} {cnt[k]:8d} {t[k]:9.3f} {100 * t[k] / tot:6.1f}% {rd.get(k, 0) / 1e6:11.1f} {wr.get(k, 0) / 1e6:11.1f}")
