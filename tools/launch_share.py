#!/usr/bin/env python
"""Kernel share of the profiled window from an `ncu --metrics gpu__time_duration.sum,... --csv` launch list (with or without
the NVTX columns).  usage: python tools/launch_share.py gpurun_out/<tag>_launches.csv [--json out.json]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
rows = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
t, rd, wr, cnt = {}, {}, {}, {}
for r in rows:
    name = r[ik].split("(")[0].replace("void ", "")[:70]
    metric, val = r[im], float(r[iv].replace(",", ""))
    if metric == "gpu__time_duration.sum":
        t[name] = t.get(name, 0.0) + val * TIME.get(r[iu], 1e-6)
        cnt[name] = cnt.get(name, 0) + 1
    elif metric == "dram__bytes_read.sum":
        rd[name] = rd.get(name, 0.0) + val * BYTES.get(r[iu], 1)
    elif metric == "dram__bytes_write.sum":
        wr[name] = wr.get(name, 0.0) + val * BYTES.get(r[iu], 1)
tot = sum(t.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {tot:.3f} ms of kernel time (ncu-serialised, cold cache)")
print(f"{'kernel':72s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
for k in sorted(t, key=lambda k: -t[k]):
    print(f"{k:72s} {cnt[k]:8d} {t[k]:9.3f} {100 * t[k] / tot:6.1f}% {rd.get(k, 0) / 1e6:11.1f} {wr.get(k, 0) / 1e6:11.1f}")
conv = [k for k in t if "conv3d_planar" in k or "conv3d_march" in k]
summary = {"conv_launches": sum(cnt[k] for k in conv), "conv_ms_under_ncu": sum(t[k] for k in conv),
           "conv_dram_bytes_read": sum(rd.get(k, 0) for k in conv), "conv_dram_bytes_written": sum(wr.get(k, 0) for k in conv),
           "all_launches": sum(cnt.values()), "all_ms_under_ncu": tot, "all_dram_bytes": sum(rd.values()) + sum(wr.values())}
print(f"# conv3d_planar_kernel + conv3d_march_kernel: {summary['conv_launches']} launches, {summary['conv_ms_under_ncu']:.3f} ms = "
      f"{100 * summary['conv_ms_under_ncu'] / tot:.1f}% of the window, DRAM {summary['conv_dram_bytes_read'] / 1e9:.2f} GB read + "
      f"{summary['conv_dram_bytes_written'] / 1e9:.2f} GB written")
if "--json" in sys.argv:
    json.dump(summary, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
