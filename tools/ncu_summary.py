#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters the roofline argument uses.
usage: python tools/ncu_summary.py gpurun_out/<file>.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "smsp__warp_issue_stalled", "launch__shared_mem_per_block_dynamic"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i, id_i = hdr.index("Kernel Name"), hdr.index("ID")
    cols = [i for i, h in enumerate(hdr) if any(h == w or h.startswith(w + ".") and h.count(".") <= w.count(".") + 1 for w in WANT)]
    print(f"# {rep}: {len(rows) - 2} launches captured with ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"\n## launch {r[id_i]}: {r[name_i][:80]}")
        for i in cols:
            if r[i] != "":
                print(f"{hdr[i]:78s} {units[i]:16s} {r[i]}")


if __name__ == "__main__":
    main()
