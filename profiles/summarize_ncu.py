#!/usr/bin/env python
"""Compact table from the `ncu --page raw --csv` exports of profiles/capture_r02.sh:
kernel, grid, duration, DRAM GB/s and % of peak, tensor-pipe % (of elapsed cycles), registers.
    python profiles/summarize_ncu.py gpurun_out/r02_*_raw.csv > profiles/r02/ncu_train_kernels.txt"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)"),
        ("launch__registers_per_thread", "regs")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    print("%-14s %-58s %-12s %8s %8s %8s %10s %7s %9s %9s %5s" % ("capture", "kernel", "grid", "us", "MB rd", "MB wr",
                                                                   "DRAM GB/s", "DRAM %", "tens%(el)", "tens%(ac)", "regs"))
    for f in sys.argv[1:]:
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            def val(name):
                if name not in ix or r[ix[name]] == "":
                    return float("nan")
                return float(r[ix[name]].replace(",", "")) * SCALE.get(units[ix[name]], 1.0)
            us = val("gpu__time_duration.sum")
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            name = r[ix["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:58]
            print("%-14s %-58s %-12s %8.1f %8.2f %8.2f %10.0f %7.1f %9.1f %9.1f %5.0f" % (
                f.split("/")[-1].replace("_raw.csv", ""), name, r[ix["Grid Size"]].replace(" ", ""), us, rd, wr,
                (rd + wr) / us * 1e3 if us == us and us > 0 else float("nan"),
                val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                val("launch__registers_per_thread")))


if __name__ == "__main__":
    main()
