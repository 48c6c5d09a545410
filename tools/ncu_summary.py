#!/usr/bin/env python
"""Summarise an .ncu-rep (one line per profiled launch) into profiles/: python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "sm__cycles_active.avg"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEYS)
        w.writerow(["(unit)"] + [units[hdr.index(k)] if k in hdr else "" for k in KEYS])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            w.writerow([d.get("Kernel Name", "")[:80]] + [d.get(k, "") for k in KEYS])
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
