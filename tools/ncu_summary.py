#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text report (run where ncu is installed)."""
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = [r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.avg$", r"^launch__registers_per_thread$",
        r"^launch__grid_size$", r"^launch__block_size$", r"^launch__occupancy_limit",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
        r"^sm__inst_executed_pipe_(xu|fma|alu|lsu|uniform|tmem|tc)\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_shared_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct",
        r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct", r"^gpu__dram_throughput",
        r"^lts__t_sectors\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_sectors_srcunit_tex_op_(red|atom)\.sum$",
        r"^lts__t_sector_hit_rate\.pct$", r"^smsp__issue_active\.avg\.pct", r"^sm__warps_active\.avg\.pct",
        r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$", r"^smsp__inst_executed\.sum$"]


def main(rep, top=30):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {kname}")
    for h, u, v in zip(hdr, units, vals):
        if any(re.search(k, h) for k in KEYS):
            try:
                if float(v) == 0 and "stalled" in h:
                    continue
            except ValueError:
                pass
            print(f"  {h} [{u}] = {v}")
    src = page(rep, "source")
    hdr = src[1]
    ia, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [(int(r[ia] or 0), r[isrc].strip(), int(r[iex] or 0), i) for i, r in enumerate(src[2:]) if len(r) > ia]
    tot = sum(d[0] for d in data) or 1
    print(f"warp-stall samples: {tot}; top {top} SASS instructions:")
    for s, text, ex, i in sorted(data, reverse=True)[:top]:
        print(f"  {100 * s / tot:5.1f}%  exec={ex:9d}  #{i:<5d} {text[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
