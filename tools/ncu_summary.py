#!/usr/bin/env python
"""Per-kernel summary table of an .ncu-rep (read here, no GPU): duration, grid, registers, smem, DRAM bytes,
tensor-pipe %, achieved occupancy.  ``python tools/ncu_summary.py rep.ncu-rep > profiles/x.txt``"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dsmem_KB"),
        ("sm__cycles_elapsed.max", "cyc_elapsed"), ("sm__cycles_active.avg", "cyc_active_avg"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("lts__t_sector_hit_rate.pct", "l2hit_%")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s  (ncu --set full --clock-control none; durations are serialised, cold-cache replays)" % rep)
    print("%-34s " % "kernel" + " ".join("%12s" % n for _, n in COLS))
    for d in data:
        name = d[idx["Kernel Name"]].replace("void ", "").split("(")[0][:34]
        vals = []
        for m, _ in COLS:
            if m not in idx:
                vals.append("-")
                continue
            v, u = d[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.2f" % f if f < 1000 else "%.0f" % f) + ({"Mbyte": "M", "Kbyte": "K", "byte": "", "Gbyte": "G"}.get(u, ""))
            except ValueError:
                pass
            vals.append(v)
        print("%-34s " % name + " ".join("%12s" % v for v in vals))


if __name__ == "__main__":
    main()
