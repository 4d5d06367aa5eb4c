#!/usr/bin/env python
"""Per-kernel summary table of an .ncu-rep (run here, no GPU): duration, grid, regs, occupancy, instructions, DRAM bytes."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print(" | ".join("%s[%s]" % (n, units[i]) for i, n in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:34] if n == "kernel" else r[i][:12]) for i, n in idx))
