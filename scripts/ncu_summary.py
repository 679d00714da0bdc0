"""Condense `ncu -i <rep> --page raw --csv` into the metric rows kept under profiles/.
usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv | python scripts/ncu_summary.py profiles/X_ncu_summary.csv"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active"]
rows = [r for r in csv.reader(sys.stdin) if r]
while rows and rows[0][0] != "ID":
    rows.pop(0)
hdr, units, data = rows[0], rows[1], rows[2:]
with open(sys.argv[1], "w", newline="") as f:
    w = csv.writer(f)
    for name in WANT:
        if name in hdr:
            i = hdr.index(name)
            w.writerow([name, units[i]] + [r[i] for r in data])
            print(name[:52].ljust(52), units[i][:10].ljust(10), [r[i][:14] for r in data])
