"""Condenses an `ncu --page raw --csv` export into the per-launch rows kept under profiles/ (selected counters):

    ncu -i steps.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv profiles/r2_ncu_raw.csv.gz

Row 0 = metric names, row 1 = units, then one row per profiled launch (the format bench.py's ncu_traffic reads)."""
import csv
import gzip
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 10]
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    op = gzip.open if dst.endswith(".gz") else open
    with op(dst, "wt", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    print(f"{len(data)} launches, {len(idx)} columns -> {dst}")


if __name__ == "__main__":
    main()
