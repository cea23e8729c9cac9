"""Markdown summary of `ncu --set full --import-source on` reports of single kernels (run here, no GPU needed):

    python tools/ncu_top_summary.py gpurun_out/r2/top_FLASH_IN.ncu-rep gpurun_out/r2/top_ATT_OUT.ncu-rep ... \
        > profiles/r2_top_kernels_ncu.md

Per report: duration, DRAM bytes, pipe utilisations, registers, the warp-stall sampling by reason (with the opcodes
that carry the samples) and the executed-instruction histogram."""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum",
       "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    print("# Top kernels at the final round-2 state: `ncu --set full --clock-control none --import-source on`\n")
    print("One launch each, B = 16 x 4 s (131 072 frames: ncu saves / restores the touched memory around every replay "
          "pass, the 64-item batch does not finish).  Durations are under the profiler: compare shares, not absolutes.\n")
    for rep in sys.argv[1:]:
        raw = [r for r in page(rep, "raw") if len(r) > 10]
        hdr, units, vals = raw[0], raw[1], raw[2]
        name = vals[hdr.index("Kernel Name")]
        print(f"## `{name[:110]}`\n\n| counter | value |\n|---|---|")
        for k in RAW:
            if k in hdr:
                print(f"| `{k}` | {vals[hdr.index(k)]} {units[hdr.index(k)]} |")
        src = page(rep, "source")
        hi = next(i for i, r in enumerate(src) if "Source" in r and "Address" in r)
        h = src[hi]
        ix = {k: i for i, k in enumerate(h)}
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot, by, ops = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
        n = ni = 0
        for r in src[hi + 1:]:
            if len(r) < len(h):
                continue
            s = r[ix["Source"]].split()
            if not s:
                continue
            op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
            n += int(r[ix["# Samples"]])
            ni += int(r[ix["Instructions Executed"]])
            ops[op] += int(r[ix["Instructions Executed"]])
            for k in stalls:
                v = int(r[ix[k]])
                tot[k] += v
                by[k][op] += v
        print(f"\nWarp-stall sampling ({n} samples), share of all samples, with the opcodes carrying them:\n")
        print("| reason | share | opcodes |\n|---|---|---|")
        for k, v in tot.most_common(9):
            print(f"| {k} | {100 * v / n:.1f} % | " + ", ".join(f"{o} {100 * c / n:.1f}" for o, c in by[k].most_common(5)) + " |")
        print(f"\nExecuted warp instructions ({ni}): " + ", ".join(f"{o} {100 * c / ni:.1f} %" for o, c in ops.most_common(12)) + "\n")


if __name__ == "__main__":
    main()
