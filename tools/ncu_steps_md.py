"""Markdown table of the per-kernel ncu counters kept in profiles/r2_ncu_raw.csv.gz (tools/ncu_steps.sh capture: every
launch step of the separator once at B = 16 x 4 s, then one fbank + embedder call on 32 utterances):

    python tools/ncu_steps_md.py profiles/r2_ncu_raw.csv.gz > profiles/r2_steps_ncu.md

DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum, given next to the measured copy
peak of MEASURED_PEAKS.json; pipe columns are ncu's pct_of_peak_sustained_active."""
import collections
import csv
import gzip
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt", newline="")))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {k: h.index(k) for k in h}
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}

    def f(r, k):
        return float(r[col[k]].replace(",", ""))

    def short(name):
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("tdz::", "")
        return name.replace("(int)", "").replace("(unsigned int)", "")

    # the separator steps come first (one launch sequence per step), the embedder call after the fbank kernel
    first_sv = next((i for i, r in enumerate(data) if "fbank_kernel" in r[col["Kernel Name"]]), len(data))
    print("# Per-kernel ncu counters at HEAD (`tools/ncu_steps.sh`, B = 16 x 4 s; embedder on 32 utterances)\n")
    print(f"Durations are under the profiler (cold caches, serialised); DRAM GB/s against the measured copy peak of "
          f"{peak:.0f} GB/s (`MEASURED_PEAKS.json`).  Raw rows: `profiles/r2_ncu_raw.csv.gz`.\n")
    hdr = "| kernel | us | DRAM MB | DRAM GB/s | of copy peak | tensor % | FMA % | XU % | issue % | regs |"
    sep = "|---|---|---|---|---|---|---|---|---|---|"

    def line(name, n, t, b, r):
        gbs = b / t / 1e9 if t else 0.0
        return (f"| `{name}`{' x' + str(n) if n > 1 else ''} | {t * 1e6:.1f} | {b / 1e6:.1f} | {gbs:.0f} | "
                f"{100 * gbs / peak:.0f} % | {r[0]:.1f} | {r[1]:.1f} | {r[2]:.1f} | {r[3]:.1f} | {r[4]} |")

    def vals(r):
        t = f(r, "gpu__time_duration.sum") * tscale[units[col["gpu__time_duration.sum"]]]
        b = (f(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]
             + f(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]])
        pipes = (f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                 f(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                 f(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                 r[col["launch__registers_per_thread"]].split(".")[0])
        return t, b, pipes

    print("## Separator: one launch sequence per step (layer-0 instance), in launch order\n")
    print(hdr)
    print(sep)
    for r in data[:first_sv]:
        t, b, p = vals(r)
        print(line(short(r[col["Kernel Name"]]), 1, t, b, p))
    print("\n## fbank + ERes2NetV2 (one call, 32 utterances of 4 s), aggregated per kernel\n")
    print(hdr)
    print(sep)
    agg = collections.OrderedDict()
    for r in data[first_sv:]:
        t, b, p = vals(r)
        a = agg.setdefault(short(r[col["Kernel Name"]]), [0, 0.0, 0.0, [0.0] * 4, p[4]])
        a[0] += 1
        a[1] += t
        a[2] += b
        for i in range(4):
            a[3][i] += p[i] * t
    tot = sum(a[1] for a in agg.values())
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(line(name, a[0], a[1], a[2], tuple(x / a[1] for x in a[3]) + (a[4],)))
    print(f"\nEmbedder total under the profiler: {tot * 1e3:.2f} ms for 32 utterances.")


if __name__ == "__main__":
    main()
