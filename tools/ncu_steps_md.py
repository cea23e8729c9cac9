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
    # ---- per launch step: DRAM traffic of its kernels against the algorithmic bytes of bench.py::STEP_TABLE
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # kernels of each step in the launch order of tools/run_all_steps.py (Separator.STEP_NAMES)
    step_kernels = [("ENCODER", 3), ("ENC1X1", 2), ("FLASH_IN", 3), ("SIM", 1), ("KV", 2), ("ATT_OUT", 1), ("TO_OUT", 2),
                    ("FSMN_C1", 1), ("FSMN_UV", 1), ("FSMN_LIN", 1), ("FSMN_PROJ", 1), ("DD1", 1), ("DD2", 2),
                    ("FSMN_TAIL", 2), ("FSMN_C2", 1), ("FINAL_LN", 2), ("FINAL_GN", 1), ("OUT1", 1), ("TANHSIG", 2),
                    ("DEC1", 2), ("DECODER", 1)]
    if sum(n for _, n in step_kernels) == first_sv:
        frames = 16 * 8192          # padded frames of the capture (B = 16, T = 64 000 -> Sp = 8 192)
        print("\n## Separator steps: measured DRAM traffic against the algorithmic bytes (`bench.py::STEP_TABLE`)\n")
        print("Algorithmic = compulsory read + write bytes per frame x 131 072 padded frames.  The capture is small enough "
              "(268 MB per 2 KB/frame tensor) for the 126 MB L2 to keep part of a producer's output for its consumer and to "
              "hold dirty lines past the end of a kernel, so measured traffic BELOW the algorithmic figure is L2 reuse, not "
              "an accounting error; traffic well above it would be wasted re-reads (none: the largest ratio is listed "
              "first).  FSMN_LIN = the back-to-back GEMM (linear + project); FSMN_PROJ alone is the two-kernel test form.\n")
        print("| step | kernels | us | measured MB | algorithmic MB | measured / algorithmic |")
        print("|---|---|---|---|---|---|")
        out, i = [], 0
        for name, n in step_kernels:
            t = sum(vals(r)[0] for r in data[i:i + n])
            b = sum(vals(r)[1] for r in data[i:i + n])
            i += n
            if name in bench.STEP_TABLE:
                alg = bench.STEP_TABLE[name]["bytes"] * frames
                out.append((b / alg, f"| {name} | {n} | {t * 1e6:.1f} | {b / 1e6:.0f} | {alg / 1e6:.0f} | {b / alg:.2f} |"))
        for _, l in sorted(out, reverse=True):
            print(l)
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
