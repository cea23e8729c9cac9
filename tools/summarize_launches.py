"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

    python tools/summarize_launches.py launches.csv [last_n_launches] > summary.md

`last_n_launches` keeps only the tail of the list (e.g. the launches of the last benchmark step)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r[iv].replace(",", ""))
        v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"^void ", "", name).replace("tdz::", "").replace("(int)", "").replace("(unsigned int)", "")
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    print(f"Total device time of the {len(rows)} launches: {total:.1f} ms\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        if v[1] / total < 0.002:
            continue
        print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / total:.1f}% |")


if __name__ == "__main__":
    main()
