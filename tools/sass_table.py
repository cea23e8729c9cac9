"""Opcode-count table of the kernels in libtdz.so (evidence that the contractions are tcgen05 / TMEM / TMA code):

    python tools/sass_table.py > profiles/r2_sass_opcodes.md

For every kernel: instructions, UTC*MMA (tcgen05.mma; `.2CTA` = cta_group::2), LDTM (tcgen05.ld), UTMALDG (TMA loads),
LDGSTS (cp.async), FFMA2 (packed fp32 FMA), MUFU, HMMA (legacy mma.sync: only the restorer's 80 x 80 attention),
ACQBULK / PREEXIT (griddepcontrol.wait / .launch_dependents: programmatic dependent launch)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "targetdiarization_b200", "libtdz.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "LDGSTS", "FFMA2", "FFMA", "MUFU", "HMMA", "ACQBULK", "PREEXIT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_n"] += 1
            if op.startswith("UTC") and "MMA" in op:
                cur["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
            for key in ("LDTM", "UTMALDG", "LDGSTS", "FFMA2", "MUFU", "HMMA", "ACQBULK", "PREEXIT"):
                if op.startswith(key):
                    cur[key] += 1
            if op == "FFMA" or op.startswith("FFMA."):
                cur["FFMA"] += 1
    print("| kernel | instr | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for name, c in sorted(kernels.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCHMMA.2CTA"])):
        short = demangle(name)
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("tdz::", "").replace("(int)", "") \
            .replace("(unsigned int)", "")
        print(f"| `{short}` | {c['_n']} | " + " | ".join(str(c[k]) for k in COLS) + " |")
    n_mma = sum(1 for c in kernels.values() if c["UTCHMMA"] + c["UTCHMMA.2CTA"])
    print(f"\n{len(kernels)} kernels, {n_mma} with tcgen05.mma; legacy HMMA instructions in the library: "
          f"{sum(c['HMMA'] for c in kernels.values())}; kernels with griddepcontrol.wait: "
          f"{sum(1 for c in kernels.values() if c['ACQBULK'])}")


if __name__ == "__main__":
    sys.exit(main())
