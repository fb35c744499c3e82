#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell-native mnemonics
(B200_PROFILING.md "What proves a Blackwell-native kernel") in `cuobjdump -sass` of libmudpt_b200.so.

    python profiles/make_sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mudpt_b200", "lib", "libmudpt_b200.so")
PATTERNS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "HMMA", "LDGSTS", "LDSM",
            "FFMA2", "FADD2", "FMUL2", "MUFU.TANH", "MUFU.EX2", "SYNCS", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_instr"] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + "."):
                counts[cur][p] += 1
    names = demangle(list(counts))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (sm_100a)")
    print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, UBLKCP = 1-D bulk copies (cp.async.bulk), HMMA = mma.sync (legacy path),")
    print("# LDL/STL = local-memory (spill) accesses")
    for fn, c in counts.items():
        short = re.sub(r"\(.*", "", names.get(fn, fn)).replace("void mudpt::", "")
        tags = " ".join(f"{p}={c[p]}" for p in PATTERNS if c[p])
        print(f"{short:<64s} instr={c['_instr']:<6d} {tags}")


if __name__ == "__main__":
    sys.exit(main())
