#!/usr/bin/env python
"""Per-kernel opcode census of libb200splat.so (cuobjdump -sass): how many TMA bulk copies (UBLKCP), mbarrier waits
(SYNCS), async copies (LDGSTS), global reductions (REDG / multimem LDGMC), warp matches / votes / shuffles, MUFU ops and
shared-memory instructions each kernel's SASS holds.  Static counts (instructions in the binary, not executed).

    python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "my_depthsplat_b200" / "libb200splat.so"
FAMILIES = ["UBLKCP", "SYNCS", "LDGSTS", "REDG", "LDGMC", "ATOMG", "ATOMS", "MATCH", "VOTE", "SHFL", "MUFU.EX2", "MUFU.RCP", "MUFU.RSQ",
            "MUFU.SQRT", "MUFU.LG2", "LDS", "STS", "LDG", "STG", "BAR", "FFMA", "FMUL", "FADD", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", name)
            name = re.sub(r"\(.*$", "", name)
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for f in FAMILIES:
                if op == f or op.startswith(f + ".") or (f.startswith("MUFU") and op.startswith(f)):
                    counts[name][f] += 1
    print(f"# SASS opcode census of {LIB.name} (sm_100a), static instruction counts per kernel")
    print(f"# {subprocess.run(['cuobjdump', '--version'], capture_output=True, text=True).stdout.strip().splitlines()[-1]}")
    used = [f for f in FAMILIES if any(c[f] for c in counts.values())]
    w = max(len(n) for n in counts) + 2
    print("kernel".ljust(w) + "total".rjust(7) + "".join(f.rjust(10) for f in used))
    for n, c in counts.items():
        print(n.ljust(w) + str(total[n]).rjust(7) + "".join((str(c[f]) if c[f] else ".").rjust(10) for f in used))


if __name__ == "__main__":
    sys.exit(main())
