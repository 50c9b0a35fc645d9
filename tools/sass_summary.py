"""Developer tool: per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass), the proof asked for in
B200_PROFILING.md: UBLKCP (bulk copies, TMA 1-D), UTMALDG/UTMASTG (tensor-map TMA), SYNCS (mbarrier), LDGSTS (cp.async),
STG/LDG (per-thread global access), LDS/STS, SHFL, BAR, DFMA/DADD/DMUL (FP64 pipe).  Writes a table to stdout.

    python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cfftpack_b200", "libcfftpack_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "DFMA", "DADD", "DMUL",
        "IMAD", "total"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur = {}, [], None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"^void ", "", cur)
        cur = re.sub(r"\(.*$", "", cur)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        c = counts[cur]
        c["total"] += 1
        if base in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "DFMA", "DADD",
                    "DMUL", "IMAD"):
            c[base] += 1
print(f"# SASS summary of {os.path.relpath(lib, ROOT)} ({len(order)} kernels, sm_100a), static instruction counts per kernel")
print("# " + " ".join(f"{k:>7s}" for k in KEYS) + "  kernel")
tot = collections.Counter()
for k in sorted(order):
    c = counts[k]
    tot.update(c)
    print("  " + " ".join(f"{c[x]:7d}" for x in KEYS) + "  " + k[:170])
print("# " + " ".join(f"{tot[x]:7d}" for x in KEYS) + "  ALL KERNELS")
