#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a built library: `cuobjdump -sass <so>` grouped by kernel, selected mnemonics.

usage: sass_histogram.py <lib.so> [kernel-name substring ...]

UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, LDGSTS = cp.async, UTMALDG = TMA
tensor loads (none: operands are written by their producing threads or arrive by cp.async), LDL / STL = local-memory (spill) traffic."""
import collections, re, subprocess, sys

lib, subs = sys.argv[1], sys.argv[2:]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
hist, cur = collections.defaultdict(collections.Counter), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "LDL", "STL", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG", "FFMA", "FMUL",
        "FADD", "F2FP", "HADD2", "MUFU", "SHFL", "DFMA", "DMUL", "DADD", "ATOMS", "ATOMG", "RED"]
print("| kernel | SASS instructions | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for mangled, name in sorted(zip(hist, names), key=lambda kv: kv[1]):
    if subs and not any(s in name for s in subs):
        continue
    c = hist[mangled]
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
