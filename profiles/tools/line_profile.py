#!/usr/bin/env python
"""Aggregate an ncu source-page capture by CUDA source line.

usage: line_profile.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top] [ncu kernel-name regex when the report holds several kernels]

ncu's `--page source --print-source sass` gives per-SASS-instruction samples; `nvdisasm -g` gives the source line of every
SASS instruction of the same cubin.  Joining them by instruction index yields samples / executed instructions / stall
reasons per source line (inlined callee lines are reported under their own file:line).
"""
import csv, re, subprocess, sys, collections

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for ln in sass:
    if ln.startswith(".text."):
        active = kern in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
sel = ["--kernel-name", "regex:" + sys.argv[5], "--launch-count", "1"] if len(sys.argv) > 5 else []
out = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
second = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(second) > 1:                      # several launches in the report: keep the first
    rows = rows[:second[1]]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
assert abs(len(body) - len(lines)) <= 2, (len(body), len(lines))
agg = collections.defaultdict(lambda: collections.Counter())
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r, src in zip(body, lines):
    a = agg[src]
    a["samples"] += int(r[ix["# Samples"]]); a["inst"] += int(r[ix["Instructions Executed"]])
    tot["samples"] += int(r[ix["# Samples"]]); tot["inst"] += int(r[ix["Instructions Executed"]])
    for c in stall_cols:
        a[c] += int(r[ix[c]]); tot[c] += int(r[ix[c]])
print(f"total samples {tot['samples']}  warp-instructions {tot['inst']}")
print("stalls:", ", ".join(f"{c[6:]} {100*tot[c]/max(1,tot['samples']):.1f}%" for c in sorted(stall_cols, key=lambda c: -tot[c])[:8]))
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s}  top stalls")
for src, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(stall_cols, key=lambda c: -a[c])[:3]
    name = f"{src[0]}:{src[1]}" if src else "?"
    print(f"{name:28s} {100*a['samples']/tot['samples']:8.2f} {100*a['inst']/tot['inst']:7.2f}  " + ", ".join(f"{c[6:]} {100*a[c]/max(1,a['samples']):.0f}%" for c in st))
