"""Opcode mix of a kernel from an `ncu --page source --csv` export: executed warp instructions per opcode and per row-warp,
plus the instructions holding most stall samples.   python tools/ncu_opmix.py source.csv <rows> [kernel-substring]"""
import collections
import csv
import re
import sys

rows_total = float(sys.argv[2])
want = sys.argv[3] if len(sys.argv) > 3 else None
kern, hdr, data = None, None, []
blocks = []
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "Kernel Name":
        if kern is not None: blocks.append((kern, hdr, data))
        kern, hdr, data = r[1], None, []
    elif kern is not None and hdr is None: hdr = r
    elif kern is not None: data.append(r)
if kern is not None: blocks.append((kern, hdr, data))
for kern, hdr, data in blocks:
    if want and want not in kern: continue
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    nw = rows_total / 32
    print(kern[:120]); print(f"  total warp instructions {tot}  per row-warp {tot / nw:.1f}  SASS lines {len(data)}")
    c, s = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        op = m.group(2).split(".")[0] if m else "?"
        c[op] += int(r[ix["Instructions Executed"]]); s[op] += int(r[ix["# Samples"]])
    for op, n in c.most_common(14): print(f"  {op:8s} {n / nw:7.2f} per row-warp {100 * n / tot:5.1f}%  samples {s[op]}")
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:8]
    for r in top: print("   ", r[ix["# Samples"]], r[ix["Source"]].strip()[:90])
