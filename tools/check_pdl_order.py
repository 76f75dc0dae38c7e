"""Static check of the shipped library: in every kernel that waits for its predecessor (griddepcontrol.wait, SASS ACQBULK),
list the global loads the compiler scheduled BEFORE the wait.  Loads of static set-up data (codes, tables, descriptors) may
sit there; a load of a vector the predecessor writes must not (a `const __restrict__` pointer makes loads `invariant` and
free to move across the wait -- that was a real bug in k_hotinj).

    python tools/check_pdl_order.py [multigrid_dolfinx_b200/libmgb200.so]
"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "multigrid_dolfinx_b200/libmgb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, early, seen_wait, out = None, [], False, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if fn and seen_wait:
            out[fn] = early
        fn, early, seen_wait = m.group(1), [], False
        continue
    if "ACQBULK" in line:
        seen_wait = True
    elif not seen_wait and re.search(r"\bLDG\b|\bLDG\.", line):
        early.append(line.split("*/")[1].strip() if "*/" in line else line.strip())
if fn and seen_wait:
    out[fn] = early
for k, v in sorted(out.items()):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{len(v):3d} loads before the wait  {name[:110]}")
    for l in v[:6]:
        print("       ", l[:100])
