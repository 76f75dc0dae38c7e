"""SASS of the shipped library, summarised on the CPU box (no GPU needed):

    python tools/sass_summary.py [profiles/r2_sass_summary.json]

* per kernel: instruction count and the mnemonics that matter here -- DFMA (must be 0 in every row-sum kernel: one accumulator,
  separately rounded multiply and add is what makes the results bit-identical to the reference's scipy row sums), DMUL / DADD,
  global loads by width, UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), ACQBULK / PREEXIT (programmatic dependent launch),
  UCGABAR (cluster barrier), and registers / stack from `cuobjdump -res-usage`;
* for the whole library: the same counts and the list of cubins (`cuobjdump -lelf`: sm_100a only).
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multigrid_dolfinx_b200", "libmgb200.so")

# kernels whose row sums must match the reference's scipy row sums bit for bit ("What the GPU path guarantees")
ROWSUM = re.compile(r"\bk_(hotrow2?|hotinj|anchrow|anchloop|rowstream|rowwin|stream|tile|seqrow|subwarp)<")
# the Gauss-Seidel epilogue divides by a_ii: the IEEE division sequence is made of DFMAs (the row sum in front of it is not)
DIVIDING = re.compile(r"EpiGaussSeidel")

KEYS = ("DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "UCGABAR", "ATOMG", "RED", "MEMBAR", "CCTL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def scan(lib=LIB):
    """-> {demangled kernel name: Counter of SASS mnemonics (first dot-component; '_n' = instructions; 'LDG.256' etc. by width)}"""
    p = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True)
    stats, cur = collections.OrderedDict(), None
    fn = re.compile(r"\s*Function : (\S+)")
    ins = re.compile(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in p.stdout:
        m = fn.match(line)
        if m:
            cur = stats.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = ins.match(line)
        if m:
            op = m.group(1)
            cur["_n"] += 1
            cur[op.split(".")[0].split("_")[0]] += 1          # UCGABAR_ARV / UCGABAR_WAIT -> UCGABAR
            if op.startswith("LDG"):
                cur["LDG.256" if ".256" in op else "LDG.128" if ".128" in op else "LDG.64" if ".64" in op else "LDG.other"] += 1
    p.wait()
    dm = demangle(list(stats))
    return {dm[k]: v for k, v in stats.items()}


def res_usage(lib=LIB):
    txt = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.split("\n")
    names = [re.match(r"\s*Function (\S+):", l).group(1) for l in txt if re.match(r"\s*Function (\S+):", l)]
    dm = demangle(names)
    out, cur = {}, None
    for l in txt:
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            cur = dm[m.group(1)]
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", l)
        if m and cur:
            out[cur] = {"regs": int(m.group(1)), "stack": int(m.group(2)), "static_smem": int(m.group(3))}
            cur = None
    return out


def fused_multiply_add_offenders(stats):
    """Row-sum kernels that contain a DFMA although their epilogue does not divide."""
    return sorted(k for k, v in stats.items() if ROWSUM.search(k) and not DIVIDING.search(k) and v.get("DFMA", 0))


def short(name):
    m = re.search(r"(k_\w+(?:<.*?>)?)\(", name)
    return (m.group(1) if m else name).replace("mgb::", "").replace("(anonymous namespace)::", "")


def main(dst):
    stats, res = scan(), res_usage()
    tot = collections.Counter()
    for v in stats.values():
        tot.update(v)
    # the kernels config 5 actually launches by default (profiles/r2_ncu_launch_list_cfg5.csv) + the TMA stream kernel (P2, config 4)
    want = ["k_hotrow<6, 128, 2, 8, false, mgb::EpiJacobiRJ>", "k_hotrow<6, 128, 2, 7, false, mgb::EpiJacobiRJFirst>",
            "k_hotrow<6, 128, 2, 8, true, mgb::EpiJacobiRJ>", "k_hotinj<15, 128, 8>", "k_anchloop<128, 4, 8, 4, mgb::EpiProlongAdd>",
            "k_anchrow<128, 4, 8, 5, false, mgb::EpiProlongAdd>", "k_stream<256, 4, 2, true, mgb::EpiJacobiRJ>",
            "k_rowstream<256, 2, 8, 3, 2, 8, 4, mgb::EpiJacobiRJ>", "k_dense_gemv(", "k_init_guess("]
    picked = [k for w in want for k in stats if w in k][:len(want) + 2]
    out = {"library": os.path.relpath(LIB, ROOT), "size_bytes": os.path.getsize(LIB),
           "cubins": [l.split(":")[-1].strip() for l in subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.split("\n") if l.strip()],
           "kernels_total": len(stats),
           "library_totals": {k: tot.get(k, 0) for k in KEYS + ("LDG.256", "LDG.128", "LDG.64")},
           "tensor_core_instructions": sum(v for k, v in tot.items() if k.startswith(("UTCMMA", "UTCHMMA", "HMMA", "DMMA", "IMMA"))),
           "row_sum_kernels": sum(1 for k in stats if ROWSUM.search(k)),
           "row_sum_kernels_with_DFMA_outside_a_division": [short(k) for k in fused_multiply_add_offenders(stats)],
           "kernels_with_DFMA": {short(k): v["DFMA"] for k, v in stats.items() if v.get("DFMA", 0)},
           "note": "DFMA only where a division (Gauss-Seidel's 1/a_ii, 1/d in getJacobiMatrices), a square root, a norm / dot product or the dense "
                   "coarsest solve is computed -- none of which the reference's bit pattern depends on through a row sum; tensor cores are not used "
                   "(fp64 sparse row sums with one right-hand side, SURVEY 7)",
           "hot_kernels": [dict({"kernel": short(k), "instructions": stats[k]["_n"]}, **res.get(k, {}),
                                **{m: stats[k].get(m, 0) for m in KEYS + ("LDG.256", "LDG.128", "LDG.64") if stats[k].get(m, 0)}) for k in picked]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["library_totals"]), "->", dst)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.json"))
