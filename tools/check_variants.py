"""Runs a structured Poisson hierarchy (generated on the device) through several option sets and checks that every
variant gives BIT-IDENTICAL iterates to the first one (the reference variant), with per-kernel event timings.

    python tools/check_variants.py --dim 3 --c 8 --finest 5 --out gpurun_out/variants.jsonl --variants base,csr,win1,win2

Used to validate new kernels against the CSR stream kernels at sizes no host oracle can hold."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_dolfinx_b200 import dist as ds  # noqa: E402

VARIANTS = {
    "csr": {"compress": 0},
    "gather": {"stage_x": 0},                      # round-1 default: row patterns, x gathered through L1/L2
    "hot1": {"stage_x": 3, "hot_cfg": 1},           # round-2 default: speculative loads at the hot pattern's offsets
    "hot2": {"stage_x": 3, "hot_cfg": 2},
    "hot3": {"stage_x": 3, "hot_cfg": 3},
    "hot4": {"stage_x": 3, "hot_cfg": 4},
    "hot1c2": {"stage_x": 3, "hot_cfg": 1, "compress": 2},      # without the anchored patterns of P
    "hot1a2": {"stage_x": 3, "hot_cfg": 1, "anch_cfg": 2},
    "hot1a3": {"stage_x": 3, "hot_cfg": 1, "anch_cfg": 3},
    "hot1a4": {"stage_x": 3, "hot_cfg": 1, "anch_cfg": 4},
    "tail300k": {"tail_rows": 300000},             # small levels in one cooperative launch (k_tail; off by default: measured slower)
    "tail50k": {"tail_rows": 50000},
    "hot1noinj": {"stage_x": 3, "hot_cfg": 1, "hot_inj": 0, "reuse_g": 0},
    "hot1pf0": {"stage_x": 3, "hot_cfg": 1, "hot_pf": 0},
    "hot1pf128k": {"stage_x": 3, "hot_cfg": 1, "hot_pf": 131072},
    "hot1pf384k": {"stage_x": 3, "hot_cfg": 1, "hot_pf": 393216},
    "hot1pf512k": {"stage_x": 3, "hot_cfg": 1, "hot_pf": 524288},
    "win1": {"stage_x": 1, "win_cfg": 1},
    "win2": {"stage_x": 1, "win_cfg": 2},
    "win3": {"stage_x": 1, "win_cfg": 3},
    "win4": {"stage_x": 1, "win_cfg": 4},
    "win1pf0": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 0},
    "win2pf0": {"stage_x": 1, "win_cfg": 2, "win_prefetch": 0},
    "win2pf8": {"stage_x": 1, "win_cfg": 2, "win_prefetch": 8},
    "win1pf8": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 8},
    "win3pf8": {"stage_x": 1, "win_cfg": 3, "win_prefetch": 8},
    "win1pf2": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 2},
    "win3pf2": {"stage_x": 1, "win_cfg": 3, "win_prefetch": 2},
    "win2pf2": {"stage_x": 1, "win_cfg": 2, "win_prefetch": 2},
    "win1pf1": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 1},
    "win1pf4": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 4},
    "win1pf16": {"stage_x": 1, "win_cfg": 1, "win_prefetch": 16},
}


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--c", type=int, default=8)
    ap.add_argument("--coarsest", type=int, default=0)
    ap.add_argument("--finest", type=int, default=5)
    ap.add_argument("--cycles", type=int, default=3)
    ap.add_argument("--variants", default="csr,gather,win1,win2,win3,win4")
    ap.add_argument("--extra", default="", help="k=v,k=v options added to every variant")
    ap.add_argument("--out", default="gpurun_out/variants.jsonl")
    args = ap.parse_args()
    extra = {k: float(v) for k, v in (kv.split("=") for kv in args.extra.split(",") if kv)}
    src = ds.StructuredSource(args.dim, args.c, args.coarsest, args.finest)
    ref = None
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    bad = 0
    with open(args.out, "a") as out:
        for name in args.variants.split(","):
            opts = dict(VARIANTS[name]); opts.update(extra)
            mg = ds.DistMG(src, device=0, r_mode="injection", smoother="jacobi", options=opts, device_gen=True)
            eng = mg.eng
            lf = mg.finest
            mg.load_rhs()
            hist = mg.cycles(args.cycles, history=True)
            v = mg.local_solution()
            sha = hashlib.sha256(v.tobytes()).hexdigest()[:16]
            st = eng.torch_stream()
            mg.cycles(3); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); mg.cycles(10); e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            eng.profile_begin(); mg.cycles(5); prof = eng.profile_end()
            if ref is None:
                ref = (sha, [float(x) for x in hist])
            same = sha == ref[0] and [float(x) for x in hist] == ref[1]
            bad += 0 if same else 1
            rec = {"dim": args.dim, "N": src.N(lf), "variant": name, "opts": opts, "sha": sha, "identical_to_first": same, "hist": [float(x) for x in hist],
                   "cycle_ms": ms,
                   "kernels": [{"k": f"{r['kind']}@{r['level']}", "us": round(r["ms_per_launch"] * 1e3, 2), "moved_gbs": round(r["moved_gbs"], 1), "n": r["launches"]}
                               for r in sorted(prof, key=lambda r: (-r["level"], r["kind"])) if r["level"] >= lf - 1]}
            out.write(json.dumps(rec) + "\n"); out.flush()
            print(name, "identical" if same else "DIFFERENT", f"cycle {ms:.3f} ms", [(k["k"], k["us"]) for k in rec["kernels"] if k["k"].endswith(f"@{lf}")], flush=True)
            if name == args.variants.split(",")[0]:
                print(eng.describe().split("level %d" % (lf - 1))[0][-900:], flush=True)
            mg.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
