"""Times every kernel variant on one workload (event-timed, per kernel kind and level) so that the
per-level choice of kernel family / tile shape is backed by measurements.  Writes JSON lines.

    python tools/variant_sweep.py --workload cfg2 --out gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multigrid_dolfinx_b200.engine import MGEngine  # noqa: E402

VARIANTS = [
    ("pattern1_256x2", {"code_cfg": 1}),
    ("pattern3_256x2x3", {"code_cfg": 3}),
    ("pattern1_pdl", {"code_cfg": 1, "pdl": 1}),
    ("pattern1_nopdl", {"code_cfg": 1, "pdl": 0}),
    ("pattern1_stage_x", {"code_cfg": 1, "stage_x": 1}),        # experimental: x staged in shared memory (unmeasured so far)
    ("coded1_256x2", {"compress": 1, "code_cfg": 1}),
    ("coded3_256x2x3", {"compress": 1, "code_cfg": 3}),
    ("csr_stream3", {"compress": 0}),
    ("tile_iter1", {"tile_iter": 1, "stream_cfg": 0, "fuse_restrict": 0}),
    ("stream1_nofuse", {"stream_cfg": 1, "fuse_restrict": 0}),
    ("tile_iter2", {"tile_iter": 2, "stream_cfg": 0}),
    ("subwarp4", {"kernel_family": 2, "lanes_per_row": 4}),
    ("subwarp8", {"kernel_family": 2, "lanes_per_row": 8}),
    ("stream1_256x8x2", {"stream_cfg": 1}),
    ("stream2_512x4x2", {"stream_cfg": 2}),
    ("stream3_256x4x2", {"stream_cfg": 3}),
    ("stream4_256x4x3", {"stream_cfg": 4}),
    ("stream5_128x8x2", {"stream_cfg": 5}),
    ("stream6_256x8x3", {"stream_cfg": 6}),
    ("tile_iter1_fused_restrict", {"tile_iter": 1, "stream_cfg": 0, "fuse_restrict": 1}),
    ("auto", {}),
    ("noauto_cfg3", {"stream_auto": 0, "stream_cfg": 3}),
    ("noauto_cfg1", {"stream_auto": 0, "stream_cfg": 1}),
    ("noauto_cfg2", {"stream_auto": 0, "stream_cfg": 2}),
    ("noauto_cfg5", {"stream_auto": 0, "stream_cfg": 5}),
    ("noauto_cfg7", {"stream_auto": 0, "stream_cfg": 7}),
    ("noauto_subwarp8", {"kernel_family": 2, "lanes_per_row": 8}),
    ("noauto_subwarp16", {"kernel_family": 2, "lanes_per_row": 16}),
]


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--cycles", type=int, default=10)
    ap.add_argument("--smoother", default="jacobi")
    ap.add_argument("--restriction", default="injection")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    H, desc = bench.build_workload(args.workload)
    lf = H.finest_level
    f = H.b_dict[lf][:, 0]
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as out:
        for name, opts in VARIANTS:
            if args.only and not any(name.startswith(o) for o in args.only.split(",")):
                continue
            eng = MGEngine.from_hierarchy(H, r_mode=args.restriction, smoother=args.smoother, options=opts)
            eng.level_buffer(lf, "f").copy_(torch.from_numpy(f)); eng.level_buffer(lf, "v").zero_()
            torch.cuda.synchronize()
            eng.vcycle_resident(lf, 3)
            st = eng.torch_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); eng.vcycle_resident(lf, args.cycles); e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.cycles
            eng.profile_begin(); eng.vcycle_resident(lf, args.cycles); prof = eng.profile_end()
            rec = {"workload": args.workload, "variant": name, "opts": opts, "cycle_ms_graph": ms, "vcycle_gbs": eng.vcycle_bytes(lf) / ms / 1e6,
                   "kernels": [{"k": f"{r['kind']}@{r['level']}", "us": round(r["ms_per_launch"] * 1e3, 2), "gbs": round(r["gbs"], 1), "moved_gbs": round(r["moved_gbs"], 1), "n": r["launches"]}
                               for r in sorted(prof, key=lambda r: (-r["level"], r["kind"]))]}
            out.write(json.dumps(rec) + "\n"); out.flush()
            top = [k for k in rec["kernels"] if k["k"].endswith(f"@{lf}") and k["k"].split("@")[0] in ("jacobi", "residual", "prolong_add")]
            print(name, f"cycle {ms:.3f} ms", top, flush=True)
            eng.close()


if __name__ == "__main__":
    main()
