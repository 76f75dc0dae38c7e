"""bench.py's arm for device-generated hierarchies: one process per GPU (torchrun for N > 1), row-sharded V-cycles, strong
scaling; N = 1 runs the same code on the whole problem (world = 1).

Timing: W warm-up cycles, barrier + synchronize, K cycles bracketed by CUDA events on every rank's engine
stream, MAX over ranks; value = global smoother DOF-updates of K cycles / that time.
Parity: the residual norms of 3 cycles from a zero guess are compared with the CPU oracle on the same hierarchy (N = 1, unless
--no-cpu) or with the committed single-GPU values (profiles/expected_resnorms.json); the run exits non-zero beyond 1e-12."""
from __future__ import annotations

import json
import os
import time

import numpy as np

# name: (dim, c, coarsest, finest, description)
DIST_WORKLOADS = {
    "cfg3": (3, 8, 0, 4, "3D Poisson P1 129^3 (2.1M DOFs), 5-level V(2,2), Jacobi, injection, row-sharded"),
    "cfg5h": (3, 8, 0, 5, "3D Poisson P1 257^3 (17M DOFs), 6-level V(2,2), Jacobi, injection, row-sharded"),
    "cfg5": (3, 8, 0, 6, "3D Poisson P1 513^3 (135M DOFs), 7-level V(2,2), Jacobi, injection, row-sharded"),
    "cfg2": (2, 32, 0, 6, "2D Poisson P1 2049^2 (4.2M DOFs), 7-level V(2,2), Jacobi, injection, row-sharded"),
    "cfg4": (3, 4, 0, 4, "3D Poisson P2 on 64^3 cells (129^3 = 2.1M DOFs, rows of 10..65), 5-level V(2,2), Jacobi, row-sharded"),
    "cfg4s": (3, 2, 0, 3, "3D Poisson P2 on 16^3 cells (33^3 DOFs), 4-level V(2,2) (small stand-in)"),
}


def run(args):
    import torch
    import torch.distributed as td
    from multigrid_dolfinx_b200 import dist as ds
    import bench as B

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if multi:
            td.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if multi:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    name = args.workload if args.workload in DIST_WORKLOADS else "cfg5h"
    dim, c, lc, lf, desc = DIST_WORKLOADS[name]
    t0 = time.perf_counter()
    if name.startswith("cfg4"):       # P2: every rank assembles the hierarchy on the host and cuts its row blocks out of it
        from multigrid_dolfinx_b200 import problems as pr
        src = ds.HierarchySource(pr.build_hierarchy_p2(c=c, coarsest_level=lc, finest_level=lf))
        args.device_gen = 0
    else:
        src = ds.StructuredSource(dim, c, lc, lf)
    mg = ds.DistMG(src, device=local_rank, r_mode=args.restriction, smoother=args.smoother, gather_threshold=args.gather_threshold,
                   options={**{"fuse_restrict": args.fuse_restrict, "stream_cfg": args.stream_cfg, "use_graph": args.use_graph,
                               "overlap_halo": args.overlap, "overlap_waves": args.overlap_waves,
                               "compress": getattr(args, "compress", 3), "code_cfg": getattr(args, "code_cfg", 1)}, **args.options},
                   device_gen=bool(args.device_gen) and (args.restriction == "injection" or not multi), p2p=bool(args.p2p))
    setup_s = time.perf_counter() - t0
    eng = mg.eng
    stream = eng.torch_stream()
    dofu = mg.dof_updates_per_cycle()
    mg.load_rhs()
    hist3 = [float(x) for x in mg.cycles(3, history=True)]        # parity: 3 cycles from a zero guess, ||f - A v||_2 after each
    mg.load_rhs()
    mg.cycles(args.warmup)
    eng.synchronize(); barrier(); torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with B.ClockSampler(local_rank) as clk:
        e0.record(stream)
        mg.cycles(args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = eng.launch_count() - l0
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)      # identical on every rank (the cycles are collective)
        if ms * args.steps < 1000.0:         # ~1 s under load for the clock sampler
            mg.cycles(int(min(2000, max(1, (1000.0 - ms * args.steps) / max(ms, 1e-3))))); torch.cuda.synchronize()

    # dominant kernel on rank 0, event-timed (all ranks run the same cycles: the halo exchanges are collective)
    ncyc = max(3, min(args.steps, 5))
    # (the first eager cycle after graph replay carries one-off costs -- event creation, the eager NCCL path of the gather level --
    # that land in whichever kernel happens to wait for a neighbour at that moment: one discarded cycle first, ranks aligned)
    eng.profile_begin(); mg.cycles(1); eng.profile_end()
    eng.synchronize(); barrier()
    eng.profile_begin(); mg.cycles(ncyc); prof = eng.profile_end()
    if multi:
        # with the exchange fused into the kernels a launch also waits for its neighbours' flags, so one rank that is late once
        # (a host hiccup in this eager pass) lands in whichever kernel waited for it: a second pass, and the pass with the smaller
        # total is reported (every rank runs both: the cycles are collective)
        eng.synchronize(); barrier()
        eng.profile_begin(); mg.cycles(ncyc); prof2 = eng.profile_end()
        if sum(r["total_ms"] for r in prof2) < sum(r["total_ms"] for r in prof):
            prof = prof2
    # end to end through the C ABI with pinned host buffers (each rank stages its own row block)
    n_loc = mg.n_local
    vp = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
    fp = torch.from_numpy(np.ascontiguousarray(mg.local_rhs())).pin_memory()
    lib, h = eng._lib, eng._h
    for _ in range(2):
        eng._ck(lib.mgb_vcycle(h, lf, vp.data_ptr(), fp.data_ptr(), 0, 1, None))
    e2e_steps = max(3, min(args.steps, 10))
    barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        eng._ck(lib.mgb_vcycle(h, lf, vp.data_ptr(), fp.data_ptr(), 0, 1, None))
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t1) / e2e_steps)
    # ---- CPU baseline (N = 1 only, rank 0): the C/OpenMP port on the SAME hierarchy, also the parity reference ----------------
    default_cfg = args.smoother == "jacobi" and args.restriction == "injection"      # what the committed norms and the C oracle's builder run
    cpu, expected, esrc = None, B.expected_from_file(name) if default_cfg else None, f"profiles/expected_resnorms.json ({name}, single-GPU run)"
    if not default_cfg:
        esrc = f"none for smoother={args.smoother}, restriction={args.restriction} (bit-compared against the host-assembled twin in tests/test_gpu_parity.py)"
    if rank == 0 and not multi and not args.no_cpu and name in B.STRUCTURED and default_cfg:
        cm, f, dofu_cpu, threads, note = B.cpu_oracle(name)
        if cm is None:
            cpu = {"value": None, "unit": B.UNIT, "cores": threads, "kind": "port", "sample": f"not run: {note}"}
        else:
            cyc = 2 if name == "cfg5" else 5
            per = B.time_cpu(cm, f, cyc, 1)
            _, hist_cpu = cm.vcycle(np.zeros_like(f), f, ncycles=3, history=True)
            cpu = {"value": dofu_cpu / per, "unit": B.UNIT, "cores": threads, "kind": "port", "ms_per_cycle": per * 1e3,
                   "sample": f"{cyc} full V-cycles of {name} after one warm-up ({note}), C/OpenMP port of the reference, {threads} threads of {os.cpu_count()} cores"}
            expected, esrc = [float(x) for x in hist_cpu], "CPU oracle (oracle/mg_oracle.c) on the same hierarchy, 3 cycles from a zero guess"
            cm.close()
    parity = B.parity_block(hist3, expected, esrc)
    if rank == 0:
        peak, peak_src = B.measured_peak()
        comp = [r for r in prof if r["kind"] not in ("halo",)]
        dom = max(comp, key=lambda r: r["total_ms"])
        halo_ms = sum(r["total_ms"] for r in prof if r["kind"] == "halo") / ncyc
        tot_ms = sum(r["total_ms"] for r in prof) / ncyc
        n_glob = src.n(lf)
        line = {"metric": B.METRIC, "value": dofu / (ms * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{name}: {desc}", "restriction": args.restriction, "smoother": args.smoother, "fine_dofs": n_glob,
                           "levels": lf - lc + 1, "mu1": src.mu1, "mu2": src.mu2, "generated_on_device": bool(mg.device_gen), "compress": getattr(args, "compress", 3),
                           "parallelism": f"row-sharded x{world}, levels <= {mg.gather_level} on rank 0" if multi else "single GPU",
                           "options": args.options,
                           "l2": ("per-rank working set of one fine-level sweep (codes + 3 vectors, %.0f MB) " % (25.0 * n_glob / world / 1e6)) +
                                 ("exceeds the 126 MB L2: no flush needed" if 25.0 * n_glob / world > 126e6 else "fits the 126 MB L2 (stated, not flushed)"),
                           "setup_s": setup_s},
                "fine_dof_cycles_per_s": n_glob / (ms * 1e-3), "parity": parity,
                "roofline": B.roofline_block(dom, peak, peak_src, name if not multi else f"{name}-n{world}", prof,
                                             {"kernel": f"{dom['kind']}@level{dom['level']}" + (" (rank 0 shard)" if multi else "")}, describe=B._describe(eng)),
                "halo_ms_per_cycle_rank0": halo_ms, "profiled_cycle_ms_rank0": tot_ms,
                "cpu_baseline": cpu,
                "e2e": {"value": dofu / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": 16 * n_glob, "d2h_bytes_per_step": 8 * n_glob, "ms_per_step": e2e_s * 1e3,
                        "api": "mgb_vcycle(mem=MGB_MEM_HOST) on every rank's row block, pinned host buffers"},
                "gpu_launches": int(launches), "clocks": clk.summary(),
                "kernels": [{"k": f"{r['kind']}@{r['level']}", "ms": round(r["ms_per_launch"], 5), "n": r["launches"], "gbs": round(r["moved_gbs"], 1), "algorithmic_gbs": round(r["gbs"], 1)}
                            for r in sorted(prof, key=lambda r: -r["total_ms"])[:10]]}
        print(json.dumps(line), flush=True)
    barrier()
    mg.close()
    if multi:
        td.destroy_process_group()
    if parity["ok"] is False:
        raise SystemExit(f"parity FAILED: residual norms differ from the expected ones by {parity['rel']:.3e} relative (> {B.PARITY_TOL})")
