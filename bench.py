#!/usr/bin/env python
"""Benchmark of the hot path: geometric-multigrid V-cycles on synthetic Poisson hierarchies.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5|cfg2|cfg3|...] [--perm lex|rcm|random]

A "step" is ONE V-cycle (V_cycle_scheme, multigrid.py:231-268) on the named hierarchy.
Metric: smoother DOF-updates per second = (mu1 + mu2) * sum_{l > coarsest} n_l / (seconds per V-cycle)
(BASELINE.md section 4); ms_per_step is the V-cycle time.  One JSON line is printed by rank 0.

Default workload at EVERY N (1, 2, 4, 8): BASELINE config 5, the north-star target -- 3-D P1 Poisson, 513^3 = 135 M DOFs,
7-level V(2,2), weighted Jacobi, injection.  It fits one B200 (55 GB), its levels are generated on the device, and the same
problem is row-sharded for N > 1, so the driver's 1 -> 8 curve is one workload (strong scaling).  The other BASELINE
configurations are `--workload cfg1|cfg2|cfg3|cfg4` (host-assembled hierarchies; their lines are committed under profiles/).

  value      device-resident: v, f and the hierarchy already in HBM, CUDA-graph replay, CUDA-event timing
  e2e        the same cycles through the C-ABI call mgb_vcycle(mem = HOST) on pinned host buffers:
             H2D of v and f and D2H of v inside the timed region, every step
  roofline   dominant kernel (finest-level weighted-Jacobi sweep): bytes the kernel STREAMS per launch / event-timed
             duration (frac = that / measured HBM copy peak); the CSR-form algorithmic figure is kept as algorithmic_gbs
  parity     residual norms of 3 cycles from a zero guess against the CPU oracle on the same hierarchy (N = 1) or against the
             committed single-GPU values (N > 1); the run exits non-zero when they differ by more than 1e-12 relative
  cpu_baseline  the C/OpenMP port of the reference's V-cycle (oracle/mg_oracle.c) on the host cores, same hierarchy
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: (dim, c, coarsest, finest, description)        cells per dim at level l = c * 2**l
WORKLOADS = {
    "cfg1": (2, 8, 0, 2, "2D Poisson P1 33x33 nodes, 3-level V-cycle (reference CPU case)"),
    "cfg2": (2, 32, 0, 6, "2D Poisson P1 2049x2049 (4.2M DOFs), 7-level V(2,2), weighted Jacobi, injection"),
    "cfg2s": (2, 32, 0, 4, "2D Poisson P1 513x513, 5-level V(2,2) (largest grid the Python reference itself can run)"),
    "cfg3": (3, 8, 0, 4, "3D Poisson P1 129^3 Kuhn mesh (2.1M DOFs), 5-level V(2,2), weighted Jacobi, injection"),
    "cfg4": (3, 4, 0, 4, "3D Poisson P2 on 64^3 cells (129^3 = 2.1M DOFs, 60.9M stored entries, rows of 10..65), 5-level V(2,2)"),
    "cfg5h": (3, 8, 0, 5, "3D Poisson P1 257^3 Kuhn mesh (17M DOFs), 6-level V(2,2), weighted Jacobi, injection"),
    "cfg5": (3, 8, 0, 6, "3D Poisson P1 513^3 Kuhn mesh (135M DOFs), 7-level V(2,2), weighted Jacobi, injection"),
}
STRUCTURED = ("cfg5", "cfg5h")                 # generated on the device, never assembled on the host
METRIC = "V-cycle smoother DOF-updates/s"
UNIT = "DOF-updates/s"
PARITY_TOL = 1e-12
EXPECTED = os.path.join(ROOT, "profiles", "expected_resnorms.json")     # single-GPU residual norms (3 cycles from zero) per workload


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_ram_gb():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    except (ValueError, OSError):
        return 0.0


def cpu_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle's thread count is set explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, mu=2, perm="lex"):
    """Host-assembled hierarchy of a BASELINE configuration.  perm: DOF numbering handed to the engine -- 'lex'
    (lexicographic), 'random' (seeded permutation per level) or 'rcm' (reverse Cuthill-McKee of a random numbering: banded
    but not lexicographic, the closest stand-in for dolfinx's own numbering)."""
    from multigrid_dolfinx_b200 import problems as pr
    if name == "cfg4":
        return pr.build_hierarchy_p2(c=4, coarsest_level=0, finest_level=4, mu1=mu, mu2=mu), WORKLOADS[name][4]
    dim, c, lc, lf, desc = WORKLOADS[name]
    kw = {}
    if perm == "random":
        kw["perm_seed"] = 7
    elif perm == "rcm":
        kw["perm_seed"] = 7
        kw["rcm"] = True
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=lc, finest_level=lf, mu1=mu, mu2=mu, with_dicts=False, **kw)
    return H, desc


def dof_updates_per_cycle(H):
    return (H.mu1 + H.mu2) * sum(H.n(l) for l in H.levels() if l > H.coarsest_level)


def structured_rhs(n):
    """The deterministic synthetic load of the device-generated workloads (dist.StructuredSource.rhs_rows)."""
    i = np.arange(n, dtype=np.float64)
    return 1.0 + np.sin(0.001 * i) + 0.5 * np.cos(0.37 * i)


def cpu_oracle(workload, H=None):
    """-> (oracle V-cycle object with .vcycle(v, f, ncycles, history), f, dof-updates per cycle, threads, note).
    Structured workloads are built inside the C oracle (no scipy matrix of 2e9 entries ever exists)."""
    from oracle import c_oracle as co
    threads = co.set_threads(cpu_threads())
    if workload in STRUCTURED:
        dim, c, lc, lf, _ = WORKLOADS[workload]
        need = 70.0 if workload == "cfg5" else 9.0
        if host_ram_gb() < need:
            return None, None, 0, threads, f"host has {host_ram_gb():.0f} GB of RAM, the {workload} hierarchy needs ~{need:.0f} GB on the CPU side"
        S = co.StructuredCOracleMG(dim, c, lc, lf)
        return S, structured_rhs(S.n[-1]), S.dof_updates_per_cycle(), threads, "hierarchy built in place by the C oracle (same arrays as the device generator)"
    cm = co.from_hierarchy(H)
    return cm, H.b_dict[H.finest_level][:, 0], dof_updates_per_cycle(H), threads, "scipy-assembled hierarchy, the one the GPU arm uploads"


def time_cpu(cm, f, cycles, warm=1):
    v = np.zeros_like(f)
    v = cm.vcycle(v, f, ncycles=warm)
    t0 = time.perf_counter()
    cm.vcycle(v, f, ncycles=cycles)
    return (time.perf_counter() - t0) / cycles


def parity_block(hist_gpu, expected, source):
    if expected is None:
        return {"resnorm": [float(x) for x in hist_gpu], "expected": None, "rel": None, "source": source, "tol": PARITY_TOL, "ok": None}
    e = np.asarray(expected, dtype=np.float64)[:len(hist_gpu)]
    rel = float(np.max(np.abs(np.asarray(hist_gpu)[:len(e)] - e) / np.abs(e)))
    return {"resnorm": [float(x) for x in hist_gpu], "expected": [float(x) for x in e], "rel": rel, "source": source, "tol": PARITY_TOL,
            "ok": bool(rel <= PARITY_TOL)}


def expected_from_file(key):
    try:
        return json.load(open(EXPECTED)).get(key)
    except Exception:
        return None


def ncu_traffic(workload, kernel_kind, describe=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed `ncu --set full` capture of THIS
    workload's fine-level sweep (profiles/r2_ncu_traffic.json names the kernel it was taken from); None when none matches.
    The figure belongs to ONE kernel: the record carries the operator, kernel family and grid it was captured with ("expects"),
    and when the engine's own description of what it will launch (mgb_describe) is at hand and disagrees, the figure is withheld
    as stale instead of being attached to another kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        rec = t.get(f"{workload}:{kernel_kind}")
        if not rec:
            return None, None, None
        exp = rec.get("expects")
        if exp and describe is not None:
            want = f"{exp['family']}("
            block = describe.split(f"level {exp['level']} ", 1)[1].split("\nlevel ", 1)[0] if f"level {exp['level']} " in describe else ""
            line = next((l for l in block.splitlines() if l.strip().startswith(exp["operator"] + " ")), "")
            tiles = re.search(r"tiles=(\d+)\s*$", line)
            # (an operator line that cannot be found or read decides nothing: only a description that names another kernel family
            # or another grid withholds the figure)
            if line and (want not in line or (tiles is not None and int(tiles.group(1)) != int(exp["tiles"]))):
                return None, "stale", f"{rec['file']} was captured on {exp['operator']}@{exp['level']} = {exp['family']}, tiles={exp['tiles']}; the engine now reports: {line.strip() or 'no such operator'}"
        return float(rec["dram_bytes"]), rec["kernel"], rec["file"]
    except Exception:
        return None, None, None


def roofline_block(dom, peak, peak_src, workload, prof, extra=None, describe=None):
    traffic, tk, tf = ncu_traffic(workload, dom["kind"], describe)
    r = {"bound": "hbm", "kernel": f"{dom['kind']}@level{dom['level']}", "achieved": dom["moved_gbs"], "peak": peak, "unit": "GB/s",
         "frac": dom["moved_gbs"] / peak, "frac_of_8TBs": dom["moved_gbs"] / 8000.0, "traffic": traffic,
         "traffic_source": (f"withheld: {tf}" if tk == "stale" else f"{tf} ({tk})") if tf else None, "peak_source": peak_src,
         "bytes_per_launch": dom["moved_bytes"], "ms_per_launch": dom["ms_per_launch"],
         "algorithmic_bytes_per_launch": dom["bytes"], "algorithmic_gbs": dom["gbs"],
         "note": "achieved/frac: bytes the kernel streams per launch (one code byte per row + the vectors for pattern-coded operators) over the "
                 "event-timed duration; algorithmic_*: the CSR-form figure of SURVEY 8d (12 B per stored entry), which the coded kernels do not move",
         "share_of_cycle": dom["total_ms"] / sum(r["total_ms"] for r in prof)}
    if extra:
        r.update(extra)
    return r


def _describe(eng):
    try:
        return eng.describe()
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric/config, on all host cores.
    The reference is pure Python over scipy; it cannot run any BASELINE configuration beyond config 1 (its coordinate keys fail
    for h < 1/512, it is 2-D only: SURVEY M4), so this arm times the oracle port (oracle/mg_oracle.c, OpenMP over rows) on the
    SAME hierarchy the GPU arm runs.  `--workload cfg1|cfg2s` additionally times the unmodified reference (C1 of BASELINE.md)
    when /root/reference is present."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    desc = WORKLOADS[wl][4]
    H = None
    if wl not in STRUCTURED:
        H, _ = build_workload(wl, perm=args.perm)
    t0 = time.perf_counter()
    cm, f, dofu, threads, note = cpu_oracle(wl, H)
    if cm is None:
        print(json.dumps({"impl": "reference", "unavailable": note}), flush=True)
        return
    build_s = time.perf_counter() - t0
    steps = max(min(args.steps, 3 if wl == "cfg5" else 10), 1)
    per = time_cpu(cm, f, steps, max(min(args.warmup, 1 if wl == "cfg5" else 2), 1))
    val = dofu / per
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": f"{wl}: {desc}"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{steps} full V-cycles of {wl} after warm-up (the hierarchy the GPU arm runs; {note}); C/OpenMP port of the "
                                       f"reference with omp_set_num_threads({threads}) of {os.cpu_count()} host cores; build {build_s:.0f} s untimed"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    c1 = python_reference_time(wl, H) if wl in ("cfg1", "cfg2s") else None
    if c1:
        line["python_reference"] = c1
    print(json.dumps(line), flush=True)


def python_reference_time(wl, H):
    """C1 of BASELINE.md: the UNMODIFIED reference (multigrid.py imported from /root/reference, one Python thread) where it can
    run.  /root/reference does not exist on the GPU box; there the figure measured in the development container is reported
    with its provenance (profiles/r2_python_reference.json)."""
    if os.path.isdir("/root/reference"):
        from multigrid_dolfinx_b200 import problems as pr
        from oracle import reference_import as ri
        dim, c, lc, lf, _ = WORKLOADS[wl]
        Hd = pr.build_hierarchy(dim=dim, c=c, coarsest_level=lc, finest_level=lf, mu1=2, mu2=2, with_dicts=True)    # the reference needs its coordinate dicts
        ri.run_reference_vcycles(Hd, 1)
        cyc = 3 if wl == "cfg1" else 1
        t0 = time.perf_counter()
        ri.run_reference_vcycles(Hd, cyc)
        per = (time.perf_counter() - t0) / cyc
        return {"kind": "reference", "where": "this process: /root/reference/multigrid.py imported unmodified, V_cycle_scheme, one Python thread",
                "ms_per_cycle": per * 1e3, "value": dof_updates_per_cycle(Hd) / per, "unit": UNIT, "cores": 1}
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_python_reference.json"))).get(wl)
    except Exception:
        return None


def run_single(args):
    """Host-assembled workloads (cfg1 .. cfg4) on one GPU."""
    import torch
    from multigrid_dolfinx_b200.engine import MGEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    t_setup = time.perf_counter()
    H, desc = build_workload(args.workload, perm=args.perm)
    lf = H.finest_level
    opts = {"fuse_restrict": args.fuse_restrict, "stream_cfg": args.stream_cfg, "compress": args.compress, "code_cfg": args.code_cfg}
    opts.update(args.options)
    eng = MGEngine.from_hierarchy(H, r_mode=args.restriction, smoother=args.smoother, device=0, options=opts,
                                  reorder=bool(args.reorder) and args.perm != "lex")
    n = H.n(lf)
    f_host = H.b_dict[lf][:, 0]
    t_setup = time.perf_counter() - t_setup
    dofu = dof_updates_per_cycle(H)
    stream = eng.torch_stream()

    # ---- parity: 3 cycles from a zero guess, residual norm after each ------------------------------------
    eng.level_buffer(lf, "f").copy_(torch.from_numpy(f_host))
    eng.level_buffer(lf, "v").zero_()
    torch.cuda.synchronize()
    hist3 = eng.vcycle_resident(lf, 3, history=True)

    # ---- device-resident timing ----------------------------------------------------------------
    eng.vcycle_resident(lf, args.warmup)
    eng.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        torch.cuda.synchronize()
        e0.record(stream)
        eng.vcycle_resident(lf, args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = eng.launch_count() - l0
        busy_ms = e0.elapsed_time(e1)
        if busy_ms < 1000.0:                 # keep the GPU busy ~1 s in total so that nvidia-smi samples it under load
            eng.vcycle_resident(lf, int(min(4000, max(1, (1000.0 - busy_ms) / max(busy_ms / args.steps, 1e-3)))))
            torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps

    # ---- dominant kernel, event-timed launch by launch (same cycles, graph off) ----------------------
    eng.profile_begin()
    eng.vcycle_resident(lf, max(3, min(args.steps, 10)))
    prof = eng.profile_end()
    dom = max(prof, key=lambda r: r["total_ms"])
    peak, peak_src = measured_peak()
    cyc_ms_prof = sum(r["total_ms"] for r in prof) / max(3, min(args.steps, 10))
    roofline = roofline_block(dom, peak, peak_src, args.workload if args.perm == "lex" else f"{args.workload}-{args.perm}", prof,
                              {"vcycle_bytes_moved": eng.vcycle_bytes_moved(lf), "vcycle_moved_gbs": eng.vcycle_bytes_moved(lf) / (ms * 1e-3) / 1e9,
                               "vcycle_algorithmic_bytes": eng.vcycle_bytes(lf)}, describe=_describe(eng))
    kernels = sorted(prof, key=lambda r: -r["total_ms"])[:8]

    # ---- end to end through the C ABI with pinned host buffers -----------------------------------------
    vp = torch.zeros(n, dtype=torch.float64).pin_memory()
    fp = torch.from_numpy(f_host.copy()).pin_memory()
    lib, h = eng._lib, eng._h
    for _ in range(max(1, args.warmup)):
        eng._ck(lib.mgb_vcycle(h, lf, vp.data_ptr(), fp.data_ptr(), 0, 1, None))
    e2e_steps = max(3, min(args.steps, 20))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng._ck(lib.mgb_vcycle(h, lf, vp.data_ptr(), fp.data_ptr(), 0, 1, None))    # returns after v is back on the host
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": dofu / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 8 * n,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "api": "mgb_vcycle(mem=MGB_MEM_HOST), pinned host v/f"}

    # ---- CPU baseline + parity against it ------------------------------------------------------------
    cpu, parity = None, parity_block(hist3, None, "no CPU oracle run (--no-cpu)")
    if not args.no_cpu:
        cm, f, _, threads, note = cpu_oracle(args.workload, H)
        cyc = 5
        per = time_cpu(cm, f, cyc, 1)
        _, hist_cpu = cm.vcycle(np.zeros_like(f), f, ncycles=3, history=True)
        cpu = {"value": dofu / per, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_cycle": per * 1e3,
               "sample": f"{cyc} full V-cycles of {args.workload} (same hierarchy), C/OpenMP port of the reference, {threads} threads of {os.cpu_count()} cores"}
        parity = parity_block(hist3, hist_cpu, "CPU oracle (oracle/mg_oracle.c) on the same hierarchy, 3 cycles from a zero guess")

    line = {"metric": METRIC, "value": dofu / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "numbering": args.perm, "lattice_numbering_handed_over": bool(args.reorder) and args.perm != "lex", "restriction": args.restriction, "smoother": args.smoother,
                       "fine_dofs": n, "levels": lf - H.coarsest_level + 1, "mu1": H.mu1, "mu2": H.mu2, "omega": H.omega,
                       "compress": args.compress, "options": args.options,
                       "l2": "fine-level working set of one sweep (codes + 3 vectors) vs the 126 MB L2: " +
                             ("exceeds it, no flush needed" if 25.0 * n > 126e6 else "FITS it -- consecutive kernels find operands in L2 (stated, not flushed: "
                              "the cycle is the unit of work and its kernels do hand vectors to each other)"), "setup_s": t_setup},
            "fine_dof_cycles_per_s": n / (ms * 1e-3), "parity": parity,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
            "kernels": [{"k": f"{r['kind']}@{r['level']}", "ms": round(r["ms_per_launch"], 5), "n": r["launches"], "gbs": round(r["moved_gbs"], 1),
                         "algorithmic_gbs": round(r["gbs"], 1)} for r in kernels],
            "profiled_cycle_ms": cyc_ms_prof}
    print(json.dumps(line), flush=True)
    eng.close()
    if parity["ok"] is False:
        raise SystemExit(f"parity FAILED: residual norms differ from the oracle by {parity['rel']:.3e} relative (> {PARITY_TOL})")


def parse_options(s):
    out = {}
    for kv in (s or "").split(","):
        if kv:
            k, v = kv.split("=")
            out[k] = float(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--perm", default="lex", choices=["lex", "rcm", "random"], help="DOF numbering of the host-assembled workloads")
    ap.add_argument("--reorder", type=int, default=1, help="hand the engine the lattice numbering of a permuted workload (mgb_set_numbering), as the "
                    "drop-in module does from the reference's coordinate dicts; 0: run in the caller's numbering")
    ap.add_argument("--gather-threshold", type=int, default=300000)
    ap.add_argument("--use-graph", type=int, default=1)
    ap.add_argument("--overlap", type=int, default=0, help="overlap the halo exchange with interior rows (sharded runs)")
    ap.add_argument("--overlap-waves", type=int, default=2)
    ap.add_argument("--p2p", type=int, default=1, help="halo exchange through NVLink peer memory (0: ncclSend/ncclRecv)")
    ap.add_argument("--device-gen", type=int, default=1, help="generate the sharded levels on the device (structured workloads)")
    ap.add_argument("--restriction", default="injection", choices=["injection", "full_weighting", "transpose"])
    ap.add_argument("--smoother", default="jacobi", choices=["jacobi", "jacobi_a", "gs", "gs_color"])
    ap.add_argument("--fuse-restrict", type=int, default=1)
    ap.add_argument("--stream-cfg", type=int, default=3)
    ap.add_argument("--compress", type=int, default=3, help="lossless operator coding: 0 CSR stream kernels only, 1 one byte per entry, 2 + row patterns, 3 + anchored row patterns (P)")
    ap.add_argument("--code-cfg", type=int, default=1)
    ap.add_argument("--options", default="", help="further engine options, k=v,k=v (mgb_set_option)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.options = parse_options(args.options)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    multi = args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1
    if args.workload == "auto":
        args.workload = "cfg5"               # the north-star configuration, at every N (one workload across the scaling run)
    if args.workload not in WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}")
    if args.impl == "reference":
        return run_reference(args)
    if multi or args.workload in STRUCTURED:      # generated on the device: same code path as the sharded arm, world = 1
        import bench_dist
        return bench_dist.run(args)
    return run_single(args)


if __name__ == "__main__":
    main()
