"""bench.py's JSON contract, checked on the CPU through the reference arm (the GPU arm prints the same keys plus
roofline / cpu_baseline / clocks and is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in r, k
    assert r["impl"] == "reference" and r["dtype"] == "f64" and r["vs_baseline"] is None and r["higher_is_better"] is True
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in r["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cfg1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_json_assembly_with_a_stub_engine(monkeypatch, capsys):
    """The GPU arm's bookkeeping (everything in run_single that is not a kernel: byte accounting, roofline / moved-bytes
    fields, e2e and cpu_baseline objects, the JSON line itself) run on the CPU against a stub engine.  Numbers are
    meaningless here; the point is that the line the driver parses keeps every contract key."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import multigrid_dolfinx_b200.engine as E

    class Ev:
        def __init__(self, enable_timing=False): pass
        def record(self, s=None): pass
        def elapsed_time(self, other): return 12.5

    class Lib:
        def mgb_vcycle(self, *a): return 0

    class Eng:
        def __init__(self, H): self.H, self._lib, self._h, self.bufs, self.count = H, Lib(), None, {}, 0
        def torch_stream(self): return None
        def level_buffer(self, l, w): return self.bufs.setdefault((l, w), torch.zeros(self.H.n(l), dtype=torch.float64))
        def vcycle_resident(self, lf, n, history=False):
            self.count += 37 * n
            return np.ones(n) if history else None
        def synchronize(self): pass
        def launch_count(self): return self.count
        def profile_begin(self): pass
        def profile_end(self):
            return [{"kind": "jacobi", "level": 2, "launches": 4, "total_ms": 0.1, "ms_per_launch": 0.025, "bytes": 3e8, "gbs": 1.2e4,
                     "moved_bytes": 1e8, "moved_gbs": 4e3},
                    {"kind": "residual", "level": 2, "launches": 1, "total_ms": 0.02, "ms_per_launch": 0.02, "bytes": 4e8, "gbs": 2e4,
                     "moved_bytes": 7e7, "moved_gbs": 3.5e3}]
        def vcycle_bytes(self, l): return 2.6e9
        def vcycle_bytes_moved(self, l): return 0.9e9
        def _ck(self, rc): assert rc == 0
        def close(self): pass

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(E.MGEngine, "from_hierarchy", classmethod(lambda cls, H, **kw: Eng(H)))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "cfg1", "--steps", "3", "--warmup", "3"])
    import pytest
    with pytest.raises(SystemExit, match="parity FAILED"):
        bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in r, k
    assert r["n_gpus"] == 1 and r["dtype"] == "f64" and r["vs_baseline"] is None and r["data"] == "synthetic" and "workload" in r["config"]
    rf = r["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "bytes_per_launch", "algorithmic_bytes_per_launch", "algorithmic_gbs"):
        assert k in rf, k
    # frac is a PHYSICAL fraction: bytes the kernel streams / time / measured peak (the CSR-form figure is algorithmic_gbs)
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["kernel"] == "jacobi@level2" and rf["bytes_per_launch"] == 1e8 and rf["achieved"] == 4e3 and rf["algorithmic_gbs"] == 1.2e4
    # the stub's residual norms (all ones) cannot match the oracle's: the parity block must say so and the run must fail
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb and cb["unit"] == r["unit"]
    e = r["e2e"]
    n = r["config"]["fine_dofs"]
    assert e["h2d_bytes_per_step"] == 16 * n and e["d2h_bytes_per_step"] == 8 * n and e["unit"] == r["unit"] and e["value"] > 0
    assert r["gpu_launches"] == 37 * 3
    p = r["parity"]
    assert p["ok"] is False and p["rel"] > 1e-12 and len(p["resnorm"]) == 3 and len(p["expected"]) == 3 and p["tol"] == 1e-12


def test_default_workload_is_config5_at_every_gpu_count():
    """One workload across the driver's 1 -> 8 scaling run (N = 1 used to run config 2 and N > 1 config 5)."""
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'args.workload = "cfg5"' in src and '"cfg2" if' not in src and "else \"cfg2\"" not in src
    assert bench.WORKLOADS["cfg5"][:4] == (3, 8, 0, 6) and "cfg5" in bench.STRUCTURED


def test_reference_arm_sets_the_thread_count_explicitly():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use the cores it reports."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    import bench
    assert r["cpu_baseline"]["cores"] == bench.cpu_threads() and f"omp_set_num_threads({bench.cpu_threads()})" in r["cpu_baseline"]["sample"]


def test_ncu_traffic_is_withheld_when_the_engine_describes_another_kernel():
    """roofline.traffic comes from a committed ncu capture of ONE kernel; it must not be attached to a different one."""
    sys.path.insert(0, ROOT)
    import bench
    good = ("level 5 n=16974593\n  A   rows=16974593 nnz=1 max_row=15 hotrow(coded mode=3: 9 row patterns; cfg=1, 128 x 2 rows) tiles=66308\n"
            "level 6 n=135005697\n  A   rows=135005697 nnz=2018775553 max_row=15 hotrow(coded mode=3: 27 row patterns, 100 table entries, hot pattern 3 of 15 entries, "
            "0 patterns take the table walk; cfg=1, 128 x 2 rows) tiles=527367\n"
            "  RJ  rows=135005697 nnz=799000000 max_row=6 hotrow(coded mode=3: 27 row patterns, 90 table entries, hot pattern 1 of 6 entries, "
            "0 patterns take the table walk; cfg=1, 128 x 2 rows) tiles=527367\n  P   rows=135005697 nnz=4 max_row=8 anchrow(coded mode=4) tiles=527367\n")
    t, kern, src = bench.ncu_traffic("cfg5", "jacobi", good)
    assert t and t > 3e9 and "k_hotrow" in kern and src.startswith("profiles/")
    assert bench.ncu_traffic("cfg5", "jacobi", None)[0] == t                        # no description at hand: the record as committed
    other_kernel = good.replace("max_row=6 hotrow(", "max_row=6 rowstream(")
    other_grid = good.replace("0 patterns take the table walk; cfg=1, 128 x 2 rows) tiles=527367\n  P ", "0 patterns take the table walk; cfg=2, 256 x 1 rows) tiles=263684\n  P ")
    assert bench.ncu_traffic("cfg5", "jacobi", "level 6 n=1\n")[0] == t            # no RJ line to compare with: nothing is decided
    for d in (other_kernel, other_grid):
        t2, tag, why = bench.ncu_traffic("cfg5", "jacobi", d)
        assert t2 is None and tag == "stale" and "captured on RJ@6" in why
    dom = {"kind": "jacobi", "level": 6, "moved_gbs": 5000.0, "moved_bytes": 3.4e9, "ms_per_launch": 0.6, "bytes": 1.3e10, "gbs": 2e4, "total_ms": 1.0}
    r = bench.roofline_block(dom, 6538.6, "measured", "cfg5", [dom], describe=other_kernel)
    assert r["traffic"] is None and r["traffic_source"].startswith("withheld: ")
    r = bench.roofline_block(dom, 6538.6, "measured", "cfg5", [dom], describe=good)
    assert r["traffic"] == t and "k_hotrow" in r["traffic_source"]
    assert bench.ncu_traffic("cfg2", "jacobi", good) == (None, None, None)          # no capture for that workload
