"""GPU parity tests (call through the C ABI).  Bars (BASELINE.json north_star):
  * sparsity patterns / level orderings / colourings: bit-exact
  * per-cycle fp64 residual norms: 1e-12 relative; final solutions: 1e-10
The tile kernels sum every row in stored order without FMA, so every operator except the coarsest
solve is in fact required to be BIT-EXACT against the golden vectors the reference produced."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_SMALL, ROOT, load_golden
from multigrid_dolfinx_b200 import _lib as L
from multigrid_dolfinx_b200 import problems as pr
from multigrid_dolfinx_b200.engine import MGEngine
from oracle import c_oracle as co
from oracle import restated as rs

pytestmark = pytest.mark.gpu

RTOL_RESNORM = 1e-12
RTOL_SOLUTION = 1e-10


def _report(name, **kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in kw.items()}}, default=str) + "\n")


def relmax(a, b):
    return float(np.abs(np.asarray(a).ravel() - np.asarray(b).ravel()).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_operators_bit_exact_vs_reference_golden(name):
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    lf = H.finest_level
    eng = MGEngine.from_hierarchy(H)
    # artefacts: R_omega pattern + values, D^-1  (getJacobiMatrices, multigrid.py:48-56)
    assert np.array_equal(eng.artifact(lf, L.ART_RJ_INDPTR), d["rj_indptr"])
    assert np.array_equal(eng.artifact(lf, L.ART_RJ_INDICES), d["rj_indices"])
    assert np.array_equal(eng.artifact(lf, L.ART_RJ_VALUES), d["rj_data"])
    assert np.array_equal(eng.artifact(lf, L.ART_DINV), d["rj_dinv"])
    x, g, e = d["in_x"], d["in_g"], d["in_e"]
    assert np.array_equal(eng.smooth(lf, x, g, 1), d["jac_1"])            # jacobiRelaxation, multigrid.py:223-228
    assert np.array_equal(eng.smooth(lf, x, g, 5), d["jac_5"])
    assert np.array_equal(eng.prolong_add(lf, e, np.zeros_like(x)), d["interp"])   # Interpolation2D, multigrid.py:59-120
    assert np.array_equal(eng.restrict(lf, x), d["inj"])                  # Restriction2D_direct, multigrid.py:123-132
    A = H.A_sp_dict[lf][0]
    assert np.array_equal(eng.spmv(lf, x), A.dot(x))                      # multigrid.py:244
    assert np.array_equal(eng.residual(lf, x, g), g - A.dot(x))
    eng.close()
    eng = MGEngine.from_hierarchy(H, r_mode="full_weighting")
    fw = eng.restrict(lf, x)                                              # Restriction2D, multigrid.py:135-198
    assert np.abs(fw - d["fw"]).max() <= 4e-16 * np.abs(d["fw"]).max()
    eng.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_vcycle_matches_reference_golden(name):
    """V_cycle_scheme (multigrid.py:231-268): K consecutive cycles from v0 = 0."""
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    lf = H.finest_level
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[lf]
    v = np.zeros_like(f)
    A = H.A_sp_dict[lf][0]
    worst_r = worst_v = 0.0
    for k in range(K):
        v = eng.vcycle(lf, v, f)
        rn = np.linalg.norm(f - A.dot(v))
        worst_r = max(worst_r, abs(rn - d["vcycle_resnorm"][k]) / d["vcycle_resnorm"][k])
        if d["vcycle_v"].shape[0] == K:
            worst_v = max(worst_v, relmax(v, d["vcycle_v"][k]))
    worst_v = max(worst_v, relmax(v, d["vcycle_v"][-1]))
    # same thing in one call with the device-side norm history
    v2, hist = eng.vcycle(lf, np.zeros_like(f), f, ncycles=K, history=True)
    worst_h = float(np.abs(hist - d["vcycle_resnorm"]).max() / d["vcycle_resnorm"].max())
    _report("vcycle_golden", case=name, resnorm_rel=worst_r, solution_rel=worst_v, hist_rel=worst_h)
    assert worst_r <= RTOL_RESNORM and worst_h <= RTOL_RESNORM
    assert worst_v <= RTOL_SOLUTION
    assert np.array_equal(v2, v)
    eng.close()


@pytest.mark.parametrize("name", ["cfg1_perm_mu2", "proto_perm_mu50"])
def test_debug_tuple_matches_reference(name):
    """test=True 4-tuple (multigrid.py:262-266) of the K-th cycle."""
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    lf = H.finest_level
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[lf]
    v = np.zeros_like(f)
    for k in range(K - 1):
        v = eng.vcycle(lf, v, f)
    v, f2h, v2h, errh = eng.vcycle_debug(lf, v, f)
    assert relmax(v, d["vcycle_v"][-1]) <= RTOL_SOLUTION
    assert relmax(f2h, d["dbg_f2h"]) <= 1e-10 and relmax(v2h, d["dbg_v2h"]) <= 1e-10 and relmax(errh, d["dbg_errh"]) <= 1e-10
    eng.close()


@pytest.mark.parametrize("dim,c,lc,lf,seed", [(2, 8, 0, 5, None), (2, 8, 0, 4, 3), (3, 2, 0, 4, None), (3, 2, 0, 3, 5)])
@pytest.mark.parametrize("r_mode", ["injection", "transpose", "full_weighting"])
def test_vcycle_vs_oracle_larger(dim, c, lc, lf, seed, r_mode):
    """Sizes the reference cannot run (3-D, > 513^2): C oracle on the same seeded inputs."""
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=lc, finest_level=lf, perm_seed=seed, with_dicts=False)
    eng = MGEngine.from_hierarchy(H, r_mode=r_mode)
    cm = co.from_hierarchy(H, r_mode=r_mode)
    f = H.b_dict[lf][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=5, history=True)
    vg, hg = eng.vcycle(lf, np.zeros_like(f), f, ncycles=5, history=True)
    r = float(np.abs(hg - ho).max() / ho.max()); s = relmax(vg, vo)
    _report("vcycle_oracle", dim=dim, n=H.n(lf), seed=seed, r_mode=r_mode, resnorm_rel=r, solution_rel=s)
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    eng.close()


@pytest.mark.parametrize("family,lanes", [(2, 0), (2, 4), (2, 32)])
def test_subwarp_family_within_tolerance(family, lanes):
    H = pr.build_hierarchy(dim=3, c=2, coarsest_level=0, finest_level=3, perm_seed=2, with_dicts=False)
    eng = MGEngine.from_hierarchy(H, options={"kernel_family": family, "lanes_per_row": lanes})
    cm = co.from_hierarchy(H)
    f = H.b_dict[3][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=4, history=True)
    vg, hg = eng.vcycle(3, np.zeros_like(f), f, ncycles=4, history=True)
    assert np.abs(hg - ho).max() / ho.max() <= RTOL_RESNORM and relmax(vg, vo) <= RTOL_SOLUTION
    assert "subwarp" in eng.describe()
    eng.close()


@pytest.mark.parametrize("opts", [{"tile_iter": 1}, {"tile_iter": 2}, {"use_graph": 0}, {"fuse_restrict": 1}, {"rj_order": 0},
                                  {"stream_cfg": 1}, {"stream_cfg": 2}, {"stream_cfg": 3}, {"stream_cfg": 4}, {"stream_cfg": 5}, {"stream_cfg": 6},
                                  {"stream_cfg": 1, "fuse_restrict": 0}, {"pdl": 1}, {"pdl": 0}, {"pdl": 1, "compress": 0}, {"stream_cfg": 7}, {"code_cfg": 3},
                                  {"anch_tiles": -2}, {"anch_tiles": -5}, {"anch_tiles": -64}])
def test_kernel_options_do_not_change_results(opts):
    for dim, c, lf, seed, r_mode in [(2, 8, 4, 1, "injection"), (3, 2, 3, None, "transpose"), (2, 5, 3, None, "full_weighting")]:
        H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
        f = H.b_dict[lf][:, 0]
        base = MGEngine.from_hierarchy(H, r_mode=r_mode, options={"stream_cfg": 0, "fuse_restrict": 0})
        v0 = base.vcycle(lf, np.zeros_like(f), f, ncycles=3)
        eng = MGEngine.from_hierarchy(H, r_mode=r_mode, options=opts)
        v1 = eng.vcycle(lf, np.zeros_like(f), f, ncycles=3)
        if "rj_order" in opts:
            assert relmax(v1, v0) <= 1e-13          # different summation order inside R_omega rows
        else:
            assert np.array_equal(v1, v0)           # bit-identical
        base.close(); eng.close()


@pytest.mark.parametrize("dim,c,lf,mu", [(2, 8, 6, (2, 2)), (3, 4, 4, (2, 2)), (2, 8, 6, (3, 5)), (3, 2, 5, (4, 1))])
@pytest.mark.parametrize("slack,tiles", [(0, 1), (7, 4), (2048, 8), (300, 64)])
def test_two_sweeps_per_launch_bit_identical(dim, c, lf, mu, slack, tiles):
    """k_hotrow2 (pairs of Jacobi sweeps in one launch, the second trailing the first through L2) against one launch per sweep:
    the same bits, for every trailing distance -- slack 0 makes almost every second-sweep tile wait for its neighbours' first-sweep
    tiles, i.e. it is the dependency tracking that is under test -- for odd sweep counts, on graph replay and eagerly, through the
    smoother entry point, and against the oracle's sweep."""
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, with_dicts=False, mu1=mu[0], mu2=mu[1])
    f = H.b_dict[lf][:, 0]
    base = MGEngine.from_hierarchy(H, options={"fuse_sweeps": 0})
    eng = MGEngine.from_hierarchy(H, options={"fuse_sweeps": 1, "s2_min_rows": 1000, "s2_slack": slack, "s2_tiles": tiles})
    v0, h0 = base.vcycle(lf, np.zeros_like(f), f, ncycles=4, history=True)
    v1, h1 = eng.vcycle(lf, np.zeros_like(f), f, ncycles=4, history=True)
    assert np.array_equal(v1, v0) and np.array_equal(h1, h0)
    kinds = {r["kind"] for r in _profile_of(eng, lf)}
    assert "jacobi2" in kinds
    eng.set_option("use_graph", 0)
    assert np.array_equal(eng.vcycle(lf, np.zeros_like(f), f, ncycles=4), v0)
    rng = np.random.default_rng(3)
    x, g = rng.standard_normal(H.n(lf)), rng.standard_normal(H.n(lf))
    for nsweeps in (2, 3, 4, 5):
        assert np.array_equal(eng.smooth(lf, x, g, nsweeps), base.smooth(lf, x, g, nsweeps)), nsweeps
    A = H.A_sp_dict[lf][0]
    RO, dinv = rs.jacobi_matrices(A)
    assert np.array_equal(eng.smooth(lf, x, g, 4), rs.jacobi_relaxation(RO, dinv, x, g, 4, H.omega))
    base.close(); eng.close()


def _profile_of(eng, lf):
    eng.profile_begin()
    eng.vcycle_resident(lf, 1)
    return eng.profile_end()


def test_jacobi_a_form_and_coarse_refine():
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=3, perm_seed=4, with_dicts=False)
    f = H.b_dict[3][:, 0]
    cm = co.from_hierarchy(H, smoother="jacobi_a")
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=4, history=True)
    eng = MGEngine.from_hierarchy(H, smoother="jacobi_a", options={"coarse_refine": 1})
    vg, hg = eng.vcycle(3, np.zeros_like(f), f, ncycles=4, history=True)
    assert np.abs(hg - ho).max() / ho.max() <= RTOL_RESNORM and relmax(vg, vo) <= RTOL_SOLUTION
    eng.close()


@pytest.mark.parametrize("dim,c,lf,seed", [(2, 8, 3, None), (2, 8, 3, 6), (3, 2, 3, None), (3, 2, 2, 7)])
def test_gauss_seidel_level_scheduled_bit_exact(dim, c, lf, seed):
    """Level-scheduled GS must reproduce the sequential natural-order sweep; artefacts bit-exact."""
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    A = H.A_sp_dict[lf][0]
    eng = MGEngine.from_hierarchy(H, smoother="gs")
    lev, order, off = co.level_sets(A)
    assert np.array_equal(eng.artifact(lf, L.ART_LEVEL_OF_ROW), lev)
    assert np.array_equal(eng.artifact(lf, L.ART_LEVEL_ORDER), order)
    assert np.array_equal(eng.artifact(lf, L.ART_LEVEL_OFFSETS), off)
    rng = np.random.default_rng(0)
    x, f = rng.standard_normal(H.n(lf)), rng.standard_normal(H.n(lf))
    xo = x.copy()
    for _ in range(3):
        xo = co.gs_forward(A, xo, f)
    assert np.array_equal(eng.smooth(lf, x, f, 3), xo)
    for variant in (0, 1):                       # grid-barrier and plain cluster variants (default 2 = cluster + ELL pipeline)
        eng2 = MGEngine.from_hierarchy(H, smoother="gs", options={"gs_cluster": variant})
        assert np.array_equal(eng2.smooth(lf, x, f, 3), xo)
        eng2.close()
    cm = co.from_hierarchy(H, smoother="gs")
    b = H.b_dict[lf][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(b), b, ncycles=3, history=True)
    vg, hg = eng.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    assert np.abs(hg - ho).max() / ho.max() <= RTOL_RESNORM and relmax(vg, vo) <= RTOL_SOLUTION
    eng.close()


@pytest.mark.parametrize("dim,c,lf,seed", [(2, 8, 3, None), (2, 8, 3, 6), (3, 2, 3, None), (3, 2, 2, 7)])
def test_gauss_seidel_multicolour(dim, c, lf, seed):
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    A = H.A_sp_dict[lf][0]
    eng = MGEngine.from_hierarchy(H, smoother="gs_color")
    col, order, off = co.greedy_colouring(A)
    assert np.array_equal(eng.artifact(lf, L.ART_COLOUR_OF_ROW), col)
    assert np.array_equal(eng.artifact(lf, L.ART_COLOUR_ORDER), order)
    assert np.array_equal(eng.artifact(lf, L.ART_COLOUR_OFFSETS), off)
    rng = np.random.default_rng(0)
    x, f = rng.standard_normal(H.n(lf)), rng.standard_normal(H.n(lf))
    xo = x.copy()
    for _ in range(2):
        xo = co.gs_forward(A, xo, f, order)
    assert np.array_equal(eng.smooth(lf, x, f, 2), xo)
    cm = co.from_hierarchy(H, smoother="gs_color")
    b = H.b_dict[lf][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(b), b, ncycles=3, history=True)
    vg, hg = eng.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    assert np.abs(hg - ho).max() / ho.max() <= RTOL_RESNORM and relmax(vg, vo) <= RTOL_SOLUTION
    eng.close()


GS_ARTS = (L.ART_LEVEL_OF_ROW, L.ART_LEVEL_ORDER, L.ART_LEVEL_OFFSETS, L.ART_COLOUR_OF_ROW, L.ART_COLOUR_ORDER, L.ART_COLOUR_OFFSETS)
R_ARTS = (L.ART_R_INDPTR, L.ART_R_INDICES, L.ART_R_VALUES)


@pytest.mark.parametrize("dim,c,lf,seed,smoother,r_mode", [
    (2, 8, 3, None, "gs", "full_weighting"), (2, 8, 3, 6, "gs_color", "transpose"), (3, 2, 3, None, "gs_color", "full_weighting"),
    (3, 2, 2, 7, "gs", "full_weighting"), (2, 5, 3, 3, "gs", "transpose")])
def test_device_setup_matches_host_setup(dim, c, lf, seed, smoother, r_mode):
    """Set-up part 2 on the device (mgb_devsetup.cu, option device_setup = 1): the transposed restriction, the level sets, the
    first-fit colouring, their stable orders and the reordered Gauss-Seidel operator are the SAME ARRAYS as the host build of
    mgb_setup.cpp -- and as the C oracle's -- on lexicographic and on scrambled (ragged) numberings; so are sweeps and cycles."""
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    host = MGEngine.from_hierarchy(H, smoother=smoother, r_mode=r_mode, options={"device_setup": 0})
    dev = MGEngine.from_hierarchy(H, smoother=smoother, r_mode=r_mode, options={"device_setup": 1})
    for l in range(1, lf + 1):
        for kind in GS_ARTS + R_ARTS:
            a, b = host.artifact(l, kind), dev.artifact(l, kind)
            assert a.size > 0 and np.array_equal(a, b), (l, kind)
        A = H.A_sp_dict[l][0]
        lev, order, off = co.level_sets(A)
        assert np.array_equal(dev.artifact(l, L.ART_LEVEL_OF_ROW), lev) and np.array_equal(dev.artifact(l, L.ART_LEVEL_ORDER), order)
        assert np.array_equal(dev.artifact(l, L.ART_LEVEL_OFFSETS), off)
        col, order, off = co.greedy_colouring(A)
        assert np.array_equal(dev.artifact(l, L.ART_COLOUR_OF_ROW), col) and np.array_equal(dev.artifact(l, L.ART_COLOUR_ORDER), order)
        assert np.array_equal(dev.artifact(l, L.ART_COLOUR_OFFSETS), off)
    rng = np.random.default_rng(2)
    x, f = rng.standard_normal(H.n(lf)), rng.standard_normal(H.n(lf))
    assert np.array_equal(host.smooth(lf, x, f, 3), dev.smooth(lf, x, f, 3))
    assert np.array_equal(host.restrict(lf, x), dev.restrict(lf, x))
    b = H.b_dict[lf][:, 0]
    v0, h0 = host.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    v1, h1 = dev.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    assert np.array_equal(v0, v1) and np.array_equal(h0, h1)
    host.close(); dev.close()


@pytest.mark.parametrize("dim,c,lf,glevel,smoother,r_mode", [(2, 8, 4, 1, "gs", "full_weighting"), (3, 2, 4, 0, "gs_color", "full_weighting"),
                                                             (3, 4, 3, 2, "gs", "injection"), (2, 16, 3, 0, "jacobi", "transpose")])
def test_generated_levels_take_gauss_seidel_and_transposed_restrictions(dim, c, lf, glevel, smoother, r_mode):
    """Levels generated on the device have no host copy: their Gauss-Seidel operators and 2^-d P^T are built from the arrays in
    HBM and must equal what the host path builds from the host-assembled twin of the same hierarchy -- artefacts and cycles."""
    from multigrid_dolfinx_b200 import dist as ds
    src = ds.StructuredSource(dim, c, 0, lf)
    mg = ds.DistMG(src, device=0, gather_level=glevel, device_gen=True, smoother=smoother, r_mode=r_mode)
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, with_dicts=False)
    ref = MGEngine.from_hierarchy(H, smoother=smoother, r_mode=r_mode)
    for l in range(glevel + 1, lf + 1):
        for kind in (GS_ARTS if smoother != "jacobi" else ()) + (R_ARTS if r_mode != "injection" else ()):
            a, b = ref.artifact(l, kind), mg.eng.artifact(l, kind)
            assert a.size > 0 and np.array_equal(a, b), (l, kind)
    f = src.rhs_rows(lf, 0, src.n(lf))
    mg.load_rhs()
    h1 = mg.cycles(3, history=True)
    v0, h0 = ref.vcycle(lf, np.zeros_like(f), f, ncycles=3, history=True)
    assert np.array_equal(mg.local_solution(), v0) and np.array_equal(h1, h0)
    assert h0[2] < h0[0]
    mg.close(); ref.close()


def test_edge_cases():
    """nw = 0 returns v (multigrid.py:223-228); mu1 = 0 / mu2 = 0 / odd sweep totals; top level = any level
    (FullMultiGrid calls the cycle with every level as top, multigrid.py:305-306); coarsest as top = direct solve."""
    H = pr.build_hierarchy(dim=2, c=4, coarsest_level=0, finest_level=3, perm_seed=9, with_dicts=False)
    for mu1, mu2 in [(0, 0), (0, 3), (3, 0), (1, 2), (2, 1)]:
        H.mu1, H.mu2 = mu1, mu2
        eng = MGEngine.from_hierarchy(H)
        mg = rs.from_hierarchy(H)
        for top in (3, 2, 1):
            f = H.b_dict[top]
            vo = mg.vcycle(top, np.zeros_like(f), f)
            vo = mg.vcycle(top, vo, f)
            vg = eng.vcycle(top, np.zeros_like(f), f, ncycles=2)
            assert relmax(vg, vo) <= RTOL_SOLUTION, (mu1, mu2, top)
        x = np.arange(H.n(3), dtype=np.float64)
        assert np.array_equal(eng.smooth(3, x, x, 0), x)
        f0 = H.b_dict[0]
        assert relmax(eng.vcycle(0, np.zeros_like(f0), f0), mg.vcycle(0, np.zeros_like(f0), f0)) <= 1e-12
        eng.close()


def test_device_pointer_path_and_resident_mode():
    import torch
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=4, with_dicts=False)
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[4][:, 0]
    v_host = eng.vcycle(4, np.zeros_like(f), f, ncycles=3)
    ft = torch.tensor(f, device="cuda:0"); vt = torch.zeros_like(ft)
    out = eng.vcycle(4, vt, ft, ncycles=3)
    assert np.array_equal(out.cpu().numpy(), v_host) and float(vt.abs().max()) == 0.0      # inputs untouched
    eng.level_buffer(4, "f").copy_(ft); eng.level_buffer(4, "v").zero_()
    torch.cuda.synchronize()
    hist = eng.vcycle_resident(4, 3, history=True)
    assert np.array_equal(eng.level_buffer(4, "v").cpu().numpy(), v_host)
    A = H.A_sp_dict[4][0]
    assert abs(hist[-1] - np.linalg.norm(f - A.dot(v_host))) <= 1e-12 * hist[0]
    assert eng.launch_count() > 0
    eng.close()


def test_device_tensor_path_matches_host_path_at_large_n():
    """The engine stream is non-blocking: the torch-side clone / contiguous copies of device tensors must be ordered before the
    engine's first kernel (the event is recorded AFTER the marshalling).  A race shows up only when the copies take long enough,
    hence 4.2 M unknowns (config 2's finest level) and fresh, non-contiguous inputs every call."""
    import torch
    H = pr.build_hierarchy(dim=2, c=32, coarsest_level=0, finest_level=6, with_dicts=False)
    lf = 6
    eng = MGEngine.from_hierarchy(H)
    n = H.n(lf)
    rng = np.random.default_rng(8)
    for trial in range(3):
        x, f = rng.standard_normal(n), rng.standard_normal(n)
        xt = torch.from_numpy(np.stack([x, x], axis=1)).cuda()[:, 0]          # a strided view: .contiguous() has to copy
        ft = torch.from_numpy(np.stack([f, f], axis=1)).cuda()[:, 1]
        assert np.array_equal(eng.smooth(lf, xt, ft, 2).cpu().numpy(), eng.smooth(lf, x, f, 2))
        assert np.array_equal(eng.residual(lf, xt, ft).cpu().numpy(), eng.residual(lf, x, f))
        assert np.array_equal(eng.vcycle(lf, xt, ft).cpu().numpy(), eng.vcycle(lf, x, f))
        e = rng.standard_normal(H.n(lf - 1))
        et = torch.from_numpy(np.stack([e, e], axis=1)).cuda()[:, 0]
        assert np.array_equal(eng.prolong_add(lf, et, xt).cpu().numpy(), eng.prolong_add(lf, e, x))
    eng.close()


def test_errors_are_reported_not_swallowed():
    H = pr.build_hierarchy(dim=2, c=4, coarsest_level=0, finest_level=1, with_dicts=False)
    eng = MGEngine(0)
    eng.set_level(0, H.A_sp_dict[0][0]); eng.set_level(1, H.A_sp_dict[1][0])
    with pytest.raises(L.MGBError):
        eng.finalize()                                    # transfer missing
    with pytest.raises(L.MGBError):
        eng.set_transfer(0, H.P[0], inj=None)             # injection list missing
    with pytest.raises(L.MGBError):
        eng._ck(eng._lib.mgb_vcycle(eng._h, 1, None, None, 0, 1, None))   # not finalized
    eng.close()


def test_full_size_properties_config2():
    """BASELINE config 2 (2049^2, 7 levels, V(2,2)): size-independent properties.
    linearity in f (exact for a power-of-two scale), determinism, graph == eager, residual reduction."""
    H = pr.build_hierarchy(dim=2, c=32, coarsest_level=0, finest_level=6, with_dicts=False)
    assert H.n(6) == 2049 ** 2 and H.A_sp_dict[6][0].nnz == 29372417
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[6][:, 0]
    v1, h1 = eng.vcycle(6, np.zeros_like(f), f, ncycles=3, history=True)
    v2, h2 = eng.vcycle(6, np.zeros_like(f), 4.0 * f, ncycles=3, history=True)
    assert np.array_equal(4.0 * v1, v2) and np.array_equal(4.0 * h1, h2)
    v3 = eng.vcycle(6, np.zeros_like(f), f, ncycles=3)
    assert np.array_equal(v1, v3)
    eng.set_option("use_graph", 0)
    v4 = eng.vcycle(6, np.zeros_like(f), f, ncycles=3)
    assert np.array_equal(v1, v4)
    assert h1[2] < h1[1] < h1[0]
    A = H.A_sp_dict[6][0]
    assert abs(np.linalg.norm(f - A.dot(v1)) - h1[2]) <= 1e-12 * h1[2]
    # the C oracle on the same input (a few seconds)
    cm = co.from_hierarchy(H)
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=3, history=True)
    r = float(np.abs(h1 - ho).max() / ho.max()); s = relmax(v1, vo)
    _report("config2_full", resnorm_rel=r, solution_rel=s, hist=[float(x) for x in h1])
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    eng.close()


@pytest.mark.parametrize("dim,c,lf,glevel", [(2, 8, 4, 1), (3, 2, 4, 0), (3, 4, 3, 2)])
def test_device_side_generation_matches_host_assembler(dim, c, lf, glevel):
    """mgb_synth_poisson_* (device-born CSR + device-built R_omega) vs problems.py + the host setup path:
    matrices, transfers, injection lists and R_omega bit-identical, hence identical V-cycles."""
    from multigrid_dolfinx_b200 import dist as ds
    src = ds.StructuredSource(dim, c, 0, lf)
    mg = ds.DistMG(src, device=0, gather_level=glevel, device_gen=True)          # world = 1, no process group
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, with_dicts=False)
    ref = MGEngine.from_hierarchy(H)
    for l in range(glevel + 1, lf + 1):
        A = H.A_sp_dict[l][0]
        assert np.array_equal(mg.eng.artifact(l, L.ART_A_INDPTR), A.indptr)
        assert np.array_equal(mg.eng.artifact(l, L.ART_A_INDICES), A.indices)
        assert np.array_equal(mg.eng.artifact(l, L.ART_A_VALUES), A.data)
        P = H.P[l - 1]
        assert np.array_equal(mg.eng.artifact(l, L.ART_P_INDPTR), P.indptr)
        assert np.array_equal(mg.eng.artifact(l, L.ART_P_INDICES), P.indices)
        assert np.array_equal(mg.eng.artifact(l, L.ART_P_VALUES), P.data)
        assert np.array_equal(mg.eng.artifact(l, L.ART_INJECTION), H.inj[l - 1])
        for kind in (L.ART_RJ_INDPTR, L.ART_RJ_INDICES, L.ART_RJ_VALUES, L.ART_DINV):
            assert np.array_equal(mg.eng.artifact(l, kind), ref.artifact(l, kind))
    f = src.rhs_rows(lf, 0, src.n(lf))
    mg.load_rhs()
    h1 = mg.cycles(3, history=True)
    v0, h0 = ref.vcycle(lf, np.zeros_like(f), f, ncycles=3, history=True)
    assert np.array_equal(mg.local_solution(), v0) and np.array_equal(h1, h0)
    mg.close(); ref.close()


@pytest.mark.parametrize("r_mode,opts", [("injection", {}), ("transpose", {}), ("transpose", {"kernel_family": 2}), ("full_weighting", {"stream_cfg": 0})])
def test_p2_hierarchy_vs_oracle(r_mode, opts):
    """BASELINE config 4 shape (3-D P2, rows of 10..65 entries, FE interpolation with negative weights), small instance."""
    H = pr.build_hierarchy_p2(c=2, coarsest_level=0, finest_level=2, perm_seed=None)
    eng = MGEngine.from_hierarchy(H, r_mode=r_mode, options=opts)
    cm = co.from_hierarchy(H, r_mode=r_mode)
    f = H.b_dict[2][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=4, history=True)
    vg, hg = eng.vcycle(2, np.zeros_like(f), f, ncycles=4, history=True)
    r = float(np.abs(hg - ho).max() / ho.max()); s = relmax(vg, vo)
    _report("p2_oracle", r_mode=r_mode, opts=str(opts), resnorm_rel=r, solution_rel=s)
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    if "kernel_family" not in opts:
        A = H.A_sp_dict[2][0]
        x = np.random.default_rng(3).standard_normal(A.shape[0])
        assert np.array_equal(eng.spmv(2, x), A.dot(x))                      # 65-entry rows, still bit-exact
    eng.close()


def _ragged(n, m, rng, max_len, empty_frac=0.15, diag=False):
    """random CSR with ragged rows (some empty unless ``diag``), unsorted columns, optional dominant diagonal"""
    import scipy.sparse as sp
    lens = rng.integers(0, max_len + 1, size=n)
    lens[rng.random(n) < empty_frac] = 0
    rows_ix, rows_ax = [], []
    for i in range(n):
        k = int(min(lens[i], m - 1 if diag else m))
        cols = rng.choice(m, size=k, replace=False) if k else np.zeros(0, dtype=np.int64)
        vals = rng.standard_normal(k)
        if diag:
            cols = cols[cols != i]; vals = vals[:len(cols)]
            pos = int(rng.integers(0, len(cols) + 1))
            cols = np.insert(cols, pos, i); vals = np.insert(vals, pos, 3.0 + np.abs(vals).sum())
        rows_ix.append(cols); rows_ax.append(vals)
    ip = np.zeros(n + 1, dtype=np.int64); np.cumsum([len(c) for c in rows_ix], out=ip[1:])
    ix = np.concatenate(rows_ix).astype(np.int32) if ip[-1] else np.zeros(0, dtype=np.int32)
    ax = np.concatenate(rows_ax) if ip[-1] else np.zeros(0)
    return sp.csr_matrix((ax, ix, ip), shape=(n, m))


@pytest.mark.parametrize("nf,nc,max_len,seed", [(1237, 311, 9, 0), (5, 3, 3, 1), (4099, 1025, 40, 2), (2050, 700, 300, 3), (1, 1, 1, 4),
                                                 (3000, 400, 1500, 5), (6000, 64, 5000, 6)])
@pytest.mark.parametrize("opts", [{}, {"stream_cfg": 0}, {"stream_cfg": 1}])
def test_ragged_operators_bit_exact(nf, nc, max_len, seed, opts):
    """Irregular inputs: empty rows, rows longer than a stream tile's row share, unsorted columns, sizes that are not
    multiples of the tile / alignment granules, explicit (non-transpose) restriction.  Everything must equal scipy bit for bit."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    Af = _ragged(nf, nf, rng, min(max_len, nf), diag=True)
    Ac = _ragged(nc, nc, rng, min(max_len, nc), diag=True)
    Af.sort_indices(); Ac.sort_indices()         # level matrices come from PETSc with sorted columns; transfers may be in any order
    P = _ragged(nf, nc, rng, min(max_len, nc))
    R = _ragged(nc, nf, rng, min(max_len, nf))
    eng = MGEngine(0)
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.set_level(0, Ac); eng.set_level(1, Af)
    eng.set_transfer(0, P, r_mode="explicit", R=R)
    eng.set_params(0.7, 2, 1, "jacobi")
    eng.finalize()
    x, f = rng.standard_normal(nf), rng.standard_normal(nf)
    e = rng.standard_normal(nc)
    assert np.array_equal(eng.spmv(1, x), Af.dot(x))
    assert np.array_equal(eng.residual(1, x, f), f - Af.dot(x))
    assert np.array_equal(eng.restrict(1, x), R.dot(x))
    assert np.array_equal(eng.prolong_add(1, e, x), x + P.dot(e))
    RO, dinv = rs.jacobi_matrices(Af)
    assert np.array_equal(eng.smooth(1, x, f, 3), rs.jacobi_relaxation(RO, dinv, x, f, 3, 0.7))
    mg = rs.RestatedMG({0: Ac, 1: Af}, {0: P}, R={0: R}, r_mode="explicit", omega=0.7, mu1=2, mu2=1)
    vo = mg.vcycle(1, x[:, None], f[:, None])
    vg = eng.vcycle(1, x, f)
    assert relmax(vg, vo) <= 1e-10
    eng.close()


def test_shard_size_limit_is_refused_not_truncated():
    """An operator beyond the int32 row-pointer range of one device must be rejected with MGB_ERR_UNSUPPORTED -- on the sizes
    alone, before any entry is read: the arrays handed over here are real but tiny, only indptr[-1] / nnz claim 2^31 + 5 entries."""
    import ctypes as C
    lib = L.load()
    h = C.c_void_p()
    assert lib.mgb_create(C.byref(h), 0) == 0
    ip = np.array([0, 2 ** 31 + 5], dtype=np.int64)
    ix = np.zeros(4, dtype=np.int32); ax = np.ones(4)
    rc = lib.mgb_set_level(h, 0, 1, 2 ** 31 + 5, ip.ctypes.data, 8, ix.ctypes.data, ax.ctypes.data)
    assert rc == L.ERR_UNSUPPORTED
    assert b"int32 row-pointer range" in lib.mgb_last_error(h)
    rc = lib.mgb_set_level(h, 0, 2 ** 31 + 5, 4, ip.ctypes.data, 8, ix.ctypes.data, ax.ctypes.data)      # too many rows
    assert rc == L.ERR_UNSUPPORTED
    rc = lib.mgb_set_level(h, 0, 1, 4, ip.ctypes.data, 8, None, None)                                     # null arrays with nnz > 0
    assert rc == L.ERR_INVALID
    lib.mgb_destroy(h)


@pytest.mark.parametrize("name", ["oracle_3d_p1_inj", "oracle_3d_p1_perm_transpose", "oracle_2d_gs", "oracle_2d_gs_color_perm", "oracle_3d_p2_transpose"])
def test_against_committed_oracle_fixtures(name):
    """Parts of the path without reference text (3-D, P2, Gauss-Seidel): GPU vs the frozen oracle outputs in tests/golden/."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_oracle_fixtures", os.path.join(ROOT, "tests", "golden", "gen_oracle_fixtures.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    kw, r_mode, smoother, K = m.CASES[name]
    H = m.build(kw)
    lf = H.finest_level
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    eng = MGEngine.from_hierarchy(H, r_mode=r_mode, smoother=smoother)
    f = H.b_dict[lf][:, 0]
    v, hist = eng.vcycle(lf, np.zeros_like(f), f, ncycles=K, history=True)
    assert np.abs(hist - d["resnorm"]).max() <= RTOL_RESNORM * d["resnorm"].max()
    assert relmax(v, d["v"]) <= RTOL_SOLUTION
    if smoother.startswith("gs"):
        assert np.array_equal(eng.artifact(lf, L.ART_LEVEL_OF_ROW), d["level_of_row"])
        assert np.array_equal(eng.artifact(lf, L.ART_LEVEL_OFFSETS), d["level_offsets"])
        assert np.array_equal(eng.artifact(lf, L.ART_COLOUR_OF_ROW), d["colour_of_row"])
        assert np.array_equal(eng.artifact(lf, L.ART_COLOUR_OFFSETS), d["colour_offsets"])
    eng.close()


def test_full_size_properties_config3():
    """BASELINE config 3 (129^3 tetrahedral P1, 5 levels): oracle parity at full size + linearity + determinism."""
    H = pr.build_hierarchy(dim=3, c=8, coarsest_level=0, finest_level=4, with_dicts=False)
    assert H.n(4) == 129 ** 3 and H.A_sp_dict[4][0].nnz == 31802497
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[4][:, 0]
    v1, h1 = eng.vcycle(4, np.zeros_like(f), f, ncycles=3, history=True)
    v2, h2 = eng.vcycle(4, np.zeros_like(f), 0.5 * f, ncycles=3, history=True)
    assert np.array_equal(0.5 * v1, v2) and np.array_equal(0.5 * h1, h2)
    assert np.array_equal(eng.vcycle(4, np.zeros_like(f), f, ncycles=3), v1)
    cm = co.from_hierarchy(H)
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=3, history=True)
    r = float(np.abs(h1 - ho).max() / ho.max()); s = relmax(v1, vo)
    _report("config3_full", resnorm_rel=r, solution_rel=s)
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    eng.close()


@pytest.mark.parametrize("dim,c,lf,seed,rcm", [(2, 8, 4, 3, False), (3, 2, 3, 5, False), (2, 8, 4, 3, True)])
def test_caller_numbering_is_bit_identical_and_unlocks_the_row_codings(dim, c, lf, seed, rcm):
    """dolfinx numbers DOFs its own way; the reference records where every DOF sits (coordinate dicts, Multigrid_prototype.py:68-74).
    With that lattice numbering handed to the engine (mgb_set_numbering) it works in lexicographic order internally: the level
    operators get the one-byte-per-row codings and the hot-row kernels, while every iterate stays BIT-IDENTICAL to the run in the
    caller's numbering (rows are moved, the entry order inside a row is kept)."""
    import torch
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False, rcm=rcm)
    n, nc = H.n(lf), H.n(lf - 1)
    plain = MGEngine.from_hierarchy(H)
    eng = MGEngine.from_hierarchy(H, reorder=True)
    dp, de = plain.describe(), eng.describe()
    fine = lambda d, tag: [ln for ln in d.splitlines() if ln.strip().startswith(tag) and f"rows={n} " in ln][0]
    assert "hotrow(" not in fine(dp, "A ")
    if rcm:     # a banded numbering stores every row's entries in the same lattice order: whole rows repeat once renumbered
        assert "hotrow(" in fine(de, "A ") and "hotrow(" in fine(de, "RJ ") and "coded mode=4" in fine(de, "P ")
    else:       # a random numbering stores every row's entries in its own order (kept: it IS the summation order): no row patterns,
        assert "coded mode=1" in fine(de, "A ")     # but the offsets are few again -> one byte per entry and coalesced gathers
    rng = np.random.default_rng(2)
    x, f, e = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(nc)
    b = H.b_dict[lf][:, 0]
    for fn in (lambda E: E.vcycle(lf, np.zeros_like(b), b, ncycles=3), lambda E: E.spmv(lf, x), lambda E: E.residual(lf, x, f),
               lambda E: E.smooth(lf, x, f, 3), lambda E: E.prolong_add(lf, e, x), lambda E: E.restrict(lf, x),
               lambda E: E.vcycle_debug(lf, np.zeros_like(b), b)[1]):
        assert np.array_equal(fn(plain), fn(eng))
    A = H.A_sp_dict[lf][0]
    assert np.array_equal(eng.spmv(lf, x), A.dot(x))
    xt, ft = torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()                 # device pointers take the same route
    assert np.array_equal(eng.smooth(lf, xt, ft, 2).cpu().numpy(), plain.smooth(lf, x, f, 2))
    assert np.array_equal(eng.vcycle(lf, torch.zeros_like(xt), torch.from_numpy(b).cuda(), ncycles=2).cpu().numpy(), plain.vcycle(lf, np.zeros_like(b), b, ncycles=2))
    _, h1 = eng.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    _, h0 = plain.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
    assert np.abs(h1 - h0).max() <= RTOL_RESNORM * h0.max()                           # (the norm's summation order follows the numbering)
    plain.close(); eng.close()


def test_full_size_config4_p2_vs_oracle():
    """BASELINE config 4 at FULL size: 3-D P2 on 64^3 cells, DOF grid 129^3 = 2,146,689, 60,859,905 stored entries, rows of
    10..65 entries, 5 levels, V(2,2) Jacobi, injection -- three cycles against the C oracle on the same hierarchy."""
    H = pr.build_hierarchy_p2(c=4, coarsest_level=0, finest_level=4)
    assert H.n(4) == 129 ** 3 and H.A_sp_dict[4][0].nnz == 60859905
    eng = MGEngine.from_hierarchy(H)
    f = H.b_dict[4][:, 0]
    v1, h1 = eng.vcycle(4, np.zeros_like(f), f, ncycles=3, history=True)
    cm = co.from_hierarchy(H)
    vo, ho = cm.vcycle(np.zeros_like(f), f, ncycles=3, history=True)
    r = float(np.abs(h1 - ho).max() / ho.max()); s = relmax(v1, vo)
    _report("config4_full", resnorm_rel=r, solution_rel=s, hist=[float(x) for x in h1])
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    A = H.A_sp_dict[4][0]
    x = np.random.default_rng(4).standard_normal(A.shape[0])
    assert np.array_equal(eng.spmv(4, x), A.dot(x))                          # 65-entry rows summed in stored order: bit-exact
    assert np.array_equal(eng.vcycle(4, np.zeros_like(f), f, ncycles=3), v1)  # deterministic
    eng.close()


@pytest.mark.parametrize("smoother", ["gs", "gs_color"])
def test_gauss_seidel_full_size_config2(smoother):
    """BASELINE config 2 names the Gauss-Seidel smoothers: 2049^2 (4.2 M DOFs, 7 levels), one V(2,2) cycle with the level-scheduled
    natural-order sweep (4097 dependency levels on the finest grid) and with the multicolour sweep, against the C oracle's
    sequential sweeps (1e-12 on the residual norm, 1e-10 on the iterate), plus one bit-exact fine-level sweep."""
    H = pr.build_hierarchy(dim=2, c=32, coarsest_level=0, finest_level=6, with_dicts=False)
    lf = 6
    assert H.n(lf) == 2049 ** 2
    eng = MGEngine.from_hierarchy(H, smoother=smoother)
    cm = co.from_hierarchy(H, smoother=smoother)
    b = H.b_dict[lf][:, 0]
    vo, ho = cm.vcycle(np.zeros_like(b), b, ncycles=1, history=True)
    vg, hg = eng.vcycle(lf, np.zeros_like(b), b, ncycles=1, history=True)
    r = float(np.abs(hg - ho).max() / ho.max()); s = relmax(vg, vo)
    _report("config2_gs_full", smoother=smoother, resnorm_rel=r, solution_rel=s)
    assert r <= RTOL_RESNORM and s <= RTOL_SOLUTION
    A = H.A_sp_dict[lf][0]
    rng = np.random.default_rng(1)
    x, f = rng.standard_normal(H.n(lf)), rng.standard_normal(H.n(lf))
    order = co.greedy_colouring(A)[1] if smoother == "gs_color" else None
    assert np.array_equal(eng.smooth(lf, x, f, 1), co.gs_forward(A, x.copy(), f, order))
    eng.close()


def test_full_size_properties_config5():
    """BASELINE config 5 (513^3, 135 M DOFs, 2.0e9 stored entries, generated on the device): the CPU oracle cannot hold it,
    so size-independent properties: V-cycles are linear in f (exact for a power-of-two scale), deterministic, identical with
    and without graph replay, the residual history decreases and equals an independently computed ||f - A v||."""
    import torch
    from multigrid_dolfinx_b200 import dist as ds
    if torch.cuda.mem_get_info(0)[0] < 80e9:
        pytest.skip("needs ~60 GB of free device memory")
    src = ds.StructuredSource(3, 8, 0, 6)
    mg = ds.DistMG(src, device=0, device_gen=True)
    eng, lf = mg.eng, 6
    assert eng.n[lf] == 513 ** 3
    mg.load_rhs()
    h1 = mg.cycles(3, history=True)
    v1 = eng.level_buffer(lf, "v").clone()
    eng.level_buffer(lf, "f").mul_(4.0); eng.level_buffer(lf, "v").zero_(); torch.cuda.synchronize()
    h2 = mg.cycles(3, history=True)
    assert torch.equal(eng.level_buffer(lf, "v"), 4.0 * v1) and np.array_equal(h2, 4.0 * h1)
    eng.set_option("use_graph", 0)
    eng.level_buffer(lf, "f").mul_(0.25); eng.level_buffer(lf, "v").zero_(); torch.cuda.synchronize()
    h3 = mg.cycles(3, history=True)
    assert torch.equal(eng.level_buffer(lf, "v"), v1) and np.array_equal(h3, h1)
    assert h1[2] < h1[1] < h1[0]
    r = eng.residual(lf, eng.level_buffer(lf, "v"), eng.level_buffer(lf, "f"))          # per-operator entry point, tile kernel path
    assert abs(float(torch.linalg.vector_norm(r)) - h1[2]) <= 1e-12 * h1[2]
    _report("config5_full", hist=[float(x) for x in h1])
    mg.close()


# ---- dictionary-coded operators (option "compress", mgb_code.cuh / k_rowstream) ---------------------------------------
@pytest.mark.parametrize("dim,c,lf,seed,r_mode", [(2, 8, 4, None, "injection"), (2, 8, 4, 1, "injection"), (3, 2, 4, None, "injection"),
                                                   (3, 2, 3, 5, "transpose"), (2, 5, 3, None, "full_weighting"), (2, 32, 2, None, "transpose")])
def test_coded_operators_bit_identical_to_uncoded(dim, c, lf, seed, r_mode):
    """The one-byte-per-entry coding is lossless: every operator and the whole cycle give the same BITS with compress = 0
    (CSR stream kernel), with the register-staged tile kernel, and with every row-stream kernel shape."""
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    n, nc = H.n(lf), H.n(lf - 1)
    rng = np.random.default_rng(17)
    x, f, e = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(nc)
    A = H.A_sp_dict[lf][0]
    outs = []
    # (the default is compress = 3, hot-row kernels (stage_x = 3), fused thread-per-coarse-row residual, g reused across cycles)
    for opts in [{"compress": 0}, {"stream_cfg": 0}, {"compress": 1}, {"compress": 2}, {"code_cfg": 3}, {"compress": 1, "code_cfg": 3},
                 {"compress": 3}, {"anch_cfg": 2}, {"stage_x": 0}, {"stage_x": 1}, {"hot_cfg": 2}, {"hot_cfg": 3}, {"hot_cfg": 4},
                 {"hot_inj": 0}, {"hot_pf": 0}, {"reuse_g": 0}, {"tail_rows": 300000}, {"tail_rows": 300},
                 {"anch_tiles": -3}, {"anch_tiles": -16, "anch_cfg": 2}, {"tail_rows": 300000, "tail_cluster": 1}, {"tail_rows": 300, "tail_cluster": 1},
                 {"compress": 2, "stage_x": 0, "hot_inj": 0, "reuse_g": 0, "tail_rows": 0}]:
        eng = MGEngine.from_hierarchy(H, r_mode=r_mode, options=opts)
        desc = eng.describe()
        coded = "coded" in desc
        level = opts.get("compress", 3)
        assert coded == (level >= 1 and opts.get("stream_cfg", 3) != 0), desc
        if coded and seed is None:              # lexicographic numbering: whole rows repeat -> row patterns (compress >= 2, the
            want = "coded mode=3" if level >= 2 else "coded mode=1"                        # default), else pair codes per entry
            assert "A   " in desc and all(want in ln for ln in desc.splitlines() if ln.strip().startswith(("A ", "RJ ")) and f"rows={n} " in ln), desc
            if level >= 2 and opts.get("stage_x", 3) == 3:
                assert all("hotrow(" in ln for ln in desc.splitlines() if ln.strip().startswith(("A ", "RJ ")) and f"rows={n} " in ln), desc
        if coded:                               # transfers: anchored row patterns on a lexicographic numbering (compress = 3), else value
            assert all("coded" in ln for ln in desc.splitlines() if ln.strip().startswith("P ")), desc      # codes (+ columns)
            wantp = "coded mode=4" if level >= 3 and seed is None else "coded mode=2"
            assert all(wantp in ln for ln in desc.splitlines() if ln.strip().startswith("P ") and f"rows={n} " in ln), desc
        b = H.b_dict[lf][:, 0]
        v, hist = eng.vcycle(lf, np.zeros_like(b), b, ncycles=3, history=True)
        outs.append((eng.spmv(lf, x), eng.residual(lf, x, f), eng.smooth(lf, x, f, 3), eng.prolong_add(lf, e, x), eng.restrict(lf, x), v, hist))
        if coded:
            eng.profile_begin(); eng.vcycle(lf, np.zeros_like(b), b); prof = eng.profile_end()
            jac = [r for r in prof if r["kind"] == "jacobi" and r["level"] == lf][0]
            assert jac["moved_bytes"] < 0.75 * jac["bytes"]
            assert eng.vcycle_bytes_moved(lf) < eng.vcycle_bytes(lf)
        eng.close()
    assert np.array_equal(outs[0][0], A.dot(x)) and np.array_equal(outs[0][1], f - A.dot(x))
    for o in outs[1:]:
        for a, b_ in zip(outs[0], o):
            assert np.array_equal(a, b_)


@pytest.mark.parametrize("nvals,expect", [(1, "coded mode=2"), (7, "coded mode=2"), (256, "coded mode=2"), (257, None)])
@pytest.mark.parametrize("banded", [False, True])
def test_coded_ragged_few_values(nvals, expect, banded):
    """Value dictionary on irregular operators: ragged rows (empty ones included), unsorted columns, values drawn from a
    small set (signed zeros, a denormal and a huge value among them).  <= 256 distinct values -> coded; 257 -> CSR kernel."""
    import scipy.sparse as sp
    rng = np.random.default_rng(nvals + 1000 * banded)
    table = np.concatenate([[0.0, -0.0, 1.0, -1.0, 2.0 ** -1070, 1e60, 0.25], rng.standard_normal(300)])[:nvals] if nvals >= 7 else np.array([0.25])
    table = np.unique(table.view(np.int64)).view(np.float64)
    if len(table) < nvals:
        table = np.concatenate([table, 10.0 + np.arange(nvals - len(table))])

    def draw(M):
        M = M.copy()
        idx = np.arange(M.nnz) % nvals                  # every table entry is used
        rng.shuffle(idx)
        M.data = table[idx]
        return M
    nf, nc = 3001, 777
    if banded:                                          # few column offsets as well -> pair codes
        offs = np.array([-40, -1, 0, 1, 40])
        rows = np.repeat(np.arange(nf), len(offs)); cols = rows + np.tile(offs, nf)
        keep = (cols >= 0) & (cols < nf)
        Af = sp.csr_matrix((np.ones(keep.sum()), (rows[keep], cols[keep])), shape=(nf, nf))
        Af.sort_indices()
    else:
        Af = _ragged(nf, nf, rng, 20, diag=True); Af.sort_indices()
    Af = draw(Af)
    rows_of = np.repeat(np.arange(nf), np.diff(Af.indptr))
    Af.data[Af.indices == rows_of] = 1.0 if nvals >= 7 else 0.25                   # a benign diagonal (1 / d stays finite)
    Ac = _ragged(nc, nc, rng, 12, diag=True); Ac.sort_indices()
    P, R = draw(_ragged(nf, nc, rng, 15)), draw(_ragged(nc, nf, rng, 22))
    eng = MGEngine(0)
    eng.set_level(0, Ac); eng.set_level(1, Af)
    eng.set_transfer(0, P, r_mode="explicit", R=R)
    eng.set_params(0.7, 2, 1, "jacobi")
    eng.finalize()
    desc = eng.describe()
    pl = [ln for ln in desc.splitlines() if ln.strip().startswith("P ")][0]
    assert (expect in pl) if expect else ("coded" not in pl), desc
    if banded and nvals <= 7:                           # few offsets and values, but rows do not repeat: pair codes per entry
        al = [ln for ln in desc.splitlines() if ln.strip().startswith("A ") and f"rows={nf} " in ln][0]
        assert ("coded mode=1" in al) or (nvals == 1 and "coded mode=3" in al), desc
    x, f, e = rng.standard_normal(nf), rng.standard_normal(nf), rng.standard_normal(nc)
    assert np.array_equal(eng.spmv(1, x), Af.dot(x))
    assert np.array_equal(eng.residual(1, x, f), f - Af.dot(x))
    assert np.array_equal(eng.restrict(1, x), R.dot(x))
    assert np.array_equal(eng.prolong_add(1, e, x), x + P.dot(e))
    RO, dinv = rs.jacobi_matrices(Af)
    assert np.array_equal(eng.smooth(1, x, f, 3), rs.jacobi_relaxation(RO, dinv, x, f, 3, 0.7))
    eng.close()


def test_row_pattern_coding_long_and_empty_rows():
    """Row-pattern codes (mode 3) on operators whose rows repeat but are neither short nor uniform: 19-entry rows (first
    8 gathers + two continuation chunks), truncated boundary rows, periodically EMPTY rows and a rectangular operator."""
    import scipy.sparse as sp
    nf, nc = 5003, 1201
    offs = np.arange(-90, 91, 10)                        # 19 offsets, 0 among them
    vals = {int(o): (4.0 if o == 0 else -1.0 / (1 + abs(o) // 10)) for o in offs}
    rows = np.repeat(np.arange(nf), len(offs)); cols = rows + np.tile(offs, nf)
    keep = (cols >= 0) & (cols < nf)
    data = np.array([vals[int(o)] for o in np.tile(offs, nf)])
    Af = sp.csr_matrix((data[keep], (rows[keep], cols[keep])), shape=(nf, nf)); Af.sort_indices()
    Ac = sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(nc, nc)).tocsr()
    # R (nc x nf): row i couples fine columns 4i + {0, 1, 2, 5, 9, 14, 20, 27, 35, 44, 54}; every 5th row is empty, every 7th short
    roffs = np.array([0, 1, 2, 5, 9, 14, 20, 27, 35, 44, 54])
    r_i, r_j, r_v = [], [], []
    for i in range(nc):
        if i % 5 == 3:
            continue
        use = roffs[:3] if i % 7 == 2 else roffs
        for q, o in enumerate(use):
            j = 4 * i + int(o)
            if j < nf:
                r_i.append(i); r_j.append(j); r_v.append(0.5 ** (q % 4))
    R = sp.csr_matrix((r_v, (r_i, r_j)), shape=(nc, nf)); R.sort_indices()
    rng = np.random.default_rng(5)
    P = _ragged(nf, nc, rng, 9)
    eng = MGEngine(0)
    eng.set_level(0, Ac); eng.set_level(1, Af)
    eng.set_transfer(0, P, r_mode="explicit", R=R)
    eng.set_params(0.7, 2, 1, "jacobi")
    eng.finalize()
    desc = eng.describe()
    for tag, rows_ in (("A ", nf), ("RJ ", nf)):
        ln = [l for l in desc.splitlines() if l.strip().startswith(tag) and f"rows={rows_} " in l][0]
        assert "coded mode=3" in ln, desc
    # R's rows are not translates of each other in (col - row) terms (columns advance by 4 per row), but they are when measured
    # from each row's first column: anchored row patterns
    assert "coded mode=4" in [l for l in desc.splitlines() if l.strip().startswith("R ")][0], desc
    x, f = rng.standard_normal(nf), rng.standard_normal(nf)
    assert np.array_equal(eng.spmv(1, x), Af.dot(x))
    assert np.array_equal(eng.residual(1, x, f), f - Af.dot(x))
    assert np.array_equal(eng.restrict(1, x), R.dot(x))
    RO, dinv = rs.jacobi_matrices(Af)
    assert np.array_equal(eng.smooth(1, x, f, 3), rs.jacobi_relaxation(RO, dinv, x, f, 3, 0.7))
    eng.close()
    # a square operator with periodically empty and short rows, used as the level matrix of a one-level "hierarchy" is not
    # possible (Jacobi needs the diagonal), so the same rows go in as the mass matrix of the FMG norm: M r through mode 3
    m_i, m_j, m_v = [], [], []
    for i in range(nf):
        if i % 5 == 3:
            continue
        for q, o in enumerate((-7, 0, 3) if i % 7 == 2 else (-50, -7, -1, 0, 1, 3, 11, 23, 40, 77)):
            if 0 <= i + o < nf:
                m_i.append(i); m_j.append(i + o); m_v.append(40.0 if o == 0 else 1.0 + 0.25 * q)     # dominant diagonal: r^T M r > 0
    M = sp.csr_matrix((m_v, (m_i, m_j)), shape=(nf, nf)); M.sort_indices()
    eng = MGEngine(0)
    eng.set_level(0, Ac); eng.set_level(1, Af)
    eng.set_transfer(0, P, r_mode="explicit", R=R)
    eng.set_params(0.7, 1, 1, "jacobi")
    eng.finalize()
    eng.set_mass_matrix(1, M)
    eng.set_rhs(0, np.ones(nc)); eng.set_rhs(1, f)
    v, hist = eng.fmg(mu0=1, tol=0.0, max_cycles=1)
    r = f - Af.dot(v)
    assert abs(hist[0] - np.sqrt(abs(r @ M.dot(r)))) <= 1e-12 * max(hist[0], 1e-300)
    eng.close()


@pytest.mark.parametrize("dim,c,lf,seed", [(2, 8, 3, None), (3, 2, 3, None), (2, 8, 3, 5)])
def test_device_built_coding_against_host_definition(dim, c, lf, seed):
    """The coding artefact: what mgb_finalize built on the device vs mgb_host_code_operator (the definition, itself pinned to the
    numpy restatement by tests/test_coding.py): mode, counts, every code, every table entry (value BITS and offset), the pattern
    heads and -- for anchored patterns -- every anchor."""
    from multigrid_dolfinx_b200.engine import host_code_operator
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    eng = MGEngine.from_hierarchy(H)
    A = H.A_sp_dict[lf][0].tocsr()
    RO, _ = rs.jacobi_matrices(A)
    for op, M in (("A", A), ("RJ", RO.tocsr()), ("P", H.P[lf - 1].tocsr())):
        dev, host = eng.code_artifact(lf, op), host_code_operator(M)
        assert dev["mode"] == host["mode"] and dev["ndict"] == host["ndict"], (op, dev["mode"], host["mode"], dev["ndict"], host["ndict"])
        if seed is None:
            assert dev["mode"] == (4 if op == "P" else 3), (op, dev["mode"])
        same_codes = bool(np.array_equal(dev["codes"], host["codes"]))
        k = min(len(dev["table"]), len(host["table"]))
        same_table = bool(len(dev["table"]) == len(host["table"]) and np.array_equal(dev["table"]["val"][:k].view(np.uint64), host["table"]["val"][:k].view(np.uint64))
                          and np.array_equal(dev["table"]["delta"][:k], host["table"]["delta"][:k]))
        same_head = bool(dev["mode"] not in (3, 4) or np.array_equal(dev["head"], host["head"]))
        same_anchor = bool(dev["mode"] != 4 or np.array_equal(dev["anchor"], host["anchor"]))
        _report("coding_artefact", dim=dim, seed=seed, op=op, mode=dev["mode"], ndict=dev["ndict"], codes_equal=same_codes,
                table_equal=same_table, head_equal=same_head, anchor_equal=same_anchor)
        assert same_codes and same_table and same_head and same_anchor, (op, same_codes, same_table, same_head, same_anchor)
    eng.close()
