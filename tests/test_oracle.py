"""The CPU oracle is pinned here: (1) the numpy restatement equals the reference bit for bit on the golden
vectors the reference itself produced (tests/golden/gen_golden.py), (2) where /root/reference is
present it is re-run live, (3) the C restatement equals scipy per operator bit for bit and the reference's
V-cycle to ~1e-15 (it replaces SuperLU by a dense LU on the coarsest level)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_SMALL, ROOT, load_golden
from multigrid_dolfinx_b200 import problems as pr
from oracle import c_oracle as co
from oracle import reference_import as ri
from oracle import restated as rs


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_restated_vcycle_equals_golden_bitwise(name):
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    mg = rs.from_hierarchy(H)
    lf = H.finest_level
    v = np.zeros((H.n(lf), 1)); f = H.b_dict[lf]
    A = H.A_sp_dict[lf][0]
    gold_v = d["vcycle_v"]
    for k in range(K):
        v = mg.vcycle(lf, v, f)
        assert np.linalg.norm(f - A.dot(v)) == d["vcycle_resnorm"][k]
        if gold_v.shape[0] == K:
            assert np.array_equal(v[:, 0], gold_v[k])
    assert np.array_equal(v[:, 0], gold_v[-1])


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_restated_operators_equal_golden_bitwise(name):
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    lf = H.finest_level
    A = H.A_sp_dict[lf][0]
    RO, dinv = rs.jacobi_matrices(A)
    assert np.array_equal(RO.indptr, d["rj_indptr"]) and np.array_equal(RO.indices, d["rj_indices"])
    assert np.array_equal(RO.data, d["rj_data"]) and np.array_equal(dinv, d["rj_dinv"])
    ip, ix, ax, di = rs.rj_pattern_from_values(A)       # independent statement: same set of entries per row
    assert np.array_equal(ip, RO.indptr)
    for i in range(0, A.shape[0], 37):
        a, b = ip[i], ip[i + 1]
        assert np.array_equal(ix[a:b][::-1], RO.indices[a:b]) and np.array_equal(ax[a:b][::-1], RO.data[a:b])
    x, g, e = d["in_x"][:, None], d["in_g"][:, None], d["in_e"][:, None]
    om = float(d["omega"])
    assert np.array_equal(rs.jacobi_relaxation(RO, dinv[:, None], x, g, 1, om)[:, 0], d["jac_1"])
    assert np.array_equal(rs.jacobi_relaxation(RO, dinv[:, None], x, g, 5, om)[:, 0], d["jac_5"])
    assert np.array_equal(H.P[lf - 1].dot(e)[:, 0], d["interp"])
    assert np.array_equal(x[H.inj[lf - 1], 0], d["inj"])
    fw = pr.full_weighting(H.P[lf - 1], 2).dot(x)[:, 0]
    assert np.abs(fw - d["fw"]).max() <= 4e-16 * np.abs(d["fw"]).max()      # Restriction2D == 1/4 P^T (sum order differs)


@pytest.mark.skipif(not ri.available(), reason="/root/reference only exists in the development container")
@pytest.mark.parametrize("seed,mu", [(None, 2), (13, 3), (14, 50)])
def test_restated_equals_live_reference(seed, mu):
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=2, perm_seed=seed, mu1=mu, mu2=mu)
    outs, dbg = ri.run_reference_vcycles(H, 3, test_tuple=True)
    mg = rs.from_hierarchy(H)
    v = np.zeros((H.n(2), 1))
    for k in range(3):
        res = mg.vcycle(2, v, H.b_dict[2], debug=True)
        v = res[0]
        assert np.array_equal(v, outs[k])
    for a, b in zip(res[1:], dbg):
        assert np.array_equal(a, b)


@pytest.mark.skipif(not ri.available(), reason="/root/reference only exists in the development container")
def test_reference_limits_documented_in_survey():
    """the reference's rounded-coordinate keys fail for non-dyadic h."""
    H = pr.build_hierarchy(dim=2, c=3, coarsest_level=0, finest_level=1)
    with pytest.raises(KeyError):
        ri.run_reference_vcycles(H, 1)


@pytest.mark.parametrize("dim,m,seed", [(2, 32, None), (2, 32, 5), (3, 8, None), (3, 8, 6)])
def test_c_oracle_operators_bitwise_equal_scipy(dim, m, seed):
    n = (m + 1) ** dim
    A = pr.stencil_p1(m, dim, pr.make_permutation(n, seed))
    rng = np.random.default_rng(0)
    x, f = rng.standard_normal(n), rng.standard_normal(n)
    assert np.array_equal(A.dot(x), co.csr_matvec(A, x))
    assert np.array_equal(f - A.dot(x), co.residual(A, f, x))
    RO, dinv = rs.jacobi_matrices(A)
    assert np.array_equal(rs.jacobi_relaxation(RO, dinv, x, f, 4, 2 / 3), co.jacobi(RO, dinv, f, x, 2 / 3, 4))
    assert np.array_equal(rs.jacobi_relaxation_aform(A, dinv, x, f, 3, 0.8), co.jacobi(A, dinv, f, x, 0.8, 3, aform=True))


@pytest.mark.parametrize("dim,m,seed", [(2, 12, None), (2, 12, 5), (3, 5, 2)])
def test_gauss_seidel_definitions_agree(dim, m, seed):
    """C and pure-Python statements of natural-order GS, level sets and greedy colouring coincide; a
    level-major / colour-major sweep equals the sequential sweep in that order."""
    n = (m + 1) ** dim
    A = pr.stencil_p1(m, dim, pr.make_permutation(n, seed))
    rng = np.random.default_rng(1)
    x, f = rng.standard_normal(n), rng.standard_normal(n)
    assert np.array_equal(rs.gs_forward_py(A, x, f), co.gs_forward(A, x, f))
    for a, b in zip(rs.level_sets_py(A), co.level_sets(A)):
        assert np.array_equal(a, b)
    for a, b in zip(rs.greedy_colouring_py(A), co.greedy_colouring(A)):
        assert np.array_equal(a, b)
    lev, order, off = co.level_sets(A)
    assert np.array_equal(co.gs_forward(A, x, f, order), co.gs_forward(A, x, f))     # level order == natural order result
    col, corder, coff = co.greedy_colouring(A)
    G = rs.sym_nonzero_graph(A)
    r = np.repeat(np.arange(n), np.diff(G.indptr))
    assert not np.any(col[r] == col[G.indices])                                      # proper colouring


def test_level_count_matches_survey_probe():
    """SURVEY 7 hard parts: lexicographic natural-order GS has ~2N-1 levels in 2-D (value-based dependencies)."""
    N = 33
    A = pr.stencil_p1(N - 1, 2)
    lev, order, off = co.level_sets(A)
    assert len(off) - 1 == 2 * (N - 2) - 1          # interior nodes only: Dirichlet rows are isolated
    assert len(co.greedy_colouring(A)[2]) - 1 == 2  # red-black


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_oracle_vcycle_matches_golden(name):
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, **kw)
    cm = co.from_hierarchy(H)
    lf = H.finest_level
    v, hist = cm.vcycle(np.zeros(H.n(lf)), H.b_dict[lf], ncycles=K, history=True)
    assert np.abs(hist - d["vcycle_resnorm"]).max() <= 1e-13 * d["vcycle_resnorm"].max()
    assert np.abs(v - d["vcycle_v"][-1]).max() <= 1e-13 * np.abs(v).max()


@pytest.mark.parametrize("r_mode", ["full_weighting", "transpose"])
def test_c_oracle_other_restrictions_match_restated(r_mode):
    H = pr.build_hierarchy(dim=3, c=2, coarsest_level=0, finest_level=2, perm_seed=3)
    mg = rs.from_hierarchy(H, r_mode=r_mode)
    cm = co.from_hierarchy(H, r_mode=r_mode)
    f = H.b_dict[2]
    v1, h1 = mg.solve_cycles(np.zeros_like(f), f, 4)
    v2, h2 = cm.vcycle(np.zeros(H.n(2)), f, ncycles=4, history=True)
    assert np.allclose(h1, h2, rtol=1e-12, atol=0)
    assert np.abs(v1[:, 0] - v2).max() < 1e-13
    assert h1[-1] < h1[0]


def _oracle_fixture_cases():
    import importlib.util
    p = os.path.join(ROOT, "tests", "golden", "gen_oracle_fixtures.py")
    spec = importlib.util.spec_from_file_location("gen_oracle_fixtures", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["oracle_3d_p1_inj", "oracle_3d_p1_perm_transpose", "oracle_2d_gs", "oracle_2d_gs_color_perm", "oracle_3d_p2_transpose"])
def test_oracle_reproduces_its_committed_fixtures(name):
    """The parts without reference text (3-D, P2, Gauss-Seidel): the oracle's outputs are frozen in tests/golden/oracle_*.npz."""
    m = _oracle_fixture_cases()
    kw, r_mode, smoother, K = m.CASES[name]
    H = m.build(kw)
    lf = H.finest_level
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    cm = co.from_hierarchy(H, r_mode=r_mode, smoother=smoother)
    f = H.b_dict[lf][:, 0]
    v, hist = cm.vcycle(np.zeros_like(f), f, ncycles=K, history=True)
    assert np.array_equal(v, d["v"]) and np.array_equal(hist, d["resnorm"])
    if smoother.startswith("gs"):
        A = H.A_sp_dict[lf][0]
        lev, order, off = co.level_sets(A)
        col, corder, coff = co.greedy_colouring(A)
        assert np.array_equal(lev, d["level_of_row"]) and np.array_equal(off, d["level_offsets"])
        assert np.array_equal(col, d["colour_of_row"]) and np.array_equal(coff, d["colour_offsets"])


@pytest.mark.parametrize("dim,c,lf", [(2, 4, 2), (3, 2, 2), (2, 1, 3)])
def test_c_structured_builder_bitwise_equal_python_builders(dim, c, lf):
    """orc_mg_build_poisson (the in-place builder that lets the 513^3 configuration run on a host) produces, array for array
    and bit for bit, what problems.stencil_p1 / prolongation / injection and the reference's getJacobiMatrices produce."""
    from multigrid_dolfinx_b200 import problems as pr
    if c * 2 ** 0 + 1 < 3:
        lc = 1
    else:
        lc = 0
    S = co.StructuredCOracleMG(dim, c, lc, lf)
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=lc, finest_level=lf, with_dicts=False)
    for k, l in enumerate(range(lc, lf + 1)):
        A = H.A_sp_dict[l][0]
        RO, dinv = rs.jacobi_matrices(A)
        assert np.array_equal(S.array(k, 0), A.indptr) and np.array_equal(S.array(k, 1), A.indices)
        assert S.array(k, 2).tobytes() == np.ascontiguousarray(A.data).tobytes()
        assert np.array_equal(S.array(k, 3), RO.indptr) and np.array_equal(S.array(k, 4), RO.indices)
        assert S.array(k, 5).tobytes() == np.ascontiguousarray(RO.data).tobytes()
        assert S.array(k, 6).tobytes() == np.ascontiguousarray(dinv).tobytes()
        if k > 0:
            P = H.P[l - 1]
            assert np.array_equal(S.array(k, 7), P.indptr) and np.array_equal(S.array(k, 8), P.indices)
            assert S.array(k, 9).tobytes() == np.ascontiguousarray(P.data).tobytes()
            assert np.array_equal(S.array(k, 10), H.inj[l - 1])
    # and the cycle on it equals the cycle on the scipy-built hierarchy
    f = H.b_dict[lf][:, 0]
    v0 = np.zeros_like(f)
    assert np.array_equal(S.vcycle(v0, f, ncycles=3), co.from_hierarchy(H).vcycle(v0, f, ncycles=3))
    assert co.set_threads(0) >= 1
    S.close()
