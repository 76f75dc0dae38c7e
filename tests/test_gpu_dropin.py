"""The drop-in module speaks the reference's protocol (Multigrid_prototype.py:135-143)."""
import numpy as np
import pytest

from conftest import load_golden
from multigrid_dolfinx_b200 import problems as pr

pytestmark = pytest.mark.gpu


def test_prototype_protocol_from_coordinate_dicts():
    """getJacobiMatrices -> initialize_problem -> FullMultiGrid_test / V_cycle_scheme with ONLY the reference's
    16 attributes (transfers are derived from the coordinate dicts like the reference does)."""
    from multigrid_dolfinx_b200 import multigrid as mg
    d, kw, K = load_golden("proto_perm_mu50")
    H = pr.build_hierarchy(dim=2, with_dicts=True, **kw)

    class Var_initializer:          # Multigrid_prototype.py:15-32
        pass
    obj = Var_initializer()
    for a in ("mesh_dof_list_dict", "element_size", "coarsest_level_elements_per_dim", "coarsest_level", "finest_level",
              "A_sp_dict", "b_dict", "mu0", "mu1", "mu2", "omega", "residual_per_V_cycle_finest",
              "error_per_V_cycle_finest", "u_exact_fine", "V_fine_dolfx"):
        setattr(obj, a, getattr(H, a))
    obj.A_jacobi_sp_dict = {k: mg.getJacobiMatrices(v) for k, v in H.A_sp_dict.items()}      # proto:135-136
    mg.initialize_problem(obj)                                                               # proto:138-140
    lf = H.finest_level
    v = np.zeros((H.n(lf), 1))
    for k in range(K):
        v = mg.V_cycle_scheme(obj.A_jacobi_sp_dict[lf], v, H.b_dict[lf])
        assert v.shape == (H.n(lf), 1)
        assert np.abs(v[:, 0] - d["vcycle_v"][k]).max() <= 1e-10 * np.abs(d["vcycle_v"][k]).max()
    out = mg.V_cycle_scheme(obj.A_jacobi_sp_dict[lf], np.zeros((H.n(lf), 1)), H.b_dict[lf], True)
    assert len(out) == 4 and out[1].shape == (H.n(lf - 1), 1)
    # lazy Jacobi tuple materialises the reference's matrices
    assert np.array_equal(obj.A_jacobi_sp_dict[lf][0].data, d["rj_data"])
    # FMG test driver (the one the prototype runs, proto:141-143)
    u, f2h, v2h, errh = mg.FullMultiGrid_test(obj.A_jacobi_sp_dict[lf], H.b_dict[lf], True)
    assert u.shape == (H.n(lf), 1) and errh.shape == (H.n(lf), 1)
    with pytest.raises(ValueError):
        mg.FullMultiGrid_test(obj.A_jacobi_sp_dict[lf], H.b_dict[lf], False)      # reference quirk, multigrid.py:331-333
    # transfer functions with the reference's signatures (test/test_restriction_interpolation.py:119-122)
    md = H.mesh_dof_list_dict
    e = d["in_e"].reshape(-1, 1); x = d["in_x"].reshape(-1, 1)
    assert np.array_equal(mg.Interpolation2D(e, md[lf - 1], md[lf], H.element_size[lf - 1], H.element_size[lf], H.n(lf))[:, 0], d["interp"])
    assert np.array_equal(mg.Restriction2D_direct(x, md[lf - 1], md[lf], H.n(lf - 1))[:, 0], d["inj"])
    fw = mg.Restriction2D(x, md[lf - 1], md[lf], H.element_size[lf - 1], H.element_size[lf], H.n(lf - 1))[:, 0]
    assert np.abs(fw - d["fw"]).max() <= 4e-16 * np.abs(d["fw"]).max()
    assert np.array_equal(mg.jacobiRelaxation(obj.A_jacobi_sp_dict[lf], x, d["in_g"].reshape(-1, 1), 5)[:, 0], d["jac_5"])


def test_fmg_driver_converges():
    from multigrid_dolfinx_b200 import multigrid as mg
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=3, mu1=2, mu2=2, with_dicts=False)
    mg.restriction = "transpose"
    try:
        H.A_jacobi_sp_dict = {k: mg.getJacobiMatrices(v) for k, v in H.A_sp_dict.items()}
        mg.initialize_problem(H)
        import os, tempfile
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.chdir(td)
            try:
                u = mg.FullMultiGrid(H.A_jacobi_sp_dict[3], H.b_dict[3])
                assert os.path.exists("iter_count_for_diff_num_elems_4_levels.csv")
            finally:
                os.chdir(cwd)
    finally:
        mg.restriction = "injection"
    A = H.A_sp_dict[3][0]
    assert np.linalg.norm(H.b_dict[3] - A.dot(u)) <= 1e-10
    assert H.residual_per_V_cycle_finest[-1] <= 1e-11


@pytest.mark.parametrize("with_mass", [False, True])
def test_device_fmg_matches_restated_fmg(with_mass):
    """mgb_fmg (FullMultiGrid, multigrid.py:271-307, on the device) vs the numpy restatement: same cycle count, same
    residual-norm history (1e-10 relative: the norms go down to 1e-11), same solution."""
    import scipy.sparse as sp
    from multigrid_dolfinx_b200.engine import MGEngine
    from oracle import restated as rs
    H = pr.build_hierarchy(dim=2, c=4, coarsest_level=0, finest_level=3, mu1=2, mu2=2, with_dicts=False)
    mg_o = rs.from_hierarchy(H, r_mode="transpose")
    n = H.n(3)
    M = sp.diags(np.linspace(0.5, 1.5, n) / n, 0).tocsr() if with_mass else None      # any SPD "mass" matrix exercises the path
    v_o, h_o = rs.fmg(mg_o, H.b_dict, H.mu0, tol=1e-11, max_cycles=200, M=M)
    eng = MGEngine.from_hierarchy(H, r_mode="transpose")
    for l in H.levels():
        eng.set_rhs(l, H.b_dict[l])
    if with_mass:
        eng.set_mass_matrix(3, M)
    v_g, h_g = eng.fmg(H.mu0, 1e-11, 200)
    assert abs(len(h_g) - len(h_o)) <= 1 and h_g[-1] <= 1e-11          # (the last norms sit at the 1e-11 threshold)
    k = min(len(h_g), len(h_o)) - 3
    assert np.abs(h_g[:k] - np.array(h_o[:k])).max() <= 1e-9 * h_o[0]
    assert np.abs(v_g - v_o[:, 0]).max() <= 1e-10 * np.abs(v_o).max()
    eng.close()


@pytest.mark.parametrize("with_mass", [False, True])
def test_fmg_records_error_and_residual_every_cycle(with_mass):
    """multigrid.py:288-302: the finest-level loop appends the error norm (:292-293) AND the residual norm (:294-295) after EVERY
    V-cycle, then writes [cells per dimension, cycle count] to the iteration CSV.  Here both norms are formed on the device in
    the norm the reference uses (sqrt(x^T M x) with the mass matrix standing in for the dolfinx form, else l2)."""
    import csv, os, tempfile
    import scipy.sparse as sp
    from multigrid_dolfinx_b200 import multigrid as mg
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=3, mu1=2, mu2=2, perm_seed=4, with_dicts=True)
    n = H.n(3)
    A = H.A_sp_dict[3][0]
    import scipy.sparse.linalg as spla
    H.u_exact_fine = spla.spsolve(A.tocsc(), H.b_dict[3][:, 0])                 # nodal values of the discrete solution
    H.V_fine_dolfx = sp.diags(np.linspace(0.5, 1.5, n) / n, 0).tocsr() if with_mass else None
    H.residual_per_V_cycle_finest, H.error_per_V_cycle_finest = [], []
    mg.restriction = "transpose"
    cwd = os.getcwd()
    try:
        H.A_jacobi_sp_dict = {k: mg.getJacobiMatrices(v) for k, v in H.A_sp_dict.items()}
        mg.initialize_problem(H)
        assert "hotrow(" in mg.engine().describe()                               # permuted numbering + coordinate dicts -> lattice numbering handed over
        with tempfile.TemporaryDirectory() as td:
            os.chdir(td)
            u = mg.FullMultiGrid(H.A_jacobi_sp_dict[3], H.b_dict[3])
            rows = list(csv.reader(open("iter_count_for_diff_num_elems_4_levels.csv")))
            os.chdir(cwd)
    finally:
        os.chdir(cwd)
        mg.restriction = "injection"
    res, err = H.residual_per_V_cycle_finest, H.error_per_V_cycle_finest
    assert len(res) == len(err) >= 3 and rows == [[str(8 * 2 ** 3), str(len(res))]]
    assert res[-1] <= 1e-11 and all(res[k + 1] < res[k] for k in range(len(res) - 1))
    assert all(err[k + 1] < err[k] for k in range(min(len(err), 6) - 1))
    # the last entries are the norms of the returned iterate
    M = H.V_fine_dolfx
    nrm = (lambda x: float(np.sqrt(x @ M.dot(x)))) if with_mass else (lambda x: float(np.linalg.norm(x)))
    assert abs(err[-1] - nrm(u[:, 0] - H.u_exact_fine)) <= 1e-9 * max(err[0], 1e-300) + 1e-14
    assert abs(res[-1] - nrm(H.b_dict[3][:, 0] - A.dot(u[:, 0]))) <= 1e-9 * res[0]
