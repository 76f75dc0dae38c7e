"""Synthetic assembler (stand-in for Multigrid_prototype.py:62-118) and transfer matrices."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

from multigrid_dolfinx_b200 import problems as pr


@pytest.mark.parametrize("dim,m", [(2, 4), (2, 8), (3, 2), (3, 4)])
@pytest.mark.parametrize("seed", [None, 3])
def test_stencil_equals_element_assembly(dim, m, seed):
    perm = pr.make_permutation((m + 1) ** dim, seed)
    A1 = pr.assemble_p1(m, dim, perm)[0]
    A2 = pr.stencil_p1(m, dim, perm)
    assert np.array_equal(A1.indptr, A2.indptr) and np.array_equal(A1.indices, A2.indices)
    assert np.array_equal(A1.data, A2.data)
    assert A2.indices.dtype == np.int32 and A2.has_sorted_indices


@pytest.mark.parametrize("dim,m,stored,nonzero", [(2, 16, 7, 5), (3, 8, 15, 7)])
def test_dolfinx_shaped_pattern(dim, m, stored, nonzero):
    """interior rows store 7 (2-D) / 15 (3-D) entries of which 5 / 7 are nonzero; nnz formulas."""
    A = pr.stencil_p1(m, dim)
    N = m + 1
    cnt = np.diff(A.indptr)
    assert cnt.max() == stored
    nzc = np.array([(A.data[A.indptr[i]:A.indptr[i + 1]] != 0).sum() for i in range(A.shape[0])])
    assert nzc.max() == nonzero
    if dim == 2:
        assert A.nnz == N * N + 4 * N * (N - 1) + 2 * (N - 1) ** 2
    else:
        assert A.nnz == N ** 3 + 2 * (3 * N * N * (N - 1) + 3 * N * (N - 1) ** 2 + (N - 1) ** 3)


def test_nnz_formula_at_config_sizes():
    """cfg2 fine level 2049^2 -> 29,372,417 stored entries; cfg3 129^3 -> 31,802,497."""
    N = 2049
    assert N * N + 4 * N * (N - 1) + 2 * (N - 1) ** 2 == 29372417
    N = 129
    assert N ** 3 + 2 * (3 * N * N * (N - 1) + 3 * N * (N - 1) ** 2 + (N - 1) ** 3) == 31802497


@pytest.mark.parametrize("dim,m", [(2, 8), (3, 4)])
def test_rhs_reproduces_manufactured_solution(dim, m):
    """u_D = 1 + x^2 + 2y^2 (+3z^2), f = -laplace(u_D) (Multigrid_prototype.py:78,90): P1 on these meshes is nodally exact."""
    A = pr.stencil_p1(m, dim)
    b = pr.rhs_p1(m, dim)
    u = spl.spsolve(A.tocsc(), b[:, 0])
    X = pr._node_coords_int(m + 1, dim) / m
    assert np.abs(u - pr.boundary_function(X)).max() < 1e-13


@pytest.mark.parametrize("dim,Nc", [(2, 5), (2, 9), (3, 3), (3, 5)])
@pytest.mark.parametrize("seed", [None, 1])
def test_transfer_identities(dim, Nc, seed):
    """P values in {1,1/2,1/4[,1/8]}, row sums 1, injection . P = I, nnz/row histogram."""
    Nf = 2 * Nc - 1
    pc = pr.make_permutation(Nc ** dim, seed)
    pf = pr.make_permutation(Nf ** dim, None if seed is None else seed + 1)
    P = pr.prolongation(Nc, dim, pc, pf)
    inj = pr.injection(Nc, dim, pc, pf)
    assert set(np.unique(P.data)) <= {1.0, 0.5, 0.25, 0.125}
    assert np.array_equal(np.asarray(P.sum(axis=1)).ravel(), np.ones(Nf ** dim))
    assert (P.tocsr()[inj] != __import__("scipy.sparse").sparse.identity(Nc ** dim)).nnz == 0
    cnt = np.bincount(np.diff(P.indptr))
    if dim == 2:
        assert cnt[1] == Nc ** 2 and cnt[2] == 2 * Nc * (Nc - 1) and cnt[4] == (Nc - 1) ** 2
    R = pr.full_weighting(P, dim)
    assert np.array_equal(R.toarray(), (0.5 ** dim) * P.toarray().T)


def test_hierarchy_shapes_match_reference_objects():
    H = pr.build_hierarchy(dim=2, c=8, coarsest_level=1, finest_level=3)     # the prototype's 17/33/65 (proto:35-46)
    assert [H.n(l) for l in H.levels()] == [17 ** 2, 33 ** 2, 65 ** 2]
    for l in H.levels():
        A, lev = H.A_sp_dict[l]
        assert lev == l and A.shape == (H.n(l), H.n(l))
        assert H.b_dict[l].shape == (H.n(l), 1)
        assert H.element_size[l] == 1 / (8 * 2 ** l)
        d = H.mesh_dof_list_dict[l]
        assert len(d) == 2 * H.n(l) and d[d[0]] == 0 and len(d[0]) == 3


# ---- 3-D P2 (BASELINE config 4) ----------------------------------------------------------------------------

def test_p2_pattern_matches_survey_probe():
    """P2 on the Kuhn mesh, n = 4 and 8 cells: average row length 23.35 / 25.81, longest row 65, row lengths
    drawn from {10,14,18,19,22,27,32,42,65}; the stiffness matrix is symmetric with zero row sums before boundary conditions."""
    for m, avg in ((4, 23.35), (8, 25.81)):
        A, bnd, load, K = pr.assemble_p2_3d(m, with_load=True)
        cnt = np.diff(A.indptr)
        assert abs(A.nnz / A.shape[0] - avg) < 0.01 and cnt.max() == 65
        assert set(np.unique(cnt)) <= {10, 14, 18, 19, 22, 27, 32, 42, 65}
        assert abs(K - K.T).max() == 0.0 and np.abs(K.sum(1)).max() < 1e-14
        assert A.shape[0] == (2 * m + 1) ** 3


def test_p2_is_exact_for_quadratics():
    m = 4
    A, bnd, load, K = pr.assemble_p2_3d(m, with_load=True)
    N = 2 * m + 1
    X = np.stack(pr.node_multi_index(N, 3), 1) / (N - 1)
    u = 1 + X[:, 0] ** 2 + 2 * X[:, 1] ** 2 + 3 * X[:, 2] ** 2
    rhs = -12.0 * load - K[:, bnd].dot(u[bnd])
    rhs[bnd] = u[bnd]
    uh = spl.spsolve(A.tocsc(), rhs)
    assert np.abs(uh - u).max() < 1e-12


@pytest.mark.parametrize("mc,seed", [(1, None), (2, None), (2, 4)])
def test_p2_prolongation_reproduces_quadratics(mc, seed):
    Nc, Nf = 2 * mc + 1, 4 * mc + 1
    pc = pr.make_permutation(Nc ** 3, seed)
    pf = pr.make_permutation(Nf ** 3, None if seed is None else seed + 1)
    P = pr.prolongation_p2_3d(mc, pc, pf)
    Xc = np.stack(pr.node_multi_index(Nc, 3), 1) / (Nc - 1)
    Xf = np.stack(pr.node_multi_index(Nf, 3), 1) / (Nf - 1)

    def q(X):
        return 1 + X[:, 0] ** 2 + 2 * X[:, 1] * X[:, 2] + 3 * X[:, 2] ** 2 - X[:, 0] * X[:, 1]
    uc, uf = q(Xc), q(Xf)
    if pc is not None:
        t = np.empty_like(uc); t[pc] = uc; uc = t
        t = np.empty_like(uf); t[pf] = uf; uf = t
    assert np.abs(P.dot(uc) - uf).max() < 1e-14
    assert np.abs(np.asarray(P.sum(1)).ravel() - 1).max() < 1e-15
    inj = pr.injection(Nc, 3, pc, pf)
    import scipy.sparse as sp
    assert (P.tocsr()[inj] != sp.identity(Nc ** 3)).nnz == 0


def test_p2_vcycle_converges_with_transpose_restriction():
    from oracle import c_oracle as co
    H = pr.build_hierarchy_p2(c=1, coarsest_level=0, finest_level=2)
    cm = co.from_hierarchy(H, r_mode="transpose")
    f = H.b_dict[2][:, 0]
    v, hist = cm.vcycle(np.zeros_like(f), f, ncycles=8, history=True)
    assert hist[-1] < 1e-2 * hist[0]
