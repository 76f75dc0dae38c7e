"""The device set-up (multigrid_dolfinx_b200/csrc/mgb_devsetup.cu) computes the Gauss-Seidel level sets and the first-fit colouring
as FIXED POINTS of relaxation passes instead of one sequential loop over the rows.  These CPU tests restate the passes in numpy and
check the claims the kernels rest on, against the sequential definitions of the C oracle (oracle/mg_oracle.c):
  * the level sets are the least fixed point of lev[i] = 1 + max{lev[j] : j ~ i, j < i} reached from zero, in at most depth + 1 passes,
    and in-place (chaotic) updates reach the same answer;
  * the colouring is the UNIQUE solution of col[i] = mex{col[j] : j ~ i, j < i}; a row of dependency level k is final after pass
    k + 1 whatever the other rows hold in the meantime;
  * the transposed restriction built by one stable sort of the stored entries by column equals the sequential transpose, duplicates
    included.
The GPU tests (tests/test_gpu_parity.py::test_device_setup_matches_host_setup) compare the kernels' output with the same oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from multigrid_dolfinx_b200 import problems as pr
from oracle import c_oracle as co


def _lower_graph(A):
    """for every row the SET of neighbours j < i in the symmetrised nonzero graph"""
    A = A.tocsr()
    n = A.shape[0]
    nb = [set() for _ in range(n)]
    ip, ix, ax = A.indptr, A.indices, A.data
    for i in range(n):
        for k in range(ip[i], ip[i + 1]):
            j = int(ix[k])
            if j == i or j >= n or ax[k] == 0.0:
                continue
            nb[max(i, j)].add(min(i, j))
    return [sorted(s) for s in nb]


def _matrices():
    H = pr.build_hierarchy(dim=2, c=4, coarsest_level=0, finest_level=2, perm_seed=5, with_dicts=False)
    yield H.A_sp_dict[2][0]
    H = pr.build_hierarchy(dim=3, c=2, coarsest_level=0, finest_level=2, perm_seed=None, with_dicts=False)
    yield H.A_sp_dict[2][0]
    rng = np.random.default_rng(4)
    M = sp.random(300, 300, density=0.02, random_state=7, format="csr")
    M = (M + sp.diags(np.full(300, 4.0))).tocsr()            # unsymmetric pattern, explicit zeros below
    M.data[rng.integers(0, M.nnz, 20)] = 0.0
    M.setdiag(4.0)
    yield M.tocsr()


@pytest.mark.parametrize("inplace", [False, True])
def test_level_sets_as_fixed_point(inplace):
    for A in _matrices():
        lev_ref, _, off = co.level_sets(A)
        nb = _lower_graph(A)
        n = A.shape[0]
        lev = np.zeros(n, dtype=np.int64)
        depth = len(off) - 1
        rng = np.random.default_rng(1)
        for p in range(depth + 2):
            src = lev if inplace else lev.copy()
            changed = False
            for i in (rng.permutation(n) if inplace else range(n)):      # in place: any order of the threads
                m = 1 + max((src[j] for j in nb[i]), default=-1)
                assert m <= lev_ref[i]                                   # never beyond the answer
                if m != lev[i]:
                    lev[i] = m; changed = True
            if not changed:
                break
        assert not changed and p <= depth + 1
        assert np.array_equal(lev, lev_ref)


def test_colouring_as_fixed_point():
    for A in _matrices():
        col_ref, _, _ = co.greedy_colouring(A)
        lev_ref, _, off = co.level_sets(A)
        nb = _lower_graph(A)
        n = A.shape[0]
        rng = np.random.default_rng(2)
        col = rng.integers(0, 5, n)                                      # any start: the fixed point is unique
        for p in range(len(off) + 1):
            new = col.copy()
            for i in range(n):
                used = {int(col[j]) for j in nb[i]}
                c = 0
                while c in used:
                    c += 1
                new[i] = c
            col = new
            assert np.array_equal(col[lev_ref <= p], col_ref[lev_ref <= p])   # level k is final after pass k + 1
            if np.array_equal(col, col_ref):
                break
        assert np.array_equal(col, col_ref)


def test_transpose_by_stable_sort():
    rng = np.random.default_rng(3)
    H = pr.build_hierarchy(dim=2, c=4, coarsest_level=0, finest_level=2, perm_seed=8, with_dicts=False)
    P = H.P[1].tocsr()
    # a duplicate entry (same row, same column, stored twice) must keep its storage order
    ip, ix, ax = P.indptr.copy(), P.indices.copy(), P.data.copy()
    r = 5
    ix = np.insert(ix, ip[r + 1], ix[ip[r]]); ax = np.insert(ax, ip[r + 1], 0.375); ip[r + 1:] += 1
    nr, nc = P.shape
    scale = 0.25
    # sequential definition (mgb_setup.cpp::transpose_scaled)
    cnt = np.bincount(ix, minlength=nc)
    tp = np.concatenate([[0], np.cumsum(cnt)])
    pos = tp[:-1].copy()
    tix = np.zeros(len(ix), dtype=np.int64); tax = np.zeros(len(ix))
    for i in range(nr):
        for k in range(ip[i], ip[i + 1]):
            d = pos[ix[k]]; pos[ix[k]] += 1
            tix[d] = i; tax[d] = ax[k] * scale
    # device form: ONE stable sort of the stored entries by column (entries of a column keep their storage order), then every
    # entry looks up the row of P it came from
    src = np.argsort(ix, kind="stable")
    dix = np.searchsorted(ip, src, side="right") - 1
    dax = ax[src] * scale
    assert np.array_equal(tix, dix) and np.array_equal(tax, dax)
    assert np.array_equal(np.concatenate([[0], np.cumsum(np.bincount(ix, minlength=nc))]), tp)
    R = pr.full_weighting(H.P[1], 2).tocsr()                              # and the definition equals scipy's 2^-d P^T
    cntP = np.bincount(H.P[1].tocsr().indices, minlength=nc)
    assert np.array_equal(np.diff(R.indptr), cntP)
