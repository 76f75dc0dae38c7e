"""The lossless operator coding as an artefact: the library's host routine (the definition the device
builder follows) against the independent numpy restatement in oracle/coding.py, on structured, permuted and ragged operators;
decoding must give the CSR arrays back BIT FOR BIT.  No GPU needed."""
import numpy as np
import pytest
import scipy.sparse as sp

from multigrid_dolfinx_b200 import problems as pr
from multigrid_dolfinx_b200.engine import host_code_operator
from oracle import coding as oc
from oracle import restated as rs


def _same(h, o):
    assert h["mode"] == o["mode"] and h["ndict"] == o["ndict"]
    assert np.array_equal(h["codes"], o["codes"])
    if h["mode"] == 0:
        return
    tv = np.array([np.float64(t[0]) for t in o["table"]]).view(np.uint64)
    td = np.array([t[1] for t in o["table"]], dtype=np.int32)
    k = len(tv)
    assert np.array_equal(h["table"]["val"][:k].view(np.uint64), tv) and np.array_equal(h["table"]["delta"][:k], td)
    assert not h["table"]["val"][k:].view(np.uint64).any() and not h["table"]["delta"][k:].any()
    if h["mode"] in (3, 4):
        assert np.array_equal(h["head"], o["head"])


def _lossless(A, h):
    if h["mode"] == 0:
        return
    coded = {"mode": h["mode"], "codes": h["codes"], "table": list(zip(h["table"]["val"], h["table"]["delta"])), "head": h["head"]}
    if h["mode"] == 4:
        assert np.array_equal(h["anchor"], oc.anchors(A))
        coded["anchor"] = h["anchor"]
    cols, bits = oc.decode(A.shape, A.indptr, A.indices, coded)
    assert np.array_equal(cols, A.indices.astype(np.int64)) and np.array_equal(bits, A.data.view(np.uint64))


def _ragged(n, m, rng, max_len, table=None):
    lens = rng.integers(0, max_len + 1, size=n)
    lens[rng.random(n) < 0.15] = 0
    ix = [rng.choice(m, size=int(min(k, m)), replace=False) for k in lens]
    ip = np.zeros(n + 1, dtype=np.int64); np.cumsum([len(c) for c in ix], out=ip[1:])
    ix = np.concatenate(ix).astype(np.int32) if ip[-1] else np.zeros(0, np.int32)
    ax = rng.standard_normal(ip[-1]) if table is None else table[rng.integers(0, len(table), size=ip[-1])]
    return sp.csr_matrix((ax, ix, ip), shape=(n, m))


@pytest.mark.parametrize("dim,c,lf,seed", [(2, 8, 3, None), (3, 2, 3, None), (2, 8, 2, 5), (3, 2, 2, 7)])
def test_hierarchy_operators_host_routine_equals_numpy_definition(dim, c, lf, seed):
    H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
    A = H.A_sp_dict[lf][0].tocsr()
    RO, _ = rs.jacobi_matrices(A)
    P = H.P[lf - 1].tocsr()
    # P: its rows repeat only when measured from their first column -> anchored row patterns (mode 4) on a lexicographic numbering
    for M, want in ((A, 3 if seed is None else None), (RO.tocsr(), 3 if seed is None else None), (P, 4 if seed is None else 2)):
        for allow in (2, 1, 0):
            h, o = host_code_operator(M, allow), oc.code_operator(M, allow)
            _same(h, o)
            _lossless(M, h)
            if want is not None and M.shape[0] > 300 and allow == 2:
                assert h["mode"] == want, (h["mode"], want)
            if allow < 2:
                assert h["mode"] != 4
            if allow < 1:
                assert h["mode"] not in (3, 4)
    if seed is None:
        assert host_code_operator(P)["ndict"] == {2: 4, 3: 8}[dim]      # one pattern per parity class of the fine node
        assert host_code_operator(P, 1)["mode"] == 2
    if seed is None:                                            # lexicographic numbering: few row patterns
        assert host_code_operator(RO.tocsr())["ndict"] == {2: 10, 3: 28}[dim]
        assert host_code_operator(A)["ndict"] == {2: 17, 3: 53}[dim]
        assert host_code_operator(A, False)["mode"] == 1        # without patterns: pair codes


@pytest.mark.parametrize("nvals", [1, 7, 256, 257])
def test_ragged_operators(nvals):
    rng = np.random.default_rng(nvals)
    table = np.unique(np.concatenate([[0.0, -0.0, 1.0, 2.0 ** -1070, 1e60], rng.standard_normal(300)]).view(np.uint64))[:nvals].view(np.float64)
    if len(table) < nvals:
        table = np.concatenate([table, 100.0 + np.arange(nvals - len(table))])
    for shape in ((900, 900), (900, 300), (300, 900)):
        M = _ragged(shape[0], shape[1], rng, 14, table)
        M.data[:len(table)] = table[:M.nnz] if M.nnz >= len(table) else M.data[:len(table)]      # every table entry occurs
        h, o = host_code_operator(M), oc.code_operator(M)
        _same(h, o)
        _lossless(M, h)
        assert h["mode"] == (0 if nvals == 257 else 2)          # random columns: neither few offsets nor repeating rows


def test_limits_and_degenerate_inputs():
    rng = np.random.default_rng(3)
    n = 4000
    for noff, want in ((5, 1), (255, 1), (256, 2), (257, 2)):   # distinct (col - row) = noff + 1 (the diagonal): 256 still fit the pair dictionary
        offs = np.arange(noff) * 3
        rows = np.arange(n); k = rows % noff
        # two entries per row: the diagonal-like one and one at a row-dependent offset -> rows do not repeat 256 ways only if noff > 256
        M = sp.csr_matrix((np.r_[np.full(n, 2.0), np.full(n, -1.0)], (np.r_[rows, rows], np.r_[rows, rows + 1 + offs[k]])), shape=(n, n + 1000))
        M.sort_indices()
        h, o = host_code_operator(M, allow_patterns=False), oc.code_operator(M, allow_patterns=False)
        _same(h, o); _lossless(M, h)
        assert h["mode"] == want
        hp = host_code_operator(M, allow_patterns=True)
        _same(hp, oc.code_operator(M, True)); _lossless(M, hp)
        # noff distinct rows: 256 patterns x 8 padded entries = 2048 is exactly the table's capacity; 257 patterns do not fit
        assert hp["mode"] == (3 if noff <= 256 else 2) and (hp["ndict"] == noff if noff <= 256 else True)
        assert len(hp["table"]) == (8 * noff if noff <= 256 else 256)
    # the all-ones NaN is the device tables' "empty" marker: such an operator is left uncoded
    M = _ragged(500, 300, rng, 6, np.array([1.0, 2.0]))
    M.data[3] = np.uint64(0xFFFFFFFFFFFFFFFF).view(np.float64)
    assert host_code_operator(M)["mode"] == 0 and oc.code_operator(M)["mode"] == 0
    # long rows on average, empty operator
    assert host_code_operator(sp.csr_matrix(np.ones((40, 40))))["mode"] == 0
    assert host_code_operator(sp.csr_matrix((50, 50)))["mode"] == 0
    # pattern table capacity: 200 patterns x 16 padded entries > 2048 -> per-entry codes instead
    rows = np.repeat(np.arange(n), 9); j = np.tile(np.arange(9), n)
    M = sp.csr_matrix((np.ones(9 * n), (rows, rows + 1 + j * (1 + rows % 200))), shape=(n, n + 2000)); M.sort_indices()
    h = host_code_operator(M); _same(h, oc.code_operator(M)); _lossless(M, h)
    assert h["mode"] == 2                                        # 9 x 200 offsets > 256 as well -> value codes
