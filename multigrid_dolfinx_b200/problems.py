"""Synthetic host assembler: stand-in for the dolfinx/PETSc assembly loop of the reference.

The reference builds, per level, a P1 Poisson stiffness matrix with dolfinx and exports its PETSc
CSR arrays into scipy (``Multigrid_prototype.py:62-118``).  dolfinx is not available where this
engine is developed or benchmarked, so this module emits the *same object shapes* from numpy:

* ``A_sp_dict[l] = (csr_matrix, l)``        (``Multigrid_prototype.py:95-99``)
* ``b_dict[l]``  (n, 1) float64              (``Multigrid_prototype.py:110``)
* ``mesh_dof_list_dict[l]``  {dof -> coord, coord -> dof}, 9-decimal keys (``Multigrid_prototype.py:68-74``)
* ``element_size[l] = 1 / (c * 2**l)``       (``Multigrid_prototype.py:63-64``)

plus the matrix form of the grid transfers the reference evaluates with Python loops
(``multigrid.py:59-120`` interpolation, ``multigrid.py:123-132`` injection,
``multigrid.py:135-198`` full weighting).

It is an INPUT GENERATOR for tests and benchmarks, not part of the product path: the engine ingests
any CSR handed to it.  Meshes: 2-D unit square, every square cut by the (x0,y0)-(x1,y1) diagonal
(dolfinx ``UnitSquareMesh`` default); 3-D unit cube, six Kuhn tetrahedra per cube sharing the
(0,0,0)-(1,1,1) diagonal.  The stored pattern is the cell-connectivity pattern (7 / 15 entries per
interior row) including the structural zeros dolfinx stores; Dirichlet rows and columns are zeroed
in place (zeros kept) with a unit diagonal, as ``assemble_matrix(a, bcs=[bc])`` does.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------------------------
# node numbering helpers
# --------------------------------------------------------------------------------------------


def lex_index(idx, N):
    """Lexicographic node index, x fastest: (i, j[, k]) -> i + N*j [+ N*N*k]."""
    out = np.zeros_like(idx[0])
    stride = 1
    for a in idx:
        out = out + a * stride
        stride *= N
    return out


def node_multi_index(N, dim):
    """All node multi-indices of an N^dim grid, x fastest. Returns list of dim int64 arrays."""
    n = N ** dim
    lin = np.arange(n, dtype=np.int64)
    out = []
    for _ in range(dim):
        out.append(lin % N)
        lin = lin // N
    return out


def make_permutation(n, seed):
    """perm[lexicographic node] = dof.  seed None -> identity (lexicographic DOFs)."""
    if seed is None:
        return None
    rng = np.random.default_rng(seed)
    return rng.permutation(n).astype(np.int64)


# --------------------------------------------------------------------------------------------
# element assembly (general, used at small sizes and to validate the stencil generator)
# --------------------------------------------------------------------------------------------

def _cells_2d(m):
    N = m + 1
    i, j = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    i = i.ravel(); j = j.ravel()
    v0 = i + N * j; v1 = v0 + 1; v2 = v0 + N; v3 = v0 + N + 1
    # two triangles per square, both containing the (v0, v3) diagonal
    return np.concatenate([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], 0)


def _cells_3d(m):
    N = m + 1
    i, j, k = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    base = (i + N * j + N * N * k).ravel()
    e = [1, N, N * N]
    tets = []
    for s in itertools.permutations(range(3)):
        a = base + e[s[0]]
        b = a + e[s[1]]
        c = b + e[s[2]]
        tets.append(np.stack([base, a, b, c], 1))
    return np.concatenate(tets, 0)


def _node_coords_int(N, dim):
    return np.stack(node_multi_index(N, dim), 1)  # integer lattice coordinates


def _p1_local_int(cells, X):
    """Integer-scaled P1 stiffness for all cells.

    Returns K (ncell, d+1, d+1) with  K_true = K * scale,  scale = 1/2 in 2-D, h/6 in 3-D.
    Working on the integer lattice keeps every entry an exact small integer so that structural
    zeros are exact 0.0 and the assembled values do not depend on summation order.
    """
    d = X.shape[1]
    P = X[cells]                                    # (nc, d+1, d)
    E = (P[:, 1:, :] - P[:, :1, :]).astype(np.float64)  # edge matrix rows
    Einv = np.linalg.inv(E)                         # entries in {-1,0,1} on Kuhn / "/" meshes
    G = np.concatenate([-Einv.sum(axis=2, keepdims=True).transpose(0, 2, 1), Einv.transpose(0, 2, 1)], 1)
    # G[c, a, :] = grad of basis a (lattice units); |det E| = 1 on these meshes
    K = np.rint(np.einsum("cad,cbd->cab", G, G))
    return K


def assemble_p1(m, dim, perm=None, bc=True):
    """Element-by-element P1 stiffness on the m^dim mesh.  Returns (A csr, boundary mask, K_full csr).

    ``A`` has Dirichlet rows/cols zeroed in place (stored zeros kept) and unit diagonal;
    ``K_full`` is the matrix before boundary conditions (needed for the lifting of the RHS).
    Columns sorted, int32 indices (PETSc 32-bit build)."""
    N = m + 1
    n = N ** dim
    h = 1.0 / m
    cells = _cells_2d(m) if dim == 2 else _cells_3d(m)
    X = _node_coords_int(N, dim)
    Kint = _p1_local_int(cells, X)
    nv = dim + 1
    rows = np.repeat(cells, nv, axis=1).ravel()
    cols = np.tile(cells, (1, nv)).ravel()
    vals = Kint.ravel()
    mi = node_multi_index(N, dim)
    bnd = np.zeros(n, dtype=bool)
    for a in mi:
        bnd |= (a == 0) | (a == N - 1)
    if perm is not None:
        rows = perm[rows]; cols = perm[cols]
        b2 = np.zeros(n, dtype=bool); b2[perm] = bnd; bnd = b2
    Kfull = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    Kfull.sum_duplicates(); Kfull.sort_indices()
    scale_num, scale_den = (1.0, 2.0) if dim == 2 else (h, 6.0)
    Kfull.data = Kfull.data * scale_num / scale_den
    A = Kfull.copy()
    if bc:
        r = np.repeat(np.arange(n), np.diff(A.indptr))
        kill = bnd[r] | bnd[A.indices]
        A.data[kill] = 0.0
        A.data[(r == A.indices) & bnd[r]] = 1.0
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    Kfull.indices = Kfull.indices.astype(np.int32); Kfull.indptr = Kfull.indptr.astype(np.int32)
    return A, bnd, Kfull


# --------------------------------------------------------------------------------------------
# structured generator (index arithmetic; identical output to assemble_p1, usable at 2049^2 / 513^3)
# --------------------------------------------------------------------------------------------

def _pattern_offsets(dim):
    """Neighbour offsets of the cell-connectivity pattern, with the stencil weight of each.
    2-D "/" mesh: 4 axis neighbours (-1) + 2 diagonal neighbours (stored 0) + self.
    3-D Kuhn mesh: 6 axis neighbours (-1) + 8 face/space-diagonal neighbours (stored 0) + self."""
    offs = []
    for o in itertools.product((-1, 0, 1), repeat=dim):
        if all(x >= 0 for x in o) or all(x <= 0 for x in o):
            nz = sum(1 for x in o if x != 0)
            w = 0.0 if nz == 0 else (-1.0 if nz == 1 else 0.0)
            offs.append((o, w))
    return offs


def stencil_p1(m, dim, perm=None, rows=None):
    """Same matrix as ``assemble_p1(m, dim, perm)[0]`` by index arithmetic (no element loop).

    ``rows`` (optional, lexicographic numbering only): a (start, stop) row range -> returns the CSR
    slab of those rows with global column indices (used to generate shards without the global matrix).
    """
    N = m + 1
    n = N ** dim
    h = 1.0 / m
    unit = 1.0 if dim == 2 else h          # off-diagonal magnitude: 1 (2-D), h (3-D)
    diag_int = (4.0 if dim == 2 else 6.0) * unit
    if rows is None:
        r0, r1 = 0, n
    else:
        assert perm is None
        r0, r1 = rows
    lin = np.arange(r0, r1, dtype=np.int64)
    mi = []
    t = lin.copy()
    for _ in range(dim):
        mi.append(t % N); t //= N
    on_bnd = np.zeros(r1 - r0, dtype=bool)
    for a in mi:
        on_bnd |= (a == 0) | (a == N - 1)
    offs = _pattern_offsets(dim)
    # sort offsets by linear displacement so lexicographic columns come out sorted
    lin_off = [sum(o[d] * N ** d for d in range(dim)) for o, _ in offs]
    order = np.argsort(lin_off)
    cols = np.empty((r1 - r0, len(offs)), dtype=np.int64)
    vals = np.empty((r1 - r0, len(offs)), dtype=np.float64)
    valid = np.empty((r1 - r0, len(offs)), dtype=bool)
    for slot, oi in enumerate(order):
        o, w = offs[oi]
        ok = np.ones(r1 - r0, dtype=bool)
        nb_bnd = np.zeros(r1 - r0, dtype=bool)
        for d in range(dim):
            c = mi[d] + o[d]
            ok &= (c >= 0) & (c <= N - 1)
            nb_bnd |= (c <= 0) | (c >= N - 1)
        valid[:, slot] = ok
        cols[:, slot] = lin + lin_off[oi]
        if all(x == 0 for x in o):
            vals[:, slot] = np.where(on_bnd, 1.0, diag_int)
        else:
            vals[:, slot] = np.where(on_bnd | nb_bnd, 0.0, w * unit)
    counts = valid.sum(1)
    indptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = cols[valid]
    data = vals[valid]
    data = data + 0.0  # normalise -0.0 -> 0.0
    if perm is None:
        A = sp.csr_matrix((data, indices.astype(np.int32), indptr.astype(np.int32 if indptr[-1] < 2**31 else np.int64)),
                          shape=(r1 - r0, n))
        return A
    rr = np.repeat(perm[lin], counts)
    cc = perm[indices]
    A = sp.coo_matrix((data, (rr, cc)), shape=(n, n)).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    return A


# --------------------------------------------------------------------------------------------
# right-hand side  (u_D = 1 + x^2 + 2 y^2 [+ 3 z^2],  f = -6 [-12]),  Multigrid_prototype.py:76-110
# --------------------------------------------------------------------------------------------

def boundary_function(X):
    w = (1.0, 2.0, 3.0)
    u = np.ones(X.shape[0])
    for d in range(X.shape[1]):
        u = u + w[d] * X[:, d] ** 2
    return u


def rhs_p1(m, dim, perm=None):
    """Load vector for f = -laplace(u_D) with lifting and set_bc, shape (n, 1) like proto:110."""
    N = m + 1
    n = N ** dim
    h = 1.0 / m
    Xi = _node_coords_int(N, dim)
    X = Xi * h
    bnd = np.zeros(n, dtype=bool)
    for d in range(dim):
        bnd |= (Xi[:, d] == 0) | (Xi[:, d] == N - 1)
    fval = -6.0 if dim == 2 else -12.0
    # sum over cells of f * |cell| / (dim+1) per vertex = f * (number of adjacent cells) * |cell|/(dim+1)
    cells = _cells_2d(m) if dim == 2 else _cells_3d(m)
    cnt = np.bincount(cells.ravel(), minlength=n).astype(np.float64)
    vol = h ** dim / (2.0 if dim == 2 else 6.0)
    b = fval * cnt * vol / (dim + 1)
    uD = boundary_function(X)
    # lifting: b -= K[:, bnd] u_D[bnd]  on free rows; the only nonzero couplings are axis neighbours
    unit = 1.0 if dim == 2 else h
    g = np.where(bnd, uD, 0.0).reshape((N,) * dim, order="F")
    lift = np.zeros_like(g)
    for d in range(dim):
        sl_lo = [slice(None)] * dim; sl_hi = [slice(None)] * dim
        sl_lo[d] = slice(0, N - 1); sl_hi[d] = slice(1, N)
        lift[tuple(sl_hi)] += g[tuple(sl_lo)]
        lift[tuple(sl_lo)] += g[tuple(sl_hi)]
    b = b + unit * lift.reshape(-1, order="F")      # off-diagonals are -unit
    b[bnd] = uD[bnd]
    if perm is not None:
        bp = np.empty_like(b); bp[perm] = b; b = bp
    return b.reshape(n, 1)


# --------------------------------------------------------------------------------------------
# grid transfers in matrix form
# --------------------------------------------------------------------------------------------

def prolongation(Nc, dim, perm_c=None, perm_f=None, rows=None):
    """P (n_f x n_c): tensor-product linear interpolation, the matrix of ``Interpolation2D``
    (multigrid.py:59-120; 3-D is its tensor-product extension).

    Row entries are stored in the reference's summation order -- (i-,j-), (i+,j-), (i-,j+), (i+,j+),
    x fastest (multigrid.py:109-118) -- NOT sorted by column, so that a sequential row sum
    reproduces the reference's result bit for bit (weights are powers of two).

    ``rows`` (lexicographic numbering only): (start, stop) -> only those rows of P (shape (stop-start, n_c))."""
    Nf = 2 * Nc - 1
    nc = Nc ** dim
    if rows is None:
        nf = Nf ** dim
        mi = node_multi_index(Nf, dim)
    else:
        assert perm_c is None and perm_f is None
        nf = rows[1] - rows[0]
        t = np.arange(rows[0], rows[1], dtype=np.int64)
        mi = []
        for _ in range(dim):
            mi.append(t % Nf); t = t // Nf
    ncomb = 2 ** dim
    cols = np.zeros((nf, ncomb), dtype=np.int64)
    wts = np.ones((nf, ncomb), dtype=np.float64)
    valid = np.ones((nf, ncomb), dtype=bool)
    for comb in range(ncomb):
        stride = 1
        for d in range(dim):
            hi = (comb >> d) & 1
            odd = (mi[d] & 1) == 1
            c = np.where(odd, (mi[d] - 1) // 2 + hi, mi[d] // 2)
            if hi:
                valid[:, comb] &= odd          # even coordinate: only the "-" choice exists
            wts[:, comb] *= np.where(odd, 0.5, 1.0)
            cols[:, comb] += c * stride
            stride *= Nc
    counts = valid.sum(1)
    indptr = np.zeros(nf + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = cols[valid]
    data = wts[valid]
    if perm_c is not None:
        indices = perm_c[indices]
    if perm_f is not None:
        # reorder rows: row perm_f[lex] <- lex row
        inv = np.empty(nf, dtype=np.int64); inv[perm_f] = np.arange(nf)
        newcounts = counts[inv]
        newptr = np.zeros(nf + 1, dtype=np.int64); np.cumsum(newcounts, out=newptr[1:])
        src = (indptr[inv][:, None] + np.arange(ncomb)[None, :])
        keep = np.arange(ncomb)[None, :] < newcounts[:, None]
        indices = indices[src[keep]]; data = data[src[keep]]; indptr = newptr
    P = sp.csr_matrix((data, indices.astype(np.int32), indptr.astype(np.int32)), shape=(nf, nc))
    return P


def injection(Nc, dim, perm_c=None, perm_f=None):
    """g[coarse dof] = fine dof at the same coordinate (multigrid.py:128-131)."""
    Nf = 2 * Nc - 1
    mi = node_multi_index(Nc, dim)
    fine = lex_index([2 * a for a in mi], Nf)
    if perm_f is not None:
        fine = perm_f[fine]
    if perm_c is not None:
        out = np.empty_like(fine); out[perm_c] = fine; fine = out
    return fine.astype(np.int32)


def full_weighting(P, dim):
    """R = 2^-dim P^T: in 2-D exactly the reference's (unused) Restriction2D (multigrid.py:135-198)."""
    R = (P.T * (0.5 ** dim)).tocsr()
    R.sort_indices()
    R.indices = R.indices.astype(np.int32); R.indptr = R.indptr.astype(np.int32)
    return R


def transpose_restriction(P):
    R = P.T.tocsr(); R.sort_indices()
    R.indices = R.indices.astype(np.int32); R.indptr = R.indptr.astype(np.int32)
    return R


# --------------------------------------------------------------------------------------------
# reference-shaped hierarchy
# --------------------------------------------------------------------------------------------

def mesh_dof_dict(m, dim, perm=None):
    """{dof -> coord tuple rounded to 9 dp, coord -> dof} as in Multigrid_prototype.py:68-74.
    2-D coordinates are (x, y, 0.0) like dolfinx's tabulate_dof_coordinates."""
    N = m + 1
    h = 1.0 / m
    Xi = _node_coords_int(N, dim)
    d = {}
    for lex in range(N ** dim):
        dof = int(perm[lex]) if perm is not None else lex
        c = [round(float(Xi[lex, k] * h), 9) for k in range(dim)]
        if dim == 2:
            c.append(0.0)
        t = tuple(c)
        d[dof] = t
        d[t] = dof
    return d


@dataclass
class Hierarchy:
    """Attribute bag with the 16 attributes ``initialize_problem`` reads (multigrid.py:30-45)
    plus the matrix-form transfers.  Any object with these attributes works as the reference's
    ``Var_initializer`` (Multigrid_prototype.py:15-32)."""
    dim: int
    coarsest_level_elements_per_dim: int
    coarsest_level: int
    finest_level: int
    mu0: int = 2
    mu1: int = 2
    mu2: int = 2
    omega: float = 2.0 / 3.0
    A_sp_dict: dict = field(default_factory=dict)
    A_jacobi_sp_dict: dict = field(default_factory=dict)
    b_dict: dict = field(default_factory=dict)
    mesh_dof_list_dict: dict = field(default_factory=dict)
    element_size: dict = field(default_factory=dict)
    residual_per_V_cycle_finest: list = field(default_factory=list)
    error_per_V_cycle_finest: list = field(default_factory=list)
    u_exact_fine: object = None
    V_fine_dolfx: object = None
    # matrix-form extras (not in the reference)
    P: dict = field(default_factory=dict)       # P[l]: level l -> l+1
    inj: dict = field(default_factory=dict)     # inj[l]: coarse level l dof -> fine level l+1 dof
    perms: dict = field(default_factory=dict)
    nodes_per_dim: dict = field(default_factory=dict)

    def levels(self):
        return range(self.coarsest_level, self.finest_level + 1)

    def n(self, l):
        return self.nodes_per_dim[l] ** self.dim


def rcm_permutation(m, dim, seed):
    """perm[lexicographic node] = dof for a banded-but-not-lexicographic numbering: the reverse Cuthill-McKee ordering
    (scipy.sparse.csgraph) of a seeded random numbering of the P1 pattern.  The closest synthetic stand-in for the numbering
    dolfinx itself produces (graph-reordered); neighbouring DOFs are close, whole rows do not repeat."""
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    n = (m + 1) ** dim
    p = make_permutation(n, seed)
    order = reverse_cuthill_mckee(stencil_p1(m, dim, p), symmetric_mode=True)      # order[k] = old dof placed at position k
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.arange(n, dtype=np.int64)
    return inv[p]


def build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=2, perm_seed=None, mu0=2, mu1=2, mu2=2,
                    omega=2.0 / 3.0, with_dicts=None, with_rhs=True, assemble="stencil", rcm=False):
    """Build the level hierarchy the reference's driver would hand to ``initialize_problem``.

    cells per dimension at level l = c * 2**l (Multigrid_prototype.py:63).  ``perm_seed`` not None
    gives every level an independent random DOF numbering (dolfinx numbering is arbitrary; the
    algorithm must be numbering-agnostic).  ``with_dicts`` builds the coordinate dicts (only needed
    to drive the imported reference; O(n) Python, default: only when n_fine <= 70k)."""
    H = Hierarchy(dim=dim, coarsest_level_elements_per_dim=c, coarsest_level=coarsest_level,
                  finest_level=finest_level, mu0=mu0, mu1=mu1, mu2=mu2, omega=omega)
    for l in H.levels():
        m = c * 2 ** l
        N = m + 1
        n = N ** dim
        H.nodes_per_dim[l] = N
        H.element_size[l] = 1 / m
        perm = make_permutation(n, None if perm_seed is None else perm_seed + 1000 * l)
        if rcm and perm_seed is not None:
            perm = rcm_permutation(m, dim, perm_seed + 1000 * l)
        H.perms[l] = perm
        if assemble == "stencil":
            A = stencil_p1(m, dim, perm)
        else:
            A = assemble_p1(m, dim, perm)[0]
        H.A_sp_dict[l] = (A, l)
        if with_rhs:
            H.b_dict[l] = rhs_p1(m, dim, perm)
        wd = with_dicts if with_dicts is not None else (dim == 2 and n <= 70000)
        if wd:
            H.mesh_dof_list_dict[l] = mesh_dof_dict(m, dim, perm)
    for l in range(coarsest_level, finest_level):
        Nc = H.nodes_per_dim[l]
        H.P[l] = prolongation(Nc, dim, H.perms[l], H.perms[l + 1])
        H.inj[l] = injection(Nc, dim, H.perms[l], H.perms[l + 1])
    return H


# --------------------------------------------------------------------------------------------
# 3-D P2 (quadratic Lagrange) on the Kuhn mesh -- BASELINE config 4.  No reference text exists for
# P2 (the reference is 2-D P1 only): element matrices and transfers below are the standard FE ones.
# --------------------------------------------------------------------------------------------
# DOFs: vertices and edge midpoints.  Every nonzero vector of {0,1}^3 is an edge direction of the Kuhn
# triangulation, so the P2 DOFs of an m^3-cell mesh are exactly the points of the (2m+1)^3 lattice of half
# cell widths ("doubled lattice"); numbering is lexicographic on that lattice, x fastest.

_P2_EDGES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def _kuhn_tets_doubled(m):
    """(ntet, 4, 3) vertex coordinates on the doubled lattice."""
    i, j, k = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], 1) * 2
    out = []
    eye = np.eye(3, dtype=np.int64) * 2
    for s in itertools.permutations(range(3)):
        v0 = base
        v1 = v0 + eye[s[0]]
        v2 = v1 + eye[s[1]]
        v3 = v2 + eye[s[2]]
        out.append(np.stack([v0, v1, v2, v3], 1))
    return np.concatenate(out, 0)


def _p2_element_int(verts2):
    """Integer-scaled P2 stiffness of every tet: K_true = E * h / 120 (10 x 10, DOF order: 4 vertices, 6 edges).
    With g_ab = grad(lambda_a).grad(lambda_b) in cell units and V = 1/6:
      vertex-vertex  (3/5 V) g_ii, (-1/5 V) g_ij ;  vertex-edge 4 V [g_ik c(i,j) + g_ij c(i,k)], c = 3/20 (same) / -1/20 ;
      edge-edge      16 V/20 [g_jl (1+d_ik) + g_jk (1+d_il) + g_il (1+d_jk) + g_ik (1+d_jl)]."""
    X = verts2.astype(np.float64) / 2.0                       # cell units
    Emat = X[:, 1:, :] - X[:, :1, :]
    Einv = np.linalg.inv(Emat)
    G = np.concatenate([-Einv.sum(axis=2, keepdims=True).transpose(0, 2, 1), Einv.transpose(0, 2, 1)], 1)   # grad lambda
    g = np.rint(np.einsum("cad,cbd->cab", G, G))              # integers on the Kuhn mesh
    nt = X.shape[0]
    E = np.zeros((nt, 10, 10))
    d = np.eye(4)
    for a in range(4):
        for b in range(4):
            E[:, a, b] = (12.0 if a == b else -4.0) * g[:, a, b]                      # x (V/20) with V factored out below
    for a in range(4):
        for e, (j, k) in enumerate(_P2_EDGES):
            cj = 3.0 if a == j else -1.0
            ck = 3.0 if a == k else -1.0
            val = 4.0 * (g[:, a, k] * cj + g[:, a, j] * ck)
            E[:, a, 4 + e] = val
            E[:, 4 + e, a] = val
    for e1, (i, j) in enumerate(_P2_EDGES):
        for e2, (k, l) in enumerate(_P2_EDGES):
            E[:, 4 + e1, 4 + e2] = 16.0 * (g[:, j, l] * (1 + d[i, k]) + g[:, j, k] * (1 + d[i, l])
                                           + g[:, i, l] * (1 + d[j, k]) + g[:, i, k] * (1 + d[j, l]))
    return E            # K_true = E * (V/20) / h^2 * h^3 = E * h / 120   (V = 1/6 in cell units)


def _p2_dofs(verts2, N):
    """(ntet, 10) lexicographic DOF ids on the doubled lattice."""
    pts = [verts2[:, a, :] for a in range(4)] + [(verts2[:, i, :] + verts2[:, j, :]) // 2 for i, j in _P2_EDGES]
    return np.stack([p[:, 0] + N * p[:, 1] + N * N * p[:, 2] for p in pts], 1)


def assemble_p2_3d(m, perm=None, bc=True, with_load=False):
    """P2 Poisson stiffness on the m^3 Kuhn mesh, dolfinx-shaped: pattern = DOF connectivity through cells (stored
    zeros kept), Dirichlet rows/columns zeroed in place with unit diagonal.  Returns (A, boundary mask[, load])
    where load_i = integral of phi_i (for a constant source)."""
    N = 2 * m + 1
    n = N ** 3
    h = 1.0 / m
    verts2 = _kuhn_tets_doubled(m)
    E = _p2_element_int(verts2)
    dofs = _p2_dofs(verts2, N)
    if perm is not None:
        dofs = perm[dofs]
    rows = np.repeat(dofs, 10, axis=1).ravel()
    cols = np.tile(dofs, (1, 10)).ravel()
    K = sp.coo_matrix((E.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates(); K.sort_indices()
    K.data = K.data * h / 120.0
    mi = node_multi_index(N, 3)
    bnd = np.zeros(n, dtype=bool)
    for a in mi:
        bnd |= (a == 0) | (a == N - 1)
    if perm is not None:
        b2 = np.zeros(n, dtype=bool); b2[perm] = bnd; bnd = b2
    A = K.copy()
    if bc:
        r = np.repeat(np.arange(n), np.diff(A.indptr))
        kill = bnd[r] | bnd[A.indices]
        A.data[kill] = 0.0
        A.data[(r == A.indices) & bnd[r]] = 1.0
    A.data = A.data + 0.0
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    if not with_load:
        return A, bnd
    V = h ** 3 / 6.0
    w = np.concatenate([np.full(4, -V / 20.0), np.full(6, V / 5.0)])      # integral of the P2 basis functions
    load = np.bincount(dofs.ravel(), weights=np.tile(w, dofs.shape[0]), minlength=n)
    return A, bnd, load, K


def prolongation_p2_3d(mc, perm_c=None, perm_f=None):
    """P2 -> P2 interpolation between the nested Kuhn meshes mc^3 -> (2 mc)^3: fine DOF value = coarse FE function
    evaluated at the fine DOF point.  Row entry order: the 4 vertex functions, then the 6 edge functions of the
    containing coarse tet (zeros dropped)."""
    Nc, Nf = 2 * mc + 1, 4 * mc + 1
    nf, nc = Nf ** 3, Nc ** 3
    mi = node_multi_index(Nf, 3)
    p = np.stack(mi, 1).astype(np.float64) / 4.0                      # coarse cell units
    cube = np.minimum(np.floor(p).astype(np.int64), mc - 1)
    r = p - cube
    order = np.argsort(-r, axis=1, kind="stable")                     # Kuhn tet: coordinates in descending order
    rs = np.take_along_axis(r, order, axis=1)
    lam = np.stack([1.0 - rs[:, 0], rs[:, 0] - rs[:, 1], rs[:, 1] - rs[:, 2], rs[:, 2]], 1)
    eye2 = np.eye(3, dtype=np.int64) * 2
    v = [cube * 2]
    for s in range(3):
        v.append(v[-1] + eye2[order[:, s]])
    pts = v + [(v[i] + v[j]) // 2 for i, j in _P2_EDGES]
    wts = [lam[:, a] * (2.0 * lam[:, a] - 1.0) for a in range(4)] + [4.0 * lam[:, i] * lam[:, j] for i, j in _P2_EDGES]
    cols = np.stack([q[:, 0] + Nc * q[:, 1] + Nc * Nc * q[:, 2] for q in pts], 1)
    W = np.stack(wts, 1)
    keep = W != 0.0
    counts = keep.sum(1)
    indptr = np.zeros(nf + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = cols[keep]
    data = W[keep]
    if perm_c is not None:
        indices = perm_c[indices]
    P = sp.csr_matrix((data, indices.astype(np.int32), indptr.astype(np.int32)), shape=(nf, nc))
    if perm_f is not None:
        inv = np.empty(nf, dtype=np.int64); inv[perm_f] = np.arange(nf)
        P = P[inv]
    return P


def build_hierarchy_p2(c=2, coarsest_level=0, finest_level=2, perm_seed=None, mu0=2, mu1=2, mu2=2, omega=2.0 / 3.0, seed_rhs=0):
    """3-D P2 hierarchy (BASELINE config 4: c=4, levels 0..4 -> DOF lattices 9^3 .. 129^3).  The right-hand side is a
    seeded random vector with homogeneous Dirichlet values (no manufactured solution is needed for V-cycle parity)."""
    H = Hierarchy(dim=3, coarsest_level_elements_per_dim=c, coarsest_level=coarsest_level, finest_level=finest_level,
                  mu0=mu0, mu1=mu1, mu2=mu2, omega=omega)
    for l in H.levels():
        m = c * 2 ** l
        N = 2 * m + 1
        n = N ** 3
        H.nodes_per_dim[l] = N
        H.element_size[l] = 1 / m
        perm = make_permutation(n, None if perm_seed is None else perm_seed + 1000 * l)
        H.perms[l] = perm
        A, bnd = assemble_p2_3d(m, perm)
        H.A_sp_dict[l] = (A, l)
        b = np.random.default_rng(seed_rhs + l).standard_normal(n)
        b[bnd] = 0.0
        H.b_dict[l] = b.reshape(n, 1)
    for l in range(coarsest_level, finest_level):
        H.P[l] = prolongation_p2_3d(c * 2 ** l, H.perms[l], H.perms[l + 1])
        H.inj[l] = injection(H.nodes_per_dim[l], 3, H.perms[l], H.perms[l + 1])
    return H
