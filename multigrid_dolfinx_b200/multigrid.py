"""Drop-in for the reference's ``multigrid.py`` -- same names, same signatures, same call protocol,
but every numerical operation runs on the GPU through libmgb200.so (no CPU fallback).

    from multigrid_dolfinx_b200.multigrid import (getJacobiMatrices, initialize_problem, V_cycle_scheme,
        FullMultiGrid, FullMultiGrid_test, Interpolation2D, Restriction2D_direct, Restriction2D,
        jacobiRelaxation, writing_error_for_mesh_to_csv, writing_residual_for_mesh_to_csv)

replaces ``from multigrid import ...`` (Multigrid_prototype.py:8, test/test_restriction_interpolation.py:14).

Protocol (Multigrid_prototype.py:135-143):
  1. ``getJacobiMatrices((csr, level)) -> (R_omega, D^-1, level)``  (multigrid.py:48-56).  Here the first two
     entries are lazy handles: the matrices are built on upload and fetched from the device only if
     somebody looks at them.
  2. ``initialize_problem(obj)`` with the 16 attributes of multigrid.py:30-45.  This is where the
     hierarchy is uploaded: coordinate dicts -> structured index maps -> injection list + CSR P.
  3. ``V_cycle_scheme(A_jacobi_sp_dict[l], v, f, test=False)`` (multigrid.py:231-268).

Like the reference the state is module-global: one hierarchy per process, not re-entrant.
Extra knobs the reference does not have are module attributes set BEFORE ``initialize_problem``:
``restriction`` ('injection' | 'full_weighting' | 'transpose'), ``smoother`` ('jacobi' | 'jacobi_a' |
'gs' | 'gs_color'), ``device``, ``max_fmg_cycles``.
"""
from __future__ import annotations

import csv

import numpy as np

from . import problems as _pr
from .engine import MGEngine

# ---- the reference's 16 module globals (multigrid.py:10-25) ---------------------------------------------
mesh_dof_list_dict = None
element_size = None
coarsest_level_elements_per_dim = None
coarsest_level = None
finest_level = None
A_sp_dict = None
A_jacobi_sp_dict = None
b_dict = None
mu0 = None
mu1 = None
mu2 = None
omega = None
residual_per_V_cycle_finest = None
error_per_V_cycle_finest = None
u_exact_fine = None
V_fine_dolfx = None

# ---- knobs that are ours ----------------------------------------------------------------------------
restriction = "injection"      # what the reference cycle executes (multigrid.py:251-252)
smoother = "jacobi"            # multigrid.py:223-228
device = 0
reorder = True                 # tell the engine the lattice numbering found in the coordinate dicts (mgb_set_numbering): it then
                               # works in lexicographic order internally, results unchanged bit for bit
max_fmg_cycles = 10000         # the reference's finest-level loop has no cap (multigrid.py:288)
engine_options = {}

_engine = None
_numbering = {}                # level -> (new_index[dof] = lexicographic node, perm[lexicographic node] = dof) handed to the engine
_dim = 2
_index_maps = {}               # level -> (N, lex->dof array)


class _LazyJacobiPart:
    """Stands in for R_omega / D^-1 in the tuple getJacobiMatrices returns; materialises the scipy
    matrix from the device artefacts on first use."""

    def __init__(self, level, which):
        self.level, self.which, self._m = level, which, None

    def _get(self):
        if self._m is None:
            import scipy.sparse as sp
            if _engine is None:
                raise RuntimeError("initialize_problem() has not been called yet")
            RO, dinv = _engine.rj_matrix(self.level)
            if self.level in _numbering:                  # artefacts come in the engine's (lexicographic) numbering: back to the caller's
                RO, dinv = _to_caller_numbering(RO, dinv, *_numbering[self.level])
            self._m = RO if self.which == 0 else sp.diags(dinv, 0)
        return self._m

    def dot(self, x):
        return self._get().dot(x)

    def __getattr__(self, name):
        return getattr(self._get(), name)


def _to_caller_numbering(RO, dinv, new_index, perm):
    """R_omega / D^-1 as the engine holds them (row new_index[i] = caller's row i) -> the caller's numbering; the stored entry
    order of every row is kept (it is the reference's, multigrid.py:52-55)."""
    import scipy.sparse as sp
    ip, ix, ax = RO.indptr.astype(np.int64), RO.indices, RO.data
    lens = (ip[1:] - ip[:-1])[new_index]
    ipu = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=ipu[1:])
    src = np.repeat(ip[:-1][new_index] - ipu[:-1], lens) + np.arange(ipu[-1], dtype=np.int64)
    return sp.csr_matrix((ax[src], perm[ix[src]].astype(np.int32), ipu.astype(np.int32)), shape=RO.shape), dinv[new_index]


def getJacobiMatrices(A):
    """(R_omega, D^-1, level) for ``A = (csr, level)`` (multigrid.py:48-56).  The matrices themselves are
    built inside ``mgb_finalize`` when ``initialize_problem`` uploads the hierarchy."""
    level = A[1]
    return (_LazyJacobiPart(level, 0), _LazyJacobiPart(level, 1), level)


def _index_map_from_dict(d, n, h, dim):
    """coordinate dict (Multigrid_prototype.py:68-74) -> (nodes per dim, lexicographic node -> dof).
    Uses integer lattice coordinates rint(x / h) instead of rounded-float dictionary keys, so it does
    not inherit the h >= 1/512 limit of the reference."""
    X = np.array([d[i][:dim] for i in range(n)], dtype=np.float64)
    I = np.rint(X / h).astype(np.int64)
    N = int(I.max()) + 1
    if N ** dim != n:
        raise KeyError(f"coordinate dict does not describe a full {N}^{dim} grid ({n} dofs)")
    lex = _pr.lex_index([I[:, k] for k in range(dim)], N)
    perm = np.empty(n, dtype=np.int64)
    perm[lex] = np.arange(n)
    return N, perm


def initialize_problem(obj):
    """multigrid.py:28-45, plus the upload of the whole hierarchy to the device."""
    global mesh_dof_list_dict, element_size, coarsest_level_elements_per_dim, coarsest_level, finest_level, A_sp_dict, \
        A_jacobi_sp_dict, b_dict, mu0, mu1, mu2, omega, residual_per_V_cycle_finest, error_per_V_cycle_finest, \
        u_exact_fine, V_fine_dolfx, _engine, _dim, _index_maps, _numbering
    mesh_dof_list_dict = obj.mesh_dof_list_dict
    element_size = obj.element_size
    coarsest_level_elements_per_dim = obj.coarsest_level_elements_per_dim
    coarsest_level = obj.coarsest_level
    finest_level = obj.finest_level
    A_sp_dict = obj.A_sp_dict
    A_jacobi_sp_dict = obj.A_jacobi_sp_dict
    b_dict = obj.b_dict
    mu0 = obj.mu0
    mu1 = obj.mu1
    mu2 = obj.mu2
    omega = obj.omega
    residual_per_V_cycle_finest = obj.residual_per_V_cycle_finest
    error_per_V_cycle_finest = obj.error_per_V_cycle_finest
    u_exact_fine = obj.u_exact_fine
    V_fine_dolfx = obj.V_fine_dolfx

    if _engine is not None:
        _engine.close()
        _engine = None
    _dim = getattr(obj, "dim", 2)
    eng = MGEngine(device)
    for k, v in engine_options.items():
        eng.set_option(k, v)
    levels = list(range(coarsest_level, finest_level + 1))
    for l in levels:
        eng.set_level(l, A_sp_dict[l][0])
    _index_maps = {}
    _numbering = {}
    if reorder and mesh_dof_list_dict is not None and all(l in mesh_dof_list_dict for l in levels):
        # dolfinx's DOF numbering is not lexicographic; the coordinate dicts (Multigrid_prototype.py:68-74) say where every DOF
        # sits on the lattice, so the engine is told to work in lexicographic order internally (bit-identical row sums: the entry
        # order of every row is kept) -- that is what lets the one-byte-per-row operator codings apply to the reference's input
        for l in levels:
            _index_maps[l] = _index_map_from_dict(mesh_dof_list_dict[l], A_sp_dict[l][0].shape[0], element_size[l], _dim)
            perm = _index_maps[l][1]                                  # perm[lex] = dof
            new_index = np.empty_like(perm)
            new_index[perm] = np.arange(len(perm), dtype=np.int64)
            eng.set_numbering(l, new_index)
            _numbering[l] = (new_index, perm)
    have_matrix_form = hasattr(obj, "P") and hasattr(obj, "inj") and all(l in obj.P for l in levels[:-1])
    for l in levels[:-1]:
        if have_matrix_form:
            P, inj = obj.P[l], obj.inj[l]
        else:
            for q in (l, l + 1):
                if q not in _index_maps:
                    _index_maps[q] = _index_map_from_dict(mesh_dof_list_dict[q], A_sp_dict[q][0].shape[0], element_size[q], _dim)
            (Nc, pc), (Nf, pf) = _index_maps[l], _index_maps[l + 1]
            if Nf != 2 * Nc - 1:
                raise KeyError(f"levels {l} and {l + 1} are not nested by a factor of two")
            P = _pr.prolongation(Nc, _dim, pc, pf)
            inj = _pr.injection(Nc, _dim, pc, pf)
        eng.set_transfer(l, P, r_mode=restriction, inj=inj if restriction == "injection" else None, dim=_dim)
    eng.set_params(omega, mu1, mu2, smoother)
    eng.finalize()
    _engine = eng


def engine():
    """The module's engine (after initialize_problem)."""
    if _engine is None:
        raise RuntimeError("initialize_problem() has not been called yet")
    return _engine


# ---- transfer operators with the reference's signatures ---------------------------------------------

_standalone = {}


def _level_pair_engine(mesh_dict_coarse, mesh_dict_fine, n_c, n_f, h_c=None, h_f=None, standalone=False):
    """Engine able to apply the transfer pair described by two coordinate dicts.  If they are the dicts of
    the initialised hierarchy the main engine is used; otherwise (test/test_restriction_interpolation.py
    calls the transfer functions without initialize_problem) a two-level engine with identity level
    matrices is built once per dict pair."""
    if _engine is not None and mesh_dof_list_dict is not None and not standalone:
        for l in range(coarsest_level, finest_level):
            if mesh_dof_list_dict.get(l) is mesh_dict_coarse and mesh_dof_list_dict.get(l + 1) is mesh_dict_fine:
                return _engine, l + 1
    key = (id(mesh_dict_coarse), id(mesh_dict_fine), restriction)
    if key not in _standalone:
        import scipy.sparse as sp
        dim = 2
        Nc = int(round(n_c ** (1.0 / dim))); Nf = int(round(n_f ** (1.0 / dim)))
        hc = h_c if h_c is not None else 1.0 / (Nc - 1)
        hf = h_f if h_f is not None else 1.0 / (Nf - 1)
        _, pc = _index_map_from_dict(mesh_dict_coarse, n_c, hc, dim)
        _, pf = _index_map_from_dict(mesh_dict_fine, n_f, hf, dim)
        eng = MGEngine(device)
        eng.set_level(0, sp.identity(n_c, format="csr"))
        eng.set_level(1, sp.identity(n_f, format="csr"))
        P = _pr.prolongation(Nc, dim, pc, pf)
        eng.set_transfer(0, P, r_mode=restriction, inj=_pr.injection(Nc, dim, pc, pf) if restriction == "injection" else None, dim=dim)
        eng.finalize()
        _standalone[key] = (eng, mesh_dict_coarse, mesh_dict_fine)     # the dicts are kept alive: the key holds their id()s
    return _standalone[key][0], 1


def Interpolation2D(vec_2h, mesh_dict_coarse, mesh_dict_fine, element_size_coarse, element_size_fine, vec_h_dim):
    """multigrid.py:59-120: bilinear interpolation of a coarse vector, returned as (vec_h_dim, 1)."""
    vec_2h = np.asarray(vec_2h, dtype=np.float64)
    eng, lf = _level_pair_engine(mesh_dict_coarse, mesh_dict_fine, vec_2h.size, vec_h_dim, element_size_coarse, element_size_fine)
    out = eng.prolong_add(lf, vec_2h.reshape(-1), np.zeros(vec_h_dim))
    return out.reshape(vec_h_dim, 1)


def Restriction2D_direct(vec_h, mesh_dict_coarse, mesh_dict_fine, vec_2h_dim):
    """multigrid.py:123-132: injection."""
    global restriction
    vec_h = np.asarray(vec_h, dtype=np.float64)
    saved = restriction
    restriction = "injection"
    try:
        # the main engine can only be used if it was built with this restriction
        eng, lf = _level_pair_engine(mesh_dict_coarse, mesh_dict_fine, vec_2h_dim, vec_h.size, standalone=saved != "injection")
    finally:
        restriction = saved
    return eng.restrict(lf, vec_h.reshape(-1)).reshape(vec_2h_dim, 1)


def Restriction2D(vec_h, mesh_dict_coarse, mesh_dict_fine, element_size_coarse, element_size_fine, vec_2h_dim):
    """multigrid.py:135-198: full weighting (= 1/4 P^T)."""
    global restriction
    vec_h = np.asarray(vec_h, dtype=np.float64)
    saved = restriction
    restriction = "full_weighting"
    try:
        eng, lf = _level_pair_engine(mesh_dict_coarse, mesh_dict_fine, vec_2h_dim, vec_h.size, element_size_coarse,
                                     element_size_fine, standalone=saved != "full_weighting")
    finally:
        restriction = saved
    return eng.restrict(lf, vec_h.reshape(-1)).reshape(vec_2h_dim, 1)


# ---- norms (multigrid.py:203-218) ---------------------------------------------------------------------

def res_calculator(res, V_space):
    """The reference assembles the L2(Omega) norm with dolfinx (multigrid.py:203-208).  Without dolfinx the
    caller may pass a mass matrix (scipy CSR) as ``V_space``: sqrt(r^T M r); with ``None`` the l2 norm."""
    r = np.asarray(res, dtype=np.float64).reshape(-1)
    if V_space is None:
        return engine().norm2(r)
    if hasattr(V_space, "dot"):
        return float(np.sqrt(r @ V_space.dot(r)))
    raise RuntimeError("res_calculator needs dolfinx in the reference; pass a mass matrix or None here")


def err_calculator(u, u_exact, V_space):
    """multigrid.py:213-218 with the same substitution as res_calculator; ``u_exact`` is a nodal vector here."""
    return res_calculator(np.asarray(u).reshape(-1) - np.asarray(u_exact).reshape(-1), V_space)


# ---- smoother and cycle ------------------------------------------------------------------------------

def jacobiRelaxation(A, v, f, nw):
    """multigrid.py:223-228 on the device; ``A`` is the tuple from getJacobiMatrices (level = A[2])."""
    if nw == 0:
        return v
    return engine().smooth(A[2], v, f, nw)


def V_cycle_scheme(A_h, v_h, f_h, test=False):
    """multigrid.py:231-268.  Returns a fresh (n, 1) array, or the 4-tuple when ``test`` and the level is
    the finest (multigrid.py:262-266)."""
    eng = engine()
    current_level = A_h[2]
    n = eng.n[current_level]
    v_in = np.zeros((n, 1)) if v_h is None else v_h
    if test and current_level == finest_level and current_level != coarsest_level:
        v, f2h, v2h, errh = eng.vcycle_debug(current_level, np.asarray(v_in).reshape(n, 1), f_h)
        return v, f2h, v2h, errh
    out = eng.vcycle(current_level, v_in, f_h, 1)
    return out if _is_device(out) else np.asarray(out).reshape(n, 1)


def _is_device(x):
    return type(x).__module__.startswith("torch")


def FullMultiGrid(A_h, f_h):
    """multigrid.py:271-307: nested iteration; at the finest level V-cycles until the residual norm is <= 1e-11
    (capped at ``max_fmg_cycles``; the reference has no cap).  Called with the finest level (the only way the
    reference calls it, Multigrid_prototype.py:148) the whole driver runs on the device (``mgb_fmg``): only one scalar
    per cycle comes back.  Norm: sqrt(r^T M r) when ``V_fine_dolfx`` is a mass matrix (the reference's L2(Omega)
    norm, multigrid.py:203-208), else the l2 norm.  Called with an intermediate level it recurses like the reference."""
    current_level = A_h[2]
    eng = engine()
    if current_level == coarsest_level:
        return np.asarray(eng.coarse_solve(np.asarray(f_h, dtype=np.float64).reshape(-1))).reshape(-1, 1)
    n = eng.n[current_level]
    if current_level == finest_level:
        for l in range(coarsest_level, finest_level):
            eng.set_rhs(l, np.asarray(b_dict[l], dtype=np.float64).reshape(-1))
        eng.set_rhs(finest_level, np.asarray(f_h, dtype=np.float64).reshape(-1))
        # the reference measures both norms with dolfinx (multigrid.py:203-218); here V_fine_dolfx is a CSR mass matrix
        # (-> sqrt(r^T M r), the same L2(Omega) norm) or None (-> l2 norm), u_exact_fine a nodal vector or None.  Anything
        # else (a live dolfinx FunctionSpace / Function) is refused BEFORE the solve instead of silently changing the norm
        # the 1e-11 stopping rule is measured in.
        if V_fine_dolfx is not None and not (hasattr(V_fine_dolfx, "dot") and hasattr(V_fine_dolfx, "indptr")):
            raise TypeError("FullMultiGrid: V_fine_dolfx must be a scipy CSR mass matrix (L2(Omega) norm, multigrid.py:203-208) or None "
                            f"(l2 norm); got {type(V_fine_dolfx).__name__}.  Export the mass matrix of the dolfinx FunctionSpace instead.")
        u_ex = None
        if u_exact_fine is not None:
            try:
                u_ex = np.asarray(u_exact_fine, dtype=np.float64).reshape(-1)
            except (TypeError, ValueError):
                u_ex = None
            if u_ex is None or u_ex.size != n:
                raise TypeError(f"FullMultiGrid: u_exact_fine must be the nodal values of the exact solution ({n} entries) or None; "
                                f"got {type(u_exact_fine).__name__}.  Pass u_exact.x.array of the dolfinx Function.")
        if V_fine_dolfx is not None:
            eng.set_mass_matrix(finest_level, V_fine_dolfx)
        eng.set_exact_solution(finest_level, u_ex)
        v, hist = eng.fmg(mu0, 1E-11, max_fmg_cycles)
        residual_per_V_cycle_finest.extend(float(x) for x in hist)                 # multigrid.py:294-295, one entry per cycle
        error_per_V_cycle_finest.extend(float(x) for x in eng.fmg_errors())        # multigrid.py:292-293, one entry per cycle
        with open(f'iter_count_for_diff_num_elems_{finest_level - coarsest_level + 1}_levels.csv', mode='a') as file1:
            csv.writer(file1, delimiter=',').writerow([coarsest_level_elements_per_dim * 2 ** finest_level, len(hist)])
        return v.reshape(n, 1)
    f_2h = b_dict[current_level - 1]
    v_2h = FullMultiGrid(A_jacobi_sp_dict[current_level - 1], f_2h)
    v_h = eng.prolong_add(current_level, v_2h.reshape(-1), np.zeros(n)).reshape(n, 1)
    for _ in range(mu0):
        v_h = V_cycle_scheme(A_h, v_h, f_h)
    return v_h


def FullMultiGrid_test(A_h, f_h, test=False):
    """multigrid.py:312-339 (the driver the prototype actually runs, Multigrid_prototype.py:141-143)."""
    current_level = A_h[2]
    eng = engine()
    dbg = (None, None, None)
    if current_level == coarsest_level:
        return np.asarray(eng.coarse_solve(np.asarray(f_h, dtype=np.float64).reshape(-1))).reshape(-1, 1)
    f_2h = b_dict[current_level - 1]
    v_2h = FullMultiGrid_test(A_jacobi_sp_dict[current_level - 1], f_2h, test)
    n = eng.n[current_level]
    v_h = eng.prolong_add(current_level, np.asarray(v_2h).reshape(-1), np.zeros(n)).reshape(n, 1)
    for _ in range(mu0):
        if current_level == finest_level:
            if not test:      # the reference unpacks a 4-tuple here even when test is False (multigrid.py:331-333)
                raise ValueError("not enough values to unpack (expected 4, got 1)")
            v_h, *dbg = V_cycle_scheme(A_h, v_h, f_h, True)
        else:
            v_h = V_cycle_scheme(A_h, v_h, f_h)
    if test and current_level == finest_level:
        return (v_h, *dbg)
    return v_h


# ---- CSV writers (multigrid.py:345-356), unchanged behaviour ------------------------------------------------

def writing_residual_for_mesh_to_csv(residual):
    with open(f'residual_for_{coarsest_level_elements_per_dim * 2 ** finest_level}_{finest_level - coarsest_level + 1}_levels.csv', mode='w') as file:
        w = csv.writer(file, delimiter=',')
        for i in range(0, len(residual)):
            w.writerow([i, residual[i]])


def writing_error_for_mesh_to_csv(error):
    with open(f'error_for_{coarsest_level_elements_per_dim * 2 ** finest_level}_{finest_level - coarsest_level + 1}_levels.csv', mode='w') as file:
        w = csv.writer(file, delimiter=',')
        for i in range(0, len(error)):
            w.writerow([i, error[i]])
