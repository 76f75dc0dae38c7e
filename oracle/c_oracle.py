"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/mg_oracle.c (the plain-C CPU oracle).

Builds ``oracle/_build/libmgoracle.so`` on demand with ``make -C oracle``.  Never imported by the
product package.  See mg_oracle.c for the reference file:line each routine follows."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return os.path.join(_HERE, "_build", "libmgoracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libmgoracle.so")
        src = os.path.join(_HERE, "mg_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.orc_csr_matvec.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p]
        L.orc_residual.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p, f64p]
        L.orc_jacobi_sweep.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p, f64p, f64p, C.c_double]
        L.orc_jacobi_sweep_aform.argtypes = L.orc_jacobi_sweep.argtypes
        L.orc_gather.argtypes = [C.c_int64, i32p, f64p, f64p]
        L.orc_prolong_add.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p, C.c_void_p]
        L.orc_gs_forward.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p, C.c_void_p]
        L.orc_norm2.argtypes = [C.c_int64, f64p]; L.orc_norm2.restype = C.c_double
        L.orc_level_sets.argtypes = [C.c_int64, i64p, i32p, f64p, i32p]; L.orc_level_sets.restype = C.c_int64
        L.orc_greedy_colouring.argtypes = [C.c_int64, i64p, i32p, f64p, i32p]; L.orc_greedy_colouring.restype = C.c_int64
        L.orc_dense_lu.argtypes = [C.c_int64, f64p, i32p]; L.orc_dense_lu.restype = C.c_int
        L.orc_dense_lu_solve.argtypes = [C.c_int64, f64p, i32p, f64p]
        L.orc_mg_create.argtypes = [C.c_int]; L.orc_mg_create.restype = C.c_void_p
        L.orc_mg_set_params.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_mg_set_level.argtypes = [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_void_p] * 4 + \
            [C.c_void_p] * 3 + [C.c_void_p, C.c_int64] + [C.c_void_p] * 3 + [C.c_void_p]
        L.orc_mg_finalize.argtypes = [C.c_void_p]; L.orc_mg_finalize.restype = C.c_int
        L.orc_mg_destroy.argtypes = [C.c_void_p]
        L.orc_mg_vcycle.argtypes = [C.c_void_p, C.c_int, f64p, f64p, C.c_int, C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]; L.orc_set_threads.restype = C.c_int
        L.orc_mg_build_poisson.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_mg_build_poisson.restype = C.c_void_p
        L.orc_mg_level_array.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]; L.orc_mg_level_array.restype = C.c_int64
        _LIB = L
    return _LIB


def _csr(A):
    return (np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_threads(n=0):
    """OpenMP threads of the C oracle: n > 0 sets them (``torchrun`` exports OMP_NUM_THREADS=1, which would otherwise
    silently make the "all cores" baseline single-threaded); returns the number in effect."""
    return int(lib().orc_set_threads(int(n)))


class StructuredCOracleMG:
    """The synthetic P1 Poisson hierarchy (lexicographic DOFs, injection) built INSIDE the C oracle by index arithmetic --
    the same arrays as problems.stencil_p1 / prolongation / injection + the reference's getJacobiMatrices, but without a
    scipy matrix ever existing, so that the 513^3 configuration (2.0e9 stored entries, ~50 GB) fits a host run."""
    _KIND = {0: np.int64, 1: np.int32, 2: np.float64, 3: np.int64, 4: np.int32, 5: np.float64, 6: np.float64,
             7: np.int64, 8: np.int32, 9: np.float64, 10: np.int32}

    def __init__(self, dim, c, coarsest_level, finest_level, omega=2.0 / 3.0, mu1=2, mu2=2):
        self.h = lib().orc_mg_build_poisson(int(dim), int(c), int(coarsest_level), int(finest_level), float(omega), int(mu1), int(mu2))
        if not self.h:
            raise RuntimeError("orc_mg_build_poisson failed")
        self.dim, self.c, self.levels = dim, c, list(range(coarsest_level, finest_level + 1))
        self.n = [(c * 2 ** l + 1) ** dim for l in self.levels]
        self.mu1, self.mu2 = mu1, mu2

    def array(self, k, what):
        """Level k (0 = coarsest) array ``what`` (see orc_mg_level_array) as a numpy VIEW."""
        p = C.c_void_p()
        cnt = lib().orc_mg_level_array(self.h, int(k), int(what), C.byref(p))
        if cnt <= 0 or not p.value:
            return np.zeros(0, dtype=self._KIND[what])
        dt = np.dtype(self._KIND[what])
        buf = (C.c_char * (cnt * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=cnt)

    def vcycle(self, v, f, ncycles=1, history=False):
        v = np.ascontiguousarray(v, dtype=np.float64).ravel().copy()
        f = np.ascontiguousarray(f, dtype=np.float64).ravel()
        hist = np.zeros(ncycles) if history else None
        lib().orc_mg_vcycle(self.h, len(self.n) - 1, v, f, ncycles, _ptr(hist))
        return (v, hist) if history else v

    def dof_updates_per_cycle(self):
        return (self.mu1 + self.mu2) * sum(self.n[1:])

    def close(self):
        if self.h:
            lib().orc_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def csr_matvec(A, x):
    ip, ix, ax = _csr(A)
    y = np.empty(A.shape[0])
    lib().orc_csr_matvec(A.shape[0], ip, ix, ax, np.ascontiguousarray(x, dtype=np.float64).ravel(), y)
    return y


def residual(A, f, v):
    ip, ix, ax = _csr(A)
    r = np.empty(A.shape[0])
    lib().orc_residual(A.shape[0], ip, ix, ax, np.ascontiguousarray(f).ravel(), np.ascontiguousarray(v).ravel(), r)
    return r


def jacobi(RO, dinv, f, v, omega, nw, aform=False):
    ip, ix, ax = _csr(RO)
    v = np.ascontiguousarray(v, dtype=np.float64).ravel().copy()
    f = np.ascontiguousarray(f, dtype=np.float64).ravel()
    dinv = np.ascontiguousarray(dinv, dtype=np.float64).ravel()
    out = np.empty_like(v)
    fn = lib().orc_jacobi_sweep_aform if aform else lib().orc_jacobi_sweep
    for _ in range(nw):
        fn(len(v), ip, ix, ax, dinv, f, v, out, omega)
        v, out = out, v
    return v


def gs_forward(A, v, f, order=None):
    ip, ix, ax = _csr(A)
    v = np.ascontiguousarray(v, dtype=np.float64).ravel().copy()
    o = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    lib().orc_gs_forward(A.shape[0], ip, ix, ax, np.ascontiguousarray(f, dtype=np.float64).ravel(), v, _ptr(o))
    return v


def level_sets(A):
    ip, ix, ax = _csr(A)
    lev = np.zeros(A.shape[0], dtype=np.int32)
    nl = lib().orc_level_sets(A.shape[0], ip, ix, ax, lev)
    order = np.argsort(lev, kind="stable").astype(np.int32)
    offsets = np.zeros(nl + 1, dtype=np.int32)
    np.cumsum(np.bincount(lev, minlength=nl), out=offsets[1:])
    return lev, order, offsets


def greedy_colouring(A):
    ip, ix, ax = _csr(A)
    col = np.zeros(A.shape[0], dtype=np.int32)
    nc = lib().orc_greedy_colouring(A.shape[0], ip, ix, ax, col)
    order = np.argsort(col, kind="stable").astype(np.int32)
    offsets = np.zeros(nc + 1, dtype=np.int32)
    np.cumsum(np.bincount(col, minlength=nc), out=offsets[1:])
    return col, order, offsets


class COracleMG:
    """Whole V-cycle in C (OpenMP over rows).  Levels ordered coarsest..finest.

    ``A``: list of raw CSR; ``RO``/``dinv``: Jacobi matrices per level (multigrid.py:48-56);
    ``P[k]``: prolongation level k-1 -> k (k >= 1); ``inj[k]``: injection list from level k to k-1;
    ``R[k]`` explicit restriction CSR or None.  smoother: 'jacobi' | 'jacobi_a' | 'gs' | 'gs_color'."""

    def __init__(self, A, RO, dinv, P, inj=None, R=None, omega=2.0 / 3.0, mu1=2, mu2=2, smoother="jacobi"):
        L = lib()
        nlev = len(A)
        self._keep = []
        self.h = L.orc_mg_create(nlev)
        self.n = [a.shape[0] for a in A]
        sm = {"jacobi": 0, "jacobi_a": 1, "gs": 2, "gs_color": 2}[smoother]
        L.orc_mg_set_params(self.h, omega, mu1, mu2, sm)
        for k in range(nlev):
            a = _csr(A[k]); r = _csr(RO[k]); d = np.ascontiguousarray(dinv[k], dtype=np.float64).ravel()
            p = _csr(P[k]) if k > 0 else (None, None, None)
            ij = np.ascontiguousarray(inj[k], dtype=np.int32) if (k > 0 and inj is not None and inj[k] is not None) else None
            rs = _csr(R[k]) if (k > 0 and R is not None and R[k] is not None) else (None, None, None)
            go = greedy_colouring(A[k])[1] if smoother == "gs_color" else None
            self._keep.append((a, r, d, p, ij, rs, go))
            L.orc_mg_set_level(self.h, k, A[k].shape[0], *[_ptr(x) for x in a], *[_ptr(x) for x in r], _ptr(d),
                               *[_ptr(x) for x in p], _ptr(ij), A[k - 1].shape[0] if k > 0 else 0,
                               *[_ptr(x) for x in rs], _ptr(go))
        rc = L.orc_mg_finalize(self.h)
        if rc != 0:
            raise RuntimeError(f"coarsest matrix singular ({rc})")

    def vcycle(self, v, f, ncycles=1, top=None, history=False):
        top = len(self.n) - 1 if top is None else top
        v = np.ascontiguousarray(v, dtype=np.float64).ravel().copy()
        f = np.ascontiguousarray(f, dtype=np.float64).ravel()
        hist = np.zeros(ncycles) if history else None
        lib().orc_mg_vcycle(self.h, top, v, f, ncycles, _ptr(hist))
        return (v, hist) if history else v

    def __del__(self):
        try:
            lib().orc_mg_destroy(self.h)
        except Exception:
            pass


def from_hierarchy(H, r_mode="injection", smoother="jacobi", jac=None):
    """Build a COracleMG from a problems.Hierarchy (R_omega built with the reference's scipy calls)."""
    from . import restated as rs
    lv = list(H.levels())
    A = [H.A_sp_dict[l][0] for l in lv]
    jm = [rs.jacobi_matrices(a) for a in A] if jac is None else jac
    P = [None] + [H.P[l] for l in lv[:-1]]
    inj = [None] + [H.inj[l] for l in lv[:-1]]
    R = None
    if r_mode == "full_weighting":
        R = [None] + [_sorted((H.P[l].T * (0.5 ** H.dim)).tocsr()) for l in lv[:-1]]
    elif r_mode == "transpose":
        R = [None] + [_sorted(H.P[l].T.tocsr()) for l in lv[:-1]]
    return COracleMG(A, [j[0] for j in jm], [j[1] for j in jm], P, inj if R is None else None, R,
                     omega=H.omega, mu1=H.mu1, mu2=H.mu2, smoother=smoother)


def _sorted(M):
    M.sort_indices()
    return M
