"""Python face of the C ABI: one ``MGEngine`` = one ``mgb_handle`` (one device, one stream).

Everything numerical happens in libmgb200.so; this class only marshals pointers.  Vectors may be
numpy arrays (host memory: staged by the library inside the call) or torch CUDA float64 tensors
(device memory: borrowed, no copy through the host)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _as_csr_arrays(M):
    indptr = np.ascontiguousarray(M.indptr)
    if indptr.dtype not in (np.int32, np.int64):
        indptr = indptr.astype(np.int64)
    indices = np.ascontiguousarray(M.indices, dtype=np.int32)
    data = np.ascontiguousarray(M.data, dtype=np.float64)
    return indptr, indices, data


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class MGEngine:
    """Device-resident level hierarchy + V-cycle.  Mirrors the protocol of the reference
    (getJacobiMatrices -> initialize_problem -> V_cycle_scheme, Multigrid_prototype.py:135-143)."""

    def __init__(self, device=0):
        self._lib = L.load()
        self._h = C.c_void_p()
        rc = self._lib.mgb_create(C.byref(self._h), int(device))
        if rc != L.OK:
            msg = self._lib.mgb_last_error(None).decode()
            self._h = None
            raise L.MGBError(rc, msg)
        self.device = int(device)
        self.n = {}
        self.finalized = False
        self._ext_stream = None

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != L.OK:
            raise L.MGBError(rc, self._lib.mgb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mgb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream_ptr(self):
        p = C.c_void_p()
        self._ck(self._lib.mgb_get_stream(self._h, C.byref(p)))
        return p.value

    def torch_stream(self):
        """The engine's stream as a torch.cuda.ExternalStream (for events / ordering)."""
        import torch
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.stream_ptr(), device=self.device)
        return self._ext_stream

    def synchronize(self):
        self._ck(self._lib.mgb_synchronize(self._h))

    # -- hierarchy upload ---------------------------------------------------------------------------
    def set_level(self, level, A):
        """A_sp_dict[level][0] as exported from PETSc (Multigrid_prototype.py:95-99)."""
        ip, ix, ax = _as_csr_arrays(A)
        n = A.shape[0]
        self._ck(self._lib.mgb_set_level(self._h, int(level), n, len(ax), ip.ctypes.data, ip.dtype.itemsize,
                                         ix.ctypes.data, ax.ctypes.data))
        self.n[int(level)] = n

    def halo_fused(self, level):
        """Whether the kernels of a row-sharded ``level`` exchange the ghost rows themselves (``mgb_halo_fused``)."""
        out = C.c_int()
        self._ck(self._lib.mgb_halo_fused(self._h, int(level), C.byref(out)))
        return bool(out.value)

    def set_halo_fused(self, level, on):
        self._ck(self._lib.mgb_set_halo_fused(self._h, int(level), int(bool(on))))

    def set_numbering(self, level, new_index):
        """Caller numbering (``mgb_set_numbering``): ``new_index[i]`` = position of dof ``i`` in the numbering the engine should
        work in -- the lexicographic lattice index for the uniform meshes of the reference, which is what makes the lossless
        row-pattern codings apply to dolfinx-ordered input.  Operators are renumbered once at ``finalize`` with the entry order of
        every row kept (bit-identical row sums); vectors are permuted on the device on the way in and out."""
        p = np.ascontiguousarray(new_index, dtype=np.int64).reshape(-1)
        self._ck(self._lib.mgb_set_numbering(self._h, int(level), len(p), p.ctypes.data))

    # -- row-sharded hierarchies (one engine per rank; see dist.py) -----------------------------------------
    def dist_unique_id(self):
        buf = C.create_string_buffer(256)
        rc = self._lib.mgb_dist_unique_id(buf, 256)
        if rc != L.OK:
            raise L.MGBError(rc, self._lib.mgb_last_error(None).decode())
        return bytes(buf.raw)

    def dist_init(self, rank, world, unique_id):
        self._ck(self._lib.mgb_dist_init(self._h, int(rank), int(world), unique_id, len(unique_id)))

    def set_level_local(self, level, A_local, n_owned, n_ghost):
        ip, ix, ax = _as_csr_arrays(A_local)
        self._ck(self._lib.mgb_set_level_local(self._h, int(level), int(n_owned), int(n_ghost), len(ax), ip.ctypes.data,
                                               ip.dtype.itemsize, ix.ctypes.data, ax.ctypes.data))
        self.n[int(level)] = int(n_owned)

    def set_halo(self, level, peers, send_idx, recv_cnt):
        peers_a = np.ascontiguousarray(peers, dtype=np.int32)
        sc = np.ascontiguousarray([len(x) for x in send_idx], dtype=np.int32)
        si = np.ascontiguousarray(np.concatenate(send_idx) if len(send_idx) else np.zeros(0), dtype=np.int32)
        rc = np.ascontiguousarray(recv_cnt, dtype=np.int32)
        self._ck(self._lib.mgb_set_halo(self._h, int(level), len(peers_a), peers_a.ctypes.data, sc.ctypes.data, si.ctypes.data, rc.ctypes.data))

    def p2p_export(self, level):
        size = C.c_int()
        self._ck(self._lib.mgb_p2p_export(self._h, int(level), None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        self._ck(self._lib.mgb_p2p_export(self._h, int(level), buf, size.value, C.byref(size)))
        return bytes(buf.raw)

    def p2p_import(self, level, peer_rank, blob):
        self._ck(self._lib.mgb_p2p_import(self._h, int(level), int(peer_rank), blob, len(blob)))

    def set_gather_level(self, level, n_global, offsets):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._ck(self._lib.mgb_set_gather_level(self._h, int(level), int(n_global), off.ctypes.data))
        self.n.setdefault(int(level), int(n_global))

    # -- device-side generation of the synthetic structured hierarchy (problems.stencil_p1 / prolongation / injection) --
    def synth_level(self, level, dim, cells_per_dim, row_begin, row_end, ghost_lo=None, ghost_hi=None):
        glo = row_begin if ghost_lo is None else ghost_lo
        ghi = row_end if ghost_hi is None else ghost_hi
        self._ck(self._lib.mgb_synth_poisson_level(self._h, int(level), int(dim), int(cells_per_dim), int(row_begin), int(row_end), int(glo), int(ghi)))
        self.n[int(level)] = int(row_end - row_begin)

    def synth_transfer(self, coarse_level, inj_coarse_begin, inj_coarse_end):
        self._ck(self._lib.mgb_synth_poisson_transfer(self._h, int(coarse_level), int(inj_coarse_begin), int(inj_coarse_end)))

    def set_restriction(self, coarse_level, r_mode, dim=2):
        """Restriction of a transfer that is already set or generated (``mgb_set_restriction``): injection, full weighting
        (2^-dim P^T, multigrid.py:135-198) or the plain transpose; the transposed operator is formed at ``finalize``."""
        mode = L.R_MODES[r_mode] if isinstance(r_mode, str) else int(r_mode)
        self._ck(self._lib.mgb_set_restriction(self._h, int(coarse_level), mode, int(dim)))

    def set_transfer(self, coarse_level, P, r_mode="injection", inj=None, R=None, dim=2, n_fine=None, n_coarse_rows=None):
        pip, pix, pax = _as_csr_arrays(P)
        mode = L.R_MODES[r_mode] if isinstance(r_mode, str) else int(r_mode)
        injp = None
        if inj is not None:
            inj = np.ascontiguousarray(inj, dtype=np.int32)
            injp = inj.ctypes.data
        rargs = (0, None, 8, None, None)
        if R is not None:
            rip, rix, rax = _as_csr_arrays(R)
            rargs = (len(rax), rip.ctypes.data, rip.dtype.itemsize, rix.ctypes.data, rax.ctypes.data)
        nf = P.shape[0] if n_fine is None else int(n_fine)
        ncr = P.shape[1] if n_coarse_rows is None else int(n_coarse_rows)
        self._ck(self._lib.mgb_set_transfer(self._h, int(coarse_level), nf, ncr, len(pax), pip.ctypes.data,
                                            pip.dtype.itemsize, pix.ctypes.data, pax.ctypes.data, mode, int(dim), injp, *rargs))

    def set_params(self, omega=2.0 / 3.0, mu1=2, mu2=2, smoother="jacobi"):
        sm = L.SMOOTHERS[smoother] if isinstance(smoother, str) else int(smoother)
        self._ck(self._lib.mgb_set_params(self._h, float(omega), int(mu1), int(mu2), sm))

    def set_option(self, key, value):
        self._ck(self._lib.mgb_set_option(self._h, key.encode(), float(value)))

    def precheck(self):
        self._ck(self._lib.mgb_precheck(self._h))

    def finalize(self):
        self._ck(self._lib.mgb_finalize(self._h))
        self.finalized = True

    @classmethod
    def from_hierarchy(cls, H, r_mode="injection", smoother="jacobi", device=0, options=None, levels=None, reorder=False):
        """Upload a ``problems.Hierarchy`` (or any object with A_sp_dict / P / inj / mu1 / mu2 / omega).
        ``reorder``: hand the engine the lattice numbering of every level (``H.perms[l][lexicographic node] = dof``, the
        information the reference keeps as coordinate dicts) so that it works in lexicographic order internally."""
        eng = cls(device)
        lv = list(H.levels()) if levels is None else list(levels)
        for k, v in (options or {}).items():
            eng.set_option(k, v)
        for l in lv:
            eng.set_level(l, H.A_sp_dict[l][0])
            if reorder:
                perm = np.asarray(H.perms[l], dtype=np.int64)            # perm[lex] = dof
                new_index = np.empty_like(perm)
                new_index[perm] = np.arange(len(perm), dtype=np.int64)   # new_index[dof] = lex
                eng.set_numbering(l, new_index)
        for l in lv[:-1]:
            eng.set_transfer(l, H.P[l], r_mode=r_mode, inj=H.inj[l] if r_mode == "injection" else None, dim=H.dim)
        eng.set_params(H.omega, H.mu1, H.mu2, smoother)
        eng.finalize()
        return eng

    # -- vector marshalling -------------------------------------------------------------------------
    def _in(self, x, n, name):
        """-> (pointer, mem kind, keepalive object)."""
        if _is_torch(x):
            import torch
            if not x.is_cuda or x.dtype != torch.float64 or x.device.index != self.device:
                raise ValueError(f"{name}: torch tensors must be float64 on cuda:{self.device}")
            t = x.contiguous().view(-1)
            if t.numel() != n:
                raise ValueError(f"{name}: expected {n} entries, got {t.numel()}")
            return t.data_ptr(), L.MEM_DEVICE, t
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if a.size != n:
            raise ValueError(f"{name}: expected {n} entries, got {a.size}")
        return a.ctypes.data, L.MEM_HOST, a

    def _out_like(self, x, n):
        if _is_torch(x):
            import torch
            t = torch.empty(n, dtype=torch.float64, device=x.device)
            return t.data_ptr(), t
        a = np.empty(n, dtype=np.float64)
        return a.ctypes.data, a

    def _fence_in(self, *xs):
        """Make the engine stream wait for torch's current stream when device tensors are passed.  Must be called AFTER all
        torch-side marshalling (clone / contiguous / output allocation): the engine stream is non-blocking, so only work
        enqueued on torch's stream BEFORE the event is recorded is ordered before the engine's kernels."""
        if any(_is_torch(x) for x in xs if x is not None):
            import torch
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.torch_stream().wait_event(ev)
            return True
        return False

    def _fence_out(self, used):
        if used:
            import torch
            ev = torch.cuda.Event()
            ev.record(self.torch_stream())
            torch.cuda.current_stream(self.device).wait_event(ev)

    @staticmethod
    def _shape_like(res, x):
        if _is_torch(x):
            return res.view(x.shape) if x.dim() > 1 else res
        x = np.asarray(x)
        return res.reshape(x.shape) if x.ndim > 1 else res

    # -- the hot path --------------------------------------------------------------------------------
    def vcycle(self, level, v, f, ncycles=1, history=False):
        """ncycles x V_cycle_scheme (multigrid.py:231-268).  Inputs are not modified; returns the new
        iterate shaped like ``v`` (and the per-cycle residual norms if ``history``)."""
        n = self.n[level]
        if _is_torch(v):
            vv = v.detach().clone().contiguous().view(-1)
            vp, mem, keep = vv.data_ptr(), L.MEM_DEVICE, vv
        else:
            vv = np.array(v, dtype=np.float64).reshape(-1).copy()
            vp, mem, keep = vv.ctypes.data, L.MEM_HOST, vv
        fp, memf, keepf = self._in(f, n, "f")
        if mem != memf:
            raise ValueError("v and f must both be numpy arrays or both be CUDA tensors")
        if (vv.numel() if _is_torch(v) else vv.size) != n:
            raise ValueError(f"v: expected {n} entries")
        hist = np.zeros(ncycles) if history else None
        fence = self._fence_in(v, f)
        self._ck(self._lib.mgb_vcycle(self._h, int(level), vp, fp, mem, int(ncycles), hist.ctypes.data if history else None))
        self._fence_out(fence)
        out = self._shape_like(vv, v)
        return (out, hist) if history else out

    def vcycle_debug(self, level, v, f):
        """One cycle + the reference's test=True tuple (multigrid.py:262-266): (v, f_2h, v_2h, err_h)."""
        n, nc = self.n[level], self.n[level - 1]
        vv = np.array(v, dtype=np.float64).reshape(-1).copy()
        ff = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        f2, v2, e = np.empty(nc), np.empty(nc), np.empty(n)
        self._ck(self._lib.mgb_vcycle_debug(self._h, int(level), vv.ctypes.data, ff.ctypes.data, L.MEM_HOST,
                                            f2.ctypes.data, v2.ctypes.data, e.ctypes.data))
        col = (lambda a: a.reshape(-1, 1)) if np.asarray(v).ndim > 1 else (lambda a: a)
        return col(vv), col(f2), col(v2), col(e)

    # -- FullMultiGrid on the device (multigrid.py:271-307) ------------------------------------------------------
    def set_rhs(self, level, b):
        bp, mem, keep = self._in(b, self.n[level], "b")
        self._ck(self._lib.mgb_set_rhs(self._h, int(level), bp, mem))

    def set_mass_matrix(self, level, M):
        ip, ix, ax = _as_csr_arrays(M)
        self._ck(self._lib.mgb_set_mass_matrix(self._h, int(level), M.shape[0], len(ax), ip.ctypes.data, ip.dtype.itemsize,
                                               ix.ctypes.data, ax.ctypes.data))

    def set_exact_solution(self, level, u):
        """Nodal values of the exact solution: ``fmg`` then records the error norm after every finest-level cycle
        (error_per_V_cycle_finest, multigrid.py:292-293).  ``None`` forgets it."""
        if u is None:
            self._ck(self._lib.mgb_set_exact_solution(self._h, int(level), None, L.MEM_HOST))
            return
        up, mem, keep = self._in(u, self.n[level], "u_exact")
        self._ck(self._lib.mgb_set_exact_solution(self._h, int(level), up, mem))

    def fmg_errors(self):
        """Error norms of the last ``fmg`` run, one per finest-level cycle (empty without an exact solution)."""
        cnt = C.c_int()
        self._ck(self._lib.mgb_fmg_error_history(self._h, None, 0, C.byref(cnt)))
        out = np.zeros(max(cnt.value, 1))
        self._ck(self._lib.mgb_fmg_error_history(self._h, out.ctypes.data, len(out), C.byref(cnt)))
        return out[:cnt.value]

    def fmg(self, mu0=2, tol=1e-11, max_cycles=10000):
        """-> (finest-level solution (n,), residual norms per finest-level cycle)."""
        lf = max(self.n)
        v = np.empty(self.n[lf])
        hist = np.zeros(max(max_cycles, 1))
        done = C.c_int()
        self._ck(self._lib.mgb_fmg(self._h, int(mu0), float(tol), int(max_cycles), v.ctypes.data, L.MEM_HOST, C.byref(done),
                                   hist.ctypes.data, len(hist)))
        return v, hist[:done.value]

    def vcycle_resident(self, level, ncycles=1, history=False):
        """Cycles on the engine's own level buffers (see ``level_buffer``): no copies at all."""
        hist = np.zeros(ncycles) if history else None
        self._ck(self._lib.mgb_vcycle_resident(self._h, int(level), int(ncycles), hist.ctypes.data if history else None))
        return hist

    def level_buffer(self, level, which="v"):
        """Engine-owned device vector as a torch tensor view (no copy)."""
        import torch
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self._lib.mgb_level_buffer(self._h, int(level), {"v": L.BUF_V, "f": L.BUF_F, "r": L.BUF_R}[which], C.byref(p), C.byref(n)))
        iface = {"shape": (n.value,), "typestr": "<f8", "data": (p.value, False), "version": 2}
        holder = type("MgbBuf", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=f"cuda:{self.device}")

    # -- per-operator entry points ---------------------------------------------------------------------
    def _binary(self, fn, level, n_in, n_out, x, name):
        xp, mem, keep = self._in(x, n_in, name)
        yp, y = self._out_like(x, n_out)
        fence = self._fence_in(x)
        self._ck(fn(self._h, int(level), xp, yp, mem))
        self._fence_out(fence)
        return y

    def spmv(self, level, x):
        return self._shape_like(self._binary(self._lib.mgb_spmv, level, self.n[level], self.n[level], x, "x"), x)

    def residual(self, level, v, f):
        n = self.n[level]
        vp, mem, k1 = self._in(v, n, "v")
        fp, mem2, k2 = self._in(f, n, "f")
        rp, r = self._out_like(v, n)
        fence = self._fence_in(v, f)
        self._ck(self._lib.mgb_residual(self._h, int(level), vp, fp, rp, mem))
        self._fence_out(fence)
        return self._shape_like(r, v)

    def smooth(self, level, v, f, nsweeps):
        """jacobiRelaxation (multigrid.py:223-228) or the selected Gauss-Seidel; inputs not modified."""
        n = self.n[level]
        if _is_torch(v):
            vv = v.detach().clone().contiguous().view(-1); vp, mem = vv.data_ptr(), L.MEM_DEVICE
        else:
            vv = np.array(v, dtype=np.float64).reshape(-1).copy(); vp, mem = vv.ctypes.data, L.MEM_HOST
        fp, memf, keep = self._in(f, n, "f")
        fence = self._fence_in(v, f)
        self._ck(self._lib.mgb_smooth(self._h, int(level), vp, fp, int(nsweeps), mem))
        self._fence_out(fence)
        return self._shape_like(vv, v)

    def restrict(self, fine_level, r):
        y = self._binary(self._lib.mgb_restrict, fine_level, self.n[fine_level], self.n[fine_level - 1], r, "r")
        return y.reshape(-1, 1) if (not _is_torch(r) and np.asarray(r).ndim > 1) else (y.view(-1, 1) if _is_torch(r) and r.dim() > 1 else y)

    def prolong_add(self, fine_level, e, v):
        n, nc = self.n[fine_level], self.n[fine_level - 1]
        if _is_torch(v):
            vv = v.detach().clone().contiguous().view(-1); vp, mem = vv.data_ptr(), L.MEM_DEVICE
        else:
            vv = np.array(v, dtype=np.float64).reshape(-1).copy(); vp, mem = vv.ctypes.data, L.MEM_HOST
        ep, meme, keep = self._in(e, nc, "e")
        fence = self._fence_in(e, v)
        self._ck(self._lib.mgb_prolong_add(self._h, int(fine_level), ep, vp, mem))
        self._fence_out(fence)
        return self._shape_like(vv, v)

    def coarse_solve(self, f):
        lc = min(self.n)
        n = self.n[lc]
        fp, mem, keep = self._in(f, n, "f")
        up, u = self._out_like(f, n)
        fence = self._fence_in(f)
        self._ck(self._lib.mgb_coarse_solve(self._h, fp, up, mem))
        self._fence_out(fence)
        return self._shape_like(u, f)

    def norm2(self, x):
        n = x.numel() if _is_torch(x) else np.asarray(x).size
        xp, mem, keep = self._in(x, n, "x")
        out = C.c_double()
        fence = self._fence_in(x)
        self._ck(self._lib.mgb_norm2(self._h, n, xp, mem, C.byref(out)))
        return out.value

    # -- introspection ---------------------------------------------------------------------------------
    def artifact(self, level, kind):
        size = C.c_int64()
        self._ck(self._lib.mgb_get_artifact(self._h, int(level), int(kind), None, 0, C.byref(size)))
        is_f64 = kind in (L.ART_RJ_VALUES, L.ART_DINV, L.ART_R_VALUES, L.ART_COARSE_INVERSE, L.ART_A_VALUES, L.ART_P_VALUES)
        out = np.empty(size.value // (8 if is_f64 else 4), dtype=np.float64 if is_f64 else np.int32)
        if size.value:
            self._ck(self._lib.mgb_get_artifact(self._h, int(level), int(kind), out.ctypes.data, size.value, None))
        return out

    def code_artifact(self, level, op):
        """The lossless coding mgb_finalize built on the device for operator ``op`` ("A", "RJ", "P", "R") of ``level``, in the
        layout of ``host_code_operator`` (module level): dict(mode, ndict, codes, table, head)."""
        base = 32 + 4 * {"A": 0, "RJ": 1, "P": 2, "R": 3}[op]

        def fetch(part, dtype):
            size = C.c_int64()
            self._ck(self._lib.mgb_get_artifact(self._h, int(level), base + part, None, 0, C.byref(size)))
            out = np.zeros(size.value // np.dtype(dtype).itemsize, dtype=dtype)
            if size.value:
                self._ck(self._lib.mgb_get_artifact(self._h, int(level), base + part, out.ctypes.data, size.value, None))
            return out
        info = fetch(0, np.int32)
        mode = int(info[0])
        out = {"mode": mode, "ndict": int(info[1]), "codes": fetch(1, np.uint8), "table": fetch(2, CODE_TABLE_DTYPE),
               "head": fetch(3, np.int32).reshape(-1, 2) if mode in (3, 4) else None}
        if mode == 4:                        # anchored row patterns: every row's first stored column
            base = 48 + {"A": 0, "RJ": 1, "P": 2, "R": 3}[op] - 0
            size = C.c_int64()
            self._ck(self._lib.mgb_get_artifact(self._h, int(level), base, None, 0, C.byref(size)))
            anc = np.zeros(size.value // 4, dtype=np.int32)
            if size.value:
                self._ck(self._lib.mgb_get_artifact(self._h, int(level), base, anc.ctypes.data, size.value, None))
            out["anchor"] = anc
        return out

    def rj_matrix(self, level):
        """R_omega of ``level`` as scipy CSR + D^-1 (what getJacobiMatrices returns, multigrid.py:56)."""
        import scipy.sparse as sp
        ip = self.artifact(level, L.ART_RJ_INDPTR); ix = self.artifact(level, L.ART_RJ_INDICES)
        ax = self.artifact(level, L.ART_RJ_VALUES); d = self.artifact(level, L.ART_DINV)
        n = self.n[level]
        return sp.csr_matrix((ax, ix, ip), shape=(n, n)), d

    def launch_count(self):
        c = C.c_int64()
        self._ck(self._lib.mgb_launch_count(self._h, C.byref(c)))
        return c.value

    def vcycle_bytes(self, level):
        b = C.c_double()
        self._ck(self._lib.mgb_vcycle_bytes(self._h, int(level), C.byref(b)))
        return b.value

    def vcycle_bytes_moved(self, level):
        """Bytes the chosen kernels stream per V-cycle (fewer than ``vcycle_bytes`` for dictionary-coded operators)."""
        b = C.c_double()
        self._ck(self._lib.mgb_vcycle_bytes_moved(self._h, int(level), C.byref(b)))
        return b.value

    def describe(self):
        buf = C.create_string_buffer(1 << 16)
        self._ck(self._lib.mgb_describe(self._h, buf, len(buf)))
        return buf.value.decode()

    def profile_begin(self):
        self._ck(self._lib.mgb_profile_begin(self._h))

    def profile_end(self):
        """-> list of dicts {kind, level, launches, total_ms, bytes, gbs, moved_bytes, moved_gbs}: ``bytes`` are the
        algorithmic bytes of the CSR form, ``moved_bytes`` what the chosen kernel streams (dictionary-coded operators)."""
        self._ck(self._lib.mgb_profile_end(self._h))
        cnt = C.c_int()
        self._ck(self._lib.mgb_profile_get(self._h, None, 0, C.byref(cnt)))
        arr = (L.ProfileRecord * max(cnt.value, 1))()
        self._ck(self._lib.mgb_profile_get(self._h, arr, cnt.value, C.byref(cnt)))
        out = []
        for i in range(cnt.value):
            r = arr[i]
            ms = r.total_ms / max(r.launches, 1)
            out.append({"kind": L.KERNEL_KINDS[r.kind], "level": r.level, "launches": r.launches, "total_ms": r.total_ms,
                        "ms_per_launch": ms, "bytes": r.bytes, "gbs": (r.bytes / (ms * 1e-3) / 1e9) if ms > 0 else 0.0,
                        "moved_bytes": r.moved_bytes, "moved_gbs": (r.moved_bytes / (ms * 1e-3) / 1e9) if ms > 0 else 0.0})
        return out


# ---- host-side setup routines (usable without a GPU) -------------------------------------------------

def host_build_rj(A, reversed_order=True):
    lib = L.load()
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64); ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    ax = np.ascontiguousarray(A.data, dtype=np.float64)
    n = A.shape[0]
    nnz = C.c_int64()
    lib.mgb_host_build_rj(n, ip.ctypes.data, ix.ctypes.data, ax.ctypes.data, int(reversed_order), C.byref(nnz), None, None, None, None)
    rip = np.empty(n + 1, dtype=np.int32); rix = np.empty(nnz.value, dtype=np.int32)
    rax = np.empty(nnz.value); dinv = np.empty(n)
    rc = lib.mgb_host_build_rj(n, ip.ctypes.data, ix.ctypes.data, ax.ctypes.data, int(reversed_order), C.byref(nnz),
                               rip.ctypes.data, rix.ctypes.data, rax.ctypes.data, dinv.ctypes.data)
    if rc != L.OK:
        raise L.MGBError(rc, "zero or missing diagonal")
    return rip, rix, rax, dinv


def _host_groups(fn, A):
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64); ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    ax = np.ascontiguousarray(A.data, dtype=np.float64)
    n = A.shape[0]
    key = np.empty(n, dtype=np.int32); order = np.empty(n, dtype=np.int32); off = np.empty(n + 2, dtype=np.int32)
    cnt = C.c_int64()
    rc = fn(n, ip.ctypes.data, ix.ctypes.data, ax.ctypes.data, key.ctypes.data, order.ctypes.data, C.byref(cnt), off.ctypes.data, n + 2)
    if rc != L.OK:
        raise L.MGBError(rc, "host artefact routine failed")
    return key, order, off[:cnt.value + 1].copy()


def host_level_sets(A):
    return _host_groups(L.load().mgb_host_level_sets, A)


def host_colouring(A):
    return _host_groups(L.load().mgb_host_colouring, A)


def host_dense_inverse(A):
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64); ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    ax = np.ascontiguousarray(A.data, dtype=np.float64)
    n = A.shape[0]
    inv = np.empty((n, n))
    rc = L.load().mgb_host_dense_inverse(n, ip.ctypes.data, ix.ctypes.data, ax.ctypes.data, inv.ctypes.data)
    if rc != L.OK:
        raise L.MGBError(rc, "singular matrix")
    return inv


CODE_TABLE_DTYPE = np.dtype([("val", "<f8"), ("delta", "<i4"), ("pad", "<i4")])


def host_code_operator(A, allow_patterns=2):
    """The lossless operator coding of DESIGN.md 4.1, computed by the library's host routine (the definition of what
    ``mgb_finalize`` builds on the device).  -> dict(mode, ndict, codes, table, head): ``codes`` uint8 per row (mode 3) or per
    stored entry (modes 1, 2); ``table`` structured array {val, delta}; ``head`` (mode 3) int32 (256, 2) {first entry, length}."""
    lib = L.load()
    ip = np.ascontiguousarray(A.indptr, dtype=np.int64); ix = np.ascontiguousarray(A.indices, dtype=np.int32)
    ax = np.ascontiguousarray(A.data, dtype=np.float64)
    n, m = A.shape
    mode, ndict, nent = C.c_int(), C.c_int(), C.c_int()
    codes = np.zeros(max(n, len(ax), 1), dtype=np.uint8)
    table = np.zeros(2048, dtype=CODE_TABLE_DTYPE)
    head = np.zeros((256, 2), dtype=np.int32)
    allow = 2 if allow_patterns is True else int(allow_patterns)          # 0 per-entry codes, 1 + row patterns, 2 + anchored row patterns
    rc = lib.mgb_host_code_operator(n, m, ip.ctypes.data, ix.ctypes.data, ax.ctypes.data, allow, C.byref(mode), C.byref(ndict),
                                    codes.ctypes.data, table.ctypes.data, C.byref(nent), head.ctypes.data)
    if rc != L.OK:
        raise L.MGBError(rc, "mgb_host_code_operator failed")
    ncodes = {0: 0, 3: n, 4: n}.get(mode.value, len(ax))
    out = {"mode": mode.value, "ndict": ndict.value, "codes": codes[:ncodes].copy(), "table": table[:nent.value].copy(),
           "head": head if mode.value in (3, 4) else None}
    if mode.value == 4:                      # the anchors are, by definition, the first stored column of every row (0: empty row)
        anc = np.zeros(n, dtype=np.int32)
        ne = ip[1:] > ip[:-1]
        anc[ne] = ix[ip[:-1][ne]]
        out["anchor"] = anc
    return out
