"""Host-side logic of the row-sharded hierarchy (multigrid_dolfinx_b200/dist.py) on CPU.

world_size-2 (and 3) gloo runs: every rank builds its local pieces, then a numpy emulation of the sharded
V-cycle (same operator order as csrc/mgb_engine.cu: exchange -> row sums on [owned | ghost] vectors, gather of
the first small level to rank 0, broadcast back) must reproduce the serial oracle BIT FOR BIT -- the local
matrices keep the stored entry order of every row, so nothing may differ."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from multigrid_dolfinx_b200 import dist as ds
from multigrid_dolfinx_b200 import problems as pr
from oracle import restated as rs


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _exchange(td, L, vec):
    """what mgb_engine.cu::exchange does: pack owned entries per neighbour, send/recv, ghosts behind the owned part"""
    import torch
    reqs, bufs = [], []
    off = L.n_owned
    for p, q in enumerate(L.peers):
        if len(L.send_idx[p]):
            reqs.append(td.isend(torch.from_numpy(np.ascontiguousarray(vec[L.send_idx[p]])), q))
        if L.recv_cnt[p]:
            t = torch.empty(L.recv_cnt[p], dtype=torch.float64)
            reqs.append(td.irecv(t, q)); bufs.append((off, t))
        off += L.recv_cnt[p]
    for r in reqs:
        r.wait()
    for o, t in bufs:
        vec[o:o + len(t)] = t.numpy()


def _local_jacobi(H, loc):
    """R_omega / D^-1 of the serial oracle (multigrid.py:48-56), rows of this rank, columns in local numbering"""
    out = {}
    for l, L in loc["levels"].items():
        RO, dinv = rs.jacobi_matrices(H.A_sp_dict[l][0])
        out[l] = (ds._localize(RO[L.s:L.e], L.s, L.e, L.ghosts, L.n_owned + L.n_ghost), dinv[L.s:L.e])
    return out


def _emulated_cycle(td, rank, world, loc, src, jac, serial_coarse, v, f):
    """one sharded V-cycle in numpy; returns the new owned part of v on the finest level"""
    import torch
    om, mu1, mu2 = src.omega, src.mu1, src.mu2
    g = loc["gather_level"]
    lf = src.levels[-1]
    vs, fs = {lf: v}, {lf: f}

    def smooth(l, x, b, nw):
        L = loc["levels"][l]
        RO, dinv = jac[l]
        for _ in range(nw):
            full = np.zeros(L.n_owned + L.n_ghost); full[:L.n_owned] = x
            _exchange(td, L, full)
            x = (1 - om) * x + om * (dinv * b) - om * RO.dot(full)
        return x

    for l in range(lf, g, -1):
        L = loc["levels"][l]
        x = vs[l] if l == lf else np.zeros(L.n_owned)
        x = smooth(l, x, fs[l], mu1)
        full = np.zeros(L.n_owned + L.n_ghost); full[:L.n_owned] = x
        _exchange(td, L, full)
        r = fs[l] - L.A.dot(full)
        if L.inj is not None:
            fc = r[L.inj]
        else:
            rfull = np.zeros(L.n_owned + L.n_ghost); rfull[:L.n_owned] = r
            _exchange(td, L, rfull)
            fc = L.R.dot(rfull)
        vs[l] = x
        fs[l - 1] = fc
    # gather the first small level to rank 0, solve the rest of the cycle there, broadcast the correction
    parts = [None] * world
    td.all_gather_object(parts, fs[g])
    e_g = serial_coarse(np.concatenate(parts)) if rank == 0 else None
    box = [e_g]
    td.broadcast_object_list(box, src=0)
    e = box[0]
    for l in range(g + 1, lf + 1):
        L = loc["levels"][l]
        if l - 1 == g:
            efull = e
        else:
            C = loc["levels"][l - 1]
            efull = np.zeros(C.n_owned + C.n_ghost); efull[:C.n_owned] = e
            _exchange(td, C, efull)
        x = vs[l] + L.P.dot(efull)
        x = smooth(l, x, fs[l], mu2)
        e = x
    return e


def _worker(rank, world, port, case, q):
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim, c, lf, seed, r_mode, glevel = case
        H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, perm_seed=seed, with_dicts=False)
        src = ds.HierarchySource(H)

        def allgather(obj):
            out = [None] * world
            td.all_gather_object(out, obj)
            return out
        loc = ds.build_local(src, rank, world, r_mode=r_mode, gather_level=glevel, allgather=allgather)
        mg = rs.from_hierarchy(H, r_mode=r_mode)

        def serial_coarse(fg):
            return mg.vcycle(glevel, np.zeros((len(fg), 1)), fg[:, None])[:, 0]
        L = loc["levels"][lf]
        jac = _local_jacobi(H, loc)
        v = np.zeros(L.n_owned); f = L.rhs.copy()
        vref = np.zeros((H.n(lf), 1))
        ok = True
        for _ in range(2):
            v = _emulated_cycle(td, rank, world, loc, src, jac, serial_coarse, v, f)
            vref = mg.vcycle(lf, vref, H.b_dict[lf])
            ok = ok and np.array_equal(v, vref[L.s:L.e, 0])
        # structural checks
        for l, LL in loc["levels"].items():
            assert LL.A.shape == (LL.n_owned, LL.n_owned + LL.n_ghost)
            assert sum(LL.recv_cnt) == LL.n_ghost
            assert np.all(np.diff(LL.ghosts) > 0)
        q.put((rank, ok, None))
    except Exception as ex:       # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))
    finally:
        td.destroy_process_group()


CASES = [
    (2, 4, 3, None, "injection", 0),       # lexicographic, aligned injection, only the coarsest level gathered
    (2, 4, 3, None, "injection", 1),
    (3, 2, 3, None, "transpose", 1),       # CSR restriction rows need fine-level ghosts
    (2, 4, 3, 5, "injection", 1),          # random numbering: injection falls back to explicit single-entry rows
    (3, 2, 2, None, "full_weighting", 0),
]


@pytest.mark.parametrize("world,case", [(2, c) for c in CASES] + [(3, CASES[1]), (3, CASES[2])])
def test_sharded_cycle_is_bit_identical_to_serial(world, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert err is None, err
        assert ok, f"rank {rank}: sharded result differs from the serial oracle"


def test_structured_source_matches_hierarchy_source():
    """Row blocks generated by index arithmetic == row blocks cut from the full matrices."""
    H = pr.build_hierarchy(dim=3, c=2, coarsest_level=0, finest_level=2, with_dicts=False)
    a, b = ds.HierarchySource(H), ds.StructuredSource(3, 2, 0, 2)
    for l in (1, 2):
        n = a.n(l)
        s, e = n // 3, 2 * n // 3
        for M1, M2 in ((a.A_rows(l, s, e), b.A_rows(l, s, e)), (a.P_rows(l - 1, s, e), b.P_rows(l - 1, s, e))):
            assert np.array_equal(M1.indptr, M2.indptr) and np.array_equal(M1.indices, M2.indices) and np.array_equal(M1.data, M2.data)
        assert np.array_equal(a.inj(l - 1), b.inj(l - 1))
        nc = a.n(l - 1)
        for mode in ("transpose", "full_weighting"):
            R1, R2 = a.R_rows(l - 1, nc // 4, nc // 2, mode), b.R_rows(l - 1, nc // 4, nc // 2, mode)
            assert (R1 != R2).nnz == 0 and np.array_equal(R1.indices, R2.indices)


def test_partition_follows_injection():
    src = ds.StructuredSource(3, 4, 0, 3)
    off, aligned = ds.plan_offsets(src, 4, 1)
    for l in (1, 2):
        inj = src.inj(l)
        for r in range(4):
            g = inj[off[l][r]:off[l][r + 1]]
            assert len(g) == 0 or (g.min() >= off[l + 1][r] and g.max() < off[l + 1][r + 1])
        assert aligned[l]
    assert ds.choose_gather_level(src, 1000) == 1 and ds.choose_gather_level(src, 10) == 0


@pytest.mark.parametrize("dim,c,lf,world", [(3, 2, 3, 4), (2, 8, 3, 3), (3, 4, 2, 8)])
def test_range_halo_plan_covers_the_real_ghosts_and_is_consistent(dim, c, lf, world):
    """The device-side generator describes ghosts as two index ranges one bandwidth wide.  They must contain every column
    the operators of that rank really reference, and what rank a sends to b must be exactly what b expects from a."""
    src = ds.StructuredSource(dim, c, 0, lf)
    g = 0
    offsets, aligned = ds.plan_offsets(src, world, g)
    needs = [None] * world

    class Stop(Exception):
        pass

    for r in range(world):                      # first pass: collect every rank's ghost requests (emulates the all-gather)
        def grab(obj, r=r):
            needs[r] = obj
            raise Stop
        try:
            ds.build_local(src, r, world, gather_level=g, allgather=grab)
        except Stop:
            pass
    for r in range(world):
        loc = ds.build_local(src, r, world, gather_level=g, allgather=lambda o: needs)
        for l, L in loc["levels"].items():
            glo, ghi, peers, send, recv = ds.range_halo_plan(offsets[l], r, src.n(l), ds.bandwidth(src.N(l), dim))
            gh = L.ghosts
            assert np.all(((gh >= glo) & (gh < L.s)) | ((gh >= L.e) & (gh < ghi)))
            assert sum(recv) == (L.s - glo) + (ghi - L.e)
    for l in range(g + 1, lf + 1):
        plans = [ds.range_halo_plan(offsets[l], r, src.n(l), ds.bandwidth(src.N(l), dim)) for r in range(world)]
        for r in range(world):
            _, _, peers, send, recv = plans[r]
            for p, q in enumerate(peers):
                _, _, pq, sq, rq = plans[q]
                j = pq.index(r)
                assert len(send[p]) == rq[j] and recv[p] == len(sq[j])
