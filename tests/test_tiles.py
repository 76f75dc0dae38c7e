"""Row tiles of the TMA stream kernels (mgb_host_make_tiles = the routine mgb_finalize uses): the invariants the kernels'
bulk copies rely on -- coverage, caps, alignment of tile starts, no tile across a breakpoint -- as properties over random
row-length distributions.  The three call shapes of finish_csr are covered: CSR stream tiles (cap = entries, quarter as many
rows, starts multiple of 4), entry-coded tiles (16-entry slack) and row-pattern tiles (rows only, starts multiple of 16)."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from multigrid_dolfinx_b200 import _lib as L


def make_tiles(ip, cap, row_cap, breaks=(), row_align=1):
    lib = L.load()
    ip = np.ascontiguousarray(ip, dtype=np.int64)
    n = len(ip) - 1
    br = np.ascontiguousarray(breaks, dtype=np.int32)
    tiles = np.zeros(n + 2, dtype=np.int32)
    bt = np.zeros(max(len(br), 1), dtype=np.int32)
    nt = C.c_int64()
    rc = lib.mgb_host_make_tiles(n, ip.ctypes.data, int(cap), int(row_cap), len(br), br.ctypes.data if len(br) else None, int(row_align),
                                 tiles.ctypes.data, len(tiles), C.byref(nt), bt.ctypes.data)
    if rc == L.ERR_UNSUPPORTED:
        return None, None
    assert rc == L.OK
    return tiles[:nt.value + 1].copy(), bt[:len(br)].copy()


def check(ip, cap, row_cap, breaks, row_align, tiles, bt):
    n = len(ip) - 1
    assert tiles[0] == 0 and tiles[-1] == n and (n == 0 or np.all(np.diff(tiles) > 0))
    for t0, t1 in zip(tiles[:-1], tiles[1:]):
        assert t1 - t0 <= row_cap
        assert ip[t1] - (ip[t0] & ~7) <= cap
        assert not any(t0 < b < t1 for b in breaks)                    # no tile straddles a breakpoint
        assert t0 % row_align == 0 or t0 in breaks                     # bulk-copy alignment of the row-wise slices
    for b, k in zip(breaks, bt):                                        # tile index at each breakpoint
        assert tiles[min(k, len(tiles) - 1)] == min(b, n) or (b <= 0 and k == 0)


row_lengths = st.one_of(
    st.lists(st.integers(0, 9), min_size=0, max_size=600),                                    # short rows, empty ones included
    st.lists(st.sampled_from([4, 4, 4, 4, 7, 15, 0, 65]), min_size=1, max_size=3000),         # the benchmark operators' row classes
    st.lists(st.integers(0, 900), min_size=1, max_size=40))                                   # some rows close to a tile's capacity


@settings(max_examples=150, deadline=None)
@given(row_lengths, st.sampled_from([(1024, 256, 4), (2048, 512, 4), (4096 - 8, 512, 4), (1 << 40, 512, 16), (1 << 40, 256, 16)]),
       st.lists(st.integers(0, 3000), max_size=2))
def test_tile_invariants(lens, shape, raw_breaks):
    cap, row_cap, align = shape
    ip = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=ip[1:])
    n = len(lens)
    breaks = sorted({min(b, n) // align * align for b in raw_breaks} - {0}) if n else []      # finish_csr rounds them the same way
    tiles, bt = make_tiles(ip, cap, row_cap, breaks, align)
    if tiles is None:                                   # refused: some row (from its 8-aligned start) exceeds cap, or a tile could not
        worst = max((ip[i + 1] - (ip[i] & ~7) for i in range(n)), default=0)                  # keep `align` rows
        assert worst > cap or align > 1
        return
    check(ip, cap, row_cap, breaks, align, tiles, bt)


def test_benchmark_shapes():
    # config 2's fine R_omega: 4 entries per interior row, boundary rows empty
    N = 257
    lens = np.full(N * N, 4); lens[:N] = 0; lens[-N:] = 0; lens[::N] = 0; lens[N - 1::N] = 0
    ip = np.r_[0, np.cumsum(lens)]
    for cap, row_cap, align in ((1024, 256, 4), (4096 - 8, 512, 4), (1 << 40, 512, 16)):
        tiles, _ = make_tiles(ip, cap, row_cap, (), align)
        check(ip, cap, row_cap, [], align, tiles, [])
        if cap == 1 << 40:
            assert np.all(np.diff(tiles)[:-1] == 512)                  # row-pattern tiles are full except the last
    # sharded operator: interior block [b0, b1) between two boundary blocks
    tiles, bt = make_tiles(ip, 1 << 40, 512, (1024, 60000), 16)
    check(ip, 1 << 40, 512, [1024, 60000], 16, tiles, bt)
    assert tiles[bt[0]] == 1024 and tiles[bt[1]] == 60000
