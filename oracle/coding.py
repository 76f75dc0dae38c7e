"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the lossless operator coding.

The reference has no such thing (it streams scipy CSR, multigrid.py:226 / :244); the coding is this engine's own artefact,
so -- like the Gauss-Seidel level sets -- its DEFINITION is stated twice, independently: in C++ (mgb_host_code_operator,
csrc/mgb_setup.cpp, mirrored on the device by csrc/mgb_code.cuh) and here.  tests/test_coding.py compares the two and checks
that decoding gives back the CSR arrays bit for bit.  Only tests may import this module.
"""
import numpy as np


def _rows(A, anchored=False):
    """row -> (row, columns measured from the row index -- or, anchored, from the row's first stored column --, value bits)"""
    ip, ix, ax = A.indptr, A.indices.astype(np.int64), A.data.view(np.uint64)
    for i in range(A.shape[0]):
        k = slice(ip[i], ip[i + 1])
        base = (ix[ip[i]] if ip[i + 1] > ip[i] else 0) if anchored else i
        yield i, (ix[k] - base), ax[k]


def anchors(A):
    """Mode 4: every row's anchor column = its first stored column (0 for an empty row)."""
    ip = np.asarray(A.indptr)
    out = np.zeros(A.shape[0], dtype=np.int32)
    ne = ip[1:] > ip[:-1]
    out[ne] = A.indices[ip[:-1][ne]]
    return out


def code_operator(A, allow_patterns=2):
    """-> dict(mode, ndict, codes, table=[(val, delta)], head) following the definition in DESIGN.md 4.1.
    allow_patterns: 0 per-entry codes only, 1 + row patterns (mode 3), 2 (or True... any value >= 2) + anchored row patterns (mode 4)."""
    n, m = A.shape
    nnz = A.nnz
    allow_patterns = 2 if allow_patterns is True else int(allow_patterns)
    none = {"mode": 0, "ndict": 0, "codes": np.zeros(0, np.uint8), "table": [], "head": None}
    if n == 0 or nnz == 0:
        return none
    for anchored in (False, True):
        if (allow_patterns < 2 or nnz / n > 32.0) if anchored else (allow_patterns < 1 or m < n or nnz / n > 24.0):
            continue
        first = {}
        for i, d, b in _rows(A, anchored):
            key = (tuple(int(x) for x in d), tuple(int(x) for x in b))
            if key not in first:
                first[key] = i
                if len(first) > 256:
                    break
        if len(first) <= 256:
            order = sorted(first.items(), key=lambda kv: kv[1])               # patterns numbered by their first row
            padded = [max(8, (len(k[0]) + 7) // 8 * 8) for k, _ in order]
            if sum(padded) <= 2048:
                table, head, off, ident = [], np.zeros((256, 2), np.int32), 0, {}
                for p, ((d, b), _) in enumerate(order):
                    head[p] = (off, len(d))
                    ent = [(np.uint64(bb).view(np.float64), int(dd)) for dd, bb in zip(d, b)]
                    fill = ent[-1] if ent else (0.0, 0)
                    table += ent + [fill] * (padded[p] - len(ent))
                    off += padded[p]
                    ident[(d, b)] = p
                codes = np.array([ident[(tuple(int(x) for x in d), tuple(int(x) for x in b))] for _, d, b in _rows(A, anchored)], dtype=np.uint8)
                return {"mode": 4 if anchored else 3, "ndict": len(order), "codes": codes, "table": table, "head": head}
    if nnz / n > 24.0:
        return none
    bits = A.data.view(np.uint64)
    V = np.unique(bits)                                                          # ascending as unsigned 64-bit patterns
    if len(V) > 256 or np.any(V == np.uint64(0xFFFFFFFFFFFFFFFF)):
        return none
    vi = np.searchsorted(V, bits)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
    delta = A.indices.astype(np.int64) - rows
    Dl = np.unique(delta)
    pair = len(Dl) <= 256 and not np.any(Dl == np.int32(-2139062144))
    if pair:
        key = np.searchsorted(Dl, delta) * 256 + vi
        K = np.unique(key)
        if len(K) <= 256:
            codes = np.searchsorted(K, key).astype(np.uint8)
            table = [(V[k & 255].view(np.float64), int(Dl[k >> 8])) for k in K.tolist()]
            return {"mode": 1, "ndict": len(K), "codes": codes, "table": table, "head": None}
    return {"mode": 2, "ndict": len(V), "codes": vi.astype(np.uint8), "table": [(v.view(np.float64), 0) for v in V], "head": None}


def decode(A_shape, indptr, indices, coded):
    """Coded operator -> (columns, value bit patterns) per stored entry, using ``indptr`` (and, in mode 2, ``indices``)."""
    n = A_shape[0]
    mode, codes, table = coded["mode"], coded["codes"], coded["table"]
    vals = np.array([np.float64(t[0]) for t in table]).view(np.uint64) if len(table) else np.zeros(0, np.uint64)
    dels = np.array([t[1] for t in table], dtype=np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    if mode in (3, 4):
        head = coded["head"]
        lens = head[codes.astype(np.int64), 1]
        assert np.array_equal(lens, np.diff(indptr))
        pos = head[codes.astype(np.int64), 0][rows] + (np.arange(len(rows)) - np.asarray(indptr)[rows])
        base = rows if mode == 3 else np.asarray(coded["anchor"], dtype=np.int64)[rows]
        return base + dels[pos], vals[pos]
    c = codes.astype(np.int64)
    if mode == 1:
        return rows + dels[c], vals[c]
    return np.asarray(indices, dtype=np.int64), vals[c]
