// Host-side setup of the level hierarchy: everything the reference does once per level before the
// V-cycle runs (getJacobiMatrices, multigrid.py:48-56) plus the artefacts this engine defines
// (restriction from P, Gauss-Seidel level sets / colours, dense coarsest inverse, row tiles).
// Pure C++ -- no CUDA here, so these routines are testable without a GPU through the mgb_host_* ABI.
#include "mgb_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>
#include <set>

#include "../../include/mgb200.h"

namespace mgb {

std::string import_csr(HostCsr& out, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr, int indptr_bytes,
                       const int32_t* indices, const double* values)
{
    if (nrows < 0 || ncols < 0 || nnz < 0) return "negative dimension";
    if (!indptr || (nnz > 0 && (!indices || !values))) return "null CSR array";
    if (indptr_bytes != 4 && indptr_bytes != 8) return "indptr_bytes must be 4 or 8";
    out.nrows = nrows; out.ncols = ncols;
    out.ip.resize((size_t)nrows + 1);
    if (indptr_bytes == 4) {
        const int32_t* p = (const int32_t*)indptr;
        for (int64_t i = 0; i <= nrows; ++i) out.ip[(size_t)i] = p[i];
    } else {
        std::memcpy(out.ip.data(), indptr, sizeof(int64_t) * (size_t)(nrows + 1));
    }
    if (out.ip[0] != 0 || out.ip[(size_t)nrows] != nnz) return "indptr[0] != 0 or indptr[n] != nnz";
    for (int64_t i = 0; i < nrows; ++i)
        if (out.ip[(size_t)i + 1] < out.ip[(size_t)i]) return "indptr not monotone";
    out.ix.assign(indices, indices + nnz);
    out.ax.assign(values, values + nnz);
    for (int64_t k = 0; k < nnz; ++k)
        if (out.ix[(size_t)k] < 0 || out.ix[(size_t)k] >= ncols) return "column index out of range";
    return "";
}

bool build_rj(const HostCsr& A, bool reversed, HostCsr& RJ, std::vector<double>& dinv)
{
    const int64_t n = A.nrows;
    dinv.assign((size_t)n, 0.0);
    RJ.nrows = n; RJ.ncols = A.ncols;
    RJ.ip.assign((size_t)n + 1, 0);
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        double d = 0.0; bool have = false; int64_t cnt = 0;
        for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) {
            if (A.ix[k] == i) { d += A.ax[k]; have = true; }     // A.diagonal() sums duplicates
            else if (A.ax[k] != 0.0) ++cnt;                      // "A - diags(d)" drops exact zeros
        }
        if (!have || d == 0.0) ok = false;
        dinv[i] = 1.0 / d;
        RJ.ip[i + 1] = RJ.ip[i] + cnt;
    }
    RJ.ix.resize((size_t)RJ.ip[n]); RJ.ax.resize((size_t)RJ.ip[n]);
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = RJ.ip[i], cnt = RJ.ip[i + 1] - RJ.ip[i], j = 0;
        for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) {
            if (A.ix[k] == i || A.ax[k] == 0.0) continue;
            int64_t dst = reversed ? o + (cnt - 1 - j) : o + j;
            RJ.ix[dst] = A.ix[k];
            RJ.ax[dst] = dinv[i] * A.ax[k];
            ++j;
        }
    }
    return ok;
}

bool split_offdiag(const HostCsr& A, HostCsr& G, std::vector<double>& diag)
{
    const int64_t n = A.nrows;
    diag.assign((size_t)n, 0.0);
    G.nrows = n; G.ncols = A.ncols; G.ip.assign((size_t)n + 1, 0);
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        int64_t cnt = 0; bool have = false;
        for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) {
            if (A.ix[k] == i) { diag[i] = A.ax[k]; have = true; }   // last stored diagonal wins (oracle does the same)
            else if (A.ax[k] != 0.0) ++cnt;
        }
        if (!have || diag[i] == 0.0) ok = false;
        G.ip[i + 1] = G.ip[i] + cnt;
    }
    G.ix.resize((size_t)G.ip[n]); G.ax.resize((size_t)G.ip[n]);
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = G.ip[i];
        for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) {
            if (A.ix[k] == i || A.ax[k] == 0.0) continue;
            G.ix[o] = A.ix[k]; G.ax[o] = A.ax[k]; ++o;
        }
    }
    return ok;
}

void transpose_scaled(const HostCsr& P, double scale, HostCsr& R)
{
    const int64_t nr = P.ncols, nc = P.nrows, nnz = P.nnz();
    R.nrows = nr; R.ncols = nc;
    R.ip.assign((size_t)nr + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) R.ip[(size_t)P.ix[k] + 1]++;
    for (int64_t i = 0; i < nr; ++i) R.ip[i + 1] += R.ip[i];
    R.ix.resize((size_t)nnz); R.ax.resize((size_t)nnz);
    std::vector<int64_t> pos(R.ip.begin(), R.ip.end() - 1);
    for (int64_t i = 0; i < nc; ++i)            // rows of P ascending -> columns of R ascending
        for (int64_t k = P.ip[i]; k < P.ip[i + 1]; ++k) {
            int64_t dst = pos[P.ix[k]]++;
            R.ix[dst] = (int32_t)i;
            R.ax[dst] = P.ax[k] * scale;
        }
}

// lower part of the symmetrised nonzero graph: for every i the neighbours j < i
static void lower_sym_graph(const HostCsr& A, std::vector<int64_t>& lp, std::vector<int32_t>& lx)
{
    const int64_t n = A.nrows;
    lp.assign((size_t)n + 1, 0);
    auto visit = [&](auto&& fn) {
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) {
                int64_t j = A.ix[k];
                if (j == i || j >= n || A.ax[k] == 0.0) continue;      // (j >= n: a ghost column of a row block -- another block's
                fn(std::max(i, j), std::min(i, j));                       //  unknown, coupled Jacobi-style, never a dependency)
            }
    };
    visit([&](int64_t hi, int64_t) { lp[hi + 1]++; });
    for (int64_t i = 0; i < n; ++i) lp[i + 1] += lp[i];
    lx.resize((size_t)lp[n]);
    std::vector<int64_t> pos(lp.begin(), lp.end() - 1);
    visit([&](int64_t hi, int64_t lo) { lx[pos[hi]++] = (int32_t)lo; });
}

static void order_from_keys(const std::vector<int32_t>& key, int32_t nkeys, std::vector<int32_t>& order, std::vector<int32_t>& offsets)
{
    const size_t n = key.size();
    offsets.assign((size_t)nkeys + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[(size_t)key[i] + 1]++;
    for (int32_t c = 0; c < nkeys; ++c) offsets[c + 1] += offsets[c];
    order.resize(n);
    std::vector<int32_t> pos(offsets.begin(), offsets.end() - 1);
    for (size_t i = 0; i < n; ++i) order[(size_t)pos[key[i]]++] = (int32_t)i;   // stable
}

void level_sets(const HostCsr& A, std::vector<int32_t>& lev, std::vector<int32_t>& order, std::vector<int32_t>& offsets)
{
    std::vector<int64_t> lp; std::vector<int32_t> lx;
    lower_sym_graph(A, lp, lx);
    const int64_t n = A.nrows;
    lev.assign((size_t)n, 0);
    int32_t mx = -1;
    for (int64_t i = 0; i < n; ++i) {
        int32_t m = -1;
        for (int64_t k = lp[i]; k < lp[i + 1]; ++k) m = std::max(m, lev[lx[k]]);
        lev[i] = m + 1;
        mx = std::max(mx, lev[i]);
    }
    order_from_keys(lev, mx + 1, order, offsets);
}

void greedy_colouring(const HostCsr& A, std::vector<int32_t>& col, std::vector<int32_t>& order, std::vector<int32_t>& offsets)
{
    std::vector<int64_t> lp; std::vector<int32_t> lx;
    lower_sym_graph(A, lp, lx);
    const int64_t n = A.nrows;
    col.assign((size_t)n, 0);
    std::vector<int64_t> mark;   // mark[c] == i  <=> colour c used by a neighbour of i
    int32_t mx = -1;
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t k = lp[i]; k < lp[i + 1]; ++k) {
            int32_t c = col[lx[k]];
            if ((size_t)c >= mark.size()) mark.resize((size_t)c + 1, -1);
            mark[c] = i;
        }
        int32_t c = 0;
        while ((size_t)c < mark.size() && mark[c] == i) ++c;
        col[i] = c;
        mx = std::max(mx, c);
    }
    order_from_keys(col, mx + 1, order, offsets);
}

void permute_rows(const HostCsr& M, const std::vector<int32_t>& order, HostCsr& out)
{
    const int64_t n = M.nrows;
    out.nrows = n; out.ncols = M.ncols;
    out.ip.assign((size_t)n + 1, 0);
    for (int64_t p = 0; p < n; ++p) out.ip[p + 1] = out.ip[p] + (M.ip[order[p] + 1] - M.ip[order[p]]);
    out.ix.resize((size_t)out.ip[n]); out.ax.resize((size_t)out.ip[n]);
    for (int64_t p = 0; p < n; ++p) {
        int64_t s = M.ip[order[p]], c = M.ip[order[p] + 1] - s;
        std::copy(M.ix.begin() + s, M.ix.begin() + s + c, out.ix.begin() + out.ip[p]);
        std::copy(M.ax.begin() + s, M.ax.begin() + s + c, out.ax.begin() + out.ip[p]);
    }
}

bool dense_inverse(const HostCsr& A, std::vector<double>& inv)
{
    const int64_t n = A.nrows;
    const int64_t w = 2 * n;
    std::vector<double> m((size_t)(n * w), 0.0);          // [A | I], row-major
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t k = A.ip[i]; k < A.ip[i + 1]; ++k) m[i * w + A.ix[k]] += A.ax[k];
        m[i * w + n + i] = 1.0;
    }
    for (int64_t c = 0; c < n; ++c) {
        int64_t p = c; double best = std::fabs(m[c * w + c]);
        for (int64_t i = c + 1; i < n; ++i) { double t = std::fabs(m[i * w + c]); if (t > best) { best = t; p = i; } }
        if (best == 0.0 || !(best == best)) return false;
        if (p != c) std::swap_ranges(m.begin() + c * w, m.begin() + (c + 1) * w, m.begin() + p * w);
        const double piv = m[c * w + c];
        // columns < c of row c are already zero; the identity block of row c is nonzero only up to
        // column n + (largest original row index swapped in), so restrict work to [c, n + n)
        double* rc = &m[c * w];
        for (int64_t j = c; j < w; ++j) rc[j] /= piv;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            if (i == c) continue;
            double* ri = &m[i * w];
            const double l = ri[c];
            if (l == 0.0) continue;
            for (int64_t j = c; j < w; ++j) ri[j] -= l * rc[j];
        }
    }
    inv.resize((size_t)(n * n));
    for (int64_t i = 0; i < n; ++i) std::memcpy(&inv[i * n], &m[i * w + n], sizeof(double) * (size_t)n);
    return true;
}

bool make_tiles(const std::vector<int64_t>& ip, int64_t cap, int64_t row_cap, const std::vector<int32_t>& breaks,
                std::vector<int32_t>& tiles, std::vector<int32_t>* break_tile_index, int64_t row_align)
{
    const int64_t n = (int64_t)ip.size() - 1;
    tiles.clear();
    if (break_tile_index) break_tile_index->clear();
    size_t bi = 0;
    int64_t r = 0;
    while (bi < breaks.size() && breaks[bi] <= 0) { if (break_tile_index) break_tile_index->push_back(0); ++bi; }
    while (r < n) {
        const int64_t next_break = bi < breaks.size() ? breaks[bi] : n;
        const int64_t base = ip[r] & ~(int64_t)7;
        int64_t e = r;
        while (e < n && e < next_break && (e - r) < row_cap && (ip[e + 1] - base) <= cap) ++e;
        if (e == r) return false;            // a single row does not fit
        if (row_align > 1 && e < n && e < next_break) {      // keep every tile start a multiple of row_align
            const int64_t ea = e - (e % row_align);
            if (ea <= r) return false;
            e = ea;
        }
        tiles.push_back((int32_t)r);
        r = e;
        while (bi < breaks.size() && breaks[bi] <= r) {
            if (break_tile_index) break_tile_index->push_back((int32_t)tiles.size());
            ++bi;
        }
    }
    tiles.push_back((int32_t)n);
    while (bi < breaks.size()) { if (break_tile_index) break_tile_index->push_back((int32_t)tiles.size() - 1); ++bi; }
    return true;
}

}  // namespace mgb

// ------------------------------------------------------------------------------------------------
// mgb_host_* ABI: the same routines, callable without a device (CPU tests compare these artefacts
// bit for bit with the oracle).
// ------------------------------------------------------------------------------------------------
using namespace mgb;

static bool wrap_csr(HostCsr& A, int64_t n, const int64_t* indptr, const int32_t* indices, const double* values)
{
    if (n < 0 || !indptr) return false;
    return import_csr(A, n, n, indptr[n], indptr, 8, indices, values).empty();
}

extern "C" int mgb_host_build_rj(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values, int reversed,
                                 int64_t* rj_nnz, int32_t* rj_indptr, int32_t* rj_indices, double* rj_values, double* dinv)
{
    HostCsr A, RJ; std::vector<double> d;
    if (!wrap_csr(A, n, indptr, indices, values) || !rj_nnz) return MGB_ERR_INVALID;
    bool ok = build_rj(A, reversed != 0, RJ, d);
    *rj_nnz = RJ.nnz();
    if (rj_indices) {
        if (!rj_indptr || !rj_values || !dinv) return MGB_ERR_INVALID;
        for (int64_t i = 0; i <= n; ++i) rj_indptr[i] = (int32_t)RJ.ip[i];
        std::copy(RJ.ix.begin(), RJ.ix.end(), rj_indices);
        std::copy(RJ.ax.begin(), RJ.ax.end(), rj_values);
        std::copy(d.begin(), d.end(), dinv);
    }
    return ok ? MGB_OK : MGB_ERR_SINGULAR;
}

static int artefact_out(const std::vector<int32_t>& key, const std::vector<int32_t>& order, const std::vector<int32_t>& offsets,
                        int32_t* key_out, int32_t* order_out, int64_t* count, int32_t* offsets_out, int64_t cap)
{
    if (count) *count = (int64_t)offsets.size() - 1;
    if (key_out) std::copy(key.begin(), key.end(), key_out);
    if (order_out) std::copy(order.begin(), order.end(), order_out);
    if (offsets_out) {
        if (cap < (int64_t)offsets.size()) return MGB_ERR_INVALID;
        std::copy(offsets.begin(), offsets.end(), offsets_out);
    }
    return MGB_OK;
}

extern "C" int mgb_host_level_sets(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                   int32_t* level_of_row, int32_t* order, int64_t* nlevels, int32_t* offsets, int64_t offsets_capacity)
{
    HostCsr A;
    if (!wrap_csr(A, n, indptr, indices, values)) return MGB_ERR_INVALID;
    std::vector<int32_t> lev, ord, off;
    level_sets(A, lev, ord, off);
    return artefact_out(lev, ord, off, level_of_row, order, nlevels, offsets, offsets_capacity);
}

extern "C" int mgb_host_colouring(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                  int32_t* colour_of_row, int32_t* order, int64_t* ncolours, int32_t* offsets, int64_t offsets_capacity)
{
    HostCsr A;
    if (!wrap_csr(A, n, indptr, indices, values)) return MGB_ERR_INVALID;
    std::vector<int32_t> col, ord, off;
    greedy_colouring(A, col, ord, off);
    return artefact_out(col, ord, off, colour_of_row, order, ncolours, offsets, offsets_capacity);
}

extern "C" int mgb_host_dense_inverse(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values, double* inv_row_major)
{
    HostCsr A;
    if (!wrap_csr(A, n, indptr, indices, values) || !inv_row_major) return MGB_ERR_INVALID;
    std::vector<double> inv;
    if (!dense_inverse(A, inv)) return MGB_ERR_SINGULAR;
    std::copy(inv.begin(), inv.end(), inv_row_major);
    return MGB_OK;
}

extern "C" int mgb_host_make_tiles(int64_t n, const int64_t* indptr, int64_t cap, int64_t row_cap, int nbreaks, const int32_t* breaks,
                                   int64_t row_align, int32_t* tiles, int64_t tiles_capacity, int64_t* ntiles, int32_t* break_tile)
{
    if (n < 0 || !indptr || !ntiles || nbreaks < 0 || (nbreaks > 0 && !breaks)) return MGB_ERR_INVALID;
    std::vector<int64_t> ip(indptr, indptr + n + 1);
    std::vector<int32_t> br(breaks, breaks + nbreaks), t, bt;
    if (!make_tiles(ip, cap, row_cap, br, t, &bt, row_align)) return MGB_ERR_UNSUPPORTED;      // a row does not fit / cannot be aligned
    *ntiles = (int64_t)t.size() - 1;
    if (tiles) {
        if (tiles_capacity < (int64_t)t.size()) return MGB_ERR_INVALID;
        std::copy(t.begin(), t.end(), tiles);
    }
    if (break_tile) std::copy(bt.begin(), bt.end(), break_tile);
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Lossless operator coding on the host: the DEFINITION of the artefact.  mgb_finalize builds the same
// dictionaries and codes on the device (csrc/mgb_code.cuh: hash tables instead of ordered maps, then an exact
// verification pass); this routine states what they must be, in plain C++, so that CPU tests can pin the definition
// against an independent numpy restatement (oracle/coding.py) without a GPU.
//   mode 3  row patterns : operators with ncols >= nrows whose rows repeat.  A pattern = a row's list of
//           (col - row, value bits); patterns are numbered by their first row; <= 256 patterns; each pattern's entries
//           are padded to a multiple of 8 (at least 8) with copies of its last entry ((0, 0.0) for an empty row); the
//           padded table holds <= 2048 entries.  codes[i] = pattern of row i.
//   mode 1  pair codes   : <= 256 distinct values (by bit pattern, all-ones excluded), <= 256 distinct (col - row)
//           (0x80808080 excluded), <= 256 distinct pairs.  Pairs are numbered by (offset rank, value rank), offsets
//           ascending, values ascending as unsigned 64-bit patterns.  codes[k] = pair of stored entry k.
//   mode 2  value codes  : <= 256 distinct values; codes[k] = rank of the value's bit pattern.
//   mode 0  otherwise (also: empty operators, more than 24 stored entries per row on average).
// ------------------------------------------------------------------------------------------------
namespace {
struct HostDictEnt { double val; int32_t delta; int32_t pad; };
static_assert(sizeof(HostDictEnt) == 16, "same layout as the device dictionary entry");
uint64_t bits_of(double v) { uint64_t b; std::memcpy(&b, &v, sizeof b); return b; }
}  // namespace

extern "C" int mgb_host_code_operator(int64_t nrows, int64_t ncols, const int64_t* indptr, const int32_t* indices, const double* values,
                                      int allow_patterns, int* mode_out, int* ndict_out, uint8_t* codes, void* table,
                                      int* table_entries, int32_t* pattern_head)
{
    if (nrows < 0 || ncols < 0 || !indptr || !mode_out) return MGB_ERR_INVALID;
    const int64_t nnz = indptr[nrows];
    if (nnz > 0 && (!indices || !values)) return MGB_ERR_INVALID;
    *mode_out = 0;
    if (ndict_out) *ndict_out = 0;
    if (table_entries) *table_entries = 0;
    if (nrows == 0 || nnz == 0) return MGB_OK;
    const double avg = (double)nnz / (double)nrows;
    HostDictEnt* tab = (HostDictEnt*)table;
    // ---- row patterns: columns measured from the row index (mode 3), then from the row's first stored column (mode 4)
    for (int anchored = 0; anchored <= 1; ++anchored) {
        if (anchored ? (allow_patterns < 2 || avg > 32.0) : (allow_patterns < 1 || ncols < nrows || avg > 24.0)) continue;
        typedef std::vector<std::pair<int32_t, uint64_t>> Row;
        std::map<Row, int64_t> first;                       // pattern -> first row showing it
        bool ok = true;
        Row key;
        auto row_key = [&](int64_t i) {
            key.clear();
            const int64_t base = anchored ? (indptr[i + 1] > indptr[i] ? (int64_t)indices[indptr[i]] : 0) : i;
            for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) key.push_back({(int32_t)(indices[k] - base), bits_of(values[k])});
        };
        for (int64_t i = 0; i < nrows && ok; ++i) {
            row_key(i);
            first.emplace(key, i);                          // rows ascend: the first insertion is the smallest row
            ok = first.size() <= 256;
        }
        if (ok) {
            std::vector<std::pair<int64_t, const Row*>> order;
            for (auto& kv : first) order.push_back({kv.second, &kv.first});
            std::sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            int total = 0;
            for (auto& o : order) total += std::max<int>(8, ((int)o.second->size() + 7) / 8 * 8);
            if (total <= 2048) {
                if (codes && tab && pattern_head) {
                    std::map<Row, int> id;
                    int off = 0;
                    for (int p = 0; p < 256; ++p) { pattern_head[2 * p] = 0; pattern_head[2 * p + 1] = 0; }
                    for (size_t p = 0; p < order.size(); ++p) {
                        const Row& r = *order[p].second;
                        const int len = (int)r.size(), padded = std::max(8, (len + 7) / 8 * 8);
                        pattern_head[2 * p] = off; pattern_head[2 * p + 1] = len;
                        for (int e = 0; e < padded; ++e) {
                            HostDictEnt d{0.0, 0, 0};
                            if (len > 0) { const auto& q = r[(size_t)std::min(e, len - 1)]; std::memcpy(&d.val, &q.second, 8); d.delta = q.first; }
                            tab[off + e] = d;
                        }
                        off += padded;
                        id[r] = (int)p;
                    }
                    for (int64_t i = 0; i < nrows; ++i) {
                        row_key(i);
                        codes[i] = (uint8_t)id[key];
                    }
                }
                *mode_out = anchored ? 4 : 3;
                if (ndict_out) *ndict_out = (int)order.size();
                if (table_entries) *table_entries = total;
                return MGB_OK;
            }
        }
    }
    if (avg > 24.0) return MGB_OK;
    // ---- per-entry codes
    std::set<uint64_t> vs;
    std::set<int32_t> dset;
    bool dfail = false;
    for (int64_t i = 0; i < nrows; ++i)
        for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
            const uint64_t b = bits_of(values[k]);
            if (b == ~(uint64_t)0) return MGB_OK;           // the device tables use this pattern as "empty"
            vs.insert(b);
            if (vs.size() > 256) return MGB_OK;
            if (!dfail) {
                const int32_t d = (int32_t)(indices[k] - i);
                if (d == (int32_t)0x80808080) dfail = true; else dset.insert(d);
                if (dset.size() > 256) dfail = true;
            }
        }
    std::vector<uint64_t> V(vs.begin(), vs.end());          // ascending as unsigned bit patterns
    std::vector<int32_t> Dl(dset.begin(), dset.end());
    auto vrank = [&](double v) { return (int)(std::lower_bound(V.begin(), V.end(), bits_of(v)) - V.begin()); };
    bool pair = !dfail && !Dl.empty();
    std::vector<int> lut;
    int npair = 0;
    if (pair) {
        std::vector<unsigned char> present(65536, 0);
        for (int64_t i = 0; i < nrows; ++i)
            for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
                const int di = (int)(std::lower_bound(Dl.begin(), Dl.end(), (int32_t)(indices[k] - i)) - Dl.begin());
                present[(size_t)(di * 256 + vrank(values[k]))] = 1;
            }
        lut.assign(65536, -1);
        for (int key = 0; key < 65536; ++key) if (present[(size_t)key]) lut[(size_t)key] = npair++;
        if (npair > 256) pair = false;
    }
    if (tab) for (int k = 0; k < 256; ++k) tab[k] = HostDictEnt{0.0, 0, 0};
    if (pair) {
        if (tab)
            for (int key = 0; key < 65536; ++key)
                if (lut[(size_t)key] >= 0) { HostDictEnt d{0.0, Dl[(size_t)(key >> 8)], 0}; std::memcpy(&d.val, &V[(size_t)(key & 255)], 8); tab[lut[(size_t)key]] = d; }
        if (codes)
            for (int64_t i = 0; i < nrows; ++i)
                for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
                    const int di = (int)(std::lower_bound(Dl.begin(), Dl.end(), (int32_t)(indices[k] - i)) - Dl.begin());
                    codes[k] = (uint8_t)lut[(size_t)(di * 256 + vrank(values[k]))];
                }
        *mode_out = 1;
        if (ndict_out) *ndict_out = npair;
    } else {
        if (tab) for (size_t k = 0; k < V.size(); ++k) { HostDictEnt d{0.0, 0, 0}; std::memcpy(&d.val, &V[k], 8); tab[k] = d; }
        if (codes)
            for (int64_t i = 0; i < nrows; ++i)
                for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) codes[k] = (uint8_t)vrank(values[k]);
        *mode_out = 2;
        if (ndict_out) *ndict_out = (int)V.size();
    }
    if (table_entries) *table_entries = 256;
    return MGB_OK;
}
