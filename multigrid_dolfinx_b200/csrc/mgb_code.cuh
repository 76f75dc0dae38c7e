// Lossless dictionary coding of device-resident CSR operators (included by mgb_engine.cu inside its anonymous
// namespace).  Set-up work only -- the coded copy is consumed by k_rowstream (mgb_kernels.cuh).
//
// The operators the reference hands over come from uniform meshes (Multigrid_prototype.py:63-66), so their stored
// entries repeat: few distinct values (R_omega of the P1 Laplacian holds a single one) and, for square operators in a
// banded numbering, few distinct column offsets.  try_encode() discovers that ON THE DEVICE, from the arrays as they
// are (nothing is assumed about geometry):
//   1. k_code_collect   distinct value bit patterns / distinct (col - row) into two small lock-free hash tables
//                       (atomicCAS on the key itself); more than 256 of either -> that dictionary is abandoned;
//   2. k_code_pairs     which (offset, value) combinations occur (byte map over 256 x 256);
//   3. k_code_encode    one byte per stored entry;
//   4. k_code_verify    decodes every entry again and compares column and value BITS with the CSR arrays; any
//                       mismatch discards the coding (the operator then simply runs through the uncoded kernels).
// Dictionaries are sorted (values by bit pattern, offsets ascending, pairs by (offset, value) index), so the coding
// is a deterministic function of the operator.
//
// Row patterns (mode 3, tried first for operators with at least as many columns as rows): on a uniform mesh whole ROWS
// repeat -- the list of (col - row, value) pairs of an interior row is the same for every interior row.  try_patterns()
// hashes every row's list on the device (k_pat_collect, lock-free table keyed by the 64-bit hash, smallest row index kept
// as the representative), builds the pattern table from the representatives' actual entries (patterns numbered by
// representative row, each padded to a multiple of 8 entries), codes every row as one byte (k_pat_encode) and verifies
// every entry of every row against the CSR arrays (k_pat_verify) -- a hash collision can therefore only cost the
// coding, never a wrong result.  The kernel then needs neither row pointers nor per-entry codes.
#pragma once

constexpr int CODE_SLOTS = 1024;                     // hash-table slots (power of two, > 256)
constexpr unsigned long long CODE_VEMPTY = ~0ULL;    // an all-ones NaN never appears as a stored value we accept
constexpr int CODE_DEMPTY = (int)0x80808080;         // memset-able "no offset" marker

__device__ __forceinline__ unsigned code_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (unsigned)k;
}

// flags: [0] values abandoned, [1] offsets abandoned, [2] distinct values, [3] distinct offsets
__device__ __forceinline__ void code_insert_value(unsigned long long* tbl, int* flags, unsigned long long key)
{
    if (key == CODE_VEMPTY) { flags[0] = 1; return; }
    unsigned s = code_hash(key) & (CODE_SLOTS - 1);
    for (int p = 0; p < CODE_SLOTS; ++p) {
        unsigned long long cur = tbl[s];                 // may be a stale "empty" from L1: the CAS below is authoritative
        if (cur == key) return;
        if (cur == CODE_VEMPTY) {
            cur = atomicCAS(tbl + s, CODE_VEMPTY, key);
            if (cur == key) return;
            if (cur == CODE_VEMPTY) { if (atomicAdd(flags + 2, 1) >= 256) flags[0] = 1; return; }
        }
        s = (s + 1) & (CODE_SLOTS - 1);
    }
    flags[0] = 1;
}
__device__ __forceinline__ void code_insert_delta(int* tbl, int* flags, int key)
{
    if (key == CODE_DEMPTY) { flags[1] = 1; return; }
    unsigned s = code_hash((unsigned long long)(unsigned)key) & (CODE_SLOTS - 1);
    for (int p = 0; p < CODE_SLOTS; ++p) {
        int cur = tbl[s];
        if (cur == key) return;
        if (cur == CODE_DEMPTY) {
            cur = atomicCAS(tbl + s, CODE_DEMPTY, key);
            if (cur == key) return;
            if (cur == CODE_DEMPTY) { if (atomicAdd(flags + 3, 1) >= 256) flags[1] = 1; return; }
        }
        s = (s + 1) & (CODE_SLOTS - 1);
    }
    flags[1] = 1;
}

__global__ void __launch_bounds__(256)
k_code_collect(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
               unsigned long long* vtbl, int* dtbl, int* flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    volatile int* vf = flags;
    if (vf[0]) return;
    const bool want_delta = vf[1] == 0;
    unsigned long long last_v = CODE_VEMPTY;
    int last_d = CODE_DEMPTY;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const unsigned long long vb = (unsigned long long)__double_as_longlong(vals[k]);
        if (vb != last_v) { code_insert_value(vtbl, flags, vb); last_v = vb; }
        if (want_delta) {
            const int dl = cols[k] - i;
            if (dl != last_d) { code_insert_delta(dtbl, flags, dl); last_d = dl; }
        }
    }
}

struct CodeDicts {                       // sorted dictionaries, passed by value (3 KB) and staged in shared memory
    unsigned long long v[256];
    int d[256];
    int nv, nd;
};

__device__ __forceinline__ int code_find_value(const unsigned long long* v, int nv, unsigned long long key)
{
    int lo = 0, hi = nv;                 // first index with v[idx] >= key
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;                            // (caller guarantees presence; k_code_verify catches anything else)
}
__device__ __forceinline__ int code_find_delta(const int* d, int nd, int key)
{
    int lo = 0, hi = nd;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (d[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// mode 1: lut == nullptr -> mark present[offset index * 256 + value index]; else codes[k] = lut[...]
// mode 2: codes[k] = value index
__global__ void __launch_bounds__(256)
k_code_encode(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
              const CodeDicts* __restrict__ dicts, int mode, unsigned char* __restrict__ present,
              const unsigned char* __restrict__ lut, unsigned char* __restrict__ codes)
{
    __shared__ unsigned long long sv[256];
    __shared__ int sd[256];
    sv[threadIdx.x] = dicts->v[threadIdx.x];
    sd[threadIdx.x] = dicts->d[threadIdx.x];
    const int nv = dicts->nv, nd = dicts->nd;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const int vi = code_find_value(sv, nv, (unsigned long long)__double_as_longlong(vals[k])) & 255;
        if (mode == 2) { codes[k] = (unsigned char)vi; continue; }
        const int di = code_find_delta(sd, nd, cols[k] - i) & 255;
        if (lut) codes[k] = lut[di * 256 + vi];
        else present[di * 256 + vi] = 1;
    }
}

__global__ void __launch_bounds__(256)
k_code_verify(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
              const unsigned char* __restrict__ codes, const DictEnt* __restrict__ dict, int mode, int* __restrict__ bad)
{
    __shared__ DictEnt sdict[256];
    sdict[threadIdx.x] = dict[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = true;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const DictEnt de = sdict[codes[k]];
        ok = ok && __double_as_longlong(de.val) == __double_as_longlong(vals[k]);
        if (mode == 1) ok = ok && (i + de.delta == cols[k]);
    }
    if (!ok) *bad = 1;
}

void free_coded(Coded& c)
{
    cudaFree(c.codes); cudaFree(c.dict); cudaFree(c.phead); cudaFree(c.dict_win); cudaFree(c.pmask); cudaFree(c.anchor);
    c = Coded();
}

// ---- row patterns --------------------------------------------------------------------------------------
constexpr int PAT_MAX_ENTRIES = 2048;                // pattern-table entries the kernel keeps in shared memory (32 KB)
constexpr int PAT_REP_EMPTY = 0x7F7F7F7F;

// base: what a row's columns are measured from -- the row index (square operators, mode 3) or the row's first stored column
// (anchored patterns, mode 4: rectangular operators such as the prolongation, whose columns do not follow the row index)
__device__ __forceinline__ int pat_base(int i, bool anchored, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols)
{
    if (!anchored) return i;
    return rp[i + 1] > rp[i] ? cols[rp[i]] : 0;
}
__device__ __forceinline__ unsigned long long pat_row_hash(int i, bool anchored, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols,
                                                           const double* __restrict__ vals)
{
    const int a = rp[i], b = rp[i + 1];
    const int base = pat_base(i, anchored, rp, cols);
    unsigned long long h = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)(unsigned)(b - a);
    for (int k = a; k < b; ++k) {
        h ^= (unsigned long long)(unsigned)(cols[k] - base);
        h *= 0xff51afd7ed558ccdULL; h ^= h >> 32;
        h ^= (unsigned long long)__double_as_longlong(vals[k]);
        h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 29;
    }
    return h == CODE_VEMPTY ? h - 1 : h;
}

// flags: [0] abandoned (more than 256 distinct hashes), [2] distinct hashes
__global__ void __launch_bounds__(256)
k_pat_collect(int n, bool anchored, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
              unsigned long long* htbl, int* rep, int* flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (((volatile int*)flags)[0]) return;
    const unsigned long long key = pat_row_hash(i, anchored, rp, cols, vals);
    unsigned s = code_hash(key) & (CODE_SLOTS - 1);
    for (int p = 0; p < CODE_SLOTS; ++p) {
        unsigned long long cur = htbl[s];
        if (cur == CODE_VEMPTY) {
            cur = atomicCAS(htbl + s, CODE_VEMPTY, key);
            if (cur == CODE_VEMPTY) { if (atomicAdd(flags + 2, 1) >= 256) flags[0] = 1; cur = key; }
        }
        if (cur == key) {                                 // representative = smallest row; rows arrive roughly in order, so a
            if (i < ((volatile int*)rep)[s]) atomicMin(rep + s, i);      // plain read filters almost every atomic (a stale value is only larger)
            return;
        }
        s = (s + 1) & (CODE_SLOTS - 1);
    }
    flags[0] = 1;
}

struct PatDicts { unsigned long long h[256]; int id[256]; int n; };      // hashes ascending -> pattern number

__global__ void __launch_bounds__(256)
k_pat_encode(int n, bool anchored, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
             const PatDicts* __restrict__ pd, unsigned char* __restrict__ rcodes, int32_t* __restrict__ anchor)
{
    __shared__ unsigned long long sh[256];
    __shared__ int sid[256];
    sh[threadIdx.x] = pd->h[threadIdx.x];
    sid[threadIdx.x] = pd->id[threadIdx.x];
    const int np = pd->n;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = code_find_value(sh, np, pat_row_hash(i, anchored, rp, cols, vals)) & 255;
    rcodes[i] = (unsigned char)sid[pos];
    if (anchored) anchor[i] = pat_base(i, true, rp, cols);
}

__global__ void __launch_bounds__(256)
k_pat_verify(int n, const int32_t* __restrict__ anchor, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
             const unsigned char* __restrict__ rcodes, const int2* __restrict__ phead, const DictEnt* __restrict__ pent, int* __restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 ph = phead[rcodes[i]];
    const int a = rp[i], len = rp[i + 1] - a;
    const int base = anchor ? anchor[i] : i;                 // (decoded exactly as the kernels decode it: from the stored anchor)
    bool ok = ph.y == len;
    for (int e = 0; ok && e < len; ++e) {
        const DictEnt de = pent[ph.x + e];
        ok = (base + de.delta == cols[a + e]) && __double_as_longlong(de.val) == __double_as_longlong(vals[a + e]);
    }
    if (!ok) *bad = 1;
}

// occurrences of every row code (grid-stride, per-block histogram in shared memory)
__global__ void __launch_bounds__(256)
k_code_hist(int n, const unsigned char* __restrict__ rcodes, unsigned long long* __restrict__ hist)
{
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[rcodes[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// Row-pattern coding of D (mode 3).  ip: the operator's row pointers on the host.  Leaves D.cd.mode == 0 when it does not apply.
// anchored: columns measured from each row's first stored column (mode 4) instead of from the row index (mode 3).
int try_patterns(mgb_handle* h, DevCsr& D, const std::vector<int64_t>& ip, bool anchored)
{
    free_coded(D.cd);
    if (h->compress < 2 || D.nrows <= 0 || D.nnz <= 0 || (!anchored && D.ncols < D.nrows)) return MGB_OK;
    if (anchored && h->compress < 3) return MGB_OK;
    if ((double)D.nnz / (double)D.nrows > (anchored ? 32.0 : 24.0)) return MGB_OK;
    const int n = (int)D.nrows;
    const int grid = (n + 255) / 256;
    unsigned long long* htbl = nullptr; int* rep = nullptr; int* flags = nullptr; PatDicts* dpd = nullptr; int* bad = nullptr;
    auto cleanup = [&] { cudaFree(htbl); cudaFree(rep); cudaFree(flags); cudaFree(dpd); cudaFree(bad); };
#define CUC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) { cleanup(); free_coded(D.cd); return fail(h, MGB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } \
    } while (0)
    CUC(cudaMalloc((void**)&htbl, CODE_SLOTS * sizeof(unsigned long long)));
    CUC(cudaMalloc((void**)&rep, CODE_SLOTS * sizeof(int)));
    CUC(cudaMalloc((void**)&flags, 4 * sizeof(int)));
    CUC(cudaMemsetAsync(htbl, 0xFF, CODE_SLOTS * sizeof(unsigned long long), h->stream));
    CUC(cudaMemsetAsync(rep, 0x7F, CODE_SLOTS * sizeof(int), h->stream));
    CUC(cudaMemsetAsync(flags, 0, 4 * sizeof(int), h->stream));
    k_pat_collect<<<grid, 256, 0, h->stream>>>(n, anchored, D.rowptr, D.cols, D.vals, htbl, rep, flags);
    CUC(cudaGetLastError());
    std::vector<unsigned long long> hh(CODE_SLOTS);
    std::vector<int> hr(CODE_SLOTS);
    int hf[4] = {0, 0, 0, 0};
    CUC(cudaMemcpyAsync(hh.data(), htbl, CODE_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaMemcpyAsync(hr.data(), rep, CODE_SLOTS * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaMemcpyAsync(hf, flags, sizeof hf, cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaStreamSynchronize(h->stream));
    std::vector<std::pair<int, unsigned long long>> pats;               // (representative row, hash)
    for (int s = 0; s < CODE_SLOTS; ++s)
        if (hh[(size_t)s] != CODE_VEMPTY && hr[(size_t)s] != PAT_REP_EMPTY) pats.push_back({hr[(size_t)s], hh[(size_t)s]});
    if (hf[0] || pats.empty() || pats.size() > 256) { cleanup(); return MGB_OK; }
    std::sort(pats.begin(), pats.end());                                  // patterns numbered by representative row
    std::vector<int2> phead(256, make_int2(0, 0));
    int total = 0;
    for (size_t p = 0; p < pats.size(); ++p) {
        const int r = pats[p].first;
        if (r < 0 || r >= n) { cleanup(); return MGB_OK; }
        const int len = (int)(ip[(size_t)r + 1] - ip[(size_t)r]);
        phead[p] = make_int2(total, len);
        total += std::max(8, (len + 7) / 8 * 8);
    }
    if (total > PAT_MAX_ENTRIES) { cleanup(); return MGB_OK; }
    std::vector<DictEnt> pent((size_t)total, DictEnt{0.0, 0, 0});
    std::vector<int32_t> rc; std::vector<double> rv;
    for (size_t p = 0; p < pats.size(); ++p) {
        const int r = pats[p].first, len = phead[p].y, off = phead[p].x;
        rc.resize((size_t)len); rv.resize((size_t)len);
        if (len > 0) {
            CUC(cudaMemcpyAsync(rc.data(), D.cols + ip[(size_t)r], sizeof(int32_t) * (size_t)len, cudaMemcpyDeviceToHost, h->stream));
            CUC(cudaMemcpyAsync(rv.data(), D.vals + ip[(size_t)r], sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, h->stream));
            CUC(cudaStreamSynchronize(h->stream));
        }
        const int base = anchored ? (len > 0 ? rc[0] : 0) : r;
        for (int e = 0; e < len; ++e) pent[(size_t)(off + e)] = DictEnt{rv[(size_t)e], rc[(size_t)e] - base, 0};
        const int padded = std::max(8, (len + 7) / 8 * 8);
        for (int e = len; e < padded; ++e)                                // copies of the last entry; (0, 0.0) for an empty row
            pent[(size_t)(off + e)] = len > 0 ? pent[(size_t)(off + len - 1)] : DictEnt{0.0, 0, 0};
    }
    PatDicts pd{};
    {
        std::vector<std::pair<unsigned long long, int>> byhash;
        for (size_t p = 0; p < pats.size(); ++p) byhash.push_back({pats[p].second, (int)p});
        std::sort(byhash.begin(), byhash.end());
        pd.n = (int)byhash.size();
        for (int k = 0; k < 256; ++k) { pd.h[k] = k < pd.n ? byhash[(size_t)k].first : ~0ULL; pd.id[k] = k < pd.n ? byhash[(size_t)k].second : 0; }
    }
    CUC(cudaMalloc((void**)&dpd, sizeof(PatDicts)));
    CUC(cudaMemcpyAsync(dpd, &pd, sizeof pd, cudaMemcpyHostToDevice, h->stream));
    CUC(cudaMalloc((void**)&D.cd.phead, 256 * sizeof(int2)));
    CUC(cudaMemcpyAsync(D.cd.phead, phead.data(), 256 * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUC(cudaMalloc((void**)&D.cd.dict, (size_t)total * sizeof(DictEnt)));
    CUC(cudaMemcpyAsync(D.cd.dict, pent.data(), (size_t)total * sizeof(DictEnt), cudaMemcpyHostToDevice, h->stream));
    const size_t cbytes = ((size_t)n + 15) / 16 * 16 + 64;
    CUC(cudaMalloc((void**)&D.cd.codes, cbytes));
    CUC(cudaMemsetAsync(D.cd.codes, 0, cbytes, h->stream));
    if (anchored) CUC(cudaMalloc((void**)&D.cd.anchor, ((size_t)n + 64) * sizeof(int32_t)));
    if (anchored) CUC(cudaMemsetAsync(D.cd.anchor, 0, ((size_t)n + 64) * sizeof(int32_t), h->stream));
    k_pat_encode<<<grid, 256, 0, h->stream>>>(n, anchored, D.rowptr, D.cols, D.vals, dpd, D.cd.codes, D.cd.anchor);
    CUC(cudaGetLastError());
    CUC(cudaMalloc((void**)&bad, sizeof(int)));
    CUC(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    k_pat_verify<<<grid, 256, 0, h->stream>>>(n, (const int32_t*)D.cd.anchor, D.rowptr, D.cols, D.vals, D.cd.codes, D.cd.phead, D.cd.dict, bad);
    CUC(cudaGetLastError());
    int hb = 0;
    CUC(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaStreamSynchronize(h->stream));
    cleanup();
#undef CUC
    if (hb) { free_coded(D.cd); return MGB_OK; }          // never trust an unverified coding
    D.cd.mode = anchored ? 4 : 3; D.cd.ndict = (int)pats.size(); D.cd.npent = total;
    if (anchored) return MGB_OK;
    {   // the most frequent ("hot") pattern -- exact histogram of the row codes (a strided sample can fall on boundary rows
        // only: 513^3 rows sampled 1024 times land on multiples of 513) -- travels to the hot-row / row-window kernels as
        // kernel parameters
        unsigned long long* dh = nullptr;
        unsigned long long hist[256];
        cudaError_t e1 = cudaMalloc((void**)&dh, 256 * sizeof(unsigned long long));
        if (e1 == cudaSuccess) e1 = cudaMemsetAsync(dh, 0, 256 * sizeof(unsigned long long), h->stream);
        if (e1 == cudaSuccess) {
            k_code_hist<<<std::min(grid, 1024), 256, 0, h->stream>>>(n, D.cd.codes, dh);
            e1 = cudaGetLastError();
        }
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(hist, dh, sizeof hist, cudaMemcpyDeviceToHost, h->stream);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(h->stream);
        cudaFree(dh);
        if (e1 != cudaSuccess) { free_coded(D.cd); return fail(h, MGB_ERR_CUDA, "histogram of the row codes failed: %s", cudaGetErrorString(e1)); }
        HotPlan& H = D.cd.hotplan;
        H = HotPlan{};
        H.hot = (int)(std::max_element(hist, hist + 256) - hist);
        H.hotlen = phead[(size_t)H.hot].y;
        for (int e = 0; e < WIN_HOT; ++e) {
            const bool in = e < H.hotlen && H.hotlen <= WIN_HOT;
            H.hd[e] = in ? pent[(size_t)(phead[(size_t)H.hot].x + e)].delta : 0;
            H.hv[e] = in ? pent[(size_t)(phead[(size_t)H.hot].x + e)].val : 0.0;
        }
    }
    if (D.cd.hotplan.hotlen >= 1 && D.cd.hotplan.hotlen <= WIN_HOT) {
        // hot-row kernel (k_hotrow): every pattern that is a subsequence of the hot one -- same (col - row, value bits) pairs in
        // the same order, some missing -- becomes a bit mask over the hot entries; anything else is marked for the table walk
        const HotPlan& HP = D.cd.hotplan;
        std::vector<uint32_t> pm(256, 0u);
        int slow = 0;
        for (size_t p = 0; p < pats.size(); ++p) {
            uint32_t m = 0;
            int at = 0;                                                     // next hot entry that may match
            bool sub = true;
            for (int e = 0; e < phead[p].y && sub; ++e) {
                const DictEnt& d = pent[(size_t)(phead[p].x + e)];
                unsigned long long vb, hb2;
                std::memcpy(&vb, &d.val, 8);
                while (at < HP.hotlen) {
                    std::memcpy(&hb2, &HP.hv[at], 8);
                    if (HP.hd[at] == d.delta && hb2 == vb) break;
                    ++at;
                }
                if (at == HP.hotlen) sub = false;
                else { m |= 1u << at; ++at; }
            }
            if (!sub) { m = HOT_SLOW; ++slow; }
            pm[p] = m;
        }
        HotArgs A{};
        A.dmin = 0; A.dmax = 0;
        for (int e = 0; e < WIN_HOT; ++e) {
            A.hd[e] = HP.hd[e]; A.hv[e] = HP.hv[e];
            if (e < HP.hotlen) { A.dmin = std::min(A.dmin, HP.hd[e]); A.dmax = std::max(A.dmax, HP.hd[e]); }
        }
        cudaError_t e1 = cudaMalloc((void**)&D.cd.pmask, 256 * sizeof(uint32_t));
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(D.cd.pmask, pm.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(h->stream);
        if (e1 == cudaSuccess) { D.cd.hot = A; D.cd.hot_ok = true; D.cd.hot_slow = slow; }
        else { cudaFree(D.cd.pmask); D.cd.pmask = nullptr; (void)cudaGetLastError(); }
    }
    if (h->stage_x == 1) {
        // row-window kernel (k_rowwin): merge the distinct offsets into windows and re-express every table entry as
        // (window, offset inside the window)
        std::vector<int> offs;
        for (size_t p = 0; p < pats.size(); ++p)
            for (int e = 0; e < phead[p].y; ++e) offs.push_back(pent[(size_t)(phead[p].x + e)].delta);
        offs.push_back(0);                                                  // x[row] itself: the aliased epilogue operand
        std::sort(offs.begin(), offs.end());
        offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
        WinPlan G{};
        bool ok = true;
        for (size_t k = 0; k < offs.size() && ok; ++k) {
            if (G.ng > 0 && offs[k] - G.gmin[G.ng - 1] <= WIN_SPAN) { G.gspan[G.ng - 1] = offs[k] - G.gmin[G.ng - 1]; continue; }
            if (G.ng == WIN_MAX) { ok = false; break; }
            G.gmin[G.ng] = offs[k] & ~1;                                    // even: slices start 16-byte aligned (tile starts are multiples of 16)
            G.gspan[G.ng] = offs[k] - G.gmin[G.ng];
            ++G.ng;
        }
        if (ok) {
            auto window_of = [&](int delta) { int g = 0; while (g + 1 < G.ng && delta >= G.gmin[g + 1]) ++g; return g; };
            std::vector<DictEnt> wt(pent);
            for (size_t p = 0; p < pats.size(); ++p) {
                const int len = phead[p].y, off = phead[p].x, padded = std::max(8, (len + 7) / 8 * 8);
                for (int e = 0; e < padded; ++e) {
                    DictEnt& d = wt[(size_t)(off + e)];
                    const int g = window_of(d.delta);                       // (empty row: delta 0, value never used)
                    d.delta = (g << WIN_GSHIFT) | (d.delta - G.gmin[g]);
                }
            }
            G.vslot = window_of(0);                                         // window holding offset 0; the slot is fixed at launch (needs the tile size)
            G.xlen = (int)((D.ncols + 16) & ~(int64_t)1);                  // engine vectors carry 16 padding entries behind ncols
            G.hot = D.cd.hotplan.hot; G.hotlen = D.cd.hotplan.hotlen;
            for (int e = 0; e < WIN_HOT; ++e) {                             // (window, offset): resolved at launch
                const bool in = e < G.hotlen && G.hotlen <= WIN_HOT;
                G.hs[e] = in ? wt[(size_t)(phead[(size_t)G.hot].x + e)].delta : 0;
                G.hv[e] = D.cd.hotplan.hv[e];
            }
            cudaError_t e1 = cudaMalloc((void**)&D.cd.dict_win, (size_t)total * sizeof(DictEnt));
            if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(D.cd.dict_win, wt.data(), (size_t)total * sizeof(DictEnt), cudaMemcpyHostToDevice, h->stream);
            if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(h->stream);
            if (e1 == cudaSuccess) D.cd.win = G;
            else { cudaFree(D.cd.dict_win); D.cd.dict_win = nullptr; (void)cudaGetLastError(); }
        }
    }
    return MGB_OK;
}

// Try to dictionary-code D (arrays already on the device).  Leaves D.cd.mode == 0 when the operator does not
// compress; only CUDA failures are errors.
int try_encode(mgb_handle* h, DevCsr& D)
{
    free_coded(D.cd);
    if (!h->compress || D.nrows <= 0 || D.nnz <= 0) return MGB_OK;
    if ((double)D.nnz / (double)D.nrows > 24.0) return MGB_OK;          // long rows: thread-per-row does not pay (P2)
    const int n = (int)D.nrows;
    const int grid = (n + 255) / 256;
    unsigned long long* vtbl = nullptr; int* dtbl = nullptr; int* flags = nullptr;
    CodeDicts* ddicts = nullptr; unsigned char* present = nullptr; unsigned char* lut = nullptr; int* bad = nullptr;
    auto cleanup = [&] { cudaFree(vtbl); cudaFree(dtbl); cudaFree(flags); cudaFree(ddicts); cudaFree(present); cudaFree(lut); cudaFree(bad); };
#define CUC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) { cleanup(); free_coded(D.cd); return fail(h, MGB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } \
    } while (0)
    CUC(cudaMalloc((void**)&vtbl, CODE_SLOTS * sizeof(unsigned long long)));
    CUC(cudaMalloc((void**)&dtbl, CODE_SLOTS * sizeof(int)));
    CUC(cudaMalloc((void**)&flags, 4 * sizeof(int)));
    CUC(cudaMemsetAsync(vtbl, 0xFF, CODE_SLOTS * sizeof(unsigned long long), h->stream));
    CUC(cudaMemsetAsync(dtbl, 0x80, CODE_SLOTS * sizeof(int), h->stream));
    CUC(cudaMemsetAsync(flags, 0, 4 * sizeof(int), h->stream));
    k_code_collect<<<grid, 256, 0, h->stream>>>(n, D.rowptr, D.cols, D.vals, vtbl, dtbl, flags);
    CUC(cudaGetLastError());
    std::vector<unsigned long long> hv(CODE_SLOTS);
    std::vector<int> hd(CODE_SLOTS);
    int hf[4] = {0, 0, 0, 0};
    CUC(cudaMemcpyAsync(hv.data(), vtbl, CODE_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaMemcpyAsync(hd.data(), dtbl, CODE_SLOTS * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaMemcpyAsync(hf, flags, sizeof hf, cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaStreamSynchronize(h->stream));
    CodeDicts dd{};
    std::vector<unsigned long long> V;
    std::vector<int> Dl;
    for (unsigned long long x : hv) if (x != CODE_VEMPTY) V.push_back(x);
    for (int x : hd) if (x != CODE_DEMPTY) Dl.push_back(x);
    if (hf[0] || V.empty() || V.size() > 256) { cleanup(); return MGB_OK; }
    bool pair = !hf[1] && !Dl.empty() && Dl.size() <= 256;
    std::sort(V.begin(), V.end());
    std::sort(Dl.begin(), Dl.end());
    dd.nv = (int)V.size(); dd.nd = pair ? (int)Dl.size() : 0;
    for (int k = 0; k < 256; ++k) { dd.v[k] = k < dd.nv ? V[(size_t)k] : ~0ULL; dd.d[k] = k < dd.nd ? Dl[(size_t)k] : INT_MAX; }
    CUC(cudaMalloc((void**)&ddicts, sizeof(CodeDicts)));
    CUC(cudaMemcpyAsync(ddicts, &dd, sizeof dd, cudaMemcpyHostToDevice, h->stream));
    std::vector<DictEnt> dict(256, DictEnt{0.0, 0, 0});
    std::vector<unsigned char> hlut;
    int ndict = 0;
    if (pair) {
        CUC(cudaMalloc((void**)&present, 65536));
        CUC(cudaMemsetAsync(present, 0, 65536, h->stream));
        k_code_encode<<<grid, 256, 0, h->stream>>>(n, D.rowptr, D.cols, D.vals, ddicts, 1, present, nullptr, nullptr);
        CUC(cudaGetLastError());
        std::vector<unsigned char> hp(65536);
        CUC(cudaMemcpyAsync(hp.data(), present, 65536, cudaMemcpyDeviceToHost, h->stream));
        CUC(cudaStreamSynchronize(h->stream));
        hlut.assign(65536, 0);
        for (int key = 0; key < 65536 && ndict <= 256; ++key) {
            if (!hp[(size_t)key]) continue;
            if (ndict < 256) {
                double val; const unsigned long long bits = V[(size_t)(key & 255)];
                std::memcpy(&val, &bits, sizeof val);
                dict[(size_t)ndict] = DictEnt{val, Dl[(size_t)(key >> 8)], 0};
                hlut[(size_t)key] = (unsigned char)ndict;
            }
            ++ndict;
        }
        if (ndict > 256) { pair = false; ndict = 0; }
    }
    const int mode = pair ? 1 : 2;
    if (!pair) {
        ndict = dd.nv;
        for (int k = 0; k < ndict; ++k) {
            double val; std::memcpy(&val, &V[(size_t)k], sizeof val);
            dict[(size_t)k] = DictEnt{val, 0, 0};
        }
    }
    CUC(cudaMalloc((void**)&D.cd.dict, 256 * sizeof(DictEnt)));
    CUC(cudaMemcpyAsync(D.cd.dict, dict.data(), 256 * sizeof(DictEnt), cudaMemcpyHostToDevice, h->stream));
    const size_t cbytes = ((size_t)D.nnz + 15) / 16 * 16 + 64;             // bulk copies read whole 16-byte groups
    CUC(cudaMalloc((void**)&D.cd.codes, cbytes));
    CUC(cudaMemsetAsync(D.cd.codes, 0, cbytes, h->stream));
    if (pair) {
        CUC(cudaMalloc((void**)&lut, 65536));
        CUC(cudaMemcpyAsync(lut, hlut.data(), 65536, cudaMemcpyHostToDevice, h->stream));
    }
    k_code_encode<<<grid, 256, 0, h->stream>>>(n, D.rowptr, D.cols, D.vals, ddicts, mode, nullptr, lut, D.cd.codes);
    CUC(cudaGetLastError());
    CUC(cudaMalloc((void**)&bad, sizeof(int)));
    CUC(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    k_code_verify<<<grid, 256, 0, h->stream>>>(n, D.rowptr, D.cols, D.vals, D.cd.codes, D.cd.dict, mode, bad);
    CUC(cudaGetLastError());
    int hb = 0;
    CUC(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUC(cudaStreamSynchronize(h->stream));
    cleanup();
#undef CUC
    if (hb) { free_coded(D.cd); return MGB_OK; }          // never trust an unverified coding
    D.cd.mode = mode; D.cd.ndict = ndict; D.cd.nvals = dd.nv; D.cd.ndeltas = (int)Dl.size();
    return MGB_OK;
}
