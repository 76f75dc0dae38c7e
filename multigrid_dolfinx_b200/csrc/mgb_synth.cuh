// Device-side setup (SURVEY 8f "next #2"): (1) the synthetic structured P1 Poisson generator -- the same
// matrices, transfers and injection lists as multigrid_dolfinx_b200/problems.py (stencil_p1 / prolongation /
// injection with lexicographic DOFs), written straight into device CSR so that a 513^3 level (2.0e9 entries)
// never exists on the host; (2) getJacobiMatrices (multigrid.py:48-56) as two kernels + a scan.
// Included by mgb_engine.cu (needs its internal types).  Input generation is a stand-in for the dolfinx/PETSc
// assembly the reference does on the host (Multigrid_prototype.py:62-118); the engine proper stays operator-agnostic.
#pragma once
// (already inside mgb_engine.cu's anonymous namespace)

struct SynthGeom {
    int dim, N, noff;
    long long n;
    double unit, diag;
    int off[15][3];          // neighbour offsets (dx, dy, dz), sorted by linear displacement
    long long lin[15];
    double w[15];            // stencil weight (in units of `unit`); the self entry has w = 0 and is handled apart
    int self;                // index of the (0,0,0) entry
};

SynthGeom make_geom(int dim, int m)
{
    SynthGeom g{};
    g.dim = dim; g.N = m + 1;
    g.n = 1; for (int d = 0; d < dim; ++d) g.n *= g.N;
    const double h = 1.0 / (double)m;
    g.unit = dim == 2 ? 1.0 : h;
    g.diag = (dim == 2 ? 4.0 : 6.0) * g.unit;
    struct E { int o[3]; long long lin; double w; };
    std::vector<E> es;
    const int lo = -1, hi = 1;
    for (int dz = (dim == 3 ? lo : 0); dz <= (dim == 3 ? hi : 0); ++dz)
        for (int dy = lo; dy <= hi; ++dy)
            for (int dx = lo; dx <= hi; ++dx) {
                const bool nonneg = dx >= 0 && dy >= 0 && dz >= 0, nonpos = dx <= 0 && dy <= 0 && dz <= 0;
                if (!nonneg && !nonpos) continue;          // cell-connectivity pattern of the "/" and Kuhn meshes
                const int nz = (dx != 0) + (dy != 0) + (dz != 0);
                E e{{dx, dy, dz}, (long long)dx + (long long)g.N * dy + (long long)g.N * g.N * dz, nz == 1 ? -1.0 : 0.0};
                es.push_back(e);
            }
    std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.lin < b.lin; });
    g.noff = (int)es.size();
    for (int k = 0; k < g.noff; ++k) {
        for (int d = 0; d < 3; ++d) g.off[k][d] = es[k].o[d];
        g.lin[k] = es[k].lin; g.w[k] = es[k].w;
        if (es[k].lin == 0) g.self = k;
    }
    return g;
}

struct ColMap {              // global column -> this rank's [owned | ghost-below | ghost-above] numbering
    long long ob, oe, glo;
    __host__ __device__ int operator()(long long c) const
    {
        if (c >= ob && c < oe) return (int)(c - ob);
        if (c < ob) return (int)((oe - ob) + (c - glo));
        return (int)((oe - ob) + (ob - glo) + (c - oe));
    }
};

__device__ __forceinline__ void lex_split(long long r, int N, int dim, int* idx)
{
    idx[0] = (int)(r % N); r /= N;
    idx[1] = (int)(r % N); r /= N;
    idx[2] = dim == 3 ? (int)r : 0;
}

__global__ void k_synth_a_count(SynthGeom g, long long row_begin, int nrows, int32_t* __restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int id[3];
    lex_split(row_begin + r, g.N, g.dim, id);
    int c = 0;
    for (int k = 0; k < g.noff; ++k) {
        bool ok = true;
        for (int d = 0; d < g.dim; ++d) { const int t = id[d] + g.off[k][d]; ok = ok && t >= 0 && t <= g.N - 1; }
        c += ok;
    }
    cnt[r] = c;
}

__global__ void k_synth_a_fill(SynthGeom g, long long row_begin, int nrows, const int32_t* __restrict__ rowptr,
                               int32_t* __restrict__ cols, double* __restrict__ vals, ColMap cm)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int id[3];
    const long long row = row_begin + r;
    lex_split(row, g.N, g.dim, id);
    bool on_bnd = false;
    for (int d = 0; d < g.dim; ++d) on_bnd = on_bnd || id[d] == 0 || id[d] == g.N - 1;
    int o = rowptr[r];
    for (int k = 0; k < g.noff; ++k) {
        bool ok = true, nb_bnd = false;
        for (int d = 0; d < g.dim; ++d) {
            const int t = id[d] + g.off[k][d];
            ok = ok && t >= 0 && t <= g.N - 1;
            nb_bnd = nb_bnd || t <= 0 || t >= g.N - 1;
        }
        if (!ok) continue;
        double v;
        if (k == g.self) v = on_bnd ? 1.0 : g.diag;
        else v = (on_bnd || nb_bnd) ? 0.0 : (g.w[k] * g.unit + 0.0);
        cols[o] = cm(row + g.lin[k]);
        vals[o] = v;
        ++o;
    }
}

// prolongation rows (multigrid.py:59-120 in matrix form, tensor-product extension in 3-D), reference entry order
__global__ void k_synth_p_count(int dim, int Nf, long long row_begin, int nrows, int32_t* __restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int id[3];
    lex_split(row_begin + r, Nf, dim, id);
    int c = 1;
    for (int d = 0; d < dim; ++d) if (id[d] & 1) c *= 2;
    cnt[r] = c;
}

__global__ void k_synth_p_fill(int dim, int Nf, int Nc, long long row_begin, int nrows, const int32_t* __restrict__ rowptr,
                               int32_t* __restrict__ cols, double* __restrict__ vals, ColMap cm)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int id[3];
    lex_split(row_begin + r, Nf, dim, id);
    double w = 1.0;
    for (int d = 0; d < dim; ++d) if (id[d] & 1) w *= 0.5;
    int o = rowptr[r];
    for (int comb = 0; comb < (1 << dim); ++comb) {
        bool ok = true;
        long long c = 0, stride = 1;
        for (int d = 0; d < dim; ++d) {
            const int hi = (comb >> d) & 1, odd = id[d] & 1;
            if (hi && !odd) ok = false;
            c += (long long)(odd ? (id[d] - 1) / 2 + hi : id[d] / 2) * stride;
            stride *= Nc;
        }
        if (!ok) continue;
        cols[o] = cm(c);
        vals[o] = w;
        ++o;
    }
}

// injection list: local fine index of the fine node at the coordinates of every coarse row in [cb, ce)
__global__ void k_synth_inj(int dim, int Nf, int Nc, long long cb, int ncoarse, long long fine_begin, int32_t* __restrict__ inj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoarse) return;
    int id[3];
    lex_split(cb + i, Nc, dim, id);
    long long f = 0, stride = 1;
    for (int d = 0; d < dim; ++d) { f += 2LL * id[d] * stride; stride *= Nf; }
    inj[i] = (int32_t)(f - fine_begin);
}

// counts (int32, n + 1 entries, last = 0) -> exclusive scan in place = row pointers
int scan_counts(mgb_handle* h, int32_t* cnt, int64_t n_plus_1)
{
    void* tmp = nullptr; size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, cnt, (int)n_plus_1, h->stream));
    CU(cudaMalloc(&tmp, bytes));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, cnt, (int)n_plus_1, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    CU(e);
    return MGB_OK;
}

int alloc_csr_from_counts(mgb_handle* h, DevCsr& D, int64_t nrows, int64_t ncols)
{
    // D.rowptr holds the scanned counts
    int32_t total = 0;
    CU(cudaMemcpyAsync(&total, D.rowptr + nrows, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    D.nrows = nrows; D.ncols = ncols; D.nnz = total;
    TRY(dev_alloc(h, &D.cols, (size_t)total + 16));
    TRY(dev_alloc(h, &D.vals, (size_t)total + 16));
    CU(cudaMemsetAsync(D.cols + total, 0, 16 * sizeof(int32_t), h->stream));
    CU(cudaMemsetAsync(D.vals + total, 0, 16 * sizeof(double), h->stream));
    return MGB_OK;
}

// ---- getJacobiMatrices on the device (multigrid.py:48-56) ------------------------------------------------
__global__ void k_rj_count(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
                           int32_t* __restrict__ cnt, double* __restrict__ dinv, int* __restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0; bool have = false; int c = 0;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        if (cols[k] == i) { d += vals[k]; have = true; }
        else if (vals[k] != 0.0) ++c;
    }
    if (!have || d == 0.0) *bad = 1;
    dinv[i] = __ddiv_rn(1.0, d);
    cnt[i] = c;
}

__global__ void k_rj_fill(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
                          const double* __restrict__ dinv, const int32_t* __restrict__ rrp, int32_t* __restrict__ rcols,
                          double* __restrict__ rvals, int reversed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = rrp[i], cnt = rrp[i + 1] - o;
    int j = 0;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        if (cols[k] == i || vals[k] == 0.0) continue;
        const int dst = reversed ? o + (cnt - 1 - j) : o + j;
        rcols[dst] = cols[k];
        rvals[dst] = __dmul_rn(dinv[i], vals[k]);
        ++j;
    }
}

int build_rj_device(mgb_handle* h, Level& L)
{
    const int n = (int)L.n;
    const size_t sn = (size_t)n;
    int* bad = nullptr;
    TRY(dev_alloc(h, &bad, 1));
    CU(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    TRY(dev_alloc(h, &L.RJ.rowptr, sn + 1 + 8));
    CU(cudaMemsetAsync(L.RJ.rowptr, 0, (sn + 1 + 8) * sizeof(int32_t), h->stream));
    TRY(dev_alloc(h, &L.dinv, sn + 16));
    CU(cudaMemsetAsync(L.dinv, 0, (sn + 16) * sizeof(double), h->stream));
    if (n > 0) k_rj_count<<<(n + 255) / 256, 256, 0, h->stream>>>(n, L.A.rowptr, L.A.cols, L.A.vals, L.RJ.rowptr, L.dinv, bad);
    TRY(scan_counts(h, L.RJ.rowptr, (int64_t)n + 1));
    TRY(alloc_csr_from_counts(h, L.RJ, n, L.A.ncols));
    if (n > 0) k_rj_fill<<<(n + 255) / 256, 256, 0, h->stream>>>(n, L.A.rowptr, L.A.cols, L.A.vals, L.dinv, L.RJ.rowptr, L.RJ.cols, L.RJ.vals, h->rj_reversed);
    int hb = 0;
    CU(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(bad);
    if (hb) return fail(h, MGB_ERR_SINGULAR, "level %d: zero or missing diagonal entry", L.level);
    std::vector<int64_t> ip;
    int64_t in[2];
    const bool sharded = h->dist && L.n_ghost > 0;
    TRY(fetch_rowptr(h, L.RJ, ip));
    if (sharded) TRY(interior_rows(h, L.RJ, L.n, in));
    return finish_csr(h, L.RJ, ip, {}, sharded ? in : nullptr);
}

