/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference V-cycle (the CPU oracle).
 *
 * Never linked into, loaded by or called from the product library (libmgb200.so).  Used by tests/
 * as the checker, by __graft_entry__.smoke() as the checker and by bench.py as the *reported*
 * CPU baseline ("port", OpenMP over rows).
 *
 * Follows /root/reference/multigrid.py:
 *   jacobi sweep        multigrid.py:226   sol = ((1-w)*v + w*(Dinv f)) - w*(R_omega v)
 *   residual            multigrid.py:244   r = f - A v   (raw A, stored zeros included)
 *   injection           multigrid.py:128-131
 *   interpolation+add   multigrid.py:59-120, :260   (matrix form, row entries in reference order)
 *   coarse zero guess   multigrid.py:253
 *   coarsest solve      multigrid.py:238-241  (reference: SuperLU via scipy spsolve; here: dense LU
 *                       with partial pivoting -- both are backward-stable direct solves)
 * CSR row sums follow scipy's csr_matvec (third-party, unpinned; scipy 1.18.1 in the dev container):
 * one accumulator per row, entries in stored order, separate multiply and add.
 * Compile with -ffp-contract=off so that no FMA is formed (scipy's wheels are built without FMA).
 *
 * No reference text exists for Gauss-Seidel, level sets and colouring: those follow the
 * definitions in DESIGN.md.
 *
 * Pinning: tests/test_oracle.py + tests/golden/ (vectors produced by the reference
 * itself in the development container).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef int32_t i32;

/* ---------------------------------------------------------------- per-op kernels */

void orc_csr_matvec(i64 n, const i64 *ip, const i32 *ix, const double *ax, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        double s = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) s += ax[k] * x[ix[k]];
        y[i] = s;
    }
}

void orc_residual(i64 n, const i64 *ip, const i32 *ix, const double *ax, const double *f, const double *v, double *r)
{
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        double s = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) s += ax[k] * v[ix[k]];
        r[i] = f[i] - s;
    }
}

/* one sweep, out-of-place.  multigrid.py:226 evaluated left to right. */
void orc_jacobi_sweep(i64 n, const i64 *ip, const i32 *ix, const double *ax, const double *dinv, const double *f,
                      const double *vin, double *vout, double omega)
{
    const double om1 = 1 - omega;
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        double s = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) s += ax[k] * vin[ix[k]];
        double t1 = om1 * vin[i];
        double t2 = omega * (dinv[i] * f[i]);
        double t3 = omega * s;
        vout[i] = (t1 + t2) - t3;
    }
}

/* single-matrix variant  v + w*(dinv*(f - A v))  (not the reference formula) */
void orc_jacobi_sweep_aform(i64 n, const i64 *ip, const i32 *ix, const double *ax, const double *dinv, const double *f,
                            const double *vin, double *vout, double omega)
{
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        double s = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) s += ax[k] * vin[ix[k]];
        vout[i] = vin[i] + omega * (dinv[i] * (f[i] - s));
    }
}

void orc_gather(i64 nc, const i32 *inj, const double *r, double *out)
{
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < nc; ++i) out[i] = r[inj[i]];
}

/* v <- v + P e ; if err != NULL also err <- P e  (multigrid.py:258-260) */
void orc_prolong_add(i64 nf, const i64 *ip, const i32 *ix, const double *ax, const double *e, double *v, double *err)
{
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < nf; ++i) {
        double s = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) s += ax[k] * e[ix[k]];
        if (err) err[i] = s;
        v[i] = v[i] + s;
    }
}

/* forward Gauss-Seidel in place; order == NULL -> natural order. Sequential by definition. */
void orc_gs_forward(i64 n, const i64 *ip, const i32 *ix, const double *ax, const double *f, double *v, const i32 *order)
{
    for (i64 t = 0; t < n; ++t) {
        i64 i = order ? order[t] : t;
        double s = 0.0, d = 0.0;
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) {
            if (ix[k] == i) d = ax[k];
            else s += ax[k] * v[ix[k]];
        }
        v[i] = (f[i] - s) / d;
    }
}

double orc_norm2(i64 n, const double *x)
{
    long double s = 0.0L;      /* extended accumulator: the norm itself must not be the noisy part of a 1e-12 comparison */
    for (i64 i = 0; i < n; ++i) s += (long double)x[i] * (long double)x[i];
    return (double)sqrtl(s);
}

/* ---------------------------------------------------------------- artefacts */

/* symmetrised nonzero graph: neighbours j<i of i with a_ij != 0 or a_ji != 0 (lower part only). */
static int build_lower_sym(i64 n, const i64 *ip, const i32 *ix, const double *ax, i64 **lp_out, i32 **lx_out)
{
    i64 *cnt = (i64 *)calloc((size_t)n + 1, sizeof(i64));
    if (!cnt) return -1;
    for (i64 i = 0; i < n; ++i)
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) {
            i64 j = ix[k];
            if (j == i || ax[k] == 0.0) continue;
            i64 hi = i > j ? i : j;
            cnt[hi + 1]++;
        }
    for (i64 i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    i32 *lx = (i32 *)malloc(sizeof(i32) * (size_t)(cnt[n] > 0 ? cnt[n] : 1));
    i64 *pos = (i64 *)malloc(sizeof(i64) * (size_t)(n + 1));
    if (!lx || !pos) return -1;
    memcpy(pos, cnt, sizeof(i64) * (size_t)(n + 1));
    for (i64 i = 0; i < n; ++i)
        for (i64 k = ip[i]; k < ip[i + 1]; ++k) {
            i64 j = ix[k];
            if (j == i || ax[k] == 0.0) continue;
            i64 hi = i > j ? i : j, lo = i > j ? j : i;
            lx[pos[hi]++] = (i32)lo;       /* duplicates (a_ij and a_ji) are harmless */
        }
    free(pos);
    *lp_out = cnt; *lx_out = lx;
    return 0;
}

/* lev[i] = 0 if no lower neighbour else 1 + max lev[j]; returns number of levels */
i64 orc_level_sets(i64 n, const i64 *ip, const i32 *ix, const double *ax, i32 *lev)
{
    i64 *lp; i32 *lx;
    if (build_lower_sym(n, ip, ix, ax, &lp, &lx)) return -1;
    i32 mx = -1;
    for (i64 i = 0; i < n; ++i) {
        i32 m = -1;
        for (i64 k = lp[i]; k < lp[i + 1]; ++k) if (lev[lx[k]] > m) m = lev[lx[k]];
        lev[i] = m + 1;
        if (lev[i] > mx) mx = lev[i];
    }
    free(lp); free(lx);
    return (i64)mx + 1;
}

/* first-fit greedy colouring in natural order; returns number of colours */
i64 orc_greedy_colouring(i64 n, const i64 *ip, const i32 *ix, const double *ax, i32 *col)
{
    i64 *lp; i32 *lx;
    if (build_lower_sym(n, ip, ix, ax, &lp, &lx)) return -1;
    i64 cap = 64;
    i64 *mark = (i64 *)malloc(sizeof(i64) * (size_t)cap);
    for (i64 c = 0; c < cap; ++c) mark[c] = -1;
    i32 mx = -1;
    for (i64 i = 0; i < n; ++i) {
        for (i64 k = lp[i]; k < lp[i + 1]; ++k) {
            i32 c = col[lx[k]];
            if (c >= cap) {
                i64 ncap = cap * 2; while (c >= ncap) ncap *= 2;
                mark = (i64 *)realloc(mark, sizeof(i64) * (size_t)ncap);
                for (i64 q = cap; q < ncap; ++q) mark[q] = -1;
                cap = ncap;
            }
            mark[c] = i;
        }
        i32 c = 0;
        while (c < cap && mark[c] == i) ++c;
        col[i] = c;
        if (c > mx) mx = c;
    }
    free(mark); free(lp); free(lx);
    return (i64)mx + 1;
}

/* ---------------------------------------------------------------- dense LU (coarsest level) */

/* in-place LU with partial pivoting of a row-major n x n matrix; returns 0 or -(k+1) if singular */
int orc_dense_lu(i64 n, double *a, i32 *piv)
{
    for (i64 k = 0; k < n; ++k) {
        i64 p = k; double mx = fabs(a[k * n + k]);
        for (i64 i = k + 1; i < n; ++i) { double t = fabs(a[i * n + k]); if (t > mx) { mx = t; p = i; } }
        piv[k] = (i32)p;
        if (mx == 0.0) return -(int)(k + 1);
        if (p != k) for (i64 j = 0; j < n; ++j) { double t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t; }
        double inv = 1.0 / a[k * n + k];
#pragma omp parallel for schedule(static) if (n - k > 256)
        for (i64 i = k + 1; i < n; ++i) {
            double l = a[i * n + k] * inv;
            a[i * n + k] = l;
            if (l != 0.0) for (i64 j = k + 1; j < n; ++j) a[i * n + j] -= l * a[k * n + j];
        }
    }
    return 0;
}

void orc_dense_lu_solve(i64 n, const double *lu, const i32 *piv, double *b)
{
    for (i64 k = 0; k < n; ++k) { i64 p = piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (i64 i = 1; i < n; ++i) { double s = b[i]; for (i64 j = 0; j < i; ++j) s -= lu[i * n + j] * b[j]; b[i] = s; }
    for (i64 i = n - 1; i >= 0; --i) { double s = b[i]; for (i64 j = i + 1; j < n; ++j) s -= lu[i * n + j] * b[j]; b[i] = s / lu[i * n + i]; }
}

/* ---------------------------------------------------------------- whole V-cycle driver (for timing and large cases) */

typedef struct {
    i64 n;
    const i64 *a_ip; const i32 *a_ix; const double *a_ax;      /* raw A                    */
    const i64 *r_ip; const i32 *r_ix; const double *r_ax;      /* R_omega                  */
    const double *dinv;
    const i64 *p_ip; const i32 *p_ix; const double *p_ax;      /* P: level-1 -> this level */
    const i32 *inj;  i64 nc;                                    /* injection list to level-1 */
    const i64 *rs_ip; const i32 *rs_ix; const double *rs_ax;   /* explicit R (or NULL)     */
    const i32 *gs_order;                                        /* NULL = natural           */
    double *v, *f, *r, *t;                                      /* work vectors             */
} orc_level;

typedef struct {
    int nlev;
    int owns;              /* level arrays were allocated by orc_mg_build_poisson and are freed with the handle */
    orc_level *L;          /* L[0] = coarsest */
    double *lu; i32 *piv;  /* dense LU of coarsest A */
    double omega; int mu1, mu2;
    int smoother;          /* 0 jacobi (reference), 1 jacobi A-form, 2 gauss-seidel (order from gs_order) */
} orc_mg;

orc_mg *orc_mg_create(int nlev)
{
    orc_mg *m = (orc_mg *)calloc(1, sizeof(orc_mg));
    m->nlev = nlev;
    m->L = (orc_level *)calloc((size_t)nlev, sizeof(orc_level));
    m->omega = 2.0 / 3.0; m->mu1 = m->mu2 = 2;
    return m;
}

void orc_mg_set_params(orc_mg *m, double omega, int mu1, int mu2, int smoother)
{ m->omega = omega; m->mu1 = mu1; m->mu2 = mu2; m->smoother = smoother; }

void orc_mg_set_level(orc_mg *m, int k, i64 n,
                      const i64 *a_ip, const i32 *a_ix, const double *a_ax,
                      const i64 *r_ip, const i32 *r_ix, const double *r_ax, const double *dinv,
                      const i64 *p_ip, const i32 *p_ix, const double *p_ax,
                      const i32 *inj, i64 nc,
                      const i64 *rs_ip, const i32 *rs_ix, const double *rs_ax,
                      const i32 *gs_order)
{
    orc_level *L = &m->L[k];
    L->n = n; L->a_ip = a_ip; L->a_ix = a_ix; L->a_ax = a_ax;
    L->r_ip = r_ip; L->r_ix = r_ix; L->r_ax = r_ax; L->dinv = dinv;
    L->p_ip = p_ip; L->p_ix = p_ix; L->p_ax = p_ax; L->inj = inj; L->nc = nc;
    L->rs_ip = rs_ip; L->rs_ix = rs_ix; L->rs_ax = rs_ax; L->gs_order = gs_order;
    L->v = (double *)malloc(sizeof(double) * (size_t)n);
    L->f = (double *)malloc(sizeof(double) * (size_t)n);
    L->r = (double *)malloc(sizeof(double) * (size_t)n);
    L->t = (double *)malloc(sizeof(double) * (size_t)n);
}

int orc_mg_finalize(orc_mg *m)
{
    orc_level *C = &m->L[0];
    i64 n = C->n;
    m->lu = (double *)calloc((size_t)(n * n), sizeof(double));
    m->piv = (i32 *)malloc(sizeof(i32) * (size_t)n);
    for (i64 i = 0; i < n; ++i)
        for (i64 k = C->a_ip[i]; k < C->a_ip[i + 1]; ++k) m->lu[i * n + C->a_ix[k]] += C->a_ax[k];
    return orc_dense_lu(n, m->lu, m->piv);
}

void orc_mg_destroy(orc_mg *m)
{
    for (int k = 0; k < m->nlev; ++k) {
        orc_level *L = &m->L[k];
        free(L->v); free(L->f); free(L->r); free(L->t);
        if (m->owns) {
            free((void *)L->a_ip); free((void *)L->a_ix); free((void *)L->a_ax);
            free((void *)L->r_ip); free((void *)L->r_ix); free((void *)L->r_ax); free((void *)L->dinv);
            free((void *)L->p_ip); free((void *)L->p_ix); free((void *)L->p_ax); free((void *)L->inj);
        }
    }
    free(m->L); free(m->lu); free(m->piv); free(m);
}

static void smooth(orc_mg *m, orc_level *L, int nw)
{
    for (int s = 0; s < nw; ++s) {
        if (m->smoother == 2) {
            orc_gs_forward(L->n, L->a_ip, L->a_ix, L->a_ax, L->f, L->v, L->gs_order);
        } else {
            if (m->smoother == 0) orc_jacobi_sweep(L->n, L->r_ip, L->r_ix, L->r_ax, L->dinv, L->f, L->v, L->t, m->omega);
            else orc_jacobi_sweep_aform(L->n, L->a_ip, L->a_ix, L->a_ax, L->dinv, L->f, L->v, L->t, m->omega);
            double *tmp = L->v; L->v = L->t; L->t = tmp;
        }
    }
}

/* one V-cycle with level index `top` as the finest level; v (in/out) and f live in L[top].v / .f */
static void vcycle_rec(orc_mg *m, int k)
{
    orc_level *L = &m->L[k];
    if (k == 0) {
        memcpy(L->v, L->f, sizeof(double) * (size_t)L->n);
        orc_dense_lu_solve(L->n, m->lu, m->piv, L->v);
        return;
    }
    orc_level *C = &m->L[k - 1];
    smooth(m, L, m->mu1);
    orc_residual(L->n, L->a_ip, L->a_ix, L->a_ax, L->f, L->v, L->r);
    if (L->rs_ip) orc_csr_matvec(C->n, L->rs_ip, L->rs_ix, L->rs_ax, L->r, C->f);
    else orc_gather(C->n, L->inj, L->r, C->f);
    memset(C->v, 0, sizeof(double) * (size_t)C->n);
    vcycle_rec(m, k - 1);
    orc_prolong_add(L->n, L->p_ip, L->p_ix, L->p_ax, C->v, L->v, NULL);
    smooth(m, L, m->mu2);
}

/* v, f: caller arrays of length L[top].n; runs ncycles cycles; hist (nullable) gets ||f - A v||_2 per cycle */
void orc_mg_vcycle(orc_mg *m, int top, double *v, const double *f, int ncycles, double *hist)
{
    orc_level *L = &m->L[top];
    memcpy(L->v, v, sizeof(double) * (size_t)L->n);
    memcpy(L->f, f, sizeof(double) * (size_t)L->n);
    for (int c = 0; c < ncycles; ++c) {
        vcycle_rec(m, top);
        if (hist) {
            orc_residual(L->n, L->a_ip, L->a_ix, L->a_ax, L->f, L->v, L->r);
            hist[c] = orc_norm2(L->n, L->r);
        }
    }
    memcpy(v, L->v, sizeof(double) * (size_t)L->n);
}


/* ---------------------------------------------------------------- threads (bench.py states how many were used) */
int orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

/* ---------------------------------------------------------------- structured P1 Poisson hierarchy, built in place
 * The synthetic stand-in for Multigrid_prototype.py:62-118 at sizes no numpy/scipy build can hold (513^3: 2.0e9 stored
 * entries): the SAME arrays as multigrid_dolfinx_b200/problems.py (stencil_p1, prolongation, injection; lexicographic
 * DOFs) and as the reference's getJacobiMatrices (multigrid.py:48-56: off-diagonal NON-ZEROS of A scaled by fl(1/a_ii),
 * rows left in descending column order by scipy's DIA x CSR product), generated row by row with OpenMP.
 * tests/test_oracle.py compares every array bit for bit with those Python builders at small sizes. */
typedef struct { int noff; i64 lin[27]; int o[27][3]; double w[27]; } orc_stencil;

static void make_stencil(int dim, i64 N, double unit, orc_stencil *S)
{
    /* neighbour offsets of the cell-connectivity pattern ("/" mesh in 2-D, Kuhn mesh in 3-D): all components >= 0 or all <= 0;
     * weight -unit for axis neighbours, stored 0.0 for the others; sorted by linear displacement */
    S->noff = 0;
    int lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
    if (dim == 2) { lo[2] = hi[2] = 0; }
    for (int c = lo[2]; c <= hi[2]; ++c)
        for (int b = -1; b <= 1; ++b)
            for (int a = -1; a <= 1; ++a) {
                int allp = a >= 0 && b >= 0 && c >= 0, alln = a <= 0 && b <= 0 && c <= 0;
                if (!allp && !alln) continue;
                int nz = (a != 0) + (b != 0) + (c != 0);
                int k = S->noff++;
                S->o[k][0] = a; S->o[k][1] = b; S->o[k][2] = c;
                S->lin[k] = a + N * b + N * N * c;
                S->w[k] = nz == 1 ? -unit : 0.0;           /* (the centre is handled separately) */
            }
    /* the triple loop above already enumerates in ascending linear displacement (x fastest) */
}

static i64 *alloc_ptr(i64 n) { return (i64 *)malloc(sizeof(i64) * (size_t)(n + 1)); }

static void prefix_counts(i64 n, i64 *ip)      /* ip[i+1] holds the count of row i on entry */
{
    ip[0] = 0;
    for (i64 i = 0; i < n; ++i) ip[i + 1] += ip[i];
}

orc_mg *orc_mg_create(int nlev);
void orc_mg_set_params(orc_mg *m, double omega, int mu1, int mu2, int smoother);
int orc_mg_finalize(orc_mg *m);

orc_mg *orc_mg_build_poisson(int dim, int c, int coarsest_level, int finest_level, double omega, int mu1, int mu2)
{
    if ((dim != 2 && dim != 3) || finest_level < coarsest_level) return NULL;
    const int nlev = finest_level - coarsest_level + 1;
    orc_mg *m = orc_mg_create(nlev);
    m->owns = 1;
    orc_mg_set_params(m, omega, mu1, mu2, 0);
    for (int k = 0; k < nlev; ++k) {
        const i64 cells = (i64)c << (coarsest_level + k), N = cells + 1;
        i64 n = 1; for (int d = 0; d < dim; ++d) n *= N;
        const double h = 1.0 / (double)cells, unit = dim == 2 ? 1.0 : h, diag_int = (dim == 2 ? 4.0 : 6.0) * unit;
        orc_stencil S; make_stencil(dim, N, unit, &S);
        orc_level *L = &m->L[k];
        L->n = n;
        /* ---- A (stored zeros kept, identity rows on the boundary) and R_omega, two passes: count, fill */
        i64 *a_ip = alloc_ptr(n), *r_ip = alloc_ptr(n);
        double *dinv = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 t = i; int mi[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) { mi[d] = (int)(t % N); t /= N; }
            int onb = 0; for (int d = 0; d < dim; ++d) onb |= (mi[d] == 0) | (mi[d] == N - 1);
            i64 ca = 0, cr = 0;
            for (int q = 0; q < S.noff; ++q) {
                int ok = 1, nbb = 0;
                for (int d = 0; d < dim; ++d) { int cc = mi[d] + S.o[q][d]; ok &= cc >= 0 && cc <= N - 1; nbb |= (cc <= 0) | (cc >= N - 1); }
                if (!ok) continue;
                ++ca;
                if (S.lin[q] != 0 && !(onb || nbb) && S.w[q] != 0.0) ++cr;
            }
            a_ip[i + 1] = ca; r_ip[i + 1] = cr;
        }
        prefix_counts(n, a_ip); prefix_counts(n, r_ip);
        i32 *a_ix = (i32 *)malloc(sizeof(i32) * (size_t)(a_ip[n] + 1)); double *a_ax = (double *)malloc(sizeof(double) * (size_t)(a_ip[n] + 1));
        i32 *r_ix = (i32 *)malloc(sizeof(i32) * (size_t)(r_ip[n] + 1)); double *r_ax = (double *)malloc(sizeof(double) * (size_t)(r_ip[n] + 1));
#pragma omp parallel for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 t = i; int mi[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) { mi[d] = (int)(t % N); t /= N; }
            int onb = 0; for (int d = 0; d < dim; ++d) onb |= (mi[d] == 0) | (mi[d] == N - 1);
            const double dii = onb ? 1.0 : diag_int, di = 1 / dii;
            dinv[i] = di;
            i64 ka = a_ip[i], kr = r_ip[i + 1];          /* R_omega rows are filled backwards: descending columns */
            for (int q = 0; q < S.noff; ++q) {
                int ok = 1, nbb = 0;
                for (int d = 0; d < dim; ++d) { int cc = mi[d] + S.o[q][d]; ok &= cc >= 0 && cc <= N - 1; nbb |= (cc <= 0) | (cc >= N - 1); }
                if (!ok) continue;
                const double val = S.lin[q] == 0 ? dii : ((onb || nbb) ? 0.0 : S.w[q] + 0.0);
                a_ix[ka] = (i32)(i + S.lin[q]); a_ax[ka] = val; ++ka;
                if (S.lin[q] != 0 && val != 0.0) { --kr; r_ix[kr] = (i32)(i + S.lin[q]); r_ax[kr] = di * val; }
            }
        }
        L->a_ip = a_ip; L->a_ix = a_ix; L->a_ax = a_ax; L->r_ip = r_ip; L->r_ix = r_ix; L->r_ax = r_ax; L->dinv = dinv;
        L->v = (double *)malloc(sizeof(double) * (size_t)n); L->f = (double *)malloc(sizeof(double) * (size_t)n);
        L->r = (double *)malloc(sizeof(double) * (size_t)n); L->t = (double *)malloc(sizeof(double) * (size_t)n);
        if (k == 0) continue;
        /* ---- P from level k-1 (multigrid.py:59-120, tensor-product extension in 3-D; entries in the reference's order:
         * "-" before "+" with x fastest) and the injection list (multigrid.py:123-132) */
        const i64 Nc = (N + 1) / 2;
        i64 nc = 1; for (int d = 0; d < dim; ++d) nc *= Nc;
        i64 *p_ip = alloc_ptr(n);
#pragma omp parallel for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 t = i, cnt = 1;
            for (int d = 0; d < dim; ++d) { if ((t % N) & 1) cnt *= 2; t /= N; }
            p_ip[i + 1] = cnt;
        }
        prefix_counts(n, p_ip);
        i32 *p_ix = (i32 *)malloc(sizeof(i32) * (size_t)(p_ip[n] + 1)); double *p_ax = (double *)malloc(sizeof(double) * (size_t)(p_ip[n] + 1));
#pragma omp parallel for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 t = i; int mi[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) { mi[d] = (int)(t % N); t /= N; }
            i64 kp = p_ip[i];
            for (int comb = 0; comb < (1 << dim); ++comb) {
                int valid = 1; double w = 1.0; i64 col = 0, stride = 1;
                for (int d = 0; d < dim; ++d) {
                    const int hi = (comb >> d) & 1, odd = mi[d] & 1;
                    if (hi && !odd) valid = 0;
                    w *= odd ? 0.5 : 1.0;
                    col += (i64)(odd ? (mi[d] - 1) / 2 + hi : mi[d] / 2) * stride;
                    stride *= Nc;
                }
                if (valid) { p_ix[kp] = (i32)col; p_ax[kp] = w; ++kp; }
            }
        }
        i32 *inj = (i32 *)malloc(sizeof(i32) * (size_t)nc);
#pragma omp parallel for schedule(static)
        for (i64 i = 0; i < nc; ++i) {
            i64 t = i, f = 0, stride = 1;
            for (int d = 0; d < dim; ++d) { f += 2 * (t % Nc) * stride; t /= Nc; stride *= N; }
            inj[i] = (i32)f;
        }
        L->p_ip = p_ip; L->p_ix = p_ix; L->p_ax = p_ax; L->inj = inj; L->nc = nc;
    }
    if (orc_mg_finalize(m) != 0) { orc_mg_destroy(m); return NULL; }
    return m;
}

/* array of a level, for the bit-for-bit comparison with the Python builders: what = 0 A.indptr, 1 A.indices, 2 A.data,
 * 3 RO.indptr, 4 RO.indices, 5 RO.data, 6 dinv, 7 P.indptr, 8 P.indices, 9 P.data, 10 inj; returns the element count */
i64 orc_mg_level_array(orc_mg *m, int k, int what, const void **out)
{
    if (!m || k < 0 || k >= m->nlev) return -1;
    orc_level *L = &m->L[k];
    switch (what) {
        case 0: *out = L->a_ip; return L->n + 1;
        case 1: *out = L->a_ix; return L->a_ip[L->n];
        case 2: *out = L->a_ax; return L->a_ip[L->n];
        case 3: *out = L->r_ip; return L->n + 1;
        case 4: *out = L->r_ix; return L->r_ip[L->n];
        case 5: *out = L->r_ax; return L->r_ip[L->n];
        case 6: *out = L->dinv; return L->n;
        case 7: *out = L->p_ip; return L->p_ip ? L->n + 1 : 0;
        case 8: *out = L->p_ix; return L->p_ip ? L->p_ip[L->n] : 0;
        case 9: *out = L->p_ax; return L->p_ip ? L->p_ip[L->n] : 0;
        case 10: *out = L->inj; return L->inj ? L->nc : 0;
        default: return -1;
    }
}
