/* cpu_solver.c -- plain C + OpenMP port of the TIMED part of the oracle (TEST INFRASTRUCTURE / CPU baseline).
 *
 * bench.py's CPU legs time the same algorithm the GPU runs: right-preconditioned GMRES (PETSc semantics,
 * oracle/krylov.py) with the 2-way block preconditioner (oracle/blockpc.py: lib/Preconditioner.py:219-246 of the
 * reference), one SA-AMG V-cycle on the solid block, Chebyshev on the fluid block and a V-cycle on the selfp
 * pressure Schur complement (oracle/amg.py).  The hierarchies are built by the numpy oracle; only the solve loop
 * is restated here so that the CPU baseline uses all host cores without Python in the loop.  Results are checked
 * against the numpy oracle in tests/test_oracle_cport.py (same iteration count, same solution).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct {
    int64_t nrows, ncols;
    const int64_t* indptr;
    const int32_t* indices;
    const double* data;
} csr_t;

typedef struct {
    csr_t A, P, R;          /* P, R unused on the last level */
    const double* dinv;
    double lmax;
} level_t;

typedef struct {
    int nlevels;
    level_t* levels;
    const double* coarse_inv;   /* dense (n x n, row-major) or NULL: smoothing only on the last level */
    int degree;
    double ratio;
    /* work vectors per level: x, b, r, d (allocated by amg_alloc) */
    double** wx; double** wb; double** wr; double** wd;
} amg_t;

static void matvec(const csr_t* A, const double* x, double* y) {
    const int64_t n = A->nrows;
#pragma omp parallel for schedule(static, 128) if (A->indptr[n] > 100000)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int64_t k = A->indptr[i]; k < A->indptr[i + 1]; ++k) s += A->data[k] * x[A->indices[k]];
        y[i] = s;
    }
}

/* y = z - A x */
static void residual(const csr_t* A, const double* x, const double* z, double* y) {
    const int64_t n = A->nrows;
#pragma omp parallel for schedule(static, 128) if (A->indptr[n] > 100000)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int64_t k = A->indptr[i]; k < A->indptr[i + 1]; ++k) s += A->data[k] * x[A->indices[k]];
        y[i] = z[i] - s;
    }
}

/* y += A x */
static void matvec_add(const csr_t* A, const double* x, double* y) {
    const int64_t n = A->nrows;
#pragma omp parallel for schedule(static, 128) if (A->indptr[n] > 100000)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int64_t k = A->indptr[i]; k < A->indptr[i + 1]; ++k) s += A->data[k] * x[A->indices[k]];
        y[i] += s;
    }
}

static double dot(int64_t n, const double* x, const double* y) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static) if (n > 50000)
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* h[i] = <V_i, w>, i < k: one pass over w in L1-sized row chunks */
static void multi_dot(int64_t n, const double* V, int k, const double* w, double* h) {
    for (int i = 0; i < k; ++i) h[i] = 0.0;
#pragma omp parallel if (n > 20000)
    {
        double* loc = (double*)calloc((size_t)k, sizeof(double));
#pragma omp for schedule(static) nowait
        for (int64_t c = 0; c < n; c += 1024) {
            const int64_t e = c + 1024 < n ? c + 1024 : n;
            for (int i = 0; i < k; ++i) {
                const double* v = V + (size_t)i * n;
                double s = 0.0;
                for (int64_t r = c; r < e; ++r) s += v[r] * w[r];
                loc[i] += s;
            }
        }
#pragma omp critical
        for (int i = 0; i < k; ++i) h[i] += loc[i];
        free(loc);
    }
}

/* Chebyshev(degree) on D^-1 A over [lmax/ratio, lmax], exactly oracle/amg.py:_cheby */
static void cheby(const amg_t* H, int l, const double* b, double* x, int zero_guess) {
    const level_t* L = &H->levels[l];
    const int64_t n = L->A.nrows;
    const double lmax = L->lmax, lmin = lmax / H->ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    double* r = H->wr[l];
    double* d = H->wd[l];
    if (zero_guess) {
#pragma omp parallel for schedule(static) if (n > 50000)
        for (int64_t i = 0; i < n; ++i) { r[i] = b[i]; d[i] = L->dinv[i] * b[i] / theta; x[i] = d[i]; }
    } else {
        residual(&L->A, x, b, r);
#pragma omp parallel for schedule(static) if (n > 50000)
        for (int64_t i = 0; i < n; ++i) { d[i] = L->dinv[i] * r[i] / theta; x[i] += d[i]; }
    }
    double* t = (double*)malloc((size_t)n * sizeof(double));
    for (int k = 1; k < H->degree; ++k) {
        matvec(&L->A, d, t);
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        const double c1 = rho_new * rho, c2 = 2.0 * rho_new / delta;
#pragma omp parallel for schedule(static) if (n > 50000)
        for (int64_t i = 0; i < n; ++i) {
            r[i] -= t[i];
            d[i] = c1 * d[i] + c2 * L->dinv[i] * r[i];
            x[i] += d[i];
        }
        rho = rho_new;
    }
    free(t);
}

static void cycle(const amg_t* H, int l, const double* b, double* x) {
    const level_t* L = &H->levels[l];
    const int64_t n = L->A.nrows;
    if (l == H->nlevels - 1) {
        if (H->coarse_inv) {
            for (int64_t i = 0; i < n; ++i) {
                double s = 0.0;
                const double* row = H->coarse_inv + (size_t)i * n;
                for (int64_t j = 0; j < n; ++j) s += row[j] * b[j];
                x[i] = s;
            }
        } else {
            cheby(H, l, b, x, 1);
            if (H->nlevels > 1) cheby(H, l, b, x, 0);
        }
        return;
    }
    cheby(H, l, b, x, 1);
    residual(&L->A, x, b, H->wr[l]);
    matvec(&L->R, H->wr[l], H->wb[l + 1]);
    cycle(H, l + 1, H->wb[l + 1], H->wx[l + 1]);
    matvec_add(&L->P, H->wx[l + 1], x);
    cheby(H, l, b, x, 0);
}

void oracle_amg_alloc(amg_t* H) {
    H->wx = (double**)calloc(H->nlevels, sizeof(double*));
    H->wb = (double**)calloc(H->nlevels, sizeof(double*));
    H->wr = (double**)calloc(H->nlevels, sizeof(double*));
    H->wd = (double**)calloc(H->nlevels, sizeof(double*));
    for (int l = 0; l < H->nlevels; ++l) {
        size_t n = (size_t)H->levels[l].A.nrows;
        H->wx[l] = (double*)calloc(n, sizeof(double));
        H->wb[l] = (double*)calloc(n, sizeof(double));
        H->wr[l] = (double*)calloc(n, sizeof(double));
        H->wd[l] = (double*)calloc(n, sizeof(double));
    }
}

void oracle_amg_free(amg_t* H) {
    for (int l = 0; l < H->nlevels; ++l) { free(H->wx[l]); free(H->wb[l]); free(H->wr[l]); free(H->wd[l]); }
    free(H->wx); free(H->wb); free(H->wr); free(H->wd);
}

void oracle_amg_apply(const amg_t* H, const double* b, double* x) { cycle(H, 0, b, x); }

/* 2-way block preconditioner with the pressure-Schur fieldsplit (split order f, p):
 *   y_s = K_s x_s ; t = x_fp - M_fps y_s ; y_f = K_f t_f ; r = t_p - A_pf y_f ; y_p = K_p r  [+ K_v r  (SchurLowerCC)]   */
typedef struct {
    int64_t ns, nf, np;
    const amg_t* Ks; const amg_t* Kf; const amg_t* Kp;
    csr_t Mfps;     /* (nf+np) x ns, may be empty */
    csr_t Apf;      /* np x nf */
    const amg_t* Kv;     /* additive viscous part of the pressure Schur preconditioner, or NULL */
} blockpc_t;

static void blockpc_apply(const blockpc_t* M, const double* x, double* y, double* work) {
    const int64_t ns = M->ns, nf = M->nf, np = M->np;
    oracle_amg_apply(M->Ks, x, y);
    double* t = work;                           /* nf + np */
    if (M->Mfps.indptr[M->Mfps.nrows] > 0) residual(&M->Mfps, y, x + ns, t);
    else memcpy(t, x + ns, (size_t)(nf + np) * sizeof(double));
    oracle_amg_apply(M->Kf, t, y + ns);
    double* tp = work + nf + np;                /* np */
    residual(&M->Apf, y + ns, t + nf, tp);
    oracle_amg_apply(M->Kp, tp, y + ns + nf);
    if (M->Kv) {
        double* tv = tp + np;                   /* np */
        oracle_amg_apply(M->Kv, tp, tv);
        double* yp = y + ns + nf;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < np; ++i) yp[i] += tv[i];
    }
}

/* right-preconditioned GMRES, zero initial guess, restart = maxit (lib/Solver.py:99-100); returns iterations */
int oracle_gmres_right(const csr_t* A, const blockpc_t* M, const double* b, double* x, double rtol, double atol,
                       int maxit, double* history, int* reason) {
    const int64_t n = A->nrows;
    const int m = maxit;
    double* V = (double*)malloc((size_t)(m + 1) * n * sizeof(double));
    double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
    double* cs = (double*)calloc(m, sizeof(double));
    double* sn = (double*)calloc(m, sizeof(double));
    double* g = (double*)calloc(m + 1, sizeof(double));
    double* z = (double*)malloc((size_t)n * sizeof(double));
    double* work = (double*)malloc((size_t)(M->nf + 3 * M->np + 8) * sizeof(double));
    double* h = (double*)malloc((size_t)(m + 1) * sizeof(double));
    memset(x, 0, (size_t)n * sizeof(double));
    double beta = sqrt(dot(n, b, b));
    history[0] = beta;
    const double ttol = fmax(rtol * beta, atol);
    int its = 0;
    *reason = 0;
    if (beta <= ttol) { *reason = beta < atol ? 3 : 2; goto done; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) V[i] = b[i] / beta;
    g[0] = beta;
    int j = 0;
    while (*reason == 0 && j < m) {
        double* vj = V + (size_t)j * n;
        double* w = V + (size_t)(j + 1) * n;
        blockpc_apply(M, vj, z, work);
        matvec(A, z, w);
        /* classical Gram-Schmidt: all projections from the un-modified w, then one update */
        multi_dot(n, V, j + 1, w, h);
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < n; ++r) {
            double s = w[r];
            for (int i = 0; i <= j; ++i) s -= h[i] * V[(size_t)i * n + r];
            w[r] = s;
        }
        const double hn = sqrt(dot(n, w, w));
        for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = h[i];
        H[(size_t)(j + 1) * m + j] = hn;
        if (hn > 0.0) {
#pragma omp parallel for schedule(static)
            for (int64_t r = 0; r < n; ++r) w[r] /= hn;
        }
        for (int i = 0; i < j; ++i) {
            const double a = H[(size_t)i * m + j], c = H[(size_t)(i + 1) * m + j];
            H[(size_t)i * m + j] = cs[i] * a + sn[i] * c;
            H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * c;
        }
        const double den = hypot(H[(size_t)j * m + j], H[(size_t)(j + 1) * m + j]);
        if (den == 0.0) { *reason = -5; break; }
        cs[j] = H[(size_t)j * m + j] / den;
        sn[j] = H[(size_t)(j + 1) * m + j] / den;
        H[(size_t)j * m + j] = den;
        H[(size_t)(j + 1) * m + j] = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
        ++its;
        ++j;
        const double res = fabs(g[j]);
        history[its] = res;
        if (res <= ttol) *reason = res < atol ? 3 : 2;
    }
    if (j > 0) {
        double* yv = (double*)calloc(j, sizeof(double));
        for (int i = j - 1; i >= 0; --i) {
            double s = g[i];
            for (int q = i + 1; q < j; ++q) s -= H[(size_t)i * m + q] * yv[q];
            yv[i] = s / H[(size_t)i * m + i];
        }
        double* dx = (double*)calloc((size_t)n, sizeof(double));
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < n; ++r) {
            double s = 0.0;
            for (int i = 0; i < j; ++i) s += yv[i] * V[(size_t)i * n + r];
            dx[r] = s;
        }
        blockpc_apply(M, dx, z, work);
        memcpy(x, z, (size_t)n * sizeof(double));
        free(dx);
        free(yv);
    }
    if (*reason == 0) *reason = -3;
done:
    free(V); free(H); free(cs); free(sn); free(g); free(z); free(work); free(h);
    return its;
}

int oracle_cport_threads(void) { return omp_get_max_threads(); }
void oracle_cport_set_threads(int t) { omp_set_num_threads(t > 0 ? t : 1); }
