/* omp_kernels.c -- OpenMP CSR matrix-vector product for the CPU baseline (TEST INFRASTRUCTURE).
 * The oracle's algorithms are numpy/scipy; scipy's CSR matvec is single-threaded, so the timed CPU
 * legs of bench.py swap it for this kernel to use all host cores (bench.py reports the thread count). */
#include <stdint.h>
#include <omp.h>

int oracle_omp_threads(void) { return omp_get_max_threads(); }

void oracle_csr_matvec(int64_t nrows, const int64_t* indptr, const int32_t* indices, const double* data,
                       const double* x, double* y) {
    /* small (coarse-level) matrices are not worth a fork/join */
#pragma omp parallel for schedule(static, 256) if (indptr[nrows] > 200000)
    for (int64_t i = 0; i < nrows; ++i) {
        double s = 0.0;
        for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) s += data[k] * x[indices[k]];
        y[i] = s;
    }
}

/* y = a*x + b*y, threaded (numpy's elementwise ops are single-threaded) */
void oracle_axpby(int64_t n, double a, const double* x, double b, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] = a * x[i] + b * y[i];
}
