// vec.cu -- fp64 BLAS-1, fused multi-dot / multi-axpy for the Arnoldi process, reductions.
//
// Replaces PETSc's Vec kernels on the reference path: Vec.axpy/aypx/scale/copy/norm
// (lib/Preconditioner.py:172-212, lib/AAR.py:54-126) and KSPGMRES's VecMDot/VecMAXPY
// (behind KSP.solve, lib/Solver.py:151).  All kernels are HBM-streaming: grid-stride,
// coalesced 8-byte accesses, warp-shuffle + shared-memory block reductions, and a second
// tiny kernel that sums the per-block partials in a fixed order (deterministic results).
#include "common.cuh"
#include "dist.cuh"
#include <algorithm>

namespace poro {

static constexpr int kBlock = 256;

// ------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kBlock) k_elementwise(int64_t n, F f) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <class F>
static void launch_elementwise(Ctx& c, int64_t n, F f) {
    if (n <= 0) return;
    k_elementwise<<<stream_grid(c, n, kBlock, 4), kBlock, 0, c.stream>>>(n, f);
    PORO_LAUNCH_CHECK(c);
}

void vec_copy(Ctx& c, double* y, const double* x, int64_t n) {
    if (n > 0 && y != x) PORO_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
}
void vec_set(Ctx& c, double* y, double a, int64_t n) {
    if (a == 0.0) { if (n > 0) PORO_CUDA(cudaMemsetAsync(y, 0, n * sizeof(double), c.stream)); return; }
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = a; });
}
void vec_scale(Ctx& c, double* y, double a, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] *= a; });
}
void vec_abs_scale(Ctx& c, double* y, double a, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = a * fabs(y[i]); });
}
void vec_axpy(Ctx& c, double* y, double a, const double* x, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = fma(a, x[i], y[i]); });
}
void vec_aypx(Ctx& c, double* y, double a, const double* x, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = fma(a, y[i], x[i]); });
}
void vec_axpby(Ctx& c, double* y, double a, const double* x, double b, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = a * x[i] + b * y[i]; });
}
void vec_waxpby(Ctx& c, double* w, double a, const double* x, double b, const double* y, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { w[i] = a * x[i] + b * y[i]; });
}
void vec_pmult(Ctx& c, double* w, const double* d, const double* x, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { w[i] = d[i] * x[i]; });
}
void vec_gather(Ctx& c, double* y, const double* x, const int* idx, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[i] = x[idx[i]]; });
}
void vec_scatter(Ctx& c, double* y, const double* x, const int* idx, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[idx[i]] = x[i]; });
}
void vec_set_idx(Ctx& c, double* y, const int* idx, double a, int64_t n) {
    launch_elementwise(c, n, [=] __device__(int64_t i) { y[idx[i]] = a; });
}

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// sums `nparts` partials per output (layout partial[part * k + j]) in a fixed order
__global__ void __launch_bounds__(kBlock) k_sum_partials(const double* __restrict__ partial, int nparts, int k,
                                                         double* __restrict__ out) {
    __shared__ double sm[kBlock / 32];
    for (int j = blockIdx.x; j < k; j += gridDim.x) {
        double s = 0.0;
        for (int p = threadIdx.x; p < nparts; p += kBlock) s += partial[(size_t)p * k + j];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) t += sm[w];
            out[j] = t;
        }
        __syncthreads();
    }
}

struct DotPairs {
    const double* x[8];
    const double* y[8];
};

template <int K>
__global__ void __launch_bounds__(kBlock) k_dots(DotPairs P, int64_t n, double* __restrict__ partial) {
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = fma(P.x[j][i], P.y[j][i], acc[j]);
    }
    __shared__ double sm[K][kBlock / 32];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = warp_sum(acc[j]);
        if ((threadIdx.x & 31) == 0) sm[j][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) t += sm[threadIdx.x][w];
        partial[(size_t)blockIdx.x * K + threadIdx.x] = t;
    }
}

void vec_dots(Ctx& c, int k, const double* const* xs, const double* const* ys, int64_t n, double* d_out) {
    PORO_REQUIRE(k >= 1 && k <= 8, "vec_dots supports 1..8 pairs");
    DotPairs P;
    for (int j = 0; j < 8; ++j) { P.x[j] = xs[j < k ? j : 0]; P.y[j] = ys[j < k ? j : 0]; }
    int grid = stream_grid(c, n, kBlock, 8, 4);
    double* partial = c.d_scal;
    switch (k) {
        case 1: k_dots<1><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 2: k_dots<2><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 3: k_dots<3><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 4: k_dots<4><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 5: k_dots<5><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 6: k_dots<6><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        case 7: k_dots<7><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
        default: k_dots<8><<<grid, kBlock, 0, c.stream>>>(P, n, partial); break;
    }
    PORO_LAUNCH_CHECK(c);
    k_sum_partials<<<k, kBlock, 0, c.stream>>>(partial, grid, k, d_out);
    PORO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------
// multi-dot: h[j] = V_j . w, one pass over the basis (classical Gram-Schmidt projection)
// ------------------------------------------------------------------------------------------
static constexpr int kRpt = 4;     // rows per thread per tile
static constexpr int kCchunk = 4;  // columns in flight

// grid.y = chunks of kMc basis columns.  A thread keeps kMc running sums in registers over ALL its rows (kRpt rows x kMc
// columns = 32 independent loads in flight) and the warp / block reduction happens ONCE at the end of the kernel -- the
// first version reduced every 1024-row tile through shuffles and was shuffle-bound (3.1 TB/s at 18 columns).  w is re-read
// once per column chunk (from L2: the chunks of one row range run concurrently).
static constexpr int kMc = 8;

__global__ void __launch_bounds__(kBlock) k_mdot(const double* __restrict__ V, int64_t ld, int ncol,
                                                 const double* __restrict__ w, int64_t n, bool with_ww,
                                                 double* __restrict__ partial) {
    const int nout = ncol + (with_ww ? 1 : 0);
    const int c0 = blockIdx.y * kMc;
    const bool do_ww = with_ww && blockIdx.y == 0;
    double s[kMc];
#pragma unroll
    for (int cc = 0; cc < kMc; ++cc) s[cc] = 0.0;
    double sww = 0.0;
    const int64_t tile = (int64_t)kBlock * kRpt;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        double wv[kRpt];
        int64_t row[kRpt];
#pragma unroll
        for (int r = 0; r < kRpt; ++r) {
            row[r] = base + (int64_t)r * kBlock + threadIdx.x;
            wv[r] = row[r] < n ? w[row[r]] : 0.0;
        }
        double v[kMc][kRpt];
#pragma unroll
        for (int cc = 0; cc < kMc; ++cc) {
            const double* col = V + (int64_t)(c0 + cc) * ld;
#pragma unroll
            for (int r = 0; r < kRpt; ++r) v[cc][r] = (c0 + cc < ncol && row[r] < n) ? col[row[r]] : 0.0;
        }
#pragma unroll
        for (int cc = 0; cc < kMc; ++cc) {
#pragma unroll
            for (int r = 0; r < kRpt; ++r) s[cc] = fma(v[cc][r], wv[r], s[cc]);
        }
        if (do_ww) {
#pragma unroll
            for (int r = 0; r < kRpt; ++r) sww = fma(wv[r], wv[r], sww);
        }
    }
    __shared__ double sm[kBlock / 32][kMc + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int cc = 0; cc < kMc; ++cc) {
        const double t = warp_sum(s[cc]);
        if (lane == 0) sm[warp][cc] = t;
    }
    {
        const double t = warp_sum(sww);
        if (lane == 0) sm[warp][kMc] = t;
    }
    __syncthreads();
    if (threadIdx.x <= kMc) {
        double t = 0.0;
#pragma unroll
        for (int wq = 0; wq < kBlock / 32; ++wq) t += sm[wq][threadIdx.x];
        if (threadIdx.x < kMc) { if (c0 + (int)threadIdx.x < ncol) partial[(size_t)blockIdx.x * nout + c0 + threadIdx.x] = t; }
        else if (do_ww) partial[(size_t)blockIdx.x * nout + ncol] = t;
    }
}

void vec_mdot(Ctx& c, const double* V, int64_t ld, int ncol, const double* w, int64_t n, double* d_h, bool with_ww) {
    int nout = ncol + (with_ww ? 1 : 0);
    if (nout == 0) return;
    int grid = stream_grid(c, n, kBlock, kRpt, 2);
    while ((int64_t)grid * nout > Ctx::kScal && grid > 1) grid /= 2;
    PORO_REQUIRE((int64_t)grid * nout <= Ctx::kScal, "mdot scratch too small");
    const int nchunk = std::max(1, (ncol + kMc - 1) / kMc);
    k_mdot<<<dim3(grid, nchunk), kBlock, 0, c.stream>>>(V, ld, ncol, w, n, with_ww, c.d_scal);
    PORO_LAUNCH_CHECK(c);
    k_sum_partials<<<nout < 64 ? nout : 64, kBlock, 0, c.stream>>>(c.d_scal, grid, nout, d_h);
    PORO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------
// multi-axpy fused with the norm of the result: w -= V h ; nrm2 = ||w||^2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_maxpy_norm(double* __restrict__ w, const double* __restrict__ V, int64_t ld,
                                                       int ncol, const double* __restrict__ h, int64_t n,
                                                       double* __restrict__ partial) {
    extern __shared__ double hs[];
    for (int j = threadIdx.x; j < ncol; j += kBlock) hs[j] = h[j];
    __syncthreads();
    double nrm = 0.0;
    const int64_t tile = (int64_t)kBlock * kRpt;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        double wv[kRpt];
        int64_t row[kRpt];
#pragma unroll
        for (int r = 0; r < kRpt; ++r) {
            row[r] = base + (int64_t)r * kBlock + threadIdx.x;
            wv[r] = row[r] < n ? w[row[r]] : 0.0;
        }
        for (int c0 = 0; c0 < ncol; c0 += kCchunk) {
            double v[kCchunk][kRpt];
#pragma unroll
            for (int cc = 0; cc < kCchunk; ++cc) {
                const double* col = V + (int64_t)(c0 + cc) * ld;
#pragma unroll
                for (int r = 0; r < kRpt; ++r) v[cc][r] = (c0 + cc < ncol && row[r] < n) ? col[row[r]] : 0.0;
            }
#pragma unroll
            for (int cc = 0; cc < kCchunk; ++cc) {
                double hc = c0 + cc < ncol ? hs[c0 + cc] : 0.0;
#pragma unroll
                for (int r = 0; r < kRpt; ++r) wv[r] = fma(-hc, v[cc][r], wv[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRpt; ++r)
            if (row[r] < n) { w[row[r]] = wv[r]; nrm = fma(wv[r], wv[r], nrm); }
    }
    __shared__ double sm[kBlock / 32];
    nrm = warp_sum(nrm);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int wq = 0; wq < kBlock / 32; ++wq) t += sm[wq];
        partial[blockIdx.x] = t;
    }
}

// w = scale * (w - V h): the Gram-Schmidt update and the normalisation of the new basis vector in one pass
__global__ void __launch_bounds__(kBlock) k_maxpy_scale(double* __restrict__ w, const double* __restrict__ V, int64_t ld,
                                                        int ncol, const double* __restrict__ h, int64_t n, double scale) {
    extern __shared__ double hs[];
    for (int j = threadIdx.x; j < ncol; j += kBlock) hs[j] = h[j];
    __syncthreads();
    const int64_t tile = (int64_t)kBlock * kRpt;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        double wv[kRpt];
        int64_t row[kRpt];
#pragma unroll
        for (int r = 0; r < kRpt; ++r) {
            row[r] = base + (int64_t)r * kBlock + threadIdx.x;
            wv[r] = row[r] < n ? w[row[r]] : 0.0;
        }
        for (int c0 = 0; c0 < ncol; c0 += kCchunk) {
            double v[kCchunk][kRpt];
#pragma unroll
            for (int cc = 0; cc < kCchunk; ++cc) {
                const double* col = V + (int64_t)(c0 + cc) * ld;
#pragma unroll
                for (int r = 0; r < kRpt; ++r) v[cc][r] = (c0 + cc < ncol && row[r] < n) ? col[row[r]] : 0.0;
            }
#pragma unroll
            for (int cc = 0; cc < kCchunk; ++cc) {
                double hc = c0 + cc < ncol ? hs[c0 + cc] : 0.0;
#pragma unroll
                for (int r = 0; r < kRpt; ++r) wv[r] = fma(-hc, v[cc][r], wv[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRpt; ++r)
            if (row[r] < n) w[row[r]] = wv[r] * scale;
    }
}

void vec_maxpy_scale(Ctx& c, double* w, const double* V, int64_t ld, int ncol, const double* d_h, int64_t n, double scale) {
    int grid = stream_grid(c, n, kBlock, kRpt, 2);
    size_t smem = (size_t)(ncol > 0 ? ncol : 1) * sizeof(double);
    k_maxpy_scale<<<grid, kBlock, smem, c.stream>>>(w, V, ld, ncol, d_h, n, scale);
    PORO_LAUNCH_CHECK(c);
}

void vec_maxpy_norm(Ctx& c, double* w, const double* V, int64_t ld, int ncol, const double* d_h, int64_t n, double* d_nrm2) {
    int grid = stream_grid(c, n, kBlock, kRpt, 2);
    size_t smem = (size_t)(ncol > 0 ? ncol : 1) * sizeof(double);
    k_maxpy_norm<<<grid, kBlock, smem, c.stream>>>(w, V, ld, ncol, d_h, n, c.d_scal);
    PORO_LAUNCH_CHECK(c);
    k_sum_partials<<<1, kBlock, 0, c.stream>>>(c.d_scal, grid, 1, d_nrm2);
    PORO_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) k_maxpy_add(double* __restrict__ y, const double* __restrict__ V, int64_t ld,
                                                      int ncol, const double* __restrict__ h, int64_t n) {
    extern __shared__ double hs[];
    for (int j = threadIdx.x; j < ncol; j += kBlock) hs[j] = h[j];
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = y[i];
        for (int j = 0; j < ncol; ++j) a = fma(hs[j], V[(int64_t)j * ld + i], a);
        y[i] = a;
    }
}

void vec_maxpy_host(Ctx& c, double* y, const double* V, int64_t ld, int ncol, const double* h_host, int64_t n) {
    if (ncol <= 0) return;
    PORO_REQUIRE(ncol <= 4096, "too many columns");
    // coefficients travel through the pinned scratch (upper half, not used by fetch())
    double* hp = c.h_pin + 4096;
    memcpy(hp, h_host, ncol * sizeof(double));
    double* dh = c.d_scal + Ctx::kScal;   // tail region reserved for coefficient vectors
    PORO_CUDA(cudaMemcpyAsync(dh, hp, ncol * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    k_maxpy_add<<<stream_grid(c, n, kBlock, 2), kBlock, ncol * sizeof(double), c.stream>>>(y, V, ld, ncol, dh, n);
    PORO_LAUNCH_CHECK(c);
    PORO_CUDA(cudaStreamSynchronize(c.stream));   // hp may be overwritten by the next call
}

// ------------------------------------------------------------------------------------------
// infinity norms of up to 4 segments of one vector (per-field residual monitor, lib/Solver.py:27-32)
// ------------------------------------------------------------------------------------------
struct Segs { int64_t off[4], len[4]; };
__global__ void __launch_bounds__(kBlock) k_amax(const double* __restrict__ x, Segs sg, double* __restrict__ out) {
    // one CTA per segment: monitors only, not a hot kernel
    const int j = blockIdx.x;
    __shared__ double sm[kBlock / 32];
    double m = 0.0;
    const double* p = x + sg.off[j];
    for (int64_t i = threadIdx.x; i < sg.len[j]; i += kBlock) m = fmax(m, fabs(p[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) t = fmax(t, sm[w]);
        out[j] = t;
    }
}
void vec_amax_segments(Ctx& c, const double* x, const int64_t* off, const int64_t* len, int nseg, double* d_out) {
    PORO_REQUIRE(nseg >= 1 && nseg <= 4, "vec_amax_segments: 1..4 segments");
    Segs sg;
    for (int j = 0; j < 4; ++j) { sg.off[j] = off[j < nseg ? j : 0]; sg.len[j] = len[j < nseg ? j : 0]; }
    k_amax<<<nseg, kBlock, 0, c.stream>>>(x, sg, d_out);
    PORO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------
// cross-rank sum + host read-back
// ------------------------------------------------------------------------------------------
void allreduce_sum(Ctx& c, double* d_vals, int k) {
    if (c.nranks > 1 && !c.local_only) dist_allreduce_sum(c, d_vals, k);
}

void allreduce_max(Ctx& c, double* d_vals, int k) {
    if (c.nranks > 1 && !c.local_only) dist_allreduce_max(c, d_vals, k);
}

void fetch(Ctx& c, const double* d_vals, int k, double* host) {
    PORO_REQUIRE(k <= 4096, "fetch: too many scalars");
    PORO_CUDA(cudaMemcpyAsync(c.h_pin, d_vals, k * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(host, c.h_pin, k * sizeof(double));
}

double dot_host(Ctx& c, const double* x, const double* y, int64_t n) {
    double* d = c.d_scal + Ctx::kScal + 4096;
    const double* xs[1] = {x};
    const double* ys[1] = {y};
    vec_dots(c, 1, xs, ys, n, d);
    allreduce_sum(c, d, 1);
    double h;
    fetch(c, d, 1, &h);
    return h;
}

double norm2_host(Ctx& c, const double* x, int64_t n) { return sqrt(dot_host(c, x, x, n)); }

}  // namespace poro
