// spmv.cu -- fp64 CSR sparse matrix-vector products (the dominant kernel of the path).
//
// Takes over MatMult_SeqAIJ/MPIAIJ behind `A.mat()` as the KSP operator (lib/Solver.py:95),
// `self.matA * x` (lib/AAR.py:56,135), and the coupling products Mfp_s.mult / Mf_p.mult /
// Ms_f.mult / Ms_p.mult of the block sweeps (lib/Preconditioner.py:180,184,192-193,232).
// The reference's "mult then aypx(-1)" pairs are fused: y = z - A x in one pass.
//
// Kernel: "vector CSR" -- a group of L lanes (L = 2..32, chosen from the mean row length)
// walks one row; lane l reads val/col at start+l, start+l+L, ... so every warp-wide load of
// the matrix streams is a contiguous, fully coalesced segment; the matrix streams use
// streaming (evict-first) loads so that L1/L2 keep the gathered x entries; partial sums
// are combined by warp shuffles.  HBM-bound: algorithmic bytes = 12 nnz + 4 (nrows+1) +
// 8 nrows + 8 ncols.
#include "common.cuh"

namespace poro {

static constexpr int kSpmvBlock = 256;

struct Epilogue {
    int mode;                 // SpmvMode, or 3 = Chebyshev step, 4 = dot
    const double* z;          // SUB/ADD source
    // Chebyshev step
    const double* d_old; double* d_new; double* r; double* xv; const double* dinv; double c1, c2;
};

template <int L>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, L);
    return v;
}

template <int L, int MODE>
__global__ void __launch_bounds__(kSpmvBlock) k_spmv(int nrows, const int* __restrict__ rowptr,
                                                     const int* __restrict__ col, const double* __restrict__ val,
                                                     const double* __restrict__ x, double* __restrict__ y, Epilogue ep,
                                                     double* __restrict__ dot_partial) {
    const int lane = threadIdx.x & (L - 1);
    const int row = (int)(((int64_t)blockIdx.x * kSpmvBlock + threadIdx.x) / L);
    double sum = 0.0;
    if (row < nrows) {
        const int start = rowptr[row], end = rowptr[row + 1];
        int k = start + lane;
        // two independent streams per lane for memory-level parallelism
        double s0 = 0.0, s1 = 0.0;
        for (; k + L < end; k += 2 * L) {
            const int c0 = __ldcs(col + k), c1 = __ldcs(col + k + L);
            const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + L);
            s0 = fma(v0, __ldg(x + c0), s0);
            s1 = fma(v1, __ldg(x + c1), s1);
        }
        if (k < end) s0 = fma(__ldcs(val + k), __ldg(x + __ldcs(col + k)), s0);
        sum = s0 + s1;
    }
    sum = group_sum<L>(sum);
    double contrib = 0.0;
    if (row < nrows && lane == 0) {
        if (MODE == SPMV_SET) y[row] = sum;
        else if (MODE == SPMV_SUB) y[row] = ep.z[row] - sum;
        else if (MODE == SPMV_ADD) y[row] = ep.z[row] + sum;
        else if (MODE == 3) {
            const double rn = ep.r[row] - sum;
            const double dn = ep.c1 * ep.d_old[row] + ep.c2 * ep.dinv[row] * rn;
            ep.r[row] = rn;
            ep.d_new[row] = dn;
            ep.xv[row] += dn;
        } else if (MODE == 4) {
            y[row] = sum;
            contrib = sum * x[row];
        }
    }
    if (MODE == 4) {
        // block-level reduction of p.w, one partial per block (summed later in a fixed order)
        __shared__ double sm[kSpmvBlock / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = contrib;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kSpmvBlock / 32; ++w) t += sm[w];
            dot_partial[blockIdx.x] = t;
        }
    }
}

void csr_choose_lanes(Csr& A) {
    double a = A.avg_row();
    A.lanes = a <= 3 ? 2 : a <= 6 ? 4 : a <= 12 ? 8 : a <= 24 ? 16 : 32;
}

template <int MODE>
static void launch_spmv(Ctx& c, const Csr& A, const double* x, double* y, const Epilogue& ep, double* dot_partial, int* grid_out) {
    if (A.nrows == 0) { if (grid_out) *grid_out = 0; return; }
    int L = A.lanes ? A.lanes : 32;
    int grid = ceil_div((int64_t)A.nrows * L, kSpmvBlock);
    if (grid_out) *grid_out = grid;
#define GO(LL) k_spmv<LL, MODE><<<grid, kSpmvBlock, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, ep, dot_partial)
    switch (L) {
        case 2: GO(2); break;
        case 4: GO(4); break;
        case 8: GO(8); break;
        case 16: GO(16); break;
        default: GO(32); break;
    }
#undef GO
    PORO_LAUNCH_CHECK(c);
}

void spmv(Ctx& c, const Csr& A, const double* x, double* y, SpmvMode mode, const double* z) {
    Epilogue ep{};
    ep.mode = mode;
    ep.z = z;
    if (mode == SPMV_SET) launch_spmv<SPMV_SET>(c, A, x, y, ep, nullptr, nullptr);
    else if (mode == SPMV_SUB) launch_spmv<SPMV_SUB>(c, A, x, y, ep, nullptr, nullptr);
    else launch_spmv<SPMV_ADD>(c, A, x, y, ep, nullptr, nullptr);
}

void spmv_cheb_step(Ctx& c, const Csr& A, const double* d_old, double* d_new, double* r, double* x,
                    const double* dinv, double c1, double c2) {
    Epilogue ep{};
    ep.mode = 3;
    ep.d_old = d_old; ep.d_new = d_new; ep.r = r; ep.xv = x; ep.dinv = dinv; ep.c1 = c1; ep.c2 = c2;
    launch_spmv<3>(c, A, d_old, nullptr, ep, nullptr, nullptr);
}

__global__ void k_sum_to(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sm[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        *out = t;
    }
}

void spmv_dot(Ctx& c, const Csr& A, const double* p, double* w, double* d_dot) {
    PORO_REQUIRE(A.nrows == A.ncols || true, "");
    Epilogue ep{};
    ep.mode = 4;
    int L = A.lanes ? A.lanes : 32;
    int grid = ceil_div((int64_t)A.nrows * L, kSpmvBlock);
    if (grid <= Ctx::kScal) {
        launch_spmv<4>(c, A, p, w, ep, c.d_scal, nullptr);
        k_sum_to<<<1, 256, 0, c.stream>>>(c.d_scal, grid, d_dot);
        PORO_LAUNCH_CHECK(c);
    } else {
        // too many blocks for the partial buffer: plain SpMV + separate dot
        launch_spmv<SPMV_SET>(c, A, p, w, ep, nullptr, nullptr);
        const double* xs[1] = {p};
        const double* ys[1] = {w};
        vec_dots(c, 1, xs, ys, A.nrows, d_dot);
    }
}

}  // namespace poro
