// Host harness for csrc/gen_stencil.cuh: runs the __host__ __device__ row logic of the device-side matrix
// generator over all rows on the CPU (g++), for tests/test_gen_stencil_host.py.
#include <stdint.h>

#include "gen_stencil.cuh"

extern "C" {

int64_t gen_host_counts(int dim, int N, int kr, int kc, int br, int bc, int diag, const int32_t* cls_ptr, const int64_t* off,
                        const double* vals, int64_t node0, int64_t nrows, int64_t* rowptr) {
    porogen::BlockTable t{dim, N, kr, kc, br, bc, diag, cls_ptr, off, vals};
    rowptr[0] = 0;
    for (int64_t r = 0; r < nrows; ++r) rowptr[r + 1] = rowptr[r] + porogen::row_entries(t, node0 + r);
    return rowptr[nrows];
}

void gen_host_fill(int dim, int N, int kr, int kc, int br, int bc, int diag, const int32_t* cls_ptr, const int64_t* off,
                   const double* vals, int64_t node0, int64_t nrows, const int64_t* rowptr, const uint8_t* bc_row,
                   int32_t* col, double* val) {
    porogen::BlockTable t{dim, N, kr, kc, br, bc, diag, cls_ptr, off, vals};
    const int bsz = br * bc;
    for (int64_t r = 0; r < nrows; ++r) {
        const int64_t node = node0 + r;
        int cls;
        int64_t base;
        porogen::row_info(t, node, cls, base);
        const int cnt = cls_ptr[cls + 1] - cls_ptr[cls];
        for (int lane = 0; lane < 32; ++lane)                  // the warp of k_gen_fill, lane by lane
            for (int q = lane; q < cnt * bsz; q += 32) {
                const int e = q / bsz, v = q - e * bsz;
                int32_t cn;
                double x;
                porogen::row_item(t, node, cls, base, e, v, bc_row, cn, x);
                val[rowptr[r] * bsz + q] = x;
                if (v == 0) col[rowptr[r] + e] = cn;
            }
    }
}
}
