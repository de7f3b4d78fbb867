// gen_stencil.cu -- kernels around gen_stencil.cuh (ROUND-2 WORK IN PROGRESS, not linked into libporo.so yet).
//   k_gen_counts : entries per block row (input of the exclusive scan that gives the BSR row pointer)
//   k_gen_fill   : one warp per block row; lanes run over the flattened (entry, value) index, so the value stores of a
//                  row are contiguous and the whole block is written once at store bandwidth
// Rows are generated for a node range [node0, node0 + nrows): a rank passes the nodes of its z-slab.
#include <cuda_runtime.h>

#include "gen_stencil.cuh"

namespace porogen {

__global__ void k_gen_counts(BlockTable t, int64_t node0, int64_t nrows, int32_t* __restrict__ counts) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < nrows) counts[r] = row_entries(t, node0 + r);
}

__global__ void k_gen_fill(BlockTable t, int64_t node0, int64_t nrows, const int64_t* __restrict__ rowptr,
                           const uint8_t* __restrict__ bc_row, int32_t* __restrict__ col, double* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= nrows) return;
    const int64_t node = node0 + r;
    int cls;
    int64_t base;
    row_info(t, node, cls, base);
    const int cnt = t.cls_ptr[cls + 1] - t.cls_ptr[cls];
    const int bsz = t.br * t.bc;
    const int64_t dst = rowptr[r];
    for (int q = lane; q < cnt * bsz; q += 32) {
        const int e = q / bsz, v = q - e * bsz;
        int32_t cn;
        double x;
        row_item(t, node, cls, base, e, v, bc_row, cn, x);
        val[dst * bsz + q] = x;
        if (v == 0) col[dst + e] = cn;
    }
}

// launch helpers (stream-ordered; the scan between the two is cub::DeviceScan in the caller)
void gen_counts(const BlockTable& t, int64_t node0, int64_t nrows, int32_t* counts, cudaStream_t s) {
    if (nrows) k_gen_counts<<<(unsigned)((nrows + 255) / 256), 256, 0, s>>>(t, node0, nrows, counts);
}
void gen_fill(const BlockTable& t, int64_t node0, int64_t nrows, const int64_t* rowptr, const uint8_t* bc_row, int32_t* col,
              double* val, cudaStream_t s) {
    if (nrows) k_gen_fill<<<(unsigned)((nrows * 32 + 255) / 256), 256, 0, s>>>(t, node0, nrows, rowptr, bc_row, col, val);
}

}  // namespace porogen
