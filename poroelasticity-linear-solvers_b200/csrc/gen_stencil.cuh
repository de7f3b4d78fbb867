// gen_stencil.cuh -- device-side generator of the structured-mesh field blocks (ROUND-2 WORK IN PROGRESS: built and
// unit-tested, NOT yet linked into libporo.so).
//
// Replaces the host assembly of lib/Assembler.py:66-221 + DirichletBC.apply (lib/Poromechanics.py:76-83) for dolfin's
// UnitSquareMesh / UnitCubeMesh: with constant coefficients the assembled block is ONE macro-cell matrix scattered
// over all cells, so the block row of a node is one of <= 4^d (P2 row lattice) or 3^d (P1 row lattice) class
// stencils, shifted (hostfem/stencil.py derives the tables and is the numpy statement of this file).
//
// The row logic lives in __host__ __device__ functions so that the SAME code is checked on the CPU
// (tests/test_gen_stencil_host.py compiles tests/gen_stencil_harness.cpp with g++ against this header);
// gen_stencil.cu wraps it in two kernels: counts (then an exclusive scan) and a warp-per-block-row fill whose
// stores are coalesced over the (entry, value) index.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PORO_GEN_HD __host__ __device__ __forceinline__
#else
#define PORO_GEN_HD inline
#endif

namespace porogen {

struct BlockTable {
    int dim, N;               // cells per side
    int kr, kc;               // row / column lattice: 2 = P2 nodes (2N+1 per axis), 1 = P1 nodes (N+1 per axis)
    int br, bc;               // block size (dofs per row node x dofs per column node)
    int diag_block;           // 1: rows and columns are the same field (a Dirichlet row gets a unit diagonal)
    const int32_t* cls_ptr;   // (ncls + 1): entries of class c are cls_ptr[c] .. cls_ptr[c+1]
    const int64_t* off;       // column-node offset of an entry relative to the row's base column node
    const double* vals;       // br*bc values per entry, row-major
};

// position class of a coordinate on one axis; the codes are the order of hostfem/stencil.py:_axis_classes
PORO_GEN_HD int axis_class(int kind, int N, int x) {
    if (kind == 2) return (x & 1) ? 0 : x == 0 ? 1 : x == 2 * N ? 2 : 3;   // odd | first plane | last plane | even interior
    return x == 0 ? 0 : x == N ? 1 : 2;                                   // first | last | interior
}
PORO_GEN_HD int lattice_extent(int kind, int N) { return kind == 2 ? 2 * N + 1 : N + 1; }
PORO_GEN_HD int64_t lattice_nodes(int kind, int N, int dim) {
    int64_t n = 1;
    for (int m = 0; m < dim; ++m) n *= lattice_extent(kind, N);
    return n;
}

// class id (axis 0 fastest) and base column node (= column node of cell c0's local origin) of a row node
PORO_GEN_HD void row_info(const BlockTable& t, int64_t node, int& cls, int64_t& base) {
    const int Lr = lattice_extent(t.kr, t.N), Lc = lattice_extent(t.kc, t.N);
    const int ncl = t.kr == 2 ? 4 : 3, sc = t.kc == 2 ? 2 : 1;
    cls = 0;
    base = 0;
    int64_t mulc = 1;
    int mulk = 1;
    for (int m = 0; m < t.dim; ++m) {
        const int x = (int)(node % Lr);
        node /= Lr;
        cls += axis_class(t.kr, t.N, x) * mulk;
        mulk *= ncl;
        const int c0 = t.kr == 2 ? x / 2 : x;
        base += (int64_t)sc * c0 * mulc;
        mulc *= Lc;
    }
}

PORO_GEN_HD int row_entries(const BlockTable& t, int64_t node) {
    int cls;
    int64_t base;
    row_info(t, node, cls, base);
    return t.cls_ptr[cls + 1] - t.cls_ptr[cls];
}

// One (entry e, value v) item of the block row of `node`: column node and value with DirichletBC.apply semantics.
// bc_row: one flag per scalar row dof of the block (node * br + i) or nullptr.
PORO_GEN_HD void row_item(const BlockTable& t, int64_t node, int cls, int64_t base, int e, int v, const uint8_t* bc_row,
                          int32_t& col_node, double& value) {
    const int p = t.cls_ptr[cls] + e;
    const int64_t c = base + t.off[p];
    col_node = (int32_t)c;
    value = t.vals[(int64_t)p * t.br * t.bc + v];
    if (bc_row) {
        const int i = v / t.bc, j = v % t.bc;
        if (bc_row[node * t.br + i]) value = (t.diag_block && c == node && i == j) ? 1.0 : 0.0;
    }
}

}  // namespace porogen
