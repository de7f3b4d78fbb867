// distamg.cu -- exchange primitives and level set-up of the DISTRIBUTED smoothed-aggregation hierarchy.
//
// The reference's inner solves are global across MPI ranks (hypre BoomerAMG / MUMPS over the communicator:
// petsc-options-inexact:16-24,88-96, paper-scripts/robustness_2d.sh:29 `mpirun -np 8`, lib/Preconditioner.py:94-118).
// This file is the CUDA + NCCL statement of oracle/distamg_rank.py (verified over gloo in
// tests/test_oracle_distamg_gloo.py); the names below are the names there.  Per level and rank:
//     aggregation + tentative prolongator T     rank-local, kernels of amg.cu ("uncoupled" aggregates)  -- no communication
//     P  = T - w D^-1 (A [T ; T_ghost])          csr_spgemm                                       -- halo_rows(T)
//     AP = A [P ; P_ghost]                       csr_spgemm                                       -- halo_rows(P)
//     Ac = (P_ext[:, owned coarse])^T [AP ; AP_ghost]   csr_transpose + csr_spgemm                -- halo_rows(AP)
//     R  = (P_ext[:, owned coarse])^T            by-product; restriction reads [owned | ghost] residuals
//     next level: localize(Ac), DistPlan from its ghost columns (one handshake)
// Matrices in flight carry GLOBAL column ids (int32: < 2^31 dofs per level) in `Csr::col` and are localised to
// [owned | ghost] only when they become level operators.  The same sparse-row exchange assembles the exact selfp Schur
// complement of the owned rows (dist_selfp_schur).
#include "distamg.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <numeric>

#include "dist.cuh"

namespace poro {

static constexpr int kB = 256;
template <class F>
__global__ void __launch_bounds__(kB) k_for_d(int64_t n, F f) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <class F>
static void pfor(Ctx& c, int64_t n, F f) {
    if (n <= 0) return;
    int64_t g = (n + kB - 1) / kB, cap = (int64_t)c.sm_count * 16;
    k_for_d<<<(int)(g < cap ? g : cap), kB, 0, c.stream>>>(n, f);
    PORO_LAUNCH_CHECK(c);
}

// out[0..n] = exclusive scan of in[0..n-1] (out has n+1 entries); returns the total (synchronises)
static int64_t scan_counts(Ctx& c, const int* in, int* out, int64_t n) {
    PORO_CUDA(cudaMemsetAsync(out, 0, sizeof(int), c.stream));
    if (n == 0) { PORO_CUDA(cudaStreamSynchronize(c.stream)); return 0; }
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, in, out + 1, n, c.stream);
    DBuf<char> tmp(tb);
    cub::DeviceScan::InclusiveSum(tmp.p, tb, in, out + 1, n, c.stream);
    c.launches++;
    int total = 0;
    PORO_CUDA(cudaMemcpyAsync(&total, out + n, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    return total;
}

// ---- halo_vec: `width` doubles per row (vectors: 1) ------------------------------------------------------------------
void dist_halo_vec(Ctx& c, DistPlan& plan, const double* x_owned, int width, double* ghost_out) {
    if (c.nranks <= 1) return;
    if (width == 1 && plan.p2p.ready) { p2p_exchange(c, plan.p2p, plan.send_idx.p, x_owned, ghost_out); return; }
    const size_t nn = plan.neigh.size();
    const int64_t nsend = nn ? plan.send_ptr[nn] : 0;
    if (plan.send_buf.n < (size_t)nsend * width) plan.send_buf.alloc((size_t)nsend * width);
    if (nsend) {
        if (width == 1) vec_gather(c, plan.send_buf.p, x_owned, plan.send_idx.p, nsend);
        else {
            const int* idx = plan.send_idx.p;
            double* sb = plan.send_buf.p;
            pfor(c, nsend * width, [=] __device__(int64_t t) { sb[t] = x_owned[(int64_t)idx[t / width] * width + t % width]; });
        }
    }
    if (!nn) return;
    dist_group_begin(c);
    for (size_t k = 0; k < nn; ++k) {
        const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
        if (ns) dist_send_bytes(c, plan.send_buf.p + plan.send_ptr[k] * width, (size_t)ns * width * sizeof(double), plan.neigh[k]);
        if (nr) dist_recv_bytes(c, ghost_out + plan.recv_ptr[k] * width, (size_t)nr * width * sizeof(double), plan.neigh[k]);
    }
    dist_group_end(c);
}

// ---- level-0 plan from the field halo plan ---------------------------------------------------------------------------
void dist_plan_from_halo(Ctx& c, const HaloField& hf, int64_t n_owned, DistPlan& plan) {
    const int R = c.nranks, me = c.rank;
    std::vector<int64_t> sizes;
    const int64_t mine = n_owned;
    dist_allgather_i64(c, &mine, 1, sizes);
    plan.offsets.assign((size_t)R + 1, 0);
    for (int r = 0; r < R; ++r) plan.offsets[r + 1] = plan.offsets[r] + sizes[r];
    PORO_REQUIRE(plan.offsets[R] < 2147483647LL, "a distributed block has more than 2^31 rows");
    plan.offset = plan.offsets[me];
    plan.n_owned = (int)n_owned;
    plan.n_ghost = (int)hf.n_halo;
    plan.neigh = c.neigh;
    plan.send_ptr = hf.send_ptr;
    plan.recv_ptr = hf.recv_ptr;
    if (plan.send_ptr.empty()) { plan.send_ptr.assign(plan.neigh.size() + 1, 0); plan.recv_ptr.assign(plan.neigh.size() + 1, 0); }
    const int64_t nsend = plan.send_ptr.back();
    plan.send_idx.alloc((size_t)nsend);
    if (nsend) PORO_CUDA(cudaMemcpyAsync(plan.send_idx.p, hf.send_idx.p, (size_t)nsend * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
    // global ids of the ghosts: the owners send (offset + local index), exact in a double below 2^53
    DBuf<double> ids((size_t)n_owned), gh((size_t)plan.n_ghost);
    {
        double* p = ids.p; const double off = (double)plan.offset;
        pfor(c, n_owned, [=] __device__(int64_t i) { p[i] = off + (double)i; });
    }
    dist_halo_vec(c, plan, ids.p, 1, gh.p);
    plan.ghost_gid.alloc((size_t)plan.n_ghost);
    {
        int* g = plan.ghost_gid.p; const double* s = gh.p;
        pfor(c, plan.n_ghost, [=] __device__(int64_t i) { g[i] = (int)(s[i] + 0.5); });
    }
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    p2p_slots_setup(c, plan.neigh, plan.send_ptr, plan.recv_ptr, plan.p2p);
}

// ---- the handshake: all-gather how many ids every rank reads from every other, then send the id lists to their owners
void dist_plan_build(Ctx& c, const std::vector<int64_t>& offsets, DBuf<int>&& ghost_gid, int n_ghost, DistPlan& plan) {
    const int R = c.nranks, me = c.rank;
    plan.offsets = offsets;
    plan.offset = offsets[me];
    plan.n_owned = (int)(offsets[me + 1] - offsets[me]);
    plan.n_ghost = n_ghost;
    plan.ghost_gid = std::move(ghost_gid);
    std::vector<int> gh((size_t)n_ghost);
    if (n_ghost) PORO_CUDA(cudaMemcpyAsync(gh.data(), plan.ghost_gid.p, (size_t)n_ghost * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<int64_t> need((size_t)R, 0), start((size_t)R + 1, 0);          // ids this rank reads from rank q
    for (int g : gh) {
        int q = (int)(std::upper_bound(offsets.begin(), offsets.end(), (int64_t)g) - offsets.begin()) - 1;
        PORO_REQUIRE(q >= 0 && q < R && q != me, "ghost id owned by this rank or out of range");
        need[q]++;
    }
    for (int q = 0; q < R; ++q) start[q + 1] = start[q] + need[q];              // gh is ascending: grouped by owner already
    std::vector<int64_t> all;
    dist_allgather_i64(c, need.data(), R, all);                                 // all[p * R + q] = ids rank p reads from rank q
    plan.neigh.clear();
    for (int q = 0; q < R; ++q)
        if (q != me && (all[(size_t)me * R + q] > 0 || all[(size_t)q * R + me] > 0)) plan.neigh.push_back(q);
    const size_t nn = plan.neigh.size();
    plan.send_ptr.assign(nn + 1, 0);
    plan.recv_ptr.assign(nn + 1, 0);
    for (size_t k = 0; k < nn; ++k) {
        plan.send_ptr[k + 1] = plan.send_ptr[k] + all[(size_t)plan.neigh[k] * R + me];
        plan.recv_ptr[k + 1] = plan.recv_ptr[k] + all[(size_t)me * R + plan.neigh[k]];
    }
    PORO_REQUIRE(plan.recv_ptr[nn] == n_ghost, "ghost list and neighbour counts disagree");
    plan.send_idx.alloc((size_t)plan.send_ptr[nn]);
    // the id lists travel as global ids and are made local on arrival
    if (nn) {
        dist_group_begin(c);
        for (size_t k = 0; k < nn; ++k) {
            const int q = plan.neigh[k];
            const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
            if (nr) dist_send_bytes(c, plan.ghost_gid.p + start[q], (size_t)nr * sizeof(int), q);
            if (ns) dist_recv_bytes(c, plan.send_idx.p + plan.send_ptr[k], (size_t)ns * sizeof(int), q);
        }
        dist_group_end(c);
    }
    {
        int* s = plan.send_idx.p;
        const int off = (int)plan.offset;
        pfor(c, plan.send_ptr[nn], [=] __device__(int64_t i) { s[i] -= off; });
    }
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    p2p_slots_setup(c, plan.neigh, plan.send_ptr, plan.recv_ptr, plan.p2p);
}

// ---- halo_rows: the sparse rows of the boundary rows (GLOBAL column ids) ---------------------------------------------
// Two grouped exchanges: row lengths, then packed column ids and values.  `ghost` gets n_ghost rows in ghost order.
void dist_halo_rows(Ctx& c, const DistPlan& plan, const Csr& M, Csr& ghost) {
    const size_t nn = plan.neigh.size();
    const int64_t nsend = nn ? plan.send_ptr[nn] : 0;
    // lengths of the rows we send, and their packed offsets
    DBuf<int> slen((size_t)nsend + 1), sptr((size_t)nsend + 1);
    {
        const int* idx = plan.send_idx.p; const int* rp = M.rowptr.p; int* L = slen.p;
        pfor(c, nsend, [=] __device__(int64_t i) { L[i] = rp[idx[i] + 1] - rp[idx[i]]; });
    }
    const int64_t send_nnz = scan_counts(c, slen.p, sptr.p, nsend);
    DBuf<int> rlen((size_t)plan.n_ghost + 1);
    if (nn) {
        dist_group_begin(c);
        for (size_t k = 0; k < nn; ++k) {
            const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
            if (ns) dist_send_bytes(c, slen.p + plan.send_ptr[k], (size_t)ns * sizeof(int), plan.neigh[k]);
            if (nr) dist_recv_bytes(c, rlen.p + plan.recv_ptr[k], (size_t)nr * sizeof(int), plan.neigh[k]);
        }
        dist_group_end(c);
    }
    ghost.nrows = plan.n_ghost;
    ghost.ncols = M.ncols;
    ghost.rowptr.alloc((size_t)plan.n_ghost + 1);
    ghost.nnz = scan_counts(c, rlen.p, ghost.rowptr.p, plan.n_ghost);
    ghost.col.alloc((size_t)ghost.nnz);
    ghost.val.alloc((size_t)ghost.nnz);
    // packed boundaries per neighbour (host needs them for the message sizes)
    std::vector<int> sp_h(nn + 1, 0), rp_h(nn + 1, 0);
    for (size_t k = 0; k <= nn; ++k) {
        PORO_CUDA(cudaMemcpyAsync(&sp_h[k], sptr.p + plan.send_ptr[k], sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        PORO_CUDA(cudaMemcpyAsync(&rp_h[k], ghost.rowptr.p + plan.recv_ptr[k], sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    }
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    DBuf<int> scol((size_t)send_nnz);
    DBuf<double> sval((size_t)send_nnz);
    {
        const int* idx = plan.send_idx.p; const int* rp = M.rowptr.p; const int* cc = M.col.p; const double* vv = M.val.p;
        const int* sp = sptr.p; int* oc = scol.p; double* ov = sval.p;
        pfor(c, nsend * 32, [=] __device__(int64_t t) {
            const int64_t i = t >> 5; const int lane = (int)(t & 31);
            const int a = rp[idx[i]], n = rp[idx[i] + 1] - a, o = sp[i];
            for (int q = lane; q < n; q += 32) { oc[o + q] = cc[a + q]; ov[o + q] = vv[a + q]; }
        });
    }
    if (nn) {
        dist_group_begin(c);
        for (size_t k = 0; k < nn; ++k) {
            const int64_t ns = sp_h[k + 1] - sp_h[k], nr = rp_h[k + 1] - rp_h[k];
            if (ns) {
                dist_send_bytes(c, scol.p + sp_h[k], (size_t)ns * sizeof(int), plan.neigh[k]);
                dist_send_bytes(c, sval.p + sp_h[k], (size_t)ns * sizeof(double), plan.neigh[k]);
            }
            if (nr) {
                dist_recv_bytes(c, ghost.col.p + rp_h[k], (size_t)nr * sizeof(int), plan.neigh[k]);
                dist_recv_bytes(c, ghost.val.p + rp_h[k], (size_t)nr * sizeof(double), plan.neigh[k]);
            }
        }
        dist_group_end(c);
    }
    PORO_CUDA(cudaStreamSynchronize(c.stream));      // the packing buffers die with this scope
    csr_choose_lanes(ghost);
}

// ---- [top ; bottom] with equal column spaces ---------------------------------------------------------------------------
void csr_vstack(Ctx& c, const Csr& top, const Csr& bot, Csr& out) {
    PORO_REQUIRE(top.ncols == bot.ncols, "csr_vstack: column spaces differ");
    PORO_REQUIRE(top.nnz + bot.nnz < 2147483647LL, "csr_vstack: more than 2^31 nonzeros");
    out.nrows = top.nrows + bot.nrows;
    out.ncols = top.ncols;
    out.nnz = top.nnz + bot.nnz;
    out.rowptr.alloc((size_t)out.nrows + 1);
    out.col.alloc((size_t)out.nnz);
    out.val.alloc((size_t)out.nnz);
    PORO_CUDA(cudaMemcpyAsync(out.rowptr.p, top.rowptr.p, ((size_t)top.nrows + 1) * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
    {
        const int* rp = bot.rowptr.p; int* o = out.rowptr.p + top.nrows; const int shift = (int)top.nnz;
        pfor(c, (int64_t)bot.nrows + 1, [=] __device__(int64_t i) { o[i] = rp[i] + shift; });
    }
    if (top.nnz) {
        PORO_CUDA(cudaMemcpyAsync(out.col.p, top.col.p, (size_t)top.nnz * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
        PORO_CUDA(cudaMemcpyAsync(out.val.p, top.val.p, (size_t)top.nnz * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    }
    if (bot.nnz) {
        PORO_CUDA(cudaMemcpyAsync(out.col.p + top.nnz, bot.col.p, (size_t)bot.nnz * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
        PORO_CUDA(cudaMemcpyAsync(out.val.p + top.nnz, bot.val.p, (size_t)bot.nnz * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    }
    csr_choose_lanes(out);
}

// ---- localize: global column ids -> [owned | ghost] of the range [a, b) -------------------------------------------------
// `extra` (may be null): further global ids that must become ghosts (the ghost columns of P next to those of A_c).
// Returns the ascending ghost id list; M.col is rewritten in place and M.ncols becomes n_owned + n_ghost.  Columns inside
// a row are no longer sorted afterwards (owned columns come first in the local numbering); nothing downstream needs that.
void dist_localize(Ctx& c, Csr& M, int a, int b, const int* extra, int64_t n_extra, DBuf<int>& ghost_gid, int& n_ghost) {
    const int64_t cand = M.nnz + n_extra;
    PORO_REQUIRE(cand < 2147483647LL, "dist_localize: too many candidates");
    DBuf<int> keys((size_t)cand + 1), sorted((size_t)cand + 1), uniq((size_t)cand + 1), d_num(1);
    {
        // owned columns collapse onto the sentinel INT_MAX so that one sort + unique yields the ghosts, ascending
        const int* cc = M.col.p; int* k = keys.p; const int64_t nnz = M.nnz;
        pfor(c, cand, [=] __device__(int64_t t) {
            const int g = t < nnz ? cc[t] : extra[t - nnz];
            k[t] = (g >= a && g < b) ? 2147483647 : g;
        });
    }
    n_ghost = 0;
    if (cand) {
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, sorted.p, (int)cand, 0, 32, c.stream);
        DBuf<char> tmp(tb);
        cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, sorted.p, (int)cand, 0, 32, c.stream);
        size_t tb2 = 0;
        cub::DeviceSelect::Unique(nullptr, tb2, sorted.p, uniq.p, d_num.p, (int)cand, c.stream);
        DBuf<char> tmp2(tb2);
        cub::DeviceSelect::Unique(tmp2.p, tb2, sorted.p, uniq.p, d_num.p, (int)cand, c.stream);
        c.launches += 2;
        int num = 0, last = 0;
        PORO_CUDA(cudaMemcpyAsync(&num, d_num.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        PORO_CUDA(cudaStreamSynchronize(c.stream));
        if (num) PORO_CUDA(cudaMemcpy(&last, uniq.p + num - 1, sizeof(int), cudaMemcpyDeviceToHost));
        n_ghost = (num && last == 2147483647) ? num - 1 : num;
    }
    ghost_gid.alloc((size_t)n_ghost);
    if (n_ghost) PORO_CUDA(cudaMemcpyAsync(ghost_gid.p, uniq.p, (size_t)n_ghost * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
    {
        int* cc = M.col.p; const int* g = ghost_gid.p; const int ng = n_ghost, no = b - a;
        pfor(c, M.nnz, [=] __device__(int64_t t) {
            const int col = cc[t];
            if (col >= a && col < b) { cc[t] = col - a; return; }
            int lo = 0, hi = ng;                                   // lower bound in the ghost list
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (g[mid] < col) lo = mid + 1; else hi = mid; }
            cc[t] = no + lo;
        });
    }
    M.ncols = (b - a) + n_ghost;
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    csr_choose_lanes(M);
}

void dist_globalize(Ctx& c, Csr& M, const DistPlan& plan, int64_t ncols_global) {
    PORO_REQUIRE(ncols_global < 2147483647LL, "dist_globalize: more than 2^31 columns");
    int* cc = M.col.p; const int* g = plan.ghost_gid.p; const int no = plan.n_owned, off = (int)plan.offset;
    pfor(c, M.nnz, [=] __device__(int64_t t) { const int col = cc[t]; cc[t] = col < no ? off + col : g[col - no]; });
    M.ncols = (int)ncols_global;
}

// local ids of global columns against an ascending ghost list
static void relabel_with_ghosts(Ctx& c, Csr& M, int a, int b, const DBuf<int>& ghosts, int n_ghost) {
    int* cc = M.col.p; const int* g = ghosts.p; const int ng = n_ghost, no = b - a;
    pfor(c, M.nnz, [=] __device__(int64_t t) {
        const int col = cc[t];
        if (col >= a && col < b) { cc[t] = col - a; return; }
        int lo = 0, hi = ng;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (g[mid] < col) lo = mid + 1; else hi = mid; }
        cc[t] = no + lo;
    });
    M.ncols = (b - a) + n_ghost;
    csr_choose_lanes(M);
}

// ---- one level of the distributed set-up (oracle/distamg_rank.py: body of RankAmg.__init__) --------------------------------
void dist_amg_level(Ctx& c, const Csr& A, const DistPlan& plan, const Csr& T, const std::vector<int64_t>& coff,
                    const double* dinv, double omega, DistLevelOut& out) {
    const int R = c.nranks, me = c.rank;
    PORO_REQUIRE(coff[R] < 2147483647LL, "coarse level has more than 2^31 dofs");
    PORO_REQUIRE(A.ncols == plan.n_owned + plan.n_ghost, "level operator and its halo plan disagree");
    const int ca = (int)coff[me], cb = (int)coff[me + 1], NC = (int)coff[R];
    PORO_REQUIRE(T.ncols == cb - ca, "tentative prolongator and coarse offsets disagree");
    // T with global coarse columns
    Csr Tg;
    csr_copy(c, T, Tg);
    Tg.ncols = NC;
    { int* cc = Tg.col.p; pfor(c, Tg.nnz, [=] __device__(int64_t t) { cc[t] += ca; }); }
    // P = T - w D^-1 A [T ; T_ghost]
    Csr Pg;
    {
        Csr Tgh, Text, AT;
        dist_halo_rows(c, plan, Tg, Tgh);
        csr_vstack(c, Tg, Tgh, Text);
        csr_spgemm(c, A, Text, AT);
        csr_add_scaled(c, Tg, AT, -omega, dinv, Pg);
    }
    // A_c = (P_ext[:, owned coarse])^T [AP ; AP_ghost]
    Csr Acg;
    {
        Csr Pgh, Pext, AP, APgh, APext, Pmine;
        dist_halo_rows(c, plan, Pg, Pgh);
        csr_vstack(c, Pg, Pgh, Pext);
        csr_spgemm(c, A, Pext, AP);
        dist_halo_rows(c, plan, AP, APgh);
        csr_vstack(c, AP, APgh, APext);
        csr_select(c, Pext, 0, Pext.nrows, ca, cb, true, Pmine);         // columns of the owned coarse dofs (renumbered from 0) ...
        csr_transpose(c, Pmine, out.R);                                   // ... transposed: the restriction of this rank
        csr_spgemm(c, out.R, APext, Acg);
    }
    {   // dead coarse dofs (rank-deficient aggregates): unit diagonal, as in Amg::setup
        const int* rp = Acg.rowptr.p; const int* cc = Acg.col.p; double* v = Acg.val.p;
        pfor(c, Acg.nrows, [=] __device__(int64_t i) {
            for (int q = rp[i]; q < rp[i + 1]; ++q) if (cc[q] == ca + (int)i && v[q] == 0.0) v[q] = 1.0;
        });
    }
    // next level: ghost coarse columns are those of A_c and of P
    DBuf<int> ghosts;
    int n_ghost = 0;
    dist_localize(c, Acg, ca, cb, Pg.col.p, Pg.nnz, ghosts, n_ghost);
    relabel_with_ghosts(c, Pg, ca, cb, ghosts, n_ghost);                  // same local numbering for P (its ghosts are a subset)
    dist_plan_build(c, coff, std::move(ghosts), n_ghost, out.coarse_plan);
    out.P = std::move(Pg);
    out.Ac = std::move(Acg);
    PORO_CUDA(cudaStreamSynchronize(c.stream));
}

// ---- exact selfp Schur complement of the owned rows -------------------------------------------------------------------
void dist_selfp_schur(Ctx& c, DistPlan& plan0, DistPlan& plan1, const Csr& A00, const Csr& A01, const Csr& A10,
                      const Csr& A11, Csr& S, DistPlan& planS, const double* diag0_owned) {
    const int n0 = plan0.n_owned, g0 = plan0.n_ghost;
    PORO_REQUIRE(A10.ncols == n0 + g0 && A01.nrows == n0, "Schur blocks and the halo plan of split 0 disagree");
    const int64_t N1 = plan1.offsets.back();
    DBuf<double> d((size_t)(n0 + g0));
    if (diag0_owned) PORO_CUDA(cudaMemcpyAsync(d.p, diag0_owned, (size_t)n0 * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    else csr_diag(c, A00, d.p);
    dist_halo_vec(c, plan0, d.p, 1, d.p + n0);
    { double* p = d.p; pfor(c, n0 + g0, [=] __device__(int64_t i) { p[i] = p[i] != 0.0 ? 1.0 / p[i] : 1.0; }); }
    Csr A01g, A01gh, A01ext, A10s, prod, A11g, Sg;
    csr_copy(c, A01, A01g);
    dist_globalize(c, A01g, plan1, N1);
    dist_halo_rows(c, plan0, A01g, A01gh);
    csr_vstack(c, A01g, A01gh, A01ext);
    csr_copy(c, A10, A10s);
    csr_scale_cols(c, A10s, d.p);
    csr_spgemm(c, A10s, A01ext, prod);
    csr_copy(c, A11, A11g);
    dist_globalize(c, A11g, plan1, N1);
    csr_add_scaled(c, A11g, prod, -1.0, nullptr, Sg);
    DBuf<int> ghosts;
    int ng = 0;
    dist_localize(c, Sg, (int)plan1.offset, (int)plan1.offset + plan1.n_owned, nullptr, 0, ghosts, ng);
    dist_plan_build(c, plan1.offsets, std::move(ghosts), ng, planS);
    S = std::move(Sg);
    PORO_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace poro
