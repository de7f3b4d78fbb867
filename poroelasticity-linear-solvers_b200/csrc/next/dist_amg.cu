// dist_amg.cu -- distributed smoothed-aggregation level set-up (ROUND-2 WORK IN PROGRESS: compiles for sm_100a, NOT
// linked into libporo.so and never run on a GPU yet).
//
// CUDA + NCCL statement of oracle/distamg_rank.py (verified on CPU over gloo, tests/test_oracle_distamg_gloo.py); the
// names below are the names there.  Per level and rank:
//     aggregation + tentative prolongator T     rank-local, existing kernels (amg.cu)            -- no communication
//     P  = T - w D^-1 (A [T ; T_ghost])          csr_spgemm                                       -- halo_rows(T)
//     AP = A [P ; P_ghost]                       csr_spgemm                                       -- halo_rows(P)
//     Ac = (P_ext[:, owned coarse])^T [AP ; AP_ghost]   csr_transpose + csr_spgemm                -- halo_rows(AP)
//     R  = (P_ext[:, owned coarse])^T            by-product; restriction reads [owned | ghost] residuals
//     next level: localize(Ac), DistPlan from its ghost columns
// Matrices in flight carry GLOBAL column ids (int32: < 2^31 dofs per level) in `Csr::col` and are localised to
// [owned | ghost] only when they become level operators.
//
// Communication goes through five calls that dist.cu will provide on top of the NCCL handle it already owns
// (grouped ncclSend / ncclRecv of bytes, ncclAllGather of int64); they are declared here so that this file builds alone.
#include <cub/cub.cuh>

#include <algorithm>
#include <numeric>

#include "../common.cuh"

namespace poro {

// ---- to be provided by dist.cu ------------------------------------------------------------------------------------
void dist_group_begin(Ctx& c);
void dist_group_end(Ctx& c);
void dist_send_bytes(Ctx& c, const void* dev, size_t bytes, int peer);
void dist_recv_bytes(Ctx& c, void* dev, size_t bytes, int peer);
// all[r * count + i] = value i of rank r (host vectors; synchronises the stream)
void dist_allgather_i64(Ctx& c, const int64_t* mine, int count, std::vector<int64_t>& all);

// ---- small device helpers -------------------------------------------------------------------------------------------
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
    if (n == 0) return 0;
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

// ---- halo plan of one level (oracle/distamg_rank.py: Plan) -------------------------------------------------------
struct DistPlan {
    int64_t offset = 0;                   // global id of the first owned row
    int n_owned = 0, n_ghost = 0;
    std::vector<int64_t> offsets;         // (nranks + 1), rank-contiguous numbering of this level
    DBuf<int> ghost_gid;                  // ascending global ids of the ghost columns = neighbour-major, owner's order
    std::vector<int> neigh;               // ranks we exchange with, ascending
    std::vector<int64_t> send_ptr, recv_ptr;   // per neighbour, sizes neigh.size() + 1
    DBuf<int> send_idx;                   // owned local indices to send, neighbour-major
};

// The handshake: all-gather how many ids every rank reads from every other, then send the id lists to their owners.
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
    dist_group_begin(c);
    for (size_t k = 0; k < nn; ++k) {
        const int q = plan.neigh[k];
        const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
        if (nr) dist_send_bytes(c, plan.ghost_gid.p + start[q], (size_t)nr * sizeof(int), q);
        if (ns) dist_recv_bytes(c, plan.send_idx.p + plan.send_ptr[k], (size_t)ns * sizeof(int), q);
    }
    dist_group_end(c);
    {
        int* s = plan.send_idx.p;
        const int off = (int)plan.offset;
        pfor(c, plan.send_ptr[nn], [=] __device__(int64_t i) { s[i] -= off; });
    }
}

// ---- halo_vec: `width` doubles per node (vectors: 1; tentative-prolongator rows: bs * k) ---------------------------
// x_owned: n_owned x width row-major; ghost_out: n_ghost x width.
void dist_halo_vec(Ctx& c, const DistPlan& plan, const double* x_owned, int width, double* ghost_out, DBuf<double>& send_buf) {
    const size_t nn = plan.neigh.size();
    const int64_t nsend = plan.send_ptr[nn];
    if (send_buf.n < (size_t)nsend * width) send_buf.alloc((size_t)nsend * width);
    {
        const int* idx = plan.send_idx.p;
        double* sb = send_buf.p;
        pfor(c, nsend * width, [=] __device__(int64_t t) { sb[t] = x_owned[(int64_t)idx[t / width] * width + t % width]; });
    }
    dist_group_begin(c);
    for (size_t k = 0; k < nn; ++k) {
        const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
        if (ns) dist_send_bytes(c, send_buf.p + plan.send_ptr[k] * width, (size_t)ns * width * sizeof(double), plan.neigh[k]);
        if (nr) dist_recv_bytes(c, ghost_out + plan.recv_ptr[k] * width, (size_t)nr * width * sizeof(double), plan.neigh[k]);
    }
    dist_group_end(c);
}

// ---- halo_rows: the sparse rows of the boundary nodes (GLOBAL column ids) --------------------------------------------
// Three grouped exchanges: row lengths, column ids, values.  `ghost` gets n_ghost rows in ghost order.
void dist_halo_rows(Ctx& c, const DistPlan& plan, const Csr& M, Csr& ghost) {
    const size_t nn = plan.neigh.size();
    const int64_t nsend = plan.send_ptr[nn];
    // lengths of the rows we send, and their packed offsets
    DBuf<int> slen((size_t)nsend), sptr((size_t)nsend + 1);
    {
        const int* idx = plan.send_idx.p; const int* rp = M.rowptr.p; int* L = slen.p;
        pfor(c, nsend, [=] __device__(int64_t i) { L[i] = rp[idx[i] + 1] - rp[idx[i]]; });
    }
    const int64_t send_nnz = scan_counts(c, slen.p, sptr.p, nsend);
    DBuf<int> rlen((size_t)plan.n_ghost);
    dist_group_begin(c);
    for (size_t k = 0; k < nn; ++k) {
        const int64_t ns = plan.send_ptr[k + 1] - plan.send_ptr[k], nr = plan.recv_ptr[k + 1] - plan.recv_ptr[k];
        if (ns) dist_send_bytes(c, slen.p + plan.send_ptr[k], (size_t)ns * sizeof(int), plan.neigh[k]);
        if (nr) dist_recv_bytes(c, rlen.p + plan.recv_ptr[k], (size_t)nr * sizeof(int), plan.neigh[k]);
    }
    dist_group_end(c);
    ghost.nrows = plan.n_ghost;
    ghost.ncols = M.ncols;
    ghost.rowptr.alloc((size_t)plan.n_ghost + 1);
    ghost.nnz = scan_counts(c, rlen.p, ghost.rowptr.p, plan.n_ghost);
    ghost.col.alloc((size_t)ghost.nnz);
    ghost.val.alloc((size_t)ghost.nnz);
    // packed boundaries per neighbour (host needs them for the message sizes)
    std::vector<int> sp_h((size_t)nn + 1, 0), rp_h((size_t)nn + 1, 0);
    for (size_t k = 0; k <= nn; ++k) {
        PORO_CUDA(cudaMemcpyAsync(&sp_h[k], sptr.p + plan.send_ptr[k], sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        PORO_CUDA(cudaMemcpyAsync(&rp_h[k], ghost.rowptr.p + plan.recv_ptr[k], sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    }
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    // pack (one warp per row would be the tuned version; rows are short: <= a few hundred entries)
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

// ---- [top ; bottom] with equal column spaces -----------------------------------------------------------------------------
void csr_vstack(Ctx& c, const Csr& top, const Csr& bot, Csr& out) {
    PORO_REQUIRE(top.ncols == bot.ncols, "csr_vstack: column spaces differ");
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
}

// ---- localize: global column ids -> [owned | ghost] of the range [a, b) -------------------------------------------------
// `extra` (may be null): further global ids that must become ghosts (the ghost columns of P next to those of A_c).
// Returns the ascending ghost id list; M.col is rewritten in place and M.ncols becomes n_owned + n_ghost.
void dist_localize(Ctx& c, Csr& M, int a, int b, const int* extra, int64_t n_extra, DBuf<int>& ghost_gid, int& n_ghost) {
    const int64_t cand = M.nnz + n_extra;
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
    // rows of M were sorted by GLOBAL column; the local numbering puts owned columns first, so re-sort inside rows
    // when the caller needs sorted columns (the SpMV kernels do not).
}

// ---- one level of the distributed set-up (oracle/distamg_rank.py: body of RankAmg.__init__) --------------------------------
struct DistLevelOut {
    Csr P;            // owned fine rows x [owned coarse | ghost coarse]
    Csr R;            // owned coarse rows x [owned fine | ghost fine]
    Csr Ac;           // owned coarse rows x [owned coarse | ghost coarse]
    DistPlan coarse_plan;
};

// A: owned rows x [owned | ghost] (level operator, plan = its halo plan); T: owned rows x local aggregates*k (rank-local
// tentative prolongator); dinv, omega as in amg.cu.
void dist_amg_level(Ctx& c, const Csr& A, const DistPlan& plan, const Csr& T, const double* dinv, double omega, DistLevelOut& out) {
    const int R = c.nranks, me = c.rank;
    // coarse offsets: one all-gather of the local coarse sizes
    std::vector<int64_t> sizes;
    const int64_t mine = T.ncols;
    dist_allgather_i64(c, &mine, 1, sizes);
    std::vector<int64_t> coff((size_t)R + 1, 0);
    for (int r = 0; r < R; ++r) coff[r + 1] = coff[r] + sizes[r];
    PORO_REQUIRE(coff[R] < 2147483647LL, "coarse level has more than 2^31 dofs");
    const int ca = (int)coff[me], cb = (int)coff[me + 1], NC = (int)coff[R];
    // T with global coarse columns
    Csr Tg;
    csr_copy(c, T, Tg);
    Tg.ncols = NC;
    { int* cc = Tg.col.p; pfor(c, Tg.nnz, [=] __device__(int64_t t) { cc[t] += ca; }); }
    // P = T - w D^-1 A [T ; T_ghost]
    Csr Tgh, Text, AT, Pg;
    dist_halo_rows(c, plan, Tg, Tgh);
    csr_vstack(c, Tg, Tgh, Text);
    csr_spgemm(c, A, Text, AT);
    csr_add_scaled(c, Tg, AT, -omega, dinv, Pg);
    // A_c = (P_ext[:, owned coarse])^T [AP ; AP_ghost]
    Csr Pgh, Pext, AP, APgh, APext, Pmine, Acg;
    dist_halo_rows(c, plan, Pg, Pgh);
    csr_vstack(c, Pg, Pgh, Pext);
    csr_spgemm(c, A, Pext, AP);
    dist_halo_rows(c, plan, AP, APgh);
    csr_vstack(c, AP, APgh, APext);
    csr_select(c, Pext, 0, Pext.nrows, ca, cb, true, Pmine);              // columns of the owned coarse dofs (renumbered from 0) ...
    csr_transpose(c, Pmine, out.R);                                        // ... transposed: the restriction of this rank
    csr_spgemm(c, out.R, APext, Acg);
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
    {   // P gets the same local numbering (its ghost columns are a subset by construction)
        int* cc = Pg.col.p; const int* g = ghosts.p; const int ng = n_ghost, no = cb - ca;
        pfor(c, Pg.nnz, [=] __device__(int64_t t) {
            const int col = cc[t];
            if (col >= ca && col < cb) { cc[t] = col - ca; return; }
            int lo = 0, hi = ng;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (g[mid] < col) lo = mid + 1; else hi = mid; }
            cc[t] = no + lo;
        });
        Pg.ncols = (cb - ca) + n_ghost;
    }
    dist_plan_build(c, coff, std::move(ghosts), n_ghost, out.coarse_plan);
    out.P = std::move(Pg);
    out.Ac = std::move(Acg);
}

}  // namespace poro
