// dist.cu -- NCCL over NVLink for the two collectives the path has:
//   * sum-allreduce of the few scalars of each dot-product batch (VecMDot/VecNorm's
//     MPI_Allreduce in the reference's PETSc path), issued on the compute stream;
//   * neighbour halo exchange before the off-diagonal part of every SpMV
//     (MatMult_MPIAIJ's VecScatter), grouped ncclSend/ncclRecv.
// NCCL is resolved with dlopen at first use so single-GPU runs have no dependency on it and
// the library binds to whichever libnccl.so.2 the process already loaded (torch's).
#include "dist.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <memory>

namespace poro {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
enum { NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t*);
    int (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int);
    int (*CommDestroy)(ncclComm_p);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, ncclComm_p, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, ncclComm_p, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
};

static NcclApi* load_nccl() {
    static NcclApi api;
    if (api.lib) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) throw Error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SYM(field, name)                                                    \
    *(void**)(&api.field) = dlsym(api.lib, name);                           \
    if (!api.field) throw Error(std::string("NCCL symbol missing: ") + name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return &api;
}

#define NCCL_OK(api, expr)                                                                       \
    do {                                                                                         \
        int r__ = (expr);                                                                        \
        if (r__ != 0) throw Error(std::string("NCCL error: ") + (api)->GetErrorString(r__));     \
    } while (0)

static void p2p_init(Ctx& c);
static void p2p_finalize(Ctx& c);

void dist_get_unique_id(unsigned char* id128) {
    NcclApi* api = load_nccl();
    ncclUniqueId_t id;
    NCCL_OK(api, api->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
}

void dist_init(Ctx& c, int rank, int nranks, const unsigned char* id128) {
    c.rank = rank;
    c.nranks = nranks;
    if (nranks <= 1) return;
    NcclApi* api = load_nccl();
    ncclUniqueId_t id;
    memcpy(id.internal, id128, 128);
    PORO_CUDA(cudaSetDevice(c.device));
    ncclComm_p comm = nullptr;
    NCCL_OK(api, api->CommInitRank(&comm, nranks, id, rank));
    c.comm = comm;
    c.nccl = api;
    p2p_init(c);
}

void dist_finalize(Ctx& c) {
    p2p_finalize(c);
    if (c.comm && c.nccl) c.nccl->CommDestroy((ncclComm_p)c.comm);
    c.comm = nullptr;
}

static bool p2p_allreduce(Ctx& c, double* d_vals, int k);

void dist_allreduce_sum(Ctx& c, double* d_vals, int k) {
    if (c.nranks <= 1) return;
    if (p2p_allreduce(c, d_vals, k)) return;
    NCCL_OK(c.nccl, c.nccl->AllReduce(d_vals, d_vals, (size_t)k, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)c.comm, c.stream));
}

void dist_allreduce_max(Ctx& c, double* d_vals, int k) {
    if (c.nranks <= 1) return;
    NCCL_OK(c.nccl, c.nccl->AllReduce(d_vals, d_vals, (size_t)k, NCCL_FLOAT64, NCCL_MAX, (ncclComm_p)c.comm, c.stream));
}

void dist_halo_exchange(Ctx& c, HaloField& hf, const double* x_owned, double* halo) {
    if (c.nranks <= 1 || c.neigh.empty()) return;
    if (hf.p2p.ready) { p2p_exchange(c, hf.p2p, hf.send_idx.p, x_owned, halo); return; }
    int64_t nsend = hf.send_ptr.empty() ? 0 : hf.send_ptr.back();
    if (nsend) vec_gather(c, hf.send_buf.p, x_owned, hf.send_idx.p, nsend);
    NcclApi* api = c.nccl;
    NCCL_OK(api, api->GroupStart());
    for (size_t k = 0; k < c.neigh.size(); ++k) {
        int64_t ns = hf.send_ptr[k + 1] - hf.send_ptr[k];
        int64_t nr = hf.recv_ptr[k + 1] - hf.recv_ptr[k];
        if (ns) NCCL_OK(api, api->Send(hf.send_buf.p + hf.send_ptr[k], (size_t)ns, NCCL_FLOAT64, c.neigh[k], (ncclComm_p)c.comm, c.stream));
        if (nr) NCCL_OK(api, api->Recv(halo + hf.recv_ptr[k], (size_t)nr, NCCL_FLOAT64, c.neigh[k], (ncclComm_p)c.comm, c.stream));
    }
    NCCL_OK(api, api->GroupEnd());
}


// =============================================================================================
// NVLink peer-store halo exchange (no NCCL on the data path)
// =============================================================================================
struct P2PState {
    unsigned char* arena = nullptr;            // this rank's receive arena (cudaMalloc, exported through CUDA IPC)
    size_t bytes = 0, used = 0;
    std::vector<unsigned char*> peer;          // mapped arenas of the other ranks (nullptr for self)
    static constexpr size_t kFlagBytes = 1 << 16;   // head of the arena: 16 Ki sequence flags
    int flags_used = 16;                            // flags 0..15: small all-reduce, indexed by source rank
    // small all-reduce over peer stores: one region per source rank (2 parities x kArMax doubles) right after the flags
    static constexpr int kArMax = 65536;            // doubles per source rank and parity (all-reduce: <= 8192 used; all-gather: a slice)
    static constexpr size_t kArRegion = (size_t)2 * kArMax * sizeof(double);
    uint32_t* ar_state = nullptr;                   // device: [seq]
    unsigned* ar_ticket = nullptr;                  // device: two CTA tickets of the multi-CTA all-gather
};

struct P2PArArgs {
    double* peer_region[8];      // where this rank writes in every other rank's arena (nullptr for self)
    uint32_t* peer_flag[8];
    const double* my_region[8];  // region of source rank r in this rank's arena
    const uint32_t* my_flag[8];
    int nranks, me;
};

static void p2p_finalize(Ctx& c) {
    if (!c.p2p) return;
    for (auto* q : c.p2p->peer) if (q) cudaIpcCloseMemHandle(q);
    if (c.p2p->arena) cudaFree(c.p2p->arena);
    delete c.p2p;
    c.p2p = nullptr;
}

// Every rank exports its arena and maps all others'.  Collective; falls back (c.p2p stays null) when any rank fails.
static void p2p_init(Ctx& c) {
    if (c.nranks <= 1 || c.nranks > 9 || !c.opt_i("-poro_p2p", 1)) return;      // P2PArgs holds up to 8 neighbours
    auto st = std::make_unique<P2PState>();
    st->bytes = (size_t)c.opt_i("-poro_p2p_arena_mb", 512) << 20;
    int64_t words[9] = {0};                       // [ok, 64-byte IPC handle]
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    bool ok = cudaMalloc(&st->arena, st->bytes) == cudaSuccess && cudaMemset(st->arena, 0, st->bytes) == cudaSuccess &&
              cudaIpcGetMemHandle(&h, st->arena) == cudaSuccess;
    if (!ok) cudaGetLastError();
    words[0] = ok ? 1 : 0;
    if (ok) memcpy(&words[1], &h, 64);
    std::vector<int64_t> all;
    dist_allgather_i64(c, words, 9, all);
    for (int r = 0; r < c.nranks; ++r) ok = ok && all[(size_t)r * 9] == 1;
    st->peer.assign((size_t)c.nranks, nullptr);
    int64_t mine = 1;
    if (ok) {
        for (int r = 0; r < c.nranks && mine; ++r) {
            if (r == c.rank) continue;
            cudaIpcMemHandle_t hr;
            memcpy(&hr, &all[(size_t)r * 9 + 1], 64);
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine = 0; }
            st->peer[r] = (unsigned char*)q;
        }
    } else mine = 0;
    std::vector<int64_t> oks;
    dist_allgather_i64(c, &mine, 1, oks);
    for (int64_t v : oks) ok = ok && v == 1;
    if (!ok) {
        for (auto* q : st->peer) if (q) cudaIpcCloseMemHandle(q);
        if (st->arena) cudaFree(st->arena);
        if (c.has_opt("-poro_verbose") && c.rank == 0) fprintf(stderr, "  [dist] peer-to-peer halo path unavailable: NCCL send/recv is used\n");
        return;
    }
    st->used = P2PState::kFlagBytes + (size_t)c.nranks * P2PState::kArRegion;
    if (c.nranks <= 8 && c.opt_i("-poro_p2p_allreduce", 1)) {
        PORO_CUDA(cudaMalloc(&st->ar_state, sizeof(uint32_t)));
        PORO_CUDA(cudaMemset(st->ar_state, 0, sizeof(uint32_t)));
        PORO_CUDA(cudaMalloc(&st->ar_ticket, 2 * sizeof(unsigned)));
        PORO_CUDA(cudaMemset(st->ar_ticket, 0, 2 * sizeof(unsigned)));
    }
    c.p2p = st.release();
    c.p2p_fused = c.opt_i("-poro_p2p_fused", 1) != 0;
    if (c.has_opt("-poro_verbose") && c.rank == 0) fprintf(stderr, "  [dist] peer-to-peer halo path: %zu MB arena per rank, %d peers mapped\n", c.p2p->bytes >> 20, c.nranks - 1);
}

void p2p_slots_setup(Ctx& c, const std::vector<int>& neigh, const std::vector<int64_t>& send_ptr,
                     const std::vector<int64_t>& recv_ptr, P2PSlots& slots) {
    slots.ready = false;
    if (!c.p2p || neigh.empty() || neigh.size() > 8) return;
    P2PState& st = *c.p2p;
    const size_t nn = neigh.size();
    // my receive regions: two buffers (sequence parity) per neighbour, 128-byte aligned, plus one flag each
    std::vector<int64_t> mine(2 * nn), theirs(2 * nn, 0);
    for (size_t k = 0; k < nn; ++k) {
        const size_t nr = (size_t)(recv_ptr[k + 1] - recv_ptr[k]);
        const size_t need = ((nr * 8 + 127) & ~(size_t)127) * 2;
        if (st.used + need > st.bytes || st.flags_used + 1 > (int)(P2PState::kFlagBytes / 4))
            throw Error("peer-to-peer halo arena exhausted: raise -poro_p2p_arena_mb (or run with -poro_p2p 0)");
        mine[2 * k] = (int64_t)st.used;
        mine[2 * k + 1] = st.flags_used;
        st.used += need;
        st.flags_used += 1;
    }
    // tell every neighbour where it writes: grouped exchange of (offset, flag index)
    DBuf<int64_t> d_mine(2 * nn), d_theirs(2 * nn);
    PORO_CUDA(cudaMemcpyAsync(d_mine.p, mine.data(), 2 * nn * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
    dist_group_begin(c);
    for (size_t k = 0; k < nn; ++k) {
        dist_send_bytes(c, d_mine.p + 2 * k, 16, neigh[k]);
        dist_recv_bytes(c, d_theirs.p + 2 * k, 16, neigh[k]);
    }
    dist_group_end(c);
    PORO_CUDA(cudaMemcpyAsync(theirs.data(), d_theirs.p, 2 * nn * sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    PORO_CUDA(cudaStreamSynchronize(c.stream));
    slots.args.nn = (int)nn;
    for (size_t k = 0; k < nn; ++k) {
        P2PNeighDev& d = slots.args.nb[k];
        const size_t nr = (size_t)(recv_ptr[k + 1] - recv_ptr[k]), ns = (size_t)(send_ptr[k + 1] - send_ptr[k]);
        const size_t my_half = (nr * 8 + 127) & ~(size_t)127, peer_half = (ns * 8 + 127) & ~(size_t)127;
        unsigned char* pa = st.peer[neigh[k]];
        PORO_REQUIRE(pa != nullptr, "peer arena not mapped");
        d.my_data[0] = (const double*)(st.arena + mine[2 * k]);
        d.my_data[1] = (const double*)(st.arena + mine[2 * k] + my_half);
        d.my_flag = (const uint32_t*)st.arena + mine[2 * k + 1];
        d.peer_data[0] = (double*)(pa + theirs[2 * k]);
        d.peer_data[1] = (double*)(pa + theirs[2 * k] + peer_half);
        d.peer_flag = (uint32_t*)pa + theirs[2 * k + 1];
        d.send_begin = (int)send_ptr[k]; d.send_end = (int)send_ptr[k + 1];
        d.recv_begin = (int)recv_ptr[k]; d.recv_end = (int)recv_ptr[k + 1];
    }
    slots.nsend = (int)send_ptr[nn];
    slots.nrecv = (int)recv_ptr[nn];
    PORO_CUDA(cudaMalloc(&slots.d_state, 4 * sizeof(uint32_t)));     // lives as long as the process (a few bytes per plan)
    PORO_CUDA(cudaMemsetAsync(slots.d_state, 0, 4 * sizeof(uint32_t), c.stream));
    slots.ready = true;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// pack + push: boundary values go straight into the neighbours' receive regions; the last CTA to finish releases the flags
__global__ void __launch_bounds__(256) k_p2p_push(P2PArgs a, const double* __restrict__ x, const int* __restrict__ send_idx,
                                                  int nsend, uint32_t* __restrict__ state) {
    const uint32_t seq = state[0] + 1u;
    const int par = (int)(seq & 1u);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nsend; i += gridDim.x * 256) {
        int k = 0;
        while (k + 1 < a.nn && i >= a.nb[k].send_end) ++k;
        a.nb[k].peer_data[par][i - a.nb[k].send_begin] = x[send_idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&state[2], 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int k = 0; k < a.nn; ++k)
                if (a.nb[k].send_end > a.nb[k].send_begin) st_release_sys(a.nb[k].peer_flag, seq);
            state[2] = 0u;
            state[0] = seq;
        }
    }
}

// wait for the neighbours' flags, then move the received values behind the owned entries
__global__ void __launch_bounds__(256) k_p2p_wait_copy(P2PArgs a, double* __restrict__ ghost, int nrecv, uint32_t* __restrict__ state) {
    const uint32_t seq = state[1] + 1u;
    const int par = (int)(seq & 1u);
    if (threadIdx.x < a.nn) {
        const P2PNeighDev& d = a.nb[threadIdx.x];
        if (d.recv_end > d.recv_begin)
            while ((int32_t)(ld_acquire_sys(d.my_flag) - seq) < 0) { }
    }
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nrecv; i += gridDim.x * 256) {
        int k = 0;
        while (k + 1 < a.nn && i >= a.nb[k].recv_end) ++k;
        ghost[i] = __ldcv(a.nb[k].my_data[par] + (i - a.nb[k].recv_begin));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&state[3], 1u);
        if (t == gridDim.x - 1) { state[3] = 0u; state[1] = seq; }
    }
}

// push, wait and copy in ONE launch: every CTA first stores its share of the boundary values into the neighbours' regions,
// the last one to finish releases the flags; then every CTA waits for the incoming flags and copies its share of the
// received values.  The CTAs only wait for REMOTE events, and the grid (<= 2 CTAs per SM) is always co-resident, so the
// pushes a neighbour waits for can never be stuck behind spinning CTAs.
__global__ void __launch_bounds__(256) k_p2p_exchange(P2PArgs a, const double* __restrict__ x, const int* __restrict__ send_idx,
                                                      int nsend, double* __restrict__ ghost, int nrecv, uint32_t* __restrict__ state) {
    const uint32_t seq = state[0] + 1u;
    const int par = (int)(seq & 1u);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nsend; i += gridDim.x * 256) {
        int k = 0;
        while (k + 1 < a.nn && i >= a.nb[k].send_end) ++k;
        a.nb[k].peer_data[par][i - a.nb[k].send_begin] = x[send_idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&state[2], 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int k = 0; k < a.nn; ++k)
                if (a.nb[k].send_end > a.nb[k].send_begin) st_release_sys(a.nb[k].peer_flag, seq);
        }
    }
    if (threadIdx.x < a.nn) {
        const P2PNeighDev& d = a.nb[threadIdx.x];
        if (d.recv_end > d.recv_begin)
            while ((int32_t)(ld_acquire_sys(d.my_flag) - seq) < 0) { }
    }
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nrecv; i += gridDim.x * 256) {
        int k = 0;
        while (k + 1 < a.nn && i >= a.nb[k].recv_end) ++k;
        ghost[i] = __ldcv(a.nb[k].my_data[par] + (i - a.nb[k].recv_begin));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&state[3], 1u);
        if (t == gridDim.x - 1) { state[2] = 0u; state[3] = 0u; state[0] = seq; state[1] = seq; }
    }
}

void p2p_exchange(Ctx& c, P2PSlots& s, const int* send_idx, const double* x_owned, double* ghost_out) {
    if (c.p2p_fused) {
        const int n = std::max(s.nsend, s.nrecv);
        const int g = std::max(1, std::min((n + 255) / 256, c.sm_count * 2));
        k_p2p_exchange<<<g, 256, 0, c.stream>>>(s.args, x_owned, send_idx, s.nsend, ghost_out, s.nrecv, s.d_state);
        PORO_LAUNCH_CHECK(c);
        return;
    }
    const int gs = std::max(1, std::min((s.nsend + 255) / 256, c.sm_count * 2));
    const int gr = std::max(1, std::min((s.nrecv + 255) / 256, c.sm_count * 2));
    k_p2p_push<<<gs, 256, 0, c.stream>>>(s.args, x_owned, send_idx, s.nsend, s.d_state);
    PORO_LAUNCH_CHECK(c);
    k_p2p_wait_copy<<<gr, 256, 0, c.stream>>>(s.args, ghost_out, s.nrecv, s.d_state);
    PORO_LAUNCH_CHECK(c);
}

// sum of k <= 8192 doubles over all ranks by peer stores: every rank writes its values into its region of every other
// rank's arena and releases a flag there; after all flags have arrived the values are added in RANK ORDER, so every rank
// gets bit-identical sums.  One single-CTA kernel, no NCCL: the per-iteration reductions of GMRES / CG and the gathered
// coarse right-hand sides are a few hundred bytes and purely latency-bound.
__global__ void __launch_bounds__(256) k_p2p_allreduce(P2PArArgs a, double* __restrict__ vals, int k, uint32_t* __restrict__ state) {
    const uint32_t seq = *state + 1u;
    const size_t par_off = (size_t)(seq & 1u) * P2PState::kArMax;
    for (int r = 0; r < a.nranks; ++r) {
        if (r == a.me) continue;
        double* dst = a.peer_region[r] + par_off;
        for (int i = threadIdx.x; i < k; i += 256) dst[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < a.nranks && threadIdx.x != a.me) {
        st_release_sys(a.peer_flag[threadIdx.x], seq);
        while ((int32_t)(ld_acquire_sys(a.my_flag[threadIdx.x]) - seq) < 0) { }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += 256) {
        double s = 0.0;
        for (int r = 0; r < a.nranks; ++r) s += r == a.me ? vals[i] : __ldcv(a.my_region[r] + par_off + i);
        vals[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *state = seq;
}

static bool p2p_allreduce(Ctx& c, double* d_vals, int k) {
    if (!c.p2p || !c.p2p->ar_state || k > 8192) return false;
    P2PState& st = *c.p2p;
    P2PArArgs a{};
    a.nranks = c.nranks;
    a.me = c.rank;
    for (int r = 0; r < c.nranks; ++r) {
        // region of source rank `src` in an arena: flags first, then the regions in source-rank order (same layout everywhere)
        a.my_region[r] = (const double*)(st.arena + P2PState::kFlagBytes + (size_t)r * P2PState::kArRegion);
        a.my_flag[r] = (const uint32_t*)st.arena + r;
        a.peer_region[r] = r == c.rank ? nullptr : (double*)(st.peer[r] + P2PState::kFlagBytes + (size_t)c.rank * P2PState::kArRegion);
        a.peer_flag[r] = r == c.rank ? nullptr : (uint32_t*)st.peer[r] + c.rank;
    }
    k_p2p_allreduce<<<1, 256, 0, c.stream>>>(a, d_vals, k, st.ar_state);
    PORO_LAUNCH_CHECK(c);
    return true;
}

// all-gather of one slice per rank (slice r = full[off[r] .. off[r+1])) by peer stores, same regions / flags / sequence as the
// small all-reduce: the gathered right-hand side of a replicated coarse sub-hierarchy (amg.cu) in one launch
struct P2PGatherOffs { int off[9]; };
__global__ void __launch_bounds__(256) k_p2p_allgather(P2PArArgs a, P2PGatherOffs o, const double* __restrict__ mine,
                                                       double* __restrict__ full, uint32_t* __restrict__ state,
                                                       unsigned* __restrict__ ticket) {
    const uint32_t seq = *state + 1u;
    const size_t par_off = (size_t)(seq & 1u) * P2PState::kArMax;
    const int n_mine = o.off[a.me + 1] - o.off[a.me];
    const int stride = gridDim.x * 256, t0 = blockIdx.x * 256 + threadIdx.x;
    for (int r = 0; r < a.nranks; ++r) {
        if (r == a.me) continue;
        double* dst = a.peer_region[r] + par_off;
        for (int i = t0; i < n_mine; i += stride) dst[i] = mine[i];
    }
    for (int i = t0; i < n_mine; i += stride) full[o.off[a.me] + i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ticket[0], 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int r = 0; r < a.nranks; ++r)
                if (r != a.me) st_release_sys(a.peer_flag[r], seq);
        }
    }
    if (threadIdx.x < a.nranks && threadIdx.x != a.me)
        while ((int32_t)(ld_acquire_sys(a.my_flag[threadIdx.x]) - seq) < 0) { }
    __syncthreads();
    for (int r = 0; r < a.nranks; ++r) {
        if (r == a.me) continue;
        const double* src = a.my_region[r] + par_off;
        const int n_r = o.off[r + 1] - o.off[r];
        for (int i = t0; i < n_r; i += stride) full[o.off[r] + i] = __ldcv(src + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ticket[1], 1u);
        if (t == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; *state = seq; }
    }
}

static void p2p_ar_args(Ctx& c, P2PArArgs& a) {
    P2PState& st = *c.p2p;
    a.nranks = c.nranks;
    a.me = c.rank;
    for (int r = 0; r < c.nranks; ++r) {
        a.my_region[r] = (const double*)(st.arena + P2PState::kFlagBytes + (size_t)r * P2PState::kArRegion);
        a.my_flag[r] = (const uint32_t*)st.arena + r;
        a.peer_region[r] = r == c.rank ? nullptr : (double*)(st.peer[r] + P2PState::kFlagBytes + (size_t)c.rank * P2PState::kArRegion);
        a.peer_flag[r] = r == c.rank ? nullptr : (uint32_t*)st.peer[r] + c.rank;
    }
}

// full[off[r] .. off[r+1]) = slice of rank r, on every rank; offsets are element offsets (nranks + 1 entries, host)
void dist_allgather_slices(Ctx& c, const double* mine, const std::vector<int64_t>& offs, double* full) {
    const int R = c.nranks, me = c.rank;
    const int64_t n_full = offs[R];
    int64_t max_slice = 0;
    for (int r = 0; r < R; ++r) max_slice = std::max(max_slice, offs[r + 1] - offs[r]);
    if (R <= 1) { vec_copy(c, full, mine, n_full); return; }
    if (c.p2p && c.p2p->ar_state && max_slice <= P2PState::kArMax && n_full < 2147483647LL) {
        P2PArArgs a{};
        p2p_ar_args(c, a);
        P2PGatherOffs o{};
        for (int r = 0; r <= R; ++r) o.off[r] = (int)offs[r];
        const int g = (int)std::max<int64_t>(1, std::min<int64_t>((max_slice + 255) / 256, 32));
        k_p2p_allgather<<<g, 256, 0, c.stream>>>(a, o, mine, full, c.p2p->ar_state, c.p2p->ar_ticket);
        PORO_LAUNCH_CHECK(c);
        return;
    }
    // fallback: sum of zero-padded vectors
    PORO_CUDA(cudaMemsetAsync(full, 0, (size_t)n_full * sizeof(double), c.stream));
    vec_copy(c, full + offs[me], mine, offs[me + 1] - offs[me]);
    NCCL_OK(c.nccl, c.nccl->AllReduce(full, full, (size_t)n_full, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)c.comm, c.stream));
}

// equal-sized raw all-gather (set-up: gathering a coarse level operator on every rank)
void dist_allgather_bytes(Ctx& c, const void* send_dev, size_t bytes_per_rank, void* recv_dev) {
    if (c.nranks <= 1) { PORO_CUDA(cudaMemcpyAsync(recv_dev, send_dev, bytes_per_rank, cudaMemcpyDeviceToDevice, c.stream)); return; }
    NCCL_OK(c.nccl, c.nccl->AllGather(send_dev, recv_dev, bytes_per_rank, NCCL_INT8, (ncclComm_p)c.comm, c.stream));
}

// ---- set-up primitives of the distributed hierarchy (distamg.cu): grouped byte send/recv, all-gather of int64 --------
void dist_group_begin(Ctx& c) {
    if (c.nranks > 1) NCCL_OK(c.nccl, c.nccl->GroupStart());
}
void dist_group_end(Ctx& c) {
    if (c.nranks > 1) NCCL_OK(c.nccl, c.nccl->GroupEnd());
}
void dist_send_bytes(Ctx& c, const void* dev, size_t bytes, int peer) {
    NCCL_OK(c.nccl, c.nccl->Send(dev, bytes, NCCL_INT8, peer, (ncclComm_p)c.comm, c.stream));
}
void dist_recv_bytes(Ctx& c, void* dev, size_t bytes, int peer) {
    NCCL_OK(c.nccl, c.nccl->Recv(dev, bytes, NCCL_INT8, peer, (ncclComm_p)c.comm, c.stream));
}
// all[r * count + i] = value i of rank r; synchronises the stream (set-up only)
void dist_allgather_i64(Ctx& c, const int64_t* mine, int count, std::vector<int64_t>& all) {
    all.assign((size_t)c.nranks * count, 0);
    if (c.nranks <= 1) {
        for (int i = 0; i < count; ++i) all[i] = mine[i];
        return;
    }
    DBuf<int64_t> d_in((size_t)count), d_out((size_t)c.nranks * count);
    PORO_CUDA(cudaMemcpyAsync(d_in.p, mine, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
    NCCL_OK(c.nccl, c.nccl->AllGather(d_in.p, d_out.p, (size_t)count, NCCL_INT64, (ncclComm_p)c.comm, c.stream));
    PORO_CUDA(cudaMemcpyAsync(all.data(), d_out.p, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    PORO_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace poro
