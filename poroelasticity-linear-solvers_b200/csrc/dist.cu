// dist.cu -- NCCL over NVLink for the two collectives the path has:
//   * sum-allreduce of the few scalars of each dot-product batch (VecMDot/VecNorm's
//     MPI_Allreduce in the reference's PETSc path), issued on the compute stream;
//   * neighbour halo exchange before the off-diagonal part of every SpMV
//     (MatMult_MPIAIJ's VecScatter), grouped ncclSend/ncclRecv.
// NCCL is resolved with dlopen at first use so single-GPU runs have no dependency on it and
// the library binds to whichever libnccl.so.2 the process already loaded (torch's).
#include "dist.cuh"
#include <dlfcn.h>

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
}

void dist_finalize(Ctx& c) {
    if (c.comm && c.nccl) c.nccl->CommDestroy((ncclComm_p)c.comm);
    c.comm = nullptr;
}

void dist_allreduce_sum(Ctx& c, double* d_vals, int k) {
    if (c.nranks <= 1) return;
    NCCL_OK(c.nccl, c.nccl->AllReduce(d_vals, d_vals, (size_t)k, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)c.comm, c.stream));
}

void dist_allreduce_max(Ctx& c, double* d_vals, int k) {
    if (c.nranks <= 1) return;
    NCCL_OK(c.nccl, c.nccl->AllReduce(d_vals, d_vals, (size_t)k, NCCL_FLOAT64, NCCL_MAX, (ncclComm_p)c.comm, c.stream));
}

void dist_halo_exchange(Ctx& c, HaloField& hf, const double* x_owned, double* halo) {
    if (c.nranks <= 1 || c.neigh.empty()) return;
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
