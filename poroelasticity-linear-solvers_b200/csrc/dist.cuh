// dist.cuh -- NCCL (dlopen'ed) all-reduce and halo exchange for row-partitioned runs.
#pragma once
#include "common.cuh"

namespace poro {
void dist_get_unique_id(unsigned char* id128);
void dist_init(Ctx& c, int rank, int nranks, const unsigned char* id128);
void dist_finalize(Ctx& c);
void dist_allreduce_sum(Ctx& c, double* d_vals, int k);
void dist_allreduce_max(Ctx& c, double* d_vals, int k);
// exchange one field's halo: gathers x[send_idx] into hf.send_buf, sends/receives, halo (n_halo doubles) filled
void dist_halo_exchange(Ctx& c, HaloField& hf, const double* x_owned, double* halo);
// NVLink peer-store halo path: slots for one plan (collective over the plan's neighbours), and the exchange itself.
// p2p_slots_setup leaves slots.ready == false when the peer path is unavailable (single rank, IPC refused, -poro_p2p 0).
void p2p_slots_setup(Ctx& c, const std::vector<int>& neigh, const std::vector<int64_t>& send_ptr,
                     const std::vector<int64_t>& recv_ptr, P2PSlots& slots);
void p2p_exchange(Ctx& c, P2PSlots& slots, const int* send_idx, const double* x_owned, double* ghost_out);
// set-up primitives (grouped point-to-point of raw bytes, all-gather of a few int64 per rank)
void dist_group_begin(Ctx& c);
void dist_group_end(Ctx& c);
void dist_send_bytes(Ctx& c, const void* dev, size_t bytes, int peer);
void dist_recv_bytes(Ctx& c, void* dev, size_t bytes, int peer);
void dist_allgather_i64(Ctx& c, const int64_t* mine, int count, std::vector<int64_t>& all);
void dist_allgather_bytes(Ctx& c, const void* send_dev, size_t bytes_per_rank, void* recv_dev);
// full[offs[r] .. offs[r+1]) = the slice of rank r, on every rank (peer stores when available, else a padded all-reduce)
void dist_allgather_slices(Ctx& c, const double* mine, const std::vector<int64_t>& offs, double* full);
}  // namespace poro
