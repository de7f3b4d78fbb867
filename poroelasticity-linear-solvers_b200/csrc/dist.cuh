// dist.cuh -- NCCL (dlopen'ed) all-reduce and halo exchange for row-partitioned runs.
#pragma once
#include "common.cuh"

namespace poro {
void dist_get_unique_id(unsigned char* id128);
void dist_init(Ctx& c, int rank, int nranks, const unsigned char* id128);
void dist_finalize(Ctx& c);
void dist_allreduce_sum(Ctx& c, double* d_vals, int k);
// exchange one field's halo: gathers x[send_idx] into hf.send_buf, sends/receives, halo (n_halo doubles) filled
void dist_halo_exchange(Ctx& c, HaloField& hf, const double* x_owned, double* halo);
}  // namespace poro
