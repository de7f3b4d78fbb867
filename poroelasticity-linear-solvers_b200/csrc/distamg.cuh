// distamg.cuh -- exchange primitives and the level set-up of the DISTRIBUTED smoothed-aggregation hierarchy.
#pragma once
#include "common.cuh"

namespace poro {

// level-0 plan from a field halo plan (poro_halo_set / poro_fields_set): all-gathers the owned sizes and learns the
// global ids of the ghosts through one vector exchange
void dist_plan_from_halo(Ctx& c, const HaloField& hf, int64_t n_owned, DistPlan& plan);
// plan of a level whose ghost columns are known by global id (ascending): one all-gather of counts + one id exchange
void dist_plan_build(Ctx& c, const std::vector<int64_t>& offsets, DBuf<int>&& ghost_gid, int n_ghost, DistPlan& plan);
// x_owned: n_owned x width row-major; ghost_out: n_ghost x width.  Collective over the plan's neighbours.
void dist_halo_vec(Ctx& c, DistPlan& plan, const double* x_owned, int width, double* ghost_out);
// sparse rows (GLOBAL column ids) of the ghost rows, in ghost order
void dist_halo_rows(Ctx& c, const DistPlan& plan, const Csr& M, Csr& ghost);
void csr_vstack(Ctx& c, const Csr& top, const Csr& bot, Csr& out);
// global column ids -> [owned | ghost] for the owned range [a, b); `extra`: further global ids that must become ghosts
void dist_localize(Ctx& c, Csr& M, int a, int b, const int* extra, int64_t n_extra, DBuf<int>& ghost_gid, int& n_ghost);
// rewrites local [owned | ghost] column ids as global ids
void dist_globalize(Ctx& c, Csr& M, const DistPlan& plan, int64_t ncols_global);

struct DistLevelOut {
    Csr P;            // owned fine rows x [owned coarse | ghost coarse]
    Csr R;            // owned coarse rows x [owned fine | ghost fine]
    Csr Ac;           // owned coarse rows x [owned coarse | ghost coarse]
    DistPlan coarse_plan;
};
// one level of the set-up: P = T - w D^-1 A [T; T_ghost], A_c = (P_ext[:, owned coarse])^T [A P_ext ; ghost rows]
void dist_amg_level(Ctx& c, const Csr& A, const DistPlan& plan, const Csr& T, const std::vector<int64_t>& coarse_offsets,
                    const double* dinv, double omega, DistLevelOut& out);
// exact selfp Schur complement of the owned rows of split 1 (oracle/distamg_rank.py: dist_selfp_schur):
//   S = A11 - A10 diag(A00)^-1 A01 with the A01 rows and the diagonal of the ghost dofs of split 0 exchanged.
// plan0 / plan1: halo plans of the two splits (local columns of A00, A10 follow plan0; of A01, A11 plan1).
// diag0_owned (optional): the diagonal to invert instead of diag(A00) (lumped mass + drag part of the velocity block, `cc`).
void dist_selfp_schur(Ctx& c, DistPlan& plan0, DistPlan& plan1, const Csr& A00, const Csr& A01, const Csr& A10,
                      const Csr& A11, Csr& S, DistPlan& planS, const double* diag0_owned = nullptr);

}  // namespace poro
