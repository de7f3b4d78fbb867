// amg.cuh -- smoothed-aggregation AMG hierarchy (device set-up + V-cycle).
#pragma once
#include "common.cuh"
#include <functional>

namespace poro {

struct AmgParams {
    double theta = 0.08;
    int max_levels = 10;
    int coarse_size = 400;
    int cheby_degree = 2;
    double cheby_ratio = 10.0;
    int power_its = 15;
    int post_smooth = 1;        // 0: V(nu,0) cycle (non-symmetric: only under GMRES/FGMRES)
};

struct AmgLevel {
    Csr A, P, R;
    DBuf<double> dinv;
    double lmax = 1.0;
    int bs = 1;
    int n_agg = 0;
    DBuf<double> x, b, r, d0, d1;
};

struct Amg {
    Ctx* ctx = nullptr;
    AmgParams par;
    std::vector<std::unique_ptr<AmgLevel>> levels;
    const Csr* A0 = nullptr;        // finest operator is borrowed
    DBuf<double> coarse_inv;
    bool coarse_direct = false;
    // row-partitioned runs: level-0 smoothing and residuals use the TRUE distributed operator (local rows x
    // [owned | halo] columns) after a halo exchange; the transfer operators and coarse levels stay rank-local
    const Csr* fine_mat = nullptr;
    std::function<const double*(const double*)> fine_extend;
    int prof_base = -1;             // phase-profile slot of level 0 (-1: not profiled)

    // B: device n x k row-major near-nullspace (may be null -> one constant per component)
    void setup(Ctx& c, const Csr& A, int bs, const double* B_dev, int k, const AmgParams& p);
    void apply(const double* b, double* x);     // x = V-cycle(b), zero initial guess
    const Csr& op(int l) const { return l == 0 ? *A0 : levels[l]->A; }
    double complexity() const;

    void cheby(int l, const double* b, double* x, bool zero_guess);
    void cycle(int l, const double* b, double* x);
};

}  // namespace poro
