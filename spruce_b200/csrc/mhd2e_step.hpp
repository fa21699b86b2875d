// mhd2e_step.hpp -- the time integrators of the IdealMHD2E path (evolution.cpp:59-124) as ONE sequence of stage calls, shared by the device
// executor (mhd2e_host.cuh: kernel launches) and the host executor of tests/hostcheck/mhd2e_host_check.cpp (plain loops over the same per-cell
// functions), so that the order of operations a GPU run performs is the order the host check proves against the CPU restatement.
//
// An executor X provides   int stage(int S, int B, int D, double coef, int kmode, int ghost_primary, int final_stage)
//   D = floor(B + (coef*step) * k(S)) for every cell (rhs_cell / apply_cell), with the integrator's K-plane rule; then the four boundary passes in
//   the reference's order (ghost_cell; open / reflect / fixed write set `ghost_primary`; open_ucnp writes D); then settle_cell on D and,
//   in the step's last stage, the dt minimum over the interior of D.
// and   void swap_sets(int a, int b).   Sets: 0 = the primary state, 1 and 2 = intermediates.
#pragma once

namespace spruce {
namespace e2 {

enum { KM2_NONE = 0, KM2_STORE_K1 = 1, KM2_STORE_K23 = 2, KM2_ADD_K23 = 3, KM2_FINAL = 4, KM2_EXPORT = 5 };
enum { TI2_EULER = 0, TI2_RK2 = 1, TI2_RK4 = 2 };

template <class X>
int advance(X &x, int integrator)
{
    int rc;
    if (integrator == TI2_EULER) {                       // evolution.cpp:84-88; D never aliases S: the new state lands in set 1, which then becomes the primary
        if ((rc = x.stage(0, 0, 1, 1.0, KM2_NONE, 1, 1))) return rc;
        x.swap_sets(0, 1);
    } else if (integrator == TI2_RK2) {                  // :90-101
        if ((rc = x.stage(0, 0, 1, 0.5, KM2_NONE, 0, 0))) return rc;
        if ((rc = x.stage(1, 0, 0, 1.0, KM2_NONE, 0, 1))) return rc;
    } else {                                             // :103-124, k = (k1 + k4)/6 + (k2 + k3)/3
        if ((rc = x.stage(0, 0, 1, 0.5, KM2_STORE_K1, 0, 0))) return rc;
        if ((rc = x.stage(1, 0, 2, 0.5, KM2_STORE_K23, 0, 0))) return rc;
        if ((rc = x.stage(2, 0, 1, 1.0, KM2_ADD_K23, 0, 0))) return rc;
        if ((rc = x.stage(1, 0, 0, 1.0, KM2_FINAL, 0, 1))) return rc;
    }
    return 0;
}

// the K-plane rule of one cell: k = this stage's right-hand side; K1, K23 = the cell's slots in the two K buffers.  Returns the k to apply.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void k_rule(int kmode, double *k, double *K1, double *K23, int nev)
{
    for (int v = 0; v < nev; v++) {
        if (kmode == KM2_STORE_K1 || kmode == KM2_EXPORT) K1[v] = k[v];
        else if (kmode == KM2_STORE_K23) K23[v] = k[v];
        else if (kmode == KM2_ADD_K23) K23[v] = K23[v] + k[v];
        else if (kmode == KM2_FINAL) k[v] = (K1[v] + k[v]) / 6.0 + K23[v] / 3.0;
    }
}

}  // namespace e2
}  // namespace spruce
