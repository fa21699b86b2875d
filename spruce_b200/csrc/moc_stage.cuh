// moc_stage.cuh -- kernels of the `open_moc` boundary (see moc_kernels.cuh for the arithmetic and the reference lines it replaces).
//
// k_moc_stage runs right after the fused stage kernel of the same Runge-Kutta stage.  The stage kernel has treated the ghost cells
// of an open_moc side like any other ghost cell (right-hand side masked to zero, idealmhd.cpp:99-103); this kernel recomputes those
// cells with the characteristic right-hand side the reference adds after the mask (char_evolution, idealmhd.cpp:88-103): one thread
// per evolved ghost cell evaluates k from the stage's input state S, applies the integrator's K-plane rule (evolution.cpp:103-124),
// writes D = B + k*s with the floors and the pointwise boundary zeroing, and -- in the step's last stage -- folds the cell's new dt into
// the running minimum, whose bounds include the ghost zone on open_moc sides (plasmadomain.cpp:155-159).
// k_moc_visc_min evaluates the minimum behind global_visc_coeff (idealmhd.cpp:90) when global_viscosity != 0.
// Slabs: cells are addressed by their GLOBAL row; a rank evolves the part of each strip it holds (the x sides belong to the first / last
// slab), reading its halo rows for the stencils along x; the minimum behind global_visc_coeff is all-gathered over the peer segments.
// STATUS: written after the round-1 GPU budget was spent.  The arithmetic is proven on the host (tests/test_moc_host_check.py);
// the launch side has not run on a GPU yet (tests/test_zz_gpu_unvalidated.py).
#pragma once
#include "mhd_kernels.cuh"
#include "moc_kernels.cuh"

namespace spruce {

struct MocArgs {
    const double *S[NEV];         // state the right-hand side is evaluated on
    const double *B[NEV];         // state the increment is added to
    double *D[NEV];               // destination
    const double *st[NSTATIC];
    double *K1[NEV], *K2[NEV];
    double *base_buf;             // [NEV][moc_threads]: B of the strip cells, saved BEFORE the stage kernel -- in the last stage D aliases B,
                                  // and the stage kernel's masked update of a ghost cell (n -> rho -> n round trip) is not the identity
    int kmode, primary;
    int dt_only;                  // after a propagateChanges: only the dt of the evolved ghost cells of D (= primary state)
    double coef;
    const double *step_ptr;
    const int *done_ptr;
    unsigned long long *dtmin_bits;
    const unsigned long long *visc_min_bits;   // minimum of visc_min_term over the dt bounds (ordered bits), null when global_viscosity == 0
    double global_viscosity;
};

__device__ __forceinline__ moc::Field moc_field(const DomainParams &P, const double *const *U, const double *const *st, double visc)
{
    moc::Field F;
    // global row indexing: shift every plane (and the x table) back by the slab's first row
    const long long sh = (long long)P.row0 * P.pitch;
    F.n = U[E_N] - sh; F.mx = U[E_MX] - sh; F.my = U[E_MY] - sh; F.mz = U[E_MZ] - sh; F.e = U[E_E] - sh; F.bix = U[E_BX] - sh; F.biy = U[E_BY] - sh; F.biz = U[E_BZ] - sh;
    F.bex = st[S_BEX] - sh; F.bey = st[S_BEY] - sh; F.bez = st[S_BEZ] - sh; F.gx = st[S_GX] - sh; F.gy = st[S_GY] - sh;
    F.dx = P.tx.d - P.row0; F.dy = P.ty.d;
    F.nx = P.gnx; F.ny = P.ny; F.pitch = P.pitch;
    F.x_halo = (P.xper && !P.xwrap) ? 1 : 0;
    F.bc[0] = P.bc_x1; F.bc[1] = P.bc_x2; F.bc[2] = P.bc_y1; F.bc[3] = P.bc_y2;
    F.m_i = P.m_i; F.gamma = P.gamma; F.gm1 = P.gm1; F.visc = visc;
    return F;
}

__device__ __forceinline__ bool moc_thread_cell(const DomainParams &P, int t, int *side, int *i, int *j) { return moc::thread_cell(P.nx, P.ny, P.row0, P.gnx, t, side, i, j); }   // (i, j): GLOBAL
__device__ __forceinline__ bool moc_thread_owns(const moc::Field &F, int side, int i, int j) { return moc::thread_owns(F, side, i, j); }

inline int moc_threads(const DomainParams &P) { return moc::n_threads(P.nx, P.ny); }     // = T in the kernels

// before the stage kernel: keep the base state of the evolved ghost cells
__global__ void __launch_bounds__(128) k_moc_save(const DomainParams P, const MocArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int side, i, j;
    if (*A.done_ptr || !moc_thread_cell(P, t, &side, &i, &j)) return;
    const moc::Field F = moc_field(P, A.S, A.st, 0.0);
    if (!moc_thread_owns(F, side, i, j)) return;
    const size_t q = (size_t)(i - P.row0) * P.pitch + j;               // local offset into the (unshifted) planes
    const int T = 2 * HALO * (P.nx + P.ny);
    for (int v = 0; v < NEV; v++) A.base_buf[(size_t)v * T + t] = A.B[v][q];
}

__global__ void __launch_bounds__(128) k_moc_visc_min(const DomainParams P, const MocArgs A, unsigned long long *out_bits)
{
    // every cell inside the dt bounds: (1/(1/dx^2 + 1/dy^2)) / dt of the stage's input state
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = P.row0 + (int)blockIdx.y;       // i: GLOBAL row of this slab's local row blockIdx.y
    double v = 1.7976931348623157e308;
    if (!*A.done_ptr && j < P.ny) {
        const moc::Field F = moc_field(P, A.S, A.st, 0.0);
        if (moc::in_dt_bounds(F, i, j)) {
            const size_t q = (size_t)i * P.pitch + j;
            const double dt = moc::cell_dt_plain(F, F.n[q], F.mx[q], F.my[q], F.e[q], F.bex[q] + F.bix[q], F.bey[q] + F.biy[q], F.bez[q] + F.biz[q], F.dx[i], F.dy[j]);
            v = moc::visc_min_term(F.dx[i], F.dy[j], dt);
        }
    }
    block_min_to_global(v, out_bits);
}

__global__ void __launch_bounds__(128) k_moc_stage(const DomainParams P, const MocArgs A)
{
    double dtc = 1.7976931348623157e308;
    int side, i, j;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (!*A.done_ptr && moc_thread_cell(P, t, &side, &i, &j)) {
        double visc = 0.0;
        if (A.visc_min_bits) visc = (A.global_viscosity * 0.5) * __longlong_as_double((long long)*A.visc_min_bits);      // idealmhd.cpp:90
        const moc::Field F = moc_field(P, A.S, A.st, visc);
        if (moc_thread_owns(F, side, i, j)) {
            const size_t q = (size_t)i * P.pitch + j;                          // into the row-shifted planes of F (global row index)
            const size_t ql = (size_t)(i - P.row0) * P.pitch + j;              // into the slab's own planes (K1, K2, D)
            if (A.dt_only) {
                if (moc::in_dt_bounds(F, i, j) && !moc::in_interior(F, i, j))
                    dtc = moc::cell_dt_plain(F, F.n[q], F.mx[q], F.my[q], F.e[q], F.bex[q] + F.bix[q], F.bey[q] + F.biy[q], F.bez[q] + F.biz[q], F.dx[i], F.dy[j]);
            } else {
                double k[NEV];
                moc::moc_cell_terms(F, i, j, k);
                // integrator K planes (evolution.cpp:103-124); the stage kernel has stored / added the masked zero for this cell already
                if (A.kmode == KM_STORE_K1 || A.kmode == KM_EXPORT) { for (int v = 0; v < NEV; v++) A.K1[v][ql] = k[v]; }
                else if (A.kmode == KM_STORE_K2) { for (int v = 0; v < NEV; v++) A.K2[v][ql] = k[v]; }
                else if (A.kmode == KM_ADD_K2) { for (int v = 0; v < NEV; v++) A.K2[v][ql] = A.K2[v][ql] + k[v]; }
                else if (A.kmode == KM_FINAL) { for (int v = 0; v < NEV; v++) k[v] = (A.K1[v][ql] + k[v]) / 6.0 + A.K2[v][ql] / 3.0; }
                if (A.kmode != KM_EXPORT) {
                    const double s = A.coef * *A.step_ptr;
                    double base[NEV];
                    const int T = 2 * HALO * (P.nx + P.ny);
                    for (int v = 0; v < NEV; v++) base[v] = A.base_buf[(size_t)v * T + t];
                    const moc::Floors fl{P.n_min, P.e_min};
                    const moc::Updated u = moc::advance_cell(F, fl, base, k, s, A.primary != 0, i, j);
                    A.D[E_N][ql] = u.n; A.D[E_MX][ql] = u.mx; A.D[E_MY][ql] = u.my; A.D[E_MZ][ql] = u.mz;
                    A.D[E_E][ql] = u.e; A.D[E_BX][ql] = u.bx; A.D[E_BY][ql] = u.by; A.D[E_BZ][ql] = u.bz;
                    if (A.primary && moc::in_dt_bounds(F, i, j) && !moc::in_interior(F, i, j))
                        dtc = moc::cell_dt_plain(F, u.n, u.mx, u.my, u.e, F.bex[q] + u.bx, F.bey[q] + u.by, F.bez[q] + u.bz, F.dx[i], F.dy[j]);
                }
            }
        }
    }
    if (A.dt_only || (A.primary && A.kmode != KM_EXPORT)) block_min_to_global(dtc, A.dtmin_bits);
}

// moc_b_limiting / moc_mom_limiting: one launch per open_moc side, in the reference's order; one thread per line of the side
struct MocLimitArgs { double *U[NEV]; const double *st[NSTATIC]; moc::Limits L; int side; const int *done_ptr; };
__global__ void __launch_bounds__(128) k_moc_limit(const DomainParams P, const MocLimitArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (*A.done_ptr) return;
    const bool xside = A.side < 2, lower = (A.side % 2) == 0;
    int a;
    if (xside) {                                           // lines = columns; only the slab that holds the side's four rows
        if (t >= P.ny || (lower ? P.row0 != 0 : P.row0 + P.nx != P.gnx)) return;
        a = t;
    } else {                                               // lines = this slab's rows
        if (t >= P.nx) return;
        a = P.row0 + t;
    }
    const double *U[NEV];
    for (int v = 0; v < NEV; v++) U[v] = A.U[v];
    const moc::Field F = moc_field(P, U, A.st, 0.0);
    const long long sh = (long long)P.row0 * P.pitch;
    const moc::Mutable M{A.U[E_MX] - sh, A.U[E_MY] - sh, A.U[E_MZ] - sh, A.U[E_BX] - sh, A.U[E_BY] - sh, A.U[E_BZ] - sh};
    moc::limit_line(F, M, A.L, A.side, a);
}
// after the limiters changed cells inside the dt bounds: drop the minimum the stage kernel accumulated and make k_dt_full evaluate every cell
__global__ void k_moc_force_full_dt(StepCtl *c)
{
    if (c->done) return;
    c->need_full = 1;
    c->dtmin_bits = 0x7FEFFFFFFFFFFFFFULL;
}

}  // namespace spruce
