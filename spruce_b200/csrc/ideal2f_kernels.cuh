// ideal2f_kernels.cuh -- the two-fluid (ion / electron) equation set on the device.
//
//   Ideal2F::computeTimeDerivativesDerived      source/equationsets/ideal2F.cpp:30-96   (use_sub_cycling = false branch; the
//                                               default `true` leaves six 1x1 grids that abort at the mask multiply)
//   Ideal2F::enforceMinimums / recomputeDerived / recomputeDT / ionTimescale            ideal2F.cpp:98-208
//   EICThermalization::computeTimeDerivativesModule                                     source/modules/ucnp/eic_thermalization.cpp:27-44
//   PlasmaDomain::ucnp/fixed/reflectBoundaryExtrapolate for two species                 source/mhd/evolution.cpp:231-333
//
// First device version of this equation set: one thread per cell, every stencil operand is read through the L1/L2 caches
// (no shared-memory staging, faces are not shared between neighbouring cells).  The arithmetic is the reference's, operation
// by operation; without EIC thermalization the results are bit-identical, with it (pow / log from CUDA's libm) within 1e-9.
#pragma once
#include "module_kernels.cuh"

namespace spruce {

constexpr int NEV2 = 14;
enum { F_IRHO = 0, F_ERHO, F_IMX, F_IMY, F_EMX, F_EMY, F_IE, F_EE, F_EX, F_EY, F_EZ, F_BX, F_BY, F_BZ };   // ideal2F.hpp:44-46
constexpr double kE = 4.80320425e-10;         // E   source/constants.hpp:19
constexpr double kC = 29979245800.0;          // C   source/constants.hpp:18 (integer literal, exact as a double)

struct TfArgs {
    const double *S[NEV2];      // state the right-hand side is evaluated on
    const double *B[NEV2];      // state the increment is added to
    double *D[NEV2];            // destination
    const double *st[NSTATIC];  // be_x, be_y, (be_z unused: ideal2F.cpp:137), grav_x, grav_y
    double *K1[NEV2], *K2[NEV2];
    int kmode, primary, eic;
    double coef;
    const double *step_ptr;
    const int *done_ptr;
    unsigned long long *dtmin_bits;
    const double *vel[4];       // i_v_x, i_v_y, e_v_x, e_v_y of the state S, every cell (k_2f_velocity), so that no stencil operand needs a division
    double m_e, rm_e;           // electron mass and RN(1/m_e)
    int curl_terms;             // !remove_curl_terms
    int fast;                   // cells whose stencil needs no index wrap take the FAST instances of the operators (SPRUCE_FAST_INTERIOR, default on; same results bit for bit)
};

// The right-hand side is evaluated at interior cells only, whose stencils (transport: +-2, central: +-1 along one axis) never leave the array on a wall side;
// only a single-rank periodic axis makes an index wrap.  Away from such an edge the operators below run as FAST instances: plain addressing, no wrap, no clamp.
__device__ __forceinline__ bool tf_no_wrap(const DomainParams &P, int r, int j)
{
    return (!P.xwrap || (r >= HALO && r <= P.nx - 1 - HALO)) && (!P.yper || (j >= HALO && j <= P.ny - 1 - HALO));
}

// transportDerivative1D of functor Q with velocity functor V along `index` at cell (r,j)  (derivs.cpp:122-162)
template <class FQ, class FV>
__device__ __forceinline__ double T1(const DomainParams &P, FQ Q, FV V, int index, int r, int j)
{
    const AxisTab &t = index == 0 ? P.tx : P.ty;
    const int i0 = index == 0 ? r : j;
    double flux[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int f = i0 + k;
        auto at = [&](int i) { return index == 0 ? Q(i, j) : Q(r, i); };
        auto vat = [&](int i) { return index == 0 ? V(i, j) : V(r, i); };
        const FaceGeom g = load_face_geom(t, f);
        const double vf = face_interp(vat(f - 1), vat(f), g.hm1, g.h0, g.fs, g.rfs);
        double d2;
        const double S = upwind_face(at(f - 2), at(f - 1), at(f), at(f + 1), vf, g, &d2);
        flux[k] = S * vf;
    }
    return ddiv(flux[1] - flux[0], t.d[i0], t.rd[i0]);
}

// Out-of-line forms of the two stencil operators on planes.  The right-hand side uses 16 transport and 20 central derivatives per
// cell; inlined that is ~12 000 instructions (190 KB) of straight-line code per thread, far beyond the instruction cache.  As real
// functions the kernel is ~2 000 instructions.  Same arithmetic, operation by operation.
__device__ __noinline__ double tf_T1(const DomainParams &P, const double *q, const double *v, int index, int r, int j)
{
    auto Q = [&](int a, int b) { return rd(P, q, a, b); };
    auto V = [&](int a, int b) { return rd(P, v, a, b); };
    return T1(P, Q, V, index, r, j);
}
// transportDerivative1D of four planes that share one velocity (a species' density, two momenta and thermal energy): the face
// velocities and the face geometry are evaluated once for the four (derivs.cpp:122-162 evaluates them once per call as well)
template <bool FAST = false>
__device__ __noinline__ void tf_T4(const DomainParams &P, const double *q0, const double *q1, const double *q2, const double *q3, const double *v,
                                   int index, int r, int j, double *out)
{
    const AxisTab &t = index == 0 ? P.tx : P.ty;
    const int i0 = index == 0 ? r : j;
    auto vat = [&](int i) { return index == 0 ? rdT<FAST>(P, v, i, j) : rdT<FAST>(P, v, r, i); };
    const FaceGeom g0 = load_face_geom(t, i0), g1 = load_face_geom(t, i0 + 1);
    const double vf0 = face_interp(vat(i0 - 1), vat(i0), g0.hm1, g0.h0, g0.fs, g0.rfs);
    const double vf1 = face_interp(vat(i0), vat(i0 + 1), g1.hm1, g1.h0, g1.fs, g1.rfs);
    const double *qs[4] = {q0, q1, q2, q3};
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const double *q = qs[k];
        auto at = [&](int i) { return index == 0 ? rdT<FAST>(P, q, i, j) : rdT<FAST>(P, q, r, i); };
        const double a = at(i0 - 2), b = at(i0 - 1), c = at(i0), d = at(i0 + 1), e = at(i0 + 2);
        double d2;
        const double S0 = upwind_face(a, b, c, d, vf0, g0, &d2);
        const double S1 = upwind_face(b, c, d, e, vf1, g1, &d2);
        out[k] = ddiv(S1 * vf1 - S0 * vf0, t.d[i0], t.rd[i0]);
    }
}

// derivative1D along `index` of the plane expression (a [+ b]) * scale   (scale = 1.0 is exact; b may be null)
template <bool FAST = false>
__device__ __noinline__ double tf_D(const DomainParams &P, const double *a, const double *b, double scale, int index, int r, int j)
{
    auto F = [&](int x, int y) { return (b ? rdT<FAST>(P, a, x, y) + rdT<FAST>(P, b, x, y) : rdT<FAST>(P, a, x, y)) * scale; };
    return index == 0 ? Dx<FAST>(P, F, r, j) : Dy<FAST>(P, F, r, j);            // FAST: the cell itself is interior (the caller's condition)
}

// Ideal2F::recomputeDT for one cell (ideal2F.cpp:169-198): electron Langmuir group speed, min with the EM Courant limit
__device__ __forceinline__ double tf_cell_dt(const DomainParams &P, const TfArgs &A, double e_rho, double emx, double emy, double e_e, double dx, double dy)
{
    const double e_n = ddiv(e_rho, A.m_e, A.rm_e);
    const double e_press = e_e * P.gm1;
    const double e_temp = e_press / (e_n * kKB);
    const double evx = emx / e_rho, evy = emy / e_rho;
    const double kx = (2. * kPI) / (dx * 2), ky = (2. * kPI) / (dy * 2);
    const double v_th = sqrt((e_temp * kKB) / A.m_e);
    const double w_pe = sqrt((((e_n * (4 * kPI)) * kE) * kE) / A.m_e);
    const double w_th_x = sqrt(((kx * kx) * 3) * (v_th * v_th)), w_th_y = sqrt(((ky * ky) * 3) * (v_th * v_th));
    const double w_L_x = sqrt(w_pe * w_pe + w_th_x * w_th_x), w_L_y = sqrt(w_pe * w_pe + w_th_y * w_th_y);
    const double v_mag_x = fabs(evx) + w_L_x / kx, v_mag_y = fabs(evy) + w_L_y / ky;
    const double dt_v = 1. / (v_mag_x / dx + v_mag_y / dy);
    if (!A.curl_terms) return dt_v;
    const double dt_EM = ((dx * dy) / (dx + dy)) / kC;
    return smin(dt_v, dt_EM);
}
__device__ __forceinline__ double tf_cell_dt_ion(const DomainParams &P, double i_rho, double imx, double imy, double i_e, double dx, double dy)
{
    const double c_s = sqrt((P.gamma * (i_e * P.gm1)) / i_rho);                    // ionTimescale :201-208
    const double vx = fabs(imx / i_rho) + c_s, vy = fabs(imy / i_rho) + c_s;
    return 1. / (vx / dx + vy / dy);
}

// fixed / reflect zero every momentum of both species in the two ghost cells and the first interior cell (primary state only)
__device__ __forceinline__ bool tf_zeroed(const DomainParams &P, int g, int j) { return zero_zones(P, g, j) != 0u; }

// computeTimeDerivativesDerived at one interior cell (ideal2F.cpp:30-96) [+ EICThermalization]: k[0..13]
template <bool FAST>
__device__ __forceinline__ void tf_rhs(const DomainParams &P, const TfArgs &A, int r, int j, size_t off, double *k)
{
    const double *ivxp = A.vel[0], *ivyp = A.vel[1], *evxp = A.vel[2], *evyp = A.vel[3];       // = i_mom_x / i_rho etc. (ideal2F.cpp:123-126), formed once per cell
    // transportDivergence2D (derivs.cpp:216-220) of rho, mom_x, mom_y, thermal_energy of each species: x term + y term
    double tix[4], tiy[4], tex[4], tey[4];
    tf_T4<FAST>(P, A.S[F_IRHO], A.S[F_IMX], A.S[F_IMY], A.S[F_IE], ivxp, 0, r, j, tix);
    tf_T4<FAST>(P, A.S[F_IRHO], A.S[F_IMX], A.S[F_IMY], A.S[F_IE], ivyp, 1, r, j, tiy);
    tf_T4<FAST>(P, A.S[F_ERHO], A.S[F_EMX], A.S[F_EMY], A.S[F_EE], evxp, 0, r, j, tex);
    tf_T4<FAST>(P, A.S[F_ERHO], A.S[F_EMX], A.S[F_EMY], A.S[F_EE], evyp, 1, r, j, tey);
    auto TDi = [&](int v) { const int k_ = v == F_IRHO ? 0 : v == F_IMX ? 1 : v == F_IMY ? 2 : 3; return tix[k_] + tiy[k_]; };
    auto TDe = [&](int v) { const int k_ = v == F_ERHO ? 0 : v == F_EMX ? 1 : v == F_EMY ? 2 : 3; return tex[k_] + tey[k_]; };
    auto Dpl = [&](const double *pl, int index) { return tf_D<FAST>(P, pl, nullptr, 1.0, index, r, j); };
    const double i_rho = A.S[F_IRHO][off], e_rho = A.S[F_ERHO][off];
    const double i_n = ddiv(i_rho, P.m_i, P.rm_i), e_n = ddiv(e_rho, A.m_e, A.rm_e);
    const double ivx = ivxp[off], ivy = ivyp[off], evx = evxp[off], evy = evyp[off];
    const double bz = A.S[F_BZ][off], Ex = A.S[F_EX][off], Ey = A.S[F_EY][off];
    const double gx = A.st[S_GX][off], gy = A.st[S_GY][off];
    const double ip_c = A.S[F_IE][off] * P.gm1, ep_c = A.S[F_EE][off] * P.gm1;
    // Lorentz forces, ideal2F.cpp:42-52
    const double icx = ivy * bz, icy = (ivx * -1.0) * bz, ecx = evy * bz, ecy = (evx * -1.0) * bz;
    const double iFx = (i_n * kE) * (Ex + icx / kC), iFy = (i_n * kE) * (Ey + icy / kC);
    const double eFx = (e_n * -kE) * (Ex + ecx / kC), eFy = (e_n * -kE) * (Ey + ecy / kC);
    k[F_IRHO] = TDi(F_IRHO) * -1.0;                                                                 // :39
    k[F_ERHO] = TDe(F_ERHO) * -1.0;                                                                 // :40
    k[F_IMX] = (((TDi(F_IMX) * -1.0) - tf_D<FAST>(P, A.S[F_IE], nullptr, P.gm1, 0, r, j)) + i_rho * gx) + iFx;   // :54-56
    k[F_IMY] = (((TDi(F_IMY) * -1.0) - tf_D<FAST>(P, A.S[F_IE], nullptr, P.gm1, 1, r, j)) + i_rho * gy) + iFy;   // :57-59
    k[F_EMX] = (((TDe(F_EMX) * -1.0) - tf_D<FAST>(P, A.S[F_EE], nullptr, P.gm1, 0, r, j)) + e_rho * gx) + eFx;   // :60-62
    k[F_EMY] = (((TDe(F_EMY) * -1.0) - tf_D<FAST>(P, A.S[F_EE], nullptr, P.gm1, 1, r, j)) + e_rho * gy) + eFy;   // :63-65
    k[F_IE] = (TDi(F_IE) * -1.0) - ip_c * (Dpl(ivxp, 0) + Dpl(ivyp, 1));                           // :67-68
    k[F_EE] = (TDe(F_EE) * -1.0) - ep_c * (Dpl(evxp, 0) + Dpl(evyp, 1));                           // :69-70
    const double jx = (i_n * kE) * ivx - (e_n * kE) * evx, jy = (i_n * kE) * ivy - (e_n * kE) * evy;   // :128-129
    if (A.curl_terms) {                                                                             // :75-80
        k[F_EX] = Dpl(A.S[F_BZ], 1) * kC - jx * (4. * kPI);
        k[F_EY] = Dpl(A.S[F_BZ], 0) * -kC - jy * (4. * kPI);
        k[F_EZ] = (tf_D<FAST>(P, A.st[S_BEY], A.S[F_BY], 1.0, 0, r, j) - tf_D<FAST>(P, A.st[S_BEX], A.S[F_BX], 1.0, 1, r, j)) * kC;
        k[F_BX] = Dpl(A.S[F_EZ], 1) * -kC;
        k[F_BY] = Dpl(A.S[F_EZ], 0) * kC;
        k[F_BZ] = (Dpl(A.S[F_EX], 1) - Dpl(A.S[F_EY], 0)) * kC;
    } else {                                                                                        // :83-88
        k[F_EX] = (jx * (4. * kPI)) * -1.0;
        k[F_EY] = (jy * (4. * kPI)) * -1.0;
    }
    if (A.eic) {                                                                                    // eic_thermalization.cpp:27-44
        const double n = i_n, Te = (A.S[F_EE][off] * P.gm1) / (e_n * kKB);                          // n = i_n (ideal2F.cpp:145), e_temp :134
        const double a = cbrt((3. / 4. / kPI) / n);                     // std::pow(x, 1./3.) to ~4e-16 relative (module held to 1e-9)
        const double w_pe = sqrt(n * (4. * kPI * kE * kE / A.m_e));
        const double Gam = ((kE * kE / kKB) / Te) / a;
        const double g15 = Gam * sqrt(Gam);                              // Gam^(3/2), within 1.5 ulp of std::pow
        const double Lam = (1. / sqrt(3.)) / g15;
        const double gam_ei = ((g15 * sqrt(2. / 3. / kPI)) * w_pe) * log(Lam);
        const double nu_ei = gam_ei * (2. * A.m_e / P.m_i);
        const double dE = nu_ei * (A.S[F_EE][off] - A.S[F_IE][off]);
        k[F_EE] = k[F_EE] - dE;                    // mask = 1 here
        k[F_IE] = k[F_IE] + dE;
    }

}

__global__ void __launch_bounds__(128) k_2f_stage(const __grid_constant__ DomainParams P, const __grid_constant__ TfArgs A)
{
    if (*A.done_ptr) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double dtc = 1.7976931348623157e308;
    if (j < P.ny) {
        const size_t off = (size_t)r * P.pitch + j;
        const int g = P.row0 + r;
        const bool interior = is_interior(P, r, j);
        double k[NEV2];
#pragma unroll
        for (int v = 0; v < NEV2; v++) k[v] = 0.0;
        if (interior) {
            if (A.fast && tf_no_wrap(P, r, j)) tf_rhs<true>(P, A, r, j, off, k);
            else tf_rhs<false>(P, A, r, j, off, k);
        }
        if (A.kmode == KM_STORE_K1 || A.kmode == KM_EXPORT) {
#pragma unroll
            for (int v = 0; v < NEV2; v++) A.K1[v][off] = k[v];
        } else if (A.kmode == KM_STORE_K2) {
#pragma unroll
            for (int v = 0; v < NEV2; v++) A.K2[v][off] = k[v];
        } else if (A.kmode == KM_ADD_K2) {
#pragma unroll
            for (int v = 0; v < NEV2; v++) A.K2[v][off] = A.K2[v][off] + k[v];
        } else if (A.kmode == KM_FINAL) {
#pragma unroll
            for (int v = 0; v < NEV2; v++) k[v] = (A.K1[v][off] + k[v]) / 6.0 + A.K2[v][off] / 3.0;
        }
        if (A.kmode != KM_EXPORT) {
            const double s = A.coef * (*A.step_ptr);
            double U[NEV2];
#pragma unroll
            for (int v = 0; v < NEV2; v++) U[v] = A.B[v][off] + k[v] * s;                                   // equationset.cpp:226-228
            // enforceMinimums, ideal2F.cpp:98-105
            U[F_IRHO] = smax(ddiv(U[F_IRHO], P.m_i, P.rm_i), P.n_min) * P.m_i;
            U[F_ERHO] = smax(ddiv(U[F_ERHO], A.m_e, A.rm_e), P.n_min) * A.m_e;
            U[F_IE] = smax(U[F_IE], P.e_min);
            U[F_EE] = smax(U[F_EE], P.e_min);
            if (A.primary && tf_zeroed(P, g, j)) { U[F_IMX] = 0.0; U[F_IMY] = 0.0; U[F_EMX] = 0.0; U[F_EMY] = 0.0; }
#pragma unroll
            for (int v = 0; v < NEV2; v++) A.D[v][off] = U[v];
            if (A.primary && interior) dtc = tf_cell_dt(P, A, U[F_ERHO], U[F_EMX], U[F_EMY], U[F_EE], P.tx.d[r], P.ty.d[j]);
        }
    }
    if (A.primary && A.kmode != KM_EXPORT) block_min_to_global(dtc, A.dtmin_bits);
}

// species velocities of one state, every cell incl. ghost cells (recomputeDerivedVarsFromEvolvedVars, ideal2F.cpp:123-126)
struct TfVelArgs { const double *U[NEV2]; double *vel[4]; const int *done_ptr; int row_off; };     // row_off = -HALO: also the halo rows of a slab
__global__ void __launch_bounds__(256) k_2f_velocity(const DomainParams P, const TfVelArgs A)
{
    if (*A.done_ptr) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)blockIdx.y + A.row_off;
    if (j >= P.ny) return;
    const long long off = (long long)r * P.pitch + j;
    const double ir = A.U[F_IRHO][off], er = A.U[F_ERHO][off];
    A.vel[0][off] = A.U[F_IMX][off] / ir; A.vel[1][off] = A.U[F_IMY][off] / ir;
    A.vel[2][off] = A.U[F_EMX][off] / er; A.vel[3][off] = A.U[F_EMY][off] / er;
}

// propagateChanges on the primary state (setup, module edits): floors, pointwise zeroing, dt minimum
struct TfPropArgs { double *U[NEV2]; const double *i_temp, *e_temp; int from_state; unsigned long long *dtmin_bits; TfArgs base; };
__global__ void __launch_bounds__(256) k_2f_propagate(const DomainParams P, const TfPropArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double dtc = 1.7976931348623157e308;
    if (j < P.ny) {
        const size_t off = (size_t)r * P.pitch + j;
        const int g = P.row0 + r;
        double i_rho = A.U[F_IRHO][off], e_rho = A.U[F_ERHO][off], i_e = A.U[F_IE][off], e_e = A.U[F_EE][off];
        if (A.from_state) {                                                                                 // ideal2F.cpp:107-117
            const double i_n = ddiv(i_rho, P.m_i, P.rm_i), e_n = ddiv(e_rho, A.base.m_e, A.base.rm_e);
            i_e = ((i_n * kKB) * A.i_temp[off]) / P.gm1;
            e_e = ((e_n * kKB) * A.e_temp[off]) / P.gm1;
        }
        i_rho = smax(ddiv(i_rho, P.m_i, P.rm_i), P.n_min) * P.m_i;
        e_rho = smax(ddiv(e_rho, A.base.m_e, A.base.rm_e), P.n_min) * A.base.m_e;
        i_e = smax(i_e, P.e_min); e_e = smax(e_e, P.e_min);
        double emx = A.U[F_EMX][off], emy = A.U[F_EMY][off];
        if (tf_zeroed(P, g, j)) { A.U[F_IMX][off] = 0.0; A.U[F_IMY][off] = 0.0; A.U[F_EMX][off] = 0.0; A.U[F_EMY][off] = 0.0; emx = 0.0; emy = 0.0; }
        A.U[F_IRHO][off] = i_rho; A.U[F_ERHO][off] = e_rho; A.U[F_IE][off] = i_e; A.U[F_EE][off] = e_e;
        if (is_interior(P, r, j)) dtc = tf_cell_dt(P, A.base, e_rho, emx, emy, e_e, P.tx.d[r], P.ty.d[j]);
    }
    block_min_to_global(dtc, A.dtmin_bits);
}

// ghost cells: open_ucnp copies the nearest interior cell of densities, thermal energies, E and momenta (evolution.cpp:321-331) on
// the set being propagated; reflect copies densities and thermal energies on the primary state (momenta are zeroed pointwise)
struct TfGhostArgs { double *U[NEV2]; int primary; };
__global__ void k_2f_ghosts(const DomainParams P, const TfGhostArgs A)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int side = blockIdx.y;
    const int n = side < 2 ? P.ny : P.nx;
    if (idx >= n) return;
    const int bc = side == 0 ? P.bc_x1 : side == 1 ? P.bc_x2 : side == 2 ? P.bc_y1 : P.bc_y2;
    if (bc != BC_OPEN_UCNP && bc != BC_REFLECT) return;
    const bool xside = side < 2;
    if (xside) { if (idx < P.yl || idx > P.yu) return; } else { const int g = P.row0 + idx; if (g < P.xl || g > P.xu) return; }
    int r1_, r2_, r3_, c1_, c2_, c3_;                                  // local rows; an x side belongs to the first / last slab only
    if (side == 0) { r1_ = 0 - P.row0; r2_ = 1 - P.row0; r3_ = 2 - P.row0; c1_ = c2_ = c3_ = idx; }
    else if (side == 1) { r1_ = P.gnx - 1 - P.row0; r2_ = P.gnx - 2 - P.row0; r3_ = P.gnx - 3 - P.row0; c1_ = c2_ = c3_ = idx; }
    else if (side == 2) { r1_ = r2_ = r3_ = idx; c1_ = 0; c2_ = 1; c3_ = 2; }
    else { r1_ = r2_ = r3_ = idx; c1_ = P.ny - 1; c2_ = P.ny - 2; c3_ = P.ny - 3; }
    if (xside && (r3_ < 0 || r3_ >= P.nx)) return;
    const size_t o1 = (size_t)r1_ * P.pitch + c1_, o2 = (size_t)r2_ * P.pitch + c2_, o3 = (size_t)r3_ * P.pitch + c3_;
    if (bc == BC_OPEN_UCNP) {
        const int vars[11] = {F_IRHO, F_ERHO, F_IE, F_EE, F_EX, F_EY, F_EZ, F_IMX, F_IMY, F_EMX, F_EMY};
#pragma unroll
        for (int k = 0; k < 11; k++) { const double x = A.U[vars[k]][o3]; A.U[vars[k]][o1] = x; A.U[vars[k]][o2] = x; }
    } else if (A.primary) {
        const int vars[4] = {F_IE, F_EE, F_IRHO, F_ERHO};
#pragma unroll
        for (int k = 0; k < 4; k++) { const double x = A.U[vars[k]][o3]; A.U[vars[k]][o1] = x; A.U[vars[k]][o2] = x; }
    }
}

// every Ideal2F variable on demand (ideal2F.hpp:32-38 numbering; recomputeDerivedVarsFromEvolvedVars :119-154, recomputeDT :169-198)
enum { W_i_rho = 0, W_e_rho, W_i_mom_x, W_i_mom_y, W_e_mom_x, W_e_mom_y, W_i_temp, W_e_temp, W_bi_x, W_bi_y, W_bi_z, W_E_x, W_E_y, W_E_z, W_grav_x, W_grav_y,
       W_i_n, W_e_n, W_i_v_x, W_i_v_y, W_e_v_x, W_e_v_y, W_j_x, W_j_y, W_i_press, W_e_press, W_press, W_i_thermal_energy, W_e_thermal_energy,
       W_rho, W_rho_c, W_n, W_dn, W_dt, W_dt_i, W_b_x, W_b_y, W_b_z, W_b_mag, W_b_mag_xy, W_b_hat_x, W_b_hat_y, W_curlE_z, W_divE, W_divB, W_i_dPdx, W_e_dPdx, W_COUNT };
struct TfDeriveArgs { const double *U[NEV2]; const double *st[NSTATIC]; double *out; int which; TfArgs base; };
__global__ void __launch_bounds__(128) k_2f_derive(const DomainParams P, const TfDeriveArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    auto F = [&](int v) { return [&, v](int a, int b) { return rd(P, A.U[v], a, b); }; };
    const double i_rho = A.U[F_IRHO][off], e_rho = A.U[F_ERHO][off];
    const double i_n = ddiv(i_rho, P.m_i, P.rm_i), e_n = ddiv(e_rho, A.base.m_e, A.base.rm_e);
    const double ivx = A.U[F_IMX][off] / i_rho, ivy = A.U[F_IMY][off] / i_rho, evx = A.U[F_EMX][off] / e_rho, evy = A.U[F_EMY][off] / e_rho;
    const double i_p = A.U[F_IE][off] * P.gm1, e_p = A.U[F_EE][off] * P.gm1;
    const double bx = A.st[S_BEX][off] + A.U[F_BX][off], by = A.st[S_BEY][off] + A.U[F_BY][off], bz = A.U[F_BZ][off];
    auto b_x = [&](int a, int b) { return rd(P, A.st[S_BEX], a, b) + rd(P, A.U[F_BX], a, b); };
    auto b_y = [&](int a, int b) { return rd(P, A.st[S_BEY], a, b) + rd(P, A.U[F_BY], a, b); };
    auto ip_f = [&](int a, int b) { return rd(P, A.U[F_IE], a, b) * P.gm1; };
    auto ep_f = [&](int a, int b) { return rd(P, A.U[F_EE], a, b) * P.gm1; };
    double o = 0.0;
    switch (A.which) {
    case W_i_rho: o = i_rho; break;            case W_e_rho: o = e_rho; break;
    case W_i_mom_x: o = A.U[F_IMX][off]; break; case W_i_mom_y: o = A.U[F_IMY][off]; break;
    case W_e_mom_x: o = A.U[F_EMX][off]; break; case W_e_mom_y: o = A.U[F_EMY][off]; break;
    case W_i_temp: o = i_p / (i_n * kKB); break; case W_e_temp: o = e_p / (e_n * kKB); break;
    case W_bi_x: o = A.U[F_BX][off]; break; case W_bi_y: o = A.U[F_BY][off]; break; case W_bi_z: o = bz; break;
    case W_E_x: o = A.U[F_EX][off]; break; case W_E_y: o = A.U[F_EY][off]; break; case W_E_z: o = A.U[F_EZ][off]; break;
    case W_grav_x: o = A.st[S_GX][off]; break; case W_grav_y: o = A.st[S_GY][off]; break;
    case W_i_n: case W_n: o = i_n; break;      case W_e_n: o = e_n; break;
    case W_i_v_x: o = ivx; break; case W_i_v_y: o = ivy; break; case W_e_v_x: o = evx; break; case W_e_v_y: o = evy; break;
    case W_j_x: o = (i_n * kE) * ivx - (e_n * kE) * evx; break;
    case W_j_y: o = (i_n * kE) * ivy - (e_n * kE) * evy; break;
    case W_i_press: o = i_p; break; case W_e_press: o = e_p; break; case W_press: o = i_p + e_p; break;
    case W_i_thermal_energy: o = A.U[F_IE][off]; break; case W_e_thermal_energy: o = A.U[F_EE][off]; break;
    case W_rho: o = i_rho + e_rho; break;      case W_rho_c: o = (i_n - e_n) * kE; break;     case W_dn: o = i_n - e_n; break;
    case W_dt: o = tf_cell_dt(P, A.base, e_rho, A.U[F_EMX][off], A.U[F_EMY][off], A.U[F_EE][off], P.tx.d[r], P.ty.d[j]); break;
    case W_dt_i: o = tf_cell_dt_ion(P, i_rho, A.U[F_IMX][off], A.U[F_IMY][off], A.U[F_IE][off], P.tx.d[r], P.ty.d[j]); break;
    case W_b_x: o = bx; break; case W_b_y: o = by; break; case W_b_z: o = bz; break;
    case W_b_mag: o = sqrt((bx * bx + by * by) + bz * bz); break;
    case W_b_mag_xy: o = sqrt(bx * bx + by * by); break;
    case W_b_hat_x: case W_b_hat_y: { const double m = sqrt(bx * bx + by * by); o = (m == 0.0) ? 0.0 : (A.which == W_b_hat_x ? bx : by) / m; } break;
    case W_curlE_z: o = (Dy(P, F(F_EX), r, j) - Dx(P, F(F_EY), r, j)) * -1.0; break;                       // :149
    case W_divE: o = Dx(P, F(F_EX), r, j) + Dy(P, F(F_EY), r, j); break;                                    // :147
    case W_divB: o = Dx(P, b_x, r, j) + Dy(P, b_y, r, j); break;                                            // :148
    case W_i_dPdx: o = Dx(P, ip_f, r, j); break; case W_e_dPdx: o = Dx(P, ep_f, r, j); break;               // :150-151
    default: break;
    }
    A.out[off] = o;
}

}  // namespace spruce
