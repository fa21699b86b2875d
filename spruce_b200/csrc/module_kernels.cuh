// module_kernels.cuh -- device kernels of the operator-split physics modules on the ideal-MHD state.
//
//   ThermalConduction   source/modules/solar/thermalconduction.cpp   (field-aligned Spitzer conduction, flux saturation,
//                                                                       sub-cycled euler / rk2 / rk4)
//   RadiativeLosses     source/modules/solar/radiativelosses.cpp     (piecewise power-law optically thin losses, sub-cycled)
//   AmbientHeating      source/modules/solar/ambientheating.cpp      (e += dt * heating)
//
// Arithmetic follows the reference operation by operation (same expression order, no FMA contraction); the only
// deviation is the libm: std::pow / std::log10 come from CUDA's math library here and from glibc in the reference,
// which differ in the last bit for some arguments.  Module runs are therefore held to the 1e-9 relative tolerance of
// the north star, not to bit equality; the sub-cycle counts are compared exactly.
//
// Every differential operator of the reference returns 0 outside [xl..xu] x [yl..yu] and nested derivatives consume
// those zeros; D_x / D_y below reproduce that by testing the interior range of the cell they are
// evaluated at.
#pragma once
#include "mhd_kernels.cuh"

namespace spruce {

constexpr double kKappa0 = 1.0e-6;          // KAPPA_0      source/constants.hpp:15
constexpr double kMElectron = 9.1094e-28;   // M_ELECTRON   source/constants.hpp:9

// periodic wrap of an index that is at most one period out of range (stencil offsets are <= 4): compare-and-add, no integer modulo
__device__ __forceinline__ int wrap_i(const DomainParams &P, int r) { return P.xwrap ? (r < 0 ? r + P.nx : (r >= P.nx ? r - P.nx : r)) : r; }
__device__ __forceinline__ int wrap_j(const DomainParams &P, int j) { return P.yper ? (j < 0 ? j + P.ny : (j >= P.ny ? j - P.ny : j)) : j; }
__device__ __forceinline__ bool is_interior(const DomainParams &P, int r, int j)
{
    const int g = P.row0 + r;
    // periodic x has no ghost rows (xl = 0, xu = gnx-1); on a slab the halo rows g = -1, gnx, ... are real (wrapped) interior rows
    return (P.xper || (g >= P.xl && g <= P.xu)) && j >= P.yl && j <= P.yu;
}
// clamped/wrapped read: rows/cols outside a non-periodic domain are never used by an in-range operator
__device__ __forceinline__ double rd(const DomainParams &P, const double *f, int r, int j)
{
    r = wrap_i(P, r); j = wrap_j(P, j);
    r = max(-HALO, min(P.nx + HALO - 1, r)); j = max(0, min(P.ny - 1, j));
    return f[(size_t)r * P.pitch + j];
}
// FAST instances of the stencil helpers below serve cells whose whole stencil (the diamond |di| + |dj| <= 2) lies inside the array AND inside the
// operators' range, so that neither an index wraps or clamps nor an operator returns its out-of-range zero: the same loads and the same rounded
// expressions without the index arithmetic and the range tests, which are two thirds of the instructions of the general instances.
// deep_interior() is the test; the general instances stay for the cells near an edge.
__device__ __forceinline__ bool deep_interior(const DomainParams &P, int r, int j)
{
    const int g = P.row0 + r;
    const bool x_ok = P.xper ? (!P.xwrap || (r >= HALO && r <= P.nx - 1 - HALO))            // periodic x on a slab: the halo rows hold the neighbours' cells
                             : (g >= P.xl + HALO && g <= P.xu - HALO);
    const bool y_ok = P.yper ? (j >= HALO && j <= P.ny - 1 - HALO) : (j >= P.yl + HALO && j <= P.yu - HALO);
    return x_ok && y_ok;
}
template <bool FAST>
__device__ __forceinline__ double rdT(const DomainParams &P, const double *f, int r, int j)
{
    if (FAST) return f[(ptrdiff_t)r * (ptrdiff_t)P.pitch + j];
    return rd(P, f, r, j);
}

// x^(3/2), x^(5/2) for the modules' std::pow(T, 1.5 | 2.5): sqrt is correctly rounded, so x*sqrt(x) and (x*x)*sqrt(x) are within 1.5 ulp of the
// exact power -- the same distance CUDA's pow keeps from glibc's -- at a tenth of the instructions.  These modules are held to a relative
// 1e-9 (with equal sub-cycle counts), not to bit equality, precisely because of libm; negative arguments give NaN and 0 gives 0 as pow does.
__device__ __forceinline__ double pow15(double x) { return x * sqrt(x); }
__device__ __forceinline__ double pow25(double x) { return (x * x) * sqrt(x); }

// temp = max((gamma-1)*e / (2 K_B n), T_min)   thermalconduction.cpp:65, radiativelosses.cpp:56
__device__ __forceinline__ double temp_of(const DomainParams &P, double e, double n) { return smax((e * P.gm1) / (n * (2.0 * kKB)), P.T_min); }

// derivative1D of an arbitrary per-cell functor F(r, j) (derivs.cpp:223-264); zero outside the interior range
template <bool FAST = false, class F>
__device__ __forceinline__ double Dx(const DomainParams &P, F f, int r, int j)
{
    if (!FAST) {
        r = wrap_i(P, r); j = wrap_j(P, j);      // periodic axes: the neighbour of the first cell is the last one (derivs.cpp:243-256)
        if (!is_interior(P, r, j)) return 0.0;
    }
    const AxisTab &t = P.tx;
    const double a = f(r - 1, j), b = f(r, j), c = f(r + 1, j);
    const double hi = face_interp(b, c, t.h[r], t.h[r + 1], t.fs[r + 1], t.rfs[r + 1]);
    const double lo = face_interp(a, b, t.h[r - 1], t.h[r], t.fs[r], t.rfs[r]);
    return ddiv(hi - lo, t.d[r], t.rd[r]);
}
template <bool FAST = false, class F>
__device__ __forceinline__ double Dy(const DomainParams &P, F f, int r, int j)
{
    if (!FAST) {
        r = wrap_i(P, r); j = wrap_j(P, j);      // periodic axes: the neighbour of the first cell is the last one (derivs.cpp:243-256)
        if (!is_interior(P, r, j)) return 0.0;
    }
    const AxisTab &t = P.ty;
    const double a = f(r, j - 1), b = f(r, j), c = f(r, j + 1);
    const double hi = face_interp(b, c, t.h[j], t.h[j + 1], t.fs[j + 1], t.rfs[j + 1]);
    const double lo = face_interp(a, b, t.h[j - 1], t.h[j], t.fs[j], t.rfs[j]);
    return ddiv(hi - lo, t.d[j], t.rd[j]);
}
// secondDerivative1D (derivs.cpp:417-455): (I(i,i+1) - 2 q + I(i-1,i)) / (0.5 d)^2
template <bool FAST = false, class F>
__device__ __forceinline__ double D2x(const DomainParams &P, F f, int r, int j)
{
    if (!FAST) {
        r = wrap_i(P, r); j = wrap_j(P, j);      // periodic axes: the neighbour of the first cell is the last one (derivs.cpp:243-256)
        if (!is_interior(P, r, j)) return 0.0;
    }
    const AxisTab &t = P.tx;
    const double a = f(r - 1, j), b = f(r, j), c = f(r + 1, j);
    const double hi = face_interp(b, c, t.h[r], t.h[r + 1], t.fs[r + 1], t.rfs[r + 1]);
    const double lo = face_interp(a, b, t.h[r - 1], t.h[r], t.fs[r], t.rfs[r]);
    return ((hi - 2.0 * b) + lo) / (t.h[r] * t.h[r]);
}
template <bool FAST = false, class F>
__device__ __forceinline__ double D2y(const DomainParams &P, F f, int r, int j)
{
    if (!FAST) {
        r = wrap_i(P, r); j = wrap_j(P, j);      // periodic axes: the neighbour of the first cell is the last one (derivs.cpp:243-256)
        if (!is_interior(P, r, j)) return 0.0;
    }
    const AxisTab &t = P.ty;
    const double a = f(r, j - 1), b = f(r, j), c = f(r, j + 1);
    const double hi = face_interp(b, c, t.h[j], t.h[j + 1], t.fs[j + 1], t.rfs[j + 1]);
    const double lo = face_interp(a, b, t.h[j - 1], t.h[j], t.fs[j], t.rfs[j]);
    return ((hi - 2.0 * b) + lo) / (t.h[j] * t.h[j]);
}

struct TcParams {
    int flux_saturation;
    double kappa;              // weakening_factor*KAPPA_0
    double dt_subcycle_min;
};

struct TcFields { const double *T, *n, *bhx, *bhy; };   // n, b_hat of the primary state at module entry; T evolves

// temp, b_hat_x, b_hat_y of the primary state in one pass (the three planes both numberSubcycles and iterateModule start from, thermalconduction.cpp:49-51, :136-138):
// the expressions of k_mhd_derive's V_temp / V_b_hat_x / V_b_hat_y cases (idealmhd.cpp:254, :262-277), 8 plane reads and 3 writes instead of 18 and 3
struct TcDeriveArgs { const double *U[NEV]; const double *st[NSTATIC]; double *T, *bhx, *bhy; int row_off; };   // row_off = -HALO: also the halo rows of a slab
__global__ void __launch_bounds__(256) k_tc_derive(const DomainParams P, const TcDeriveArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)blockIdx.y + A.row_off;
    if (j >= P.ny) return;
    const long long off = (long long)r * P.pitch + j;
    const double n_ = A.U[E_N][off];
    const double p = A.U[E_E][off] * P.gm1;
    A.T[off] = smax(p / (n_ * (2 * kKB)), P.T_min);
    const double bx = A.st[S_BEX][off] + A.U[E_BX][off], by = A.st[S_BEY][off] + A.U[E_BY][off], bz = A.st[S_BEZ][off] + A.U[E_BZ][off];
    const double bm = sqrt((bx * bx + by * by) + bz * bz);
    A.bhx[off] = (bm == 0.0) ? 0.0 : bx / bm;                           // catchNullFieldDirection
    A.bhy[off] = (bm == 0.0) ? 0.0 : by / bm;
}

// fieldAlignedConductiveFlux at one cell (thermalconduction.cpp:154-178); zero outside the interior
template <bool FAST = false>
__device__ __forceinline__ void tc_raw_flux(const DomainParams &P, const TcParams &C, const TcFields &F, int r, int j, double *fx, double *fy)
{
    *fx = 0.0; *fy = 0.0;
    if (!FAST) {
        r = wrap_i(P, r); j = wrap_j(P, j);
        if (!is_interior(P, r, j)) return;
    }
    auto T = [&](int a, int b) { return rdT<FAST>(P, F.T, a, b); };
    const double Tc = T(r, j);
    const double rho = rdT<FAST>(P, F.n, r, j) * P.m_i;
    const double kmax = (((P.tx.d[r] * P.ty.d[j]) * kKB) * ddiv(rho, P.m_i, P.rm_i)) / C.dt_subcycle_min;     // :155
    const double kap = smin(pow25(Tc) * C.kappa, kmax);                                            // :156
    const double cx = (kap * -1.0) * Dx<FAST>(P, T, r, j), cy = (kap * -1.0) * Dy<FAST>(P, T, r, j);
    const double bx = rdT<FAST>(P, F.bhx, r, j), by = rdT<FAST>(P, F.bhy, r, j);
    const double fm = cx * bx + cy * by;                                                                    // :173-175
    *fx = fm * bx; *fy = fm * by;
}
// saturateConductiveFlux scale factor at one cell (thermalconduction.cpp:182-188); acts on every cell of the plane
template <bool FAST = false>
__device__ __forceinline__ void tc_saturate(const DomainParams &P, const TcFields &F, int r, int j, double *fx, double *fy)
{
    const double c1 = (1.0 / 6.0) * (3.0 / 2.0);
    if (!FAST) { r = wrap_i(P, r); j = wrap_j(P, j); }
    const double rho = rdT<FAST>(P, F.n, r, j) * P.m_i;
    const double sat = ((ddiv(rho, P.m_i, P.rm_i) * c1) * pow15(rdT<FAST>(P, F.T, r, j) * kKB)) / sqrt(kMElectron);
    const double fm = sqrt((*fx) * (*fx) + (*fy) * (*fy));
    const double sc = sat / sqrt(sat * sat + fm * fm);
    *fx *= sc; *fy *= sc;
}
// saturation coefficient of saturationTerms at one cell (thermalconduction.cpp:211-224)
// (out of line: the saturation terms evaluate it at five points; inlined five times the kernel no longer fits the instruction cache)
template <bool FAST = false>
__device__ __noinline__ double tc_coefficient(const DomainParams &P, const TcParams &C, const TcFields &F, int r, int j)
{
    double fx, fy;
    tc_raw_flux<FAST>(P, C, F, r, j, &fx, &fy);
    const double fm = sqrt(fx * fx + fy * fy);
    tc_saturate<FAST>(P, F, r, j, &fx, &fy);
    const double sfm = sqrt(fx * fx + fy * fy);
    return (fm != 0.0) ? sfm / fm : 1.0;
}

// planes of the two-pass form of saturated conduction (k_tc_coef below): saturation coefficient and raw field-aligned flux of every cell, for the temperature plane F.T
struct TcSat { const double *coef = nullptr, *rx = nullptr, *ry = nullptr; };

// thermalEnergyDerivative at one cell (thermalconduction.cpp:114-132)
template <bool FAST = false>
__device__ double tc_energy_derivative(const DomainParams &P, const TcParams &C, const TcFields &F, int r, int j, const TcSat S = TcSat())
{
    auto T = [&](int a, int b) { return rdT<FAST>(P, F.T, a, b); };
    auto BX = [&](int a, int b) { return rdT<FAST>(P, F.bhx, a, b); };
    auto BY = [&](int a, int b) { return rdT<FAST>(P, F.bhy, a, b); };
    auto TX = [&](int a, int b) { return Dx<FAST>(P, T, a, b); };
    auto TY = [&](int a, int b) { return Dy<FAST>(P, T, a, b); };
    const double Tx = TX(r, j), Ty = TY(r, j);
    const double Txx = D2x<FAST>(P, T, r, j), Tyy = D2y<FAST>(P, T, r, j);
    const double Txy = Dy<FAST>(P, TX, r, j), Tyx = Dx<FAST>(P, TY, r, j);            // nested: derivative1D(dtemp_dx,1), derivative1D(dtemp_dy,0)
    const double bxx = Dx<FAST>(P, BX, r, j), bxy = Dy<FAST>(P, BX, r, j), byx = Dx<FAST>(P, BY, r, j), byy = Dy<FAST>(P, BY, r, j);
    const double bhx = BX(r, j), bhy = BY(r, j), Tc = T(r, j);
    const double bg = bhx * Tx + bhy * Ty;
    const double t1x = bhx * Txx + bhy * Txy, t1y = bhx * Tyx + bhy * Tyy;
    const double t2x = Tx * bxx + Ty * bxy, t2y = Tx * byx + Ty * byy;
    const double cu = byx - bxy;
    const double t3x = ((cu * -1.0) * -1.0) * Ty, t3y = (cu * -1.0) * Tx;
    const double p15 = pow15(Tc), p25 = pow25(Tc);
    const double ttx = ((p15 * (5.0 / 2.0)) * Tx) * bg + p25 * ((t1x + t2x) + t3x);
    const double tty = ((p15 * (5.0 / 2.0)) * Ty) * bg + p25 * ((t1y + t2y) + t3y);
    double out = (((p25 * bg) * (bxx + byy)) + (bhx * ttx + bhy * tty)) * C.kappa;
    if (C.flux_saturation) {
        if (S.coef) {
            // two-pass form: every cell's coefficient and raw flux were written by k_tc_coef (the same functions, evaluated once per cell instead of at each of the five
            // points of every neighbour's stencil); saturationTerms (thermalconduction.cpp:211-224) differentiates the plane
            auto CP = [&](int a, int b) { return rdT<FAST>(P, S.coef, a, b); };
            const double coef = CP(r, j), rx = rdT<FAST>(P, S.rx, r, j), ry = rdT<FAST>(P, S.ry, r, j);
            const double mask = (FAST || is_interior(P, r, j)) ? 1.0 : 0.0;
            const double add = (mask * -1.0) * (Dx<FAST>(P, CP, r, j) * rx + Dy<FAST>(P, CP, r, j) * ry);
            return coef * out + add;
        }
        auto CO = [&](int a, int b) { return tc_coefficient<FAST>(P, C, F, a, b); };
        if (FAST) {
            // the general form below evaluates the coefficient seven times (the centre once by itself and once inside each derivative) and the centre's raw flux twice;
            // with direct indices the five distinct coefficients and the one raw flux can be named and reused: the same rounded expressions, 5 + 0 evaluations instead of 7 + 1
            double rx, ry;
            tc_raw_flux<true>(P, C, F, r, j, &rx, &ry);
            double sx = rx, sy = ry;                                            // tc_coefficient's body at the centre, on the flux just computed
            const double fm = sqrt(sx * sx + sy * sy);
            tc_saturate<true>(P, F, r, j, &sx, &sy);
            const double sfm = sqrt(sx * sx + sy * sy);
            const double coef = (fm != 0.0) ? sfm / fm : 1.0;
            const double cxm = CO(r - 1, j), cxp = CO(r + 1, j), cym = CO(r, j - 1), cyp = CO(r, j + 1);
            auto COX = [&](int a, int) { return a < r ? cxm : (a > r ? cxp : coef); };
            auto COY = [&](int, int b) { return b < j ? cym : (b > j ? cyp : coef); };
            const double add = (1.0 * -1.0) * (Dx<true>(P, COX, r, j) * rx + Dy<true>(P, COY, r, j) * ry);
            return coef * out + add;
        }
        const double coef = CO(r, j);
        double rx, ry;
        tc_raw_flux<FAST>(P, C, F, r, j, &rx, &ry);
        const double mask = (FAST || is_interior(P, r, j)) ? 1.0 : 0.0;
        const double add = (mask * -1.0) * (Dx<FAST>(P, CO, r, j) * rx + Dy<FAST>(P, CO, r, j) * ry);
        out = coef * out + add;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------
// Device-resident sub-cycle plan (opt-in, SPRUCE_DEVICE_SUBCYCLES=1): the sub-cycle counts of thermal conduction and radiative losses
// (numberSubcycles, thermalconduction.cpp:135-149, radiativelosses.cpp:161-166) are taken from the count kernels' reductions by ONE device thread
// instead of the host, so that a step with these modules is enqueued without a host round trip.  The host enqueues a BUDGET of conduction
// sub-cycles; launches beyond the planned count return at once.  A count above the budget stops the run before the step has changed anything
// (StepCtl::done = 3): spruce_advance raises the budget and re-enqueues from that step.
// STATUS: written after the round-2 GPU budget was spent; executed on the host from this source (tests/test_capi_hooks_emulated.py), first device run at the round's end.
// ---------------------------------------------------------------------------------------------------------
constexpr int DONE_SUBCYCLE_BUDGET = 3;     // StepCtl::done: 1 = max_time reached, 2 = a peer stopped answering, 3 = more conduction sub-cycles needed than were enqueued
struct SubPlan { int tc_nsub, rl_nsub; double tc_dts; int tc_need, tc_max; int tc_last, rl_last; };     // tc_last / rl_last: counts of the last step that was planned (curr_num_subcycles)
struct SubPlanArgs { StepCtl *ctl; const unsigned long long *red; SubPlan *plan; int tc_on, tc_sat; double tc_eps, tc_dtmin; int rl_on; double rl_eps; int tc_budget; };
__global__ void k_sub_plan(const SubPlanArgs A)
{
    SubPlan *p = A.plan;
    p->tc_nsub = 0; p->rl_nsub = 0; p->tc_dts = 0.0;
    if (A.ctl->done) return;
    const double dt = A.ctl->step;
    int tc = 0, rl = 0;
    if (A.tc_on && !(A.tc_sat && __longlong_as_double((long long)A.red[1]) == 0.0)) {              // thermalconduction.cpp:143
        const double a = A.tc_eps * __longlong_as_double((long long)A.red[0]), b = A.tc_dtmin;
        const double md = (a < b) ? b : a;                                                          // std::max :147
        tc = (int)(dt / md) + 1;                                                                    // :148
    }
    if (A.rl_on && !(__longlong_as_double((long long)A.red[3]) == 0.0)) {                           // radiativelosses.cpp:162
        const double sdt = A.rl_eps * __longlong_as_double((long long)A.red[2]);
        rl = (int)(dt / sdt) + 1;                                                                   // :165
    }
    if (tc > p->tc_max) p->tc_max = tc;
    if (tc > A.tc_budget) { p->tc_need = tc; A.ctl->done = DONE_SUBCYCLE_BUDGET; return; }          // nothing of this step has been applied yet
    p->tc_nsub = tc; p->rl_nsub = rl; p->tc_last = tc; p->rl_last = rl;
    p->tc_dts = dt / (double)tc;                                                                    // thermalconduction.cpp:60
}

enum { TC_FINAL = 0, TC_INTERMEDIATE = 1, TC_RK4_FINAL = 2 };
struct TcStageArgs {
    TcFields F;
    TcParams C;
    const double *e_base;      // thermal energy at the start of the sub-cycle
    double *e_out;             // TC_FINAL / TC_RK4_FINAL: new thermal energy (may alias e_base: own cell only)
    double *T_out;             // temperature of the produced energy
    double *K_store;           // optional: store dE of this evaluation (rk4 k1..k3)
    const double *K1, *K2, *K3;
    int mode;
    double c;                  // 0.5*dt_sub or dt_sub
    int fast;                  // deep-interior cells take the FAST instance (SPRUCE_FAST_INTERIOR, default on; same results bit for bit)
    const SubPlan *plan;       // device-resident plan (or null): sub-cycle `sub` runs only below plan->tc_nsub, c = plan->tc_dts (times 0.5 when `half`)
    int sub, half;
    TcSat S;                   // two-pass form of saturated conduction (or nulls)
};

// first pass of the two-pass form: saturation coefficient (tc_coefficient's body) and raw flux of every cell of the plane -- ghost cells (no flux: coefficient 1) and, on a
// slab, one halo row per side included (row_off = -1), which the neighbours' temperature rows make computable
struct TcCoefArgs { TcFields F; TcParams C; double *coef, *rx, *ry; int fast, row_off; const SubPlan *plan; int sub; };
__global__ void __launch_bounds__(128) k_tc_coef(const __grid_constant__ DomainParams P, const __grid_constant__ TcCoefArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)blockIdx.y + A.row_off;
    if (j >= P.ny || (A.plan && A.sub >= A.plan->tc_nsub)) return;
    const ptrdiff_t off = (ptrdiff_t)r * (ptrdiff_t)P.pitch + j;
    const bool fast = A.fast && deep_interior(P, r, j);
    double fx, fy;
    if (fast) tc_raw_flux<true>(P, A.C, A.F, r, j, &fx, &fy); else tc_raw_flux<false>(P, A.C, A.F, r, j, &fx, &fy);
    A.rx[off] = fx; A.ry[off] = fy;
    const double fm = sqrt(fx * fx + fy * fy);
    if (fast) tc_saturate<true>(P, A.F, r, j, &fx, &fy); else tc_saturate<false>(P, A.F, r, j, &fx, &fy);
    const double sfm = sqrt(fx * fx + fy * fy);
    A.coef[off] = (fm != 0.0) ? sfm / fm : 1.0;
}

__global__ void __launch_bounds__(128) k_tc_stage(const __grid_constant__ DomainParams P, const __grid_constant__ TcStageArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    double cstep = A.c;
    if (A.plan) {
        if (A.sub >= A.plan->tc_nsub) return;
        cstep = A.half ? 0.5 * A.plan->tc_dts : A.plan->tc_dts;
    }
    const size_t off = (size_t)r * P.pitch + j;
    const bool in = is_interior(P, r, j);
    const double mask = in ? 1.0 : 0.0;
    double dE = 0.0;                                                    // ghost cells: multiplied by mask = 0
    if (in) dE = (A.fast && deep_interior(P, r, j)) ? tc_energy_derivative<true>(P, A.C, A.F, r, j, A.S) : tc_energy_derivative<false>(P, A.C, A.F, r, j, A.S);
    if (A.K_store) A.K_store[off] = dE;
    const double e0 = A.e_base[off];
    double e1;
    if (A.mode == TC_RK4_FINAL) {
        const double ks = ((A.K1[off] + A.K2[off] * 2.0) + A.K3[off] * 2.0) + dE;
        e1 = smax(e0 + ((mask * cstep) * ks) / 6.0, P.e_min);                     // :96-97
    } else {
        e1 = smax(e0 + (mask * cstep) * dE, P.e_min);                             // :63-64, :70-71, :84 ...
    }
    if (A.mode != TC_INTERMEDIATE) A.e_out[off] = e1;
    A.T_out[off] = temp_of(P, e1, A.F.n[off]);
}

// numberSubcycles reductions (thermalconduction.cpp:135-149): out[0] = min over the interior of dt_subcycle (ordered bits),
// out[1] = max over the whole plane of |b_hat . grad T| (ordered bits; saturated case only)
__global__ void __launch_bounds__(128) k_tc_count(const DomainParams P, const TcParams C, const TcFields F, unsigned long long *out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double v = 1.7976931348623157e308;
    unsigned long long gmax = 0ULL;
    if (j < P.ny) {
        const size_t off = (size_t)r * P.pitch + j;
        const double rho = F.n[off] * P.m_i;
        const double nd = ddiv(rho, P.m_i, P.rm_i);
        const bool in = is_interior(P, r, j);
        if (!C.flux_saturation) {
            if (in) v = (((nd * (kKB / C.kappa)) * P.tx.d[r]) * P.ty.d[j]) / pow25(F.T[off]);      // :139
        } else {
            auto T = [&](int a, int b) { return rd(P, F.T, a, b); };
            const double ftg = Dx(P, T, r, j) * F.bhx[off] + Dy(P, T, r, j) * F.bhy[off];            // :141-142
            gmax = (unsigned long long)__double_as_longlong(fabs(ftg));
            if (ftg != ftg) gmax = 0ULL;
            if (in) {
                double fx, fy;
                tc_raw_flux(P, C, F, r, j, &fx, &fy);
                tc_saturate(P, F, r, j, &fx, &fy);
                const double km = fabs(sqrt(fx * fx + fy * fy) / ftg);                                // :201-202
                v = (((kKB / km) * nd) * P.tx.d[r]) * P.ty.d[j];                                      // :145
            }
        }
    }
    block_min_to_global(v, out);
    // max of |ftg| (non-negative doubles order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, gmax, o); gmax = t > gmax ? t : gmax; }
    if ((threadIdx.x & 31) == 0 && gmax) atomicMax(out + 1, gmax);
}

// ---------------------------------------------------------------------------------------------------------
// Radiative losses: the loss rate depends on the own cell only, so the whole sub-cycling loop runs in registers.
// ---------------------------------------------------------------------------------------------------------
struct RlParams { int integrator, prevent_subcycling; double cutoff_ramp, cutoff_temp, epsilon; };

// computeLosses at one interior cell (radiativelosses.cpp:110-158)
__device__ __forceinline__ double rl_loss(const RlParams &R, double T, double n, double e_primary, double dt_primary, double eps_domain)
{
    if (T < R.cutoff_temp) return 0.0;
    const double lt = log10(T);
    double chi, alpha;
    if (lt <= 4.97) { chi = 1.09e-31; alpha = 2.0; }
    else if (lt <= 5.67) { chi = 8.87e-17; alpha = -1.0; }
    else if (lt <= 6.18) { chi = 1.90e-22; alpha = 0.0; }
    else if (lt <= 6.55) { chi = 3.53e-13; alpha = -1.5; }
    else if (lt <= 6.90) { chi = 3.46e-25; alpha = 1.0 / 3.0; }
    else if (lt <= 7.63) { chi = 5.49e-16; alpha = -1.0; }
    else { chi = 1.96e-27; alpha = 0.5; }
    double r = (n * n) * chi * pow(T, alpha);      // pow(n, 2.0) is exactly RN(n*n)
    if (T < R.cutoff_temp + R.cutoff_ramp) { const double ramp = (T - R.cutoff_temp) / R.cutoff_ramp; r *= ramp; }
    if (R.prevent_subcycling) {
        if (0.1 * R.epsilon * (e_primary / r) < eps_domain * dt_primary) r = 0.1 * R.epsilon * e_primary / (eps_domain * dt_primary);
    }
    return r;
}

struct RlArgs {
    RlParams R;
    const double *U[NEV];
    const double *st[NSTATIC];
    double *e_out;
    int n_sub;
    double dt;                 // step size of this iteration
    unsigned long long *red;   // count mode: red[0] = min |e/L| over the whole plane, red[1] = max L
    int count_mode;
    const SubPlan *plan;       // device-resident plan (or null): n_sub = plan->rl_nsub, dt = *dt_ptr, nothing is written once *done_ptr is set
    const double *dt_ptr; const int *done_ptr;
};

__global__ void __launch_bounds__(128) k_rl(const DomainParams P, const RlArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double vmin = 1.7976931348623157e308;
    unsigned long long lmax = 0ULL;
    if (j < P.ny) {
        const size_t off = (size_t)r * P.pitch + j;
        const bool in = is_interior(P, r, j);
        const double n = A.U[E_N][off];
        const double e0 = A.U[E_E][off];
        double dtp = 0.0;
        if (A.R.prevent_subcycling && in)
            dtp = cell_dt(P, n * P.m_i, A.U[E_MX][off], A.U[E_MY][off], e0, A.st[S_BEX][off] + A.U[E_BX][off], A.st[S_BEY][off] + A.U[E_BY][off],
                          A.st[S_BEZ][off] + A.U[E_BZ][off], P.tx.d[r], P.tx.rd[r], P.ty.d[j], P.ty.rd[j]);
        auto L = [&](double T) { return in ? rl_loss(A.R, T, n, e0, dtp, P.epsilon) : 0.0; };
        double e = e0, T = temp_of(P, e0, n);
        if (A.count_mode) {
            const double l = L(T);
            const double q = fabs(e0 / l);                                        // radiativelosses.cpp:164 (inf where l == 0)
            vmin = q;
            lmax = (l == l && l > 0.0) ? (unsigned long long)__double_as_longlong(l) : 0ULL;
        } else {
            const double mask = in ? 1.0 : 0.0;
            if (A.plan && *A.done_ptr) return;                                    // apply mode has no block-wide reduction below
            const int n_sub = A.plan ? A.plan->rl_nsub : A.n_sub;
            const double dts = (A.plan ? *A.dt_ptr : A.dt) / (double)n_sub;       // :51
            for (int s = 0; s < n_sub; s++) {
                if (A.R.integrator == 0) {                                        // euler :53-57
                    e = smax(e - (mask * dts) * L(T), P.e_min);
                } else if (A.R.integrator == 1) {                                 // rk2 :58-69
                    const double half = smax(e - (mask * (0.5 * dts)) * L(T), P.e_min);
                    e = smax(e - (mask * dts) * L(temp_of(P, half, n)), P.e_min);
                } else {                                                          // rk4 :70-91
                    const double k1 = L(T) * -1.0;
                    double im = smax(e + (mask * (0.5 * dts)) * k1, P.e_min);
                    const double k2 = L(temp_of(P, im, n)) * -1.0;
                    im = smax(e + (mask * (0.5 * dts)) * k2, P.e_min);
                    const double k3 = L(temp_of(P, im, n)) * -1.0;
                    im = smax(e + (mask * dts) * k3, P.e_min);
                    const double k4 = L(temp_of(P, im, n)) * -1.0;
                    e = smax(e + ((mask * dts) * (((k1 + k2 * 2.0) + k3 * 2.0) + k4)) / 6.0, P.e_min);
                }
                T = temp_of(P, e, n);
            }
            A.e_out[off] = e;
        }
    }
    if (A.count_mode) {
        block_min_to_global(vmin, A.red);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = t > lmax ? t : lmax; }
        if ((threadIdx.x & 31) == 0 && lmax) atomicMax(A.red + 1, lmax);
    }
}

// AmbientHeating::postIterateModule (ambientheating.cpp:43): e += dt*heating  (tmp = heating*dt, then +=)
__global__ void __launch_bounds__(256) k_ambient_heating(const DomainParams P, double *e, const double *heating, const double *step_ptr, const int *done_ptr)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny || (done_ptr && *done_ptr)) return;
    const size_t off = (size_t)r * P.pitch + j;
    e[off] = e[off] + heating[off] * (*step_ptr);
}


// ---------------------------------------------------------------------------------------------------------
// Diagnostic planes of output_to_file = true (ThermalConduction::fileOutput thermalconduction.cpp:226-237, RadiativeLosses::fileOutput
// radiativelosses.cpp:172-179): the module's average rate of change of the thermal energy over the step, (e_after - e_before)/dt, taken before the
// closing propagateChanges (:101-103, :93), and the saturation coefficient of the step's first temperature field (:53-58, :103).
// STATUS: written after the round-1 GPU budget was spent; not yet run on a GPU.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_plane_copy(const DomainParams P, double *dst, const double *src)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    dst[off] = src[off];
}
__global__ void __launch_bounds__(256) k_plane_product(const DomainParams P, double *out, const double *a, const double *b)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    out[off] = a[off] * b[off];
}
__global__ void __launch_bounds__(256) k_avg_change(const DomainParams P, double *out, const double *e, const double *old, double dt)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    out[off] = (e[off] - old[off]) / dt;
}
// the same with the step size read from the device step control; keeps the previous step's plane once the run has stopped
__global__ void __launch_bounds__(256) k_avg_change_dev(const DomainParams P, double *out, const double *e, const double *old, const double *dt_ptr, const int *done_ptr)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny || *done_ptr) return;
    const size_t off = (size_t)r * P.pitch + j;
    out[off] = (e[off] - old[off]) / (*dt_ptr);
}
// multispecies_mode (plasmadomain.hpp:134-135): a module's energy input w joins the cumulative planes as  ion (+/-)= (1 - f) w  when f < 1,  electron (+/-)= f w  when f > 0,
// f = the module's ms_electron_heating_fraction.  w per cell, the fraction factor fr placed where the reference places it:
//   MS_DIFF   (a - b) fr              thermalconduction.cpp:105-108, radiativelosses.cpp:94-97        (a = sub-cycled energy, b = energy before)
//   MS_RATE   ((mask fr) a) dt        ambientheating.cpp:45-48, ambientheatingsink.cpp:38-41 (sign -1)
//   MS_PULSE  (mask fr) (a dt)        localizedheating.cpp:63-66 (without the ramp factor, as there)
//   MS_JOULE  joule += a - b          anomalousresistivity.cpp:168-170  (cum_i = the joule plane, no fraction)
enum { MS_DIFF = 0, MS_RATE = 1, MS_PULSE = 2, MS_JOULE = 3 };
struct MsArgs { double *cum_i, *cum_e; const double *a, *b; double f, sign; int mode; const double *dt_ptr; const int *done_ptr; };
__global__ void __launch_bounds__(256) k_ms_feed(const DomainParams P, const MsArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny || (A.done_ptr && *A.done_ptr)) return;
    const size_t off = (size_t)r * P.pitch + j;
    if (A.mode == MS_JOULE) { A.cum_i[off] = A.cum_i[off] + (A.a[off] - A.b[off]); return; }
    const double mask = is_interior(P, r, j) ? 1.0 : 0.0, dt = A.dt_ptr ? *A.dt_ptr : 0.0;
    auto w = [&](double fr) {
        if (A.mode == MS_DIFF) return (A.a[off] - A.b[off]) * fr;
        if (A.mode == MS_RATE) return ((mask * fr) * A.a[off]) * dt;
        return (mask * fr) * (A.a[off] * dt);
    };
    if (A.f < 1.0) A.cum_i[off] = A.sign < 0.0 ? A.cum_i[off] - w(1.0 - A.f) : A.cum_i[off] + w(1.0 - A.f);
    if (A.f > 0.0) A.cum_e[off] = A.sign < 0.0 ? A.cum_e[off] - w(A.f) : A.cum_e[off] + w(A.f);
}
__global__ void __launch_bounds__(128) k_tc_saturation_plane(const DomainParams P, const TcParams C, const TcFields F, double *out, const int *done_ptr)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny || (done_ptr && *done_ptr)) return;                   // a stopped run keeps the plane of its last step
    out[(size_t)r * P.pitch + j] = tc_coefficient(P, C, F, r, j);
}

// ---------------------------------------------------------------------------------------------------------
// Pointwise solar source terms with a static spatial template (postIterateModule hooks, evolution.cpp:74):
//   SRC_SINK      AmbientHeatingSink  ambientheatingsink.cpp:36   thermal_energy -= dt*reduction          (the mask is part of the plane)
//   SRC_HEATING   LocalizedHeating    localizedheating.cpp:60     thermal_energy += mask*((dt*ramp)*template)
//   SRC_MASS      MassInjection       massinjection.cpp:49        rho += mask*((dt*template)*m_i)         (E_N then holds rho: raw_rho propagate)
//   SRC_MOMENTUM  MomentumInjection   momentuminjection.cpp:71-72 mom_k += mask*(((dt*(osc*max_accel))*template_k)*rho)
// f = the host-evaluated scalar of this step (dt, dt*ramp, dt, dt*(osc*max_accel)).
// STATUS: written after the round-1 GPU budget was spent; not yet run on a GPU (tests/test_zz_gpu_unvalidated.py).
// ---------------------------------------------------------------------------------------------------------
enum { SRC_SINK = 0, SRC_HEATING = 1, SRC_MASS = 2, SRC_MOMENTUM = 3 };
struct SrcArgs { double *U[NEV]; const double *p0, *p1; int kind; double f; const int *done_ptr; };

__global__ void __launch_bounds__(256) k_source_term(const DomainParams P, const SrcArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny || *A.done_ptr) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double mask = is_interior(P, r, j) ? 1.0 : 0.0;
    if (A.kind == SRC_SINK) A.U[E_E][off] = A.U[E_E][off] - A.f * A.p0[off];
    else if (A.kind == SRC_HEATING) A.U[E_E][off] = A.U[E_E][off] + mask * (A.f * A.p0[off]);
    else if (A.kind == SRC_MASS) A.U[E_N][off] = (A.U[E_N][off] * P.m_i) + mask * ((A.f * A.p0[off]) * P.m_i);
    else {
        const double rho = A.U[E_N][off] * P.m_i;
        A.U[E_MX][off] = A.U[E_MX][off] + mask * ((A.f * A.p0[off]) * rho);
        A.U[E_MY][off] = A.U[E_MY][off] + mask * ((A.f * A.p1[off]) * rho);
    }
}

// ---------------------------------------------------------------------------------------------------------
// DivCleaning (source/modules/solar/divcleaning.cpp:21-47): sub-cycled diffusion of div(b) out of bi_x, bi_y.  The mixed and second
// derivatives are whole-plane passes of k_operator (derivative1D returns zero outside the interior, as the reference's does, so the
// composition d/dx(d/dy b_y) sees zeros in the ghost rows exactly like the reference); this kernel is the update of one component:
//   bi += ((mask*dt_sub) * coeff) * (mixed + second),   coeff = ((1/(1/dx^2 + 1/dy^2))/2)/time_scale      (:24, :36-40)
// STATUS: not yet run on a GPU (tests/test_zz_gpu_unvalidated.py).
// ---------------------------------------------------------------------------------------------------------
struct DcArgs { double *bi; const double *mixed, *second; double dts, time_scale; };
__global__ void __launch_bounds__(256) k_dc_update(const DomainParams P, const DcArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double mask = is_interior(P, r, j) ? 1.0 : 0.0;
    const double dx = P.tx.d[r], dy = P.ty.d[j];
    const double coeff = ((1.0 / (1.0 / (dx * dx) + 1.0 / (dy * dy))) / 2.0) / A.time_scale;
    A.bi[off] = A.bi[off] + ((mask * A.dts) * coeff) * (A.mixed[off] + A.second[off]);
}
__global__ void __launch_bounds__(256) k_plane_sum(const DomainParams P, double *out, const double *a, const double *b)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    out[off] = a[off] + b[off];
}

// ---------------------------------------------------------------------------------------------------------
// FieldHeating (source/modules/solar/fieldheating.cpp:30-58): heating = coeff * (c/4pi |curl b|)^current_pow * |b|^b_pow * n^n_pow / roc^roc_pow,
// roc = 1/max(|(b_hat . grad) b_hat|, 1e-16), evaluated in preIterateModule on the state BEFORE any module iterates (:30-46);
// iterateModule (:48-58) applies thermal_energy += mask*(dt*heating).  pow with run-time exponents: the libm tolerance class.
// Operands: derived planes b_x, b_y, b_hat_x, b_hat_y, b_mag materialised by k_mhd_derive.  STATUS: not yet run on a GPU.
// ---------------------------------------------------------------------------------------------------------
constexpr double kCLight = 29979245800.0;            // C, source/constants.hpp:18
struct FhArgs { const double *bx, *by, *bhx, *bhy, *bmag, *n; double *H; double *e; double coeff, current_pow, b_pow, n_pow, roc_pow, dt; int inactive; };
__global__ void __launch_bounds__(128) k_fh_compute(const DomainParams P, const FhArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    double h = A.coeff;
    if (A.coeff == 0.0) { A.H[off] = 0.0; return; }
    if (A.current_pow != 0.0) {
        auto BX = [&](int a, int b) { return rd(P, A.bx, a, b); };
        auto BY = [&](int a, int b) { return rd(P, A.by, a, b); };
        const double curl = Dx(P, BY, r, j) - Dy(P, BX, r, j);                                  // curl2D, derivs.cpp:472-474
        h = h * pow((kCLight / (4.0 * kPI)) * fabs(curl), A.current_pow);
    }
    if (A.b_pow != 0.0) h = h * pow(A.bmag[off], A.b_pow);
    if (A.n_pow != 0.0) h = h * pow(A.n[off], A.n_pow);
    if (A.roc_pow != 0.0) {
        auto HX = [&](int a, int b) { return rd(P, A.bhx, a, b); };
        auto HY = [&](int a, int b) { return rd(P, A.bhy, a, b); };
        const double hx = A.bhx[off], hy = A.bhy[off];
        const double cx = hx * Dx(P, HX, r, j) + hy * Dy(P, HX, r, j), cy = hx * Dx(P, HY, r, j) + hy * Dy(P, HY, r, j);
        const double roc = 1.0 / smax(sqrt(cx * cx + cy * cy), 1.0e-16);
        h = h / pow(roc, A.roc_pow);
    }
    A.H[off] = h;
}
__global__ void __launch_bounds__(256) k_fh_apply(const DomainParams P, const FhArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double mask = is_interior(P, r, j) ? 1.0 : 0.0;
    const double h = mask * (A.dt * A.H[off]);
    A.H[off] = h;
    if (!A.inactive) A.e[off] = A.e[off] + h;
}

// ---------------------------------------------------------------------------------------------------------
// BoundaryOutflow (source/modules/solar/boundaryoutflow.cpp:39-63, 140-236): an acceleration near one boundary, along x / y or along
// the field, optionally steered (dynamic_mode) towards a target outflow speed by the largest field-aligned outflow found in a
// window next to that boundary.  k_bo_mean: that maximum (std::max semantics: a NaN candidate is never selected), as an order-preserving integer key; k_bo_apply:
//   mom += (dt*(accel*template)) [* b_hat_k] * rho .   b_hat, v are the derived variables of idealmhd.cpp:248-277, formed per cell.
// STATUS: not yet run on a GPU.
// ---------------------------------------------------------------------------------------------------------
struct BoArgs { double *U[NEV]; const double *st[NSTATIC]; const double *tmpl; int xl, xu, yl, yu; /* GLOBAL index window */ int boundary, field_aligned; double dt, accel;
                unsigned long long *max_key; /* running maximum as an order-preserving key (bo_key), so that slabs can be combined by an integer max */ };
// order-preserving map double -> unsigned: a < b  <=>  bo_key(a) < bo_key(b) (no NaN); 0 is below every key
__host__ __device__ inline unsigned long long bo_key(double x)
{
    unsigned long long b;
    memcpy(&b, &x, sizeof(b));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ inline double bo_unkey(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
    double x;
    memcpy(&x, &b, sizeof(x));
    return x;
}
__device__ __forceinline__ void bo_bhat(const DomainParams &P, const BoArgs &A, size_t off, double *hx, double *hy)
{
    const double bx = A.st[S_BEX][off] + A.U[E_BX][off], by = A.st[S_BEY][off] + A.U[E_BY][off], bz = A.st[S_BEZ][off] + A.U[E_BZ][off];
    const double bm = sqrt((bx * bx + by * by) + bz * bz);
    if (bm == 0.0) { *hx = 0.0; *hy = 0.0; } else { *hx = bx / bm; *hy = by / bm; }
}
__global__ void __launch_bounds__(128) k_bo_mean(const DomainParams P, const BoArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int g = P.row0 + r;
    if (j >= P.ny || g < A.xl || g > A.xu || j < A.yl || j > A.yu) return;
    const size_t off = (size_t)r * P.pitch + j;
    double hx, hy;
    bo_bhat(P, A, off, &hx, &hy);
    const double rho = A.U[E_N][off] * P.m_i;
    double cur = hx * (A.U[E_MX][off] / rho) + hy * (A.U[E_MY][off] / rho);
    if (A.boundary == 0 && hx > 0.0) cur *= -1.0;
    else if (A.boundary == 1 && hx < 0.0) cur *= -1.0;
    else if (A.boundary == 3 && hy < 0.0) cur *= -1.0;
    else if (A.boundary == 2 && hy > 0.0) cur *= -1.0;
    if (cur == cur) atomicMax(A.max_key, bo_key(cur));                           // std::max(max, curr) never selects a NaN candidate
}
__global__ void __launch_bounds__(256) k_bo_apply(const DomainParams P, const BoArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double a = A.dt * (A.accel * A.tmpl[off]);                               // dt*accel, accel = curr_accel*accel_template  (:47, :50-61)
    const double rho = A.U[E_N][off] * P.m_i;
    if (A.field_aligned) {
        double hx, hy;
        bo_bhat(P, A, off, &hx, &hy);
        A.U[E_MX][off] = A.U[E_MX][off] + (a * hx) * rho;
        A.U[E_MY][off] = A.U[E_MY][off] + (a * hy) * rho;
    } else if (A.boundary < 2) A.U[E_MX][off] = A.U[E_MX][off] + a * rho;
    else A.U[E_MY][off] = A.U[E_MY][off] + a * rho;
}

// ---------------------------------------------------------------------------------------------------------
// Artificial viscosity (source/modules/viscosity.cpp:185-267): one term  dq = visc_coeff * laplacian(q) * scale_fac
// (+ gradient correction), q = the variable to differentiate of the grid set the RHS is evaluated on (materialised by
// k_mhd_derive when it is a derived variable), timescale from the PRIMARY state's dt plane / its minimum.
// ---------------------------------------------------------------------------------------------------------
struct ViscArgs {
    const double *q;             // variable to differentiate
    const double *n;             // number density of the same grid set (scale factor)
    const double *dt_plane;      // primary dt plane (local / boundary) or nullptr
    const unsigned long long *dt_min_bits;   // primary dt minimum over the dt bounds (global / boundary_global)
    const double *strength_plane;            // boundary profiles, or nullptr
    double strength;
    int scale_mode;              // 0: 1 ; 1: n*m_i (momentum <- velocity) ; 2: n*(K_B/(gamma-1)) (thermal energy <- temperature)
    int gradient_correction;
    int masked;                  // multiply by the ghost-zone mask (RHS form, viscosity.cpp:117)
    double *out;                 // may be null (an evaluation made only for the planes below)
    // the planes Viscosity::fileOutput appends (viscosity.cpp:351-376), each the leftover of the term's last evaluation; null = not kept
    double *lap_out;             // m_grids_lap[i]  (:225)
    double *dtg_out;             // m_grids_dt[i]   (:209-210: the primary dt plane, or its minimum everywhere)
    double *dq_out;              // m_grids_dqdt[i] (:116 masked RHS term; :148 second rk2 evaluation)
};

__global__ void __launch_bounds__(128) k_visc_term(const DomainParams P, const ViscArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double dtm = __longlong_as_double((long long)*A.dt_min_bits);
    auto Q = [&](int a, int b) { return rd(P, A.q, a, b); };
    auto coef = [&](int a, int b) {                                                   // :213
        a = wrap_i(P, a); b = wrap_j(P, b);
        const double str = A.strength_plane ? rd(P, A.strength_plane, a, b) : A.strength;
        const double dtg = A.dt_plane ? rd(P, A.dt_plane, a, b) : dtm;
        const double dx = P.tx.d[a], dy = P.ty.d[b];
        return (((str * 1.0) / (1.0 / (dx * dx) + 1.0 / (dy * dy))) / 2.) / dtg;
    };
    auto scale = [&](int a, int b) {                                                  // :228-258
        if (A.scale_mode == 1) return rd(P, A.n, a, b) * P.m_i;
        if (A.scale_mode == 2) return rd(P, A.n, a, b) * (kKB / (P.gamma - 1));
        return 1.0;
    };
    const double lap = D2x(P, Q, r, j) + D2y(P, Q, r, j);                             // laplacian, derivs.cpp:458-462
    double out = (coef(r, j) * lap) * scale(r, j);                                    // :266
    if (A.gradient_correction) {                                                      // :261-265
        auto CS = [&](int a, int b) { return coef(a, b) * scale(a, b); };
        out = (out + Dx(P, CS, r, j) * Dx(P, Q, r, j)) + Dy(P, CS, r, j) * Dy(P, Q, r, j);
    }
    if (A.masked) out = out * (is_interior(P, r, j) ? 1.0 : 0.0);
    if (A.out) A.out[off] = out;
    if (A.lap_out) A.lap_out[off] = lap;
    if (A.dtg_out) A.dtg_out[off] = A.dt_plane ? A.dt_plane[off] : dtm;
    if (A.dq_out) A.dq_out[off] = out;
}

// grid_to_evol = base + (mask*c)*term   (viscosity.cpp:135,145,150,...) ; rk4 combination when d2..d4 are given (:173-174)
struct AxpyArgs { const double *base; const double *t1, *t2, *t3, *t4; double c; double *out; int base_is_n; double *comb_out; /* rk4: m_grids_dqdt[i] (:172), or null */ };
__global__ void __launch_bounds__(256) k_visc_apply(const DomainParams P, const AxpyArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double mask = is_interior(P, r, j) ? 1.0 : 0.0;
    double t = A.t1[off];
    if (A.t4) t = (((A.t1[off] + A.t2[off] * 2.0) + A.t3[off] * 2.0) + A.t4[off]) / 6.0;
    if (A.t4 && A.comb_out) A.comb_out[off] = t;
    const double b = A.base_is_n ? A.base[off] * P.m_i : A.base[off];
    A.out[off] = b + (mask * A.c) * t;
}

// ---------------------------------------------------------------------------------------------------------
// PlasmaDomain differential operators on a plane (source/mhd/derivs.cpp), for host-side modules that are not ported.
// op: 0 derivative1D, 1 secondDerivative1D, 2 laplacian, 3 transportDerivative1D (needs vel)
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// PhysicalViscosity (source/modules/solar/physicalviscosity.cpp): Braginskii eta_0 viscous heating and force, sub-cycled.
// One kernel per sub-cycle stage.  The force is a derivative of products of first derivatives; everything it needs lives on the
// plus-shaped stencil {(r,j), (r+-1,j), (r,j+-1)}: the six velocity derivatives, T^(5/2) and b_hat are evaluated once at each of
// those five points (derivatives are zero outside the interior) and shared by the 3 x 2 tensor terms.
// ---------------------------------------------------------------------------------------------------------
struct PvArgs {
    const double *v[3], *T, *bh[3], *n, *cg;       // velocity and temperature of this stage; b_hat, n of the primary state; coefficient plane
    double *e, *mom[3];                             // primary thermal_energy / momenta (read; written in the final stage)
    double *v_out[3], *T_out;                       // velocity and temperature after this stage (never alias v, T)
    double coeff, dt, half;                         // half = 0.5 for the RK2 half step (e, mom stay untouched), 1.0 otherwise
    int heating_on, force_on, gc, final_stage;
    unsigned long long *red;                        // count mode: red[0] = min timescale
    int fast;                                       // deep-interior cells take the FAST instances (SPRUCE_FAST_INTERIOR, default on; same results bit for bit)
    // output_to_file (physicalviscosity.cpp:151-152, 166, 170, 218, 222, 292-308): avg[0] = viscous_heating, avg[1..3] = viscous_force_x/y/z -- zeroed by the host before the
    // sub-cycles, every FULL-step stage adds its heating / force divided by the sub-cycle count.  diag_only (inactive_mode): nothing else is written, and the same
    // stage value is added `repeat` times (the state never changes between the reference's sub-cycles then).
    double *avg[4];
    double nsub;
    int diag_only, repeat;
    double *ms_i, *ms_e; double ms_f;               // multispecies_mode: cumulative ion / electron heating planes and the module's electron fraction (:174-177, :226-229); null = off
};
struct PvPoint { double dxv[3], dyv[3], t25, b[3], cg; };

template <bool FAST = false>
__device__ __noinline__ void pv_point(const DomainParams &P, const PvArgs &A, int a, int b, PvPoint &o)
{
#pragma unroll
    for (int k = 0; k < 3; k++) {
        auto V = [&](int x, int y) { return rdT<FAST>(P, A.v[k], x, y); };
        o.dxv[k] = Dx<FAST>(P, V, a, b);
        o.dyv[k] = Dy<FAST>(P, V, a, b);
        o.b[k] = rdT<FAST>(P, A.bh[k], a, b);
    }
    o.t25 = pow25(rdT<FAST>(P, A.T, a, b));
    o.cg = rdT<FAST>(P, A.cg, a, b);
}
// same_factor_off_diag (physicalviscosity.cpp:90-93)
__device__ __forceinline__ double pv_sfo(const PvPoint &p)
{
    return ((p.b[0] * (p.b[1] * p.dyv[0]) + p.b[1] * (p.b[0] * p.dxv[1])) + p.b[2] * (p.b[0] * p.dxv[2] + p.b[1] * p.dyv[2])) - (p.dxv[0] + p.dyv[1]) / 3.0;
}
// derivative1D of a plane whose values at (lower, centre, upper) along one axis are given (derivs.cpp:223-264)
__device__ __forceinline__ double pv_d3(const AxisTab &t, int i, double lo, double c, double hi)
{
    const double up = face_interp(c, hi, t.h[i], t.h[i + 1], t.fs[i + 1], t.rfs[i + 1]);
    const double dn = face_interp(lo, c, t.h[i - 1], t.h[i], t.fs[i], t.rfs[i]);
    return ddiv(up - dn, t.d[i], t.rd[i]);
}

// heating rate and viscous force at one cell (computeHeating :47-64, computeViscousForce :82-140)
template <bool FAST>
__device__ __forceinline__ void pv_cell(const DomainParams &P, const PvArgs &A, int r, int j, bool in, double mask, double &heating, double *force)
{
    PvPoint c;
    pv_point<FAST>(P, A, r, j, c);
    if (A.heating_on && A.coeff != 0.0) {                                                          // computeHeating :47-64
        const double X = ((c.b[0] * (c.b[0] * c.dxv[0] + c.b[1] * c.dyv[0]) + c.b[1] * (c.b[0] * c.dxv[1] + c.b[1] * c.dyv[1]))
                          + c.b[2] * (c.b[0] * c.dxv[2] + c.b[1] * c.dyv[2])) - (c.dxv[0] + c.dyv[1]) / 3.0;
        heating = mask * (((c.cg * 3.0) * c.t25) * (X * X));
    }
    if (A.force_on && in) {                                                                        // computeViscousForce :82-140 (zero outside: mask)
        PvPoint xm, xp, ym, yp;
        pv_point<FAST>(P, A, r - 1, j, xm); pv_point<FAST>(P, A, r + 1, j, xp);
        pv_point<FAST>(P, A, r, j - 1, ym); pv_point<FAST>(P, A, r, j + 1, yp);
        const PvPoint *lo[2] = {&xm, &ym}, *hi[2] = {&xp, &yp};
        const double sfo_c = pv_sfo(c), sfo_lo[2] = {pv_sfo(xm), pv_sfo(ym)}, sfo_hi[2] = {pv_sfo(xp), pv_sfo(yp)};
        const double sfd = c.b[0] * (c.b[0] * c.dxv[0]) + c.b[1] * (c.b[1] * c.dyv[1]);             // same_factor_diag :94-95
        auto BX = [&](int x, int y) { return rdT<FAST>(P, A.bh[0], x, y); };
        auto BY = [&](int x, int y) { return rdT<FAST>(P, A.bh[1], x, y); };
        auto VX = [&](int x, int y) { return rdT<FAST>(P, A.v[0], x, y); };
        auto VY = [&](int x, int y) { return rdT<FAST>(P, A.v[1], x, y); };
        const double dxbx = Dx<FAST>(P, BX, r, j), dybx = Dy<FAST>(P, BX, r, j), dxby = Dx<FAST>(P, BY, r, j), dyby = Dy<FAST>(P, BY, r, j);
        // grad_b_terms_diag :96-101; derivative1D(del_y_v_y,0) and derivative1D(del_x_v_x,1) from the neighbours' derivatives
        const double gbd[2] = {
            ((((c.b[0] * 2.0) * dxbx) * c.dxv[0] + (c.b[0] * c.b[0]) * D2x<FAST>(P, VX, r, j)) + ((c.b[1] * 2.0) * dxby) * c.dyv[1]) + (c.b[1] * c.b[1]) * pv_d3(P.tx, r, xm.dyv[1], c.dyv[1], xp.dyv[1]),
            ((((c.b[0] * 2.0) * dybx) * c.dxv[0] + (c.b[0] * c.b[0]) * pv_d3(P.ty, j, ym.dxv[0], c.dxv[0], yp.dxv[0])) + ((c.b[1] * 2.0) * dyby) * c.dyv[1]) + (c.b[1] * c.b[1]) * D2y<FAST>(P, VY, r, j)};
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
            double res = 0.0;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double delta = (i == jj) ? 1.0 / 3.0 : 0.0;
                const AxisTab &t = i == 0 ? P.tx : P.ty;
                const int idx = i == 0 ? r : j;
                auto G = [&](const PvPoint &p) { return A.gc ? (p.t25 * p.cg) * (delta - p.b[i] * p.b[jj]) : p.t25 * (delta - p.b[i] * p.b[jj]); };
                const double g_lo = G(*lo[i]), g_c = G(c), g_hi = G(*hi[i]);
                res = res + pv_d3(t, idx, g_lo * sfo_lo[i], g_c * sfo_c, g_hi * sfo_hi[i]);                               // :110-113 / :121-124
                res = res + (A.gc ? ((gbd[i] * c.t25) * c.cg) * (delta - c.b[i] * c.b[jj]) : (gbd[i] * c.t25) * (delta - c.b[i] * c.b[jj]));   // :114-115 / :125-126
                res = res + sfd * pv_d3(t, idx, g_lo, g_c, g_hi);                                                          // :116-118 / :127-129
            }
            force[jj] = A.gc ? res * (mask * -3.0) : res * ((c.cg * -3.0) * mask);                                         // :132-136
        }
    }
}

__global__ void __launch_bounds__(128, 4) k_pv_stage(const __grid_constant__ DomainParams P, const __grid_constant__ PvArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const bool in = is_interior(P, r, j);
    const double mask = in ? 1.0 : 0.0;
    const double rho = A.n[off] * P.m_i;
    double heating = 0.0, force[3] = {0.0, 0.0, 0.0};
    if (A.fast && deep_interior(P, r, j)) pv_cell<true>(P, A, r, j, in, mask, heating, force);
    else pv_cell<false>(P, A, r, j, in, mask, heating, force);
    if (A.avg[0] && A.final_stage) {
        const int reps = A.diag_only ? A.repeat : 1;
        if (A.heating_on) {
            double a = A.avg[0][off];
            const double q = heating / A.nsub;
            for (int s = 0; s < reps; s++) a = a + q;
            A.avg[0][off] = a;
        }
        if (A.force_on) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double a = A.avg[1 + k][off];
                const double q = force[k] / A.nsub;
                for (int s = 0; s < reps; s++) a = a + q;
                A.avg[1 + k][off] = a;
            }
        }
    }
    if (A.diag_only) return;
    if (A.ms_i && A.final_stage && A.heating_on) {                                                  // heating applied in this sub-cycle: (1 - f) heating dt_sub to the ions, f ... to the electrons
        if (A.ms_f < 1.0) A.ms_i[off] = A.ms_i[off] + (heating * (1.0 - A.ms_f)) * A.dt;
        if (A.ms_f > 0.0) A.ms_e[off] = A.ms_e[off] + (heating * A.ms_f) * A.dt;
    }
    const double hs = A.half;
    double e1 = A.e[off], T1 = A.T[off];
    if (A.heating_on) {                                                                            // :176-185 / :214-218
        e1 = smax(e1 + (hs == 1.0 ? heating * A.dt : (heating * 0.5) * A.dt), P.e_min);
        T1 = smax((e1 * P.gm1) / (A.n[off] * (2.0 * kKB)), P.T_min);
    }
    A.T_out[off] = T1;
    if (A.final_stage && A.heating_on) A.e[off] = e1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double v1 = A.v[k][off];
        if (A.force_on) {                                                                          // :186-193 / :219-226
            const double m1 = A.mom[k][off] + (hs == 1.0 ? force[k] * A.dt : (force[k] * 0.5) * A.dt);
            v1 = m1 / rho;
            if (A.final_stage) A.mom[k][off] = m1;
        }
        A.v_out[k][off] = v1;
    }
}

// computeViscousSubcycles (physicalviscosity.cpp:66-80): min over the interior of dx dy rho / (3 f coeff T^2.5), f = 3 cos^2 + 1
__global__ void __launch_bounds__(128) k_pv_count(const DomainParams P, const PvArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double ts = 1.7976931348623157e308;
    if (j < P.ny && is_interior(P, r, j)) {
        const size_t off = (size_t)r * P.pitch + j;
        const double vx = A.v[0][off], vy = A.v[1][off];
        const double vm = sqrt(vx * vx + vy * vy);
        const double s_ = (A.bh[0][off] * vx) / vm + (A.bh[1][off] * vy) / vm;
        const double df = (vm == 0.0) ? 4.0 : (s_ * s_) * 3.0 + 1.0;
        ts = ((P.tx.d[r] * P.ty.d[j]) * (A.n[off] * P.m_i)) / (((df * 3.0) * smax(A.cg[off], 0.000001 * A.coeff)) * pow25(A.T[off]));
    }
    block_min_to_global(ts, A.red);
}

struct OpArgs { const double *q, *vel; double *out; int op, index; };
__global__ void __launch_bounds__(128) k_operator(const DomainParams P, const OpArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    auto Q = [&](int a, int b) { return rd(P, A.q, a, b); };
    double out = 0.0;
    if (A.op == 0) out = A.index == 0 ? Dx(P, Q, r, j) : Dy(P, Q, r, j);
    else if (A.op == 1) out = A.index == 0 ? D2x(P, Q, r, j) : D2y(P, Q, r, j);
    else if (A.op == 2) out = D2x(P, Q, r, j) + D2y(P, Q, r, j);
    else if (is_interior(P, r, j)) {
        // transportDerivative1D (derivs.cpp:122-162): (S[f+1]*vf[f+1] - S[f]*vf[f]) / d
        auto V = [&](int a, int b) { return rd(P, A.vel, a, b); };
        const AxisTab &t = A.index == 0 ? P.tx : P.ty;
        const int i0 = A.index == 0 ? r : j;
        double flux[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int f = i0 + k;
            auto at = [&](int i) { return A.index == 0 ? Q(i, j) : Q(r, i); };
            auto vat = [&](int i) { return A.index == 0 ? V(i, j) : V(r, i); };
            const FaceGeom g = load_face_geom(t, f);
            const double vf = face_interp(vat(f - 1), vat(f), g.hm1, g.h0, g.fs, g.rfs);
            double d2;
            const double S = upwind_face(at(f - 2), at(f - 1), at(f), at(f + 1), vf, g, &d2);
            flux[k] = S * vf;
        }
        out = ddiv(flux[1] - flux[0], t.d[i0], t.rd[i0]);
    }
    A.out[off] = out;
}

}  // namespace spruce
