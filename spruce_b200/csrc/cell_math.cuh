// cell_math.cuh -- the domain parameters and the per-cell scalar functions of the ideal-MHD kernels that touch no thread state: the density
// floor round trip, recomputeDT for one cell, and the division-free test that lets the stage kernel skip cells whose dt cannot be the minimum.
// Split out of mhd_kernels.cuh so that tests/hostcheck/cell_math_check.cpp can compile them for the host (with exact_math.cuh) and check
// cell_dt against the reference's formula and the soundness of dt_can_skip on random states without a GPU.
#pragma once
#include "exact_math.cuh"

namespace spruce {

// 1-D cell-size tables of one axis; pointers are pre-offset so that index -TAB_APRON is the first element.
struct AxisTab {
    const double *h;     // 0.5*d
    const double *fs;    // fs[f] = h[f] + h[f-1]
    const double *rfs;   // RN(1/fs[f])
    const double *ep;    // ep[f] = h[f-2] + 2*h[f-1]
    const double *em;    // em[f] = h[f+1] + 2*h[f]
    const double *d;     // d
    const double *rd;    // RN(1/d)
};

struct DomainParams {
    int nx, ny, pitch;            // local rows, columns, doubles per row
    int gnx, row0;                // global xdim, global index of local row 0
    int xl, xu, yl, yu;           // GLOBAL interior bounds (computeIterationBounds, plasmadomain.cpp:138-161)
    int xper, yper;               // both sides periodic along that axis
    int xwrap;                    // single-rank periodic x: wrap the local row index
    int bc_x1, bc_x2, bc_y1, bc_y2;
    double m_i, rm_i;             // ion mass, RN(1/m_i)
    double gamma, gm1;            // adiabatic index, gamma - 1.0
    double n_min, T_min, e_min;
    double fourpi, rfourpi;       // 4.0*PI, RN(1/(4.0*PI))
    double epsilon;
    AxisTab tx, ty;               // tx indexed by LOCAL row
};

// enforceMinimums + recomputeDerivedVarsFromEvolvedVars for rho (idealmhd.cpp:237,246-247). Returns n; *r1 = post-floor rho.
__device__ __forceinline__ double density_floor(const DomainParams &P, double rho_u, double *r1)
{
    const double n1 = smax(ddiv(rho_u, P.m_i, P.rm_i), P.n_min);
    const double rr = n1 * P.m_i;
    *r1 = rr;
    return smax(ddiv(rr, P.m_i, P.rm_i), P.n_min);
}

// recomputeDT for one cell (idealmhd.cpp:279-304)
__device__ __forceinline__ double cell_dt(const DomainParams &P, double rho, double mx, double my, double e,
                                          double bx, double by, double bz, double dx, double rdx, double dy, double rdy)
{
    // three quotients share the divisor rho: one IEEE reciprocal + the exact-division correction each (exact_math.cuh)
    const double rr = 1.0 / rho;
    const double vx = ddiv(mx, rho, rr), vy = ddiv(my, rho, rr);
    const double p = e * P.gm1;
    const double bm = sqrt((bx * bx + by * by) + bz * bz);
    const double cs = sqrt(ddiv(p * P.gamma, rho, rr));
    const double cs2 = cs * cs;
    const double va = bm / sqrt(rho * P.fourpi);
    const double va2 = va * va;
    const double s = cs2 + va2;
    const double delta = sqrt(1.0 - ((cs2 * 4.0) * va2) / (s * s));
    const double vfast = sqrt((s * 0.5) * (1.0 + delta));
    const double vslow = sqrt((s * 0.5) * (1.0 - delta));
    // sqrt(RN(v*v)) == |v| in binary64 round-to-nearest whenever v*v neither underflows nor overflows (Boldo 2015); zero maps to zero
    const double ax = fabs(vx), ay = fabs(vy);
    const double vmx = (ax == 0.0 || (ax > 1.0e-140 && ax < 1.0e140)) ? ax : sqrt(vx * vx);
    const double vmy = (ay == 0.0 || (ay > 1.0e-140 && ay < 1.0e140)) ? ay : sqrt(vy * vy);
    const double M = smax(smax(smax(cs, va), vfast), vslow);
    return 1.0 / (ddiv(vmx + M, dx, rdx) + ddiv(vmy + M, dy, rdy));
}

// Only the MINIMUM of dt over the cells is ever used (evolution.cpp:62).  A cell whose dt is certainly above thr = F * (previous
// global minimum) cannot be the new minimum as long as the new minimum turns out <= thr (k_dt_validate checks that afterwards and
// k_dt_full re-evaluates every cell when it does not hold).  Certain means: with M <= sqrt(c_s^2 + v_A^2) (v_fast^2 = s/2 (1+delta) <= s)
//   1/dt <= (|v_x| + sqrt(s))/dx + (|v_y| + sqrt(s))/dy  <  R = 1/thr
// which, multiplied through by rho and squared, needs neither a division nor a square root:
//   t = R rho - (|m_x|/dx + |m_y|/dy) > 0   and   (gamma p + b^2/4pi) rho (1/dx + 1/dy)^2 (1 + 1e-3) < t^2 .
// The 1e-3 margin covers every rounding in this test by twelve orders of magnitude; NaN or underflow makes the test fail (= evaluate).
__device__ __forceinline__ bool dt_can_skip(const DomainParams &P, double R, double rho, double mx, double my, double e,
                                            double bx, double by, double bz, double rdx, double rdy)
{
    const double t = R * rho - (fabs(mx) * rdx + fabs(my) * rdy);
    const double S = (e * P.gm1) * P.gamma + ((bx * bx + by * by) + bz * bz) * P.rfourpi;
    const double g = rdx + rdy;
    return (t > 0.0) && (((S * rho) * (g * g)) * 1.001 < t * t);
}

}  // namespace spruce
