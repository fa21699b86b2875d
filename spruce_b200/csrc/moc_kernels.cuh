// moc_kernels.cuh -- method-of-characteristics open boundary (`open_moc`): the ghost cells of such a side are EVOLVED, by
// the characteristic form of the ideal-MHD equations, instead of being overwritten by a boundary pass.
//
// Replaces   IdealMHD::computeTimeDerivativesCharacteristicBoundary / singleBoundaryTermsMOC   source/equationsets/idealmhd.cpp:306-615
//            PlasmaDomain::characteristicBartonDerivative1D (+ upwindSurfaceDirected)           source/mhd/derivs.cpp:77-117, 167-214
//            PlasmaDomain::derivative1DBackward / secondDerivative1DBackward                    source/mhd/derivs.cpp:322-404
//            the ranged derivative1D / transportDerivative1D / secondDerivative1D calls          source/mhd/derivs.cpp:122-162, 223-264, 417-455
//
// The reference evaluates ~45 whole-plane temporaries per side; every one of them is non-zero only inside the two ghost layers of
// that side, and each value there depends on a plus-shaped neighbourhood (3 cells towards the interior, +-2 cells along the side).
// Here ONE thread owns one ghost cell and evaluates all of it in registers, operation for operation in the reference's order
// (every +,-,*,/ and sqrt individually rounded: this file is compiled with -fmad=false like the rest of the library and uses plain
// IEEE division -- the strips are O(perimeter) work, speed is irrelevant next to the stage kernel).  std::pow(x, 2.0) of the
// reference (idealmhd.cpp:377) is x*x here: glibc's pow is correctly rounded for all but ~1e-5 of arguments, so this path is held
// to the "libm" tolerance class of DESIGN.md section 2 (<= 1e-9), although it is bit-identical in every case tested so far.
//
// The per-cell arithmetic (namespace spruce::moc, everything up to moc_cell_terms) is plain C++ with no CUDA dependence so that
// tests/hostcheck/moc_host_check.cpp can compile THE SAME SOURCE with g++ and check it, bit for bit, on a machine without a GPU.
#pragma once
#include <cmath>
#include <cstddef>

#if defined(__CUDACC__)
#define MOC_HD __host__ __device__ inline
#else
#define MOC_HD static inline
#endif

namespace spruce {
namespace moc {

constexpr int NG = 2;                                  // N_GHOST
constexpr double kPi = 3.14159265358979323846;         // source/constants.hpp:16
enum { MBC_PERIODIC = 0, MBC_OPEN_MOC = 4 };

// The state the characteristic terms are evaluated on.  Planes are row-major, `pitch` doubles per row, and are indexed by the GLOBAL row
// (x) index: on a slab the caller passes plane pointers shifted back by row0 rows (and dx shifted by row0 entries), so that only rows the
// slab holds -- its own plus two halo rows on each side -- are ever touched.  Plane `n` holds the number density (rho = n*m_i, idealmhd.cpp:247).
struct Field {
    const double *n, *mx, *my, *mz, *e, *bix, *biy, *biz;
    const double *bex, *bey, *bez, *gx, *gy;
    const double *dx, *dy;                             // cell sizes: dx[i], dy[j]
    int nx, ny, pitch;                                 // GLOBAL extent
    int x_halo;                                        // slab of a periodic-x domain: the rows -2,-1 / nx,nx+1 are resident halo rows, x indices are not wrapped
    int bc[4];                                         // x1, x2, y1, y2
    double m_i, gamma, gm1;
    double visc;                                       // global_visc_coeff (idealmhd.cpp:90)
};

MOC_HD double smin_(double a, double b) { return (b < a) ? b : a; }      // std::min / std::max (NaN behaviour: SURVEY Q22)
MOC_HD double smax_(double a, double b) { return (a < b) ? b : a; }

// One side: (a, b) = (index along the boundary normal, index along the side).
struct Side {
    int bidx, lower;          // 0: an x side (normal = x), 1: a y side; lower: the side at index 0
    int N, M;                 // extent of the normal / parallel axis
    int ppar;                 // parallel axis periodic
    int Flo, Fhi;             // parallel range of the evolved ghost cells            (x_idx / y_idx of the reference)
    int Ilo, Ihi;             //  ... of the part that does not overlap another ghost zone (xi / yi)
    int Plo, Phi;             //  ... "padded" range of the central alternative gradients  (xp / yp)
    int LClo, LChi, UClo, UChi;   // corner ranges of the one-sided alternative gradients  (xlc / xuc ...)
    int alo, ahi;             // normal range: the two ghost layers
};

// idealmhd.cpp:334-372
MOC_HD Side make_side(const Field &F, int s)
{
    Side S;
    S.bidx = s / 2; S.lower = (s % 2) == 0;
    S.N = S.bidx == 0 ? F.nx : F.ny; S.M = S.bidx == 0 ? F.ny : F.nx;
    const int lo_bc = S.bidx == 0 ? F.bc[2] : F.bc[0], hi_bc = S.bidx == 0 ? F.bc[3] : F.bc[1];
    S.alo = S.lower ? 0 : S.N - NG; S.ahi = S.lower ? NG - 1 : S.N - 1;
    S.ppar = (lo_bc == MBC_PERIODIC);
    const int f0 = 0, f1 = S.M - 1;
    S.Flo = S.Ilo = S.Plo = S.LClo = S.UClo = f0;
    S.Fhi = S.Ihi = S.Phi = S.LChi = S.UChi = f1;
    if (!S.ppar) {
        S.Ilo = f0 + NG; S.Ihi = f1 - NG; S.Plo = f0 + 1; S.Phi = f1 - 1;
        if (lo_bc != MBC_OPEN_MOC) { S.Flo = S.Ilo; S.Plo = S.Ilo; }
        if (hi_bc != MBC_OPEN_MOC) { S.Fhi = S.Ihi; S.Phi = S.Ihi; }
        S.LClo = S.Flo; S.LChi = S.Plo - 1; S.UClo = S.Phi + 1; S.UChi = S.Fhi;
    }
    return S;
}

struct Ctx { const Field *F; Side S; };

MOC_HD size_t cidx(const Ctx &c, int a, int b) { return c.S.bidx == 0 ? (size_t)a * c.F->pitch + b : (size_t)b * c.F->pitch + a; }
MOC_HD double dn(const Ctx &c, int a) { return c.S.bidx == 0 ? c.F->dx[a] : c.F->dy[a]; }
MOC_HD double dp(const Ctx &c, int b) { return c.S.bidx == 0 ? c.F->dy[b] : c.F->dx[b]; }
MOC_HD int wrapb(const Ctx &c, int b) { return (c.S.ppar && !(c.S.bidx == 1 && c.F->x_halo)) ? (b + c.S.M) % c.S.M : b; }

enum Role { R_RHO = 0, R_EN, R_PRESS, R_VPERP, R_VPARA, R_VGUIDE, R_BEPERP, R_BEPARA, R_BEGUIDE, R_BIPERP, R_BIPARA, R_BIGUIDE,
            R_BPERP, R_RHOVPARA, R_RHOVGUIDE, R_RHOVPERP };

// derived variables of one cell, exactly as recomputeDerivedVarsFromEvolvedVars forms them (idealmhd.cpp:241-277)
MOC_HD double val(const Ctx &c, int role, int a, int b)
{
    const Field &F = *c.F;
    const size_t k = cidx(c, a, b);
    const bool xs = c.S.bidx == 0;
    switch (role) {
    case R_RHO: return F.n[k] * F.m_i;
    case R_EN: return F.e[k];
    case R_PRESS: return F.e[k] * F.gm1;
    case R_VPERP: return (xs ? F.mx[k] : F.my[k]) / (F.n[k] * F.m_i);
    case R_VPARA: return (xs ? F.my[k] : F.mx[k]) / (F.n[k] * F.m_i);
    case R_VGUIDE: return F.mz[k] / (F.n[k] * F.m_i);
    case R_BEPERP: return xs ? F.bex[k] : F.bey[k];
    case R_BEPARA: return xs ? F.bey[k] : F.bex[k];
    case R_BEGUIDE: return F.bez[k];
    case R_BIPERP: return xs ? F.bix[k] : F.biy[k];
    case R_BIPARA: return xs ? F.biy[k] : F.bix[k];
    case R_BIGUIDE: return F.biz[k];
    case R_BPERP: return xs ? F.bex[k] + F.bix[k] : F.bey[k] + F.biy[k];
    case R_RHOVPARA: { const double rho = F.n[k] * F.m_i; return rho * ((xs ? F.my[k] : F.mx[k]) / rho); }
    case R_RHOVGUIDE: { const double rho = F.n[k] * F.m_i; return rho * (F.mz[k] / rho); }
    default: { const double rho = F.n[k] * F.m_i; return rho * ((xs ? F.mx[k] : F.my[k]) / rho); }
    }
}

// boundaryInterpolate / boundaryExtrapolate (derivs.cpp:477-499) between two cells of one line; da, db = half cell sizes
MOC_HD double interp2(double qa, double qb, double da, double db) { return (qa * db + qb * da) / (db + da); }
MOC_HD double extrap2(double qa, double qb, double da, double db) { return qa + (qb - qa) * (da + 2.0 * db) / (da + db); }
MOC_HD double interp_n(const Ctx &c, int role, int a1, int a2, int b) { return interp2(val(c, role, a1, b), val(c, role, a2, b), 0.5 * dn(c, a1), 0.5 * dn(c, a2)); }
MOC_HD double extrap_n(const Ctx &c, int role, int a1, int a2, int b) { return extrap2(val(c, role, a1, b), val(c, role, a2, b), 0.5 * dn(c, a1), 0.5 * dn(c, a2)); }
MOC_HD double interp_p(const Ctx &c, int role, int a, int b1, int b2) { return interp2(val(c, role, a, b1), val(c, role, a, b2), 0.5 * dp(c, b1), 0.5 * dp(c, b2)); }
MOC_HD double extrap_p(const Ctx &c, int role, int a, int b1, int b2) { return extrap2(val(c, role, a, b1), val(c, role, a, b2), 0.5 * dp(c, b1), 0.5 * dp(c, b2)); }

// upwindSurfaceDirected (derivs.cpp:77-117): the face value ghost cell `ac` contributes -- upwinded from the interior side
MOC_HD double char_face(const Ctx &c, int role, int ac, int b)
{
    const int vs = c.S.lower ? -1 : 1;
    const int a1 = ac - vs, a0 = ac - 2 * vs;
    const double d2 = interp_n(c, role, a1, ac, b), d3 = val(c, role, a1, b), d1 = extrap_n(c, role, a0, a1, b);
    return (val(c, role, ac, b) <= d3) ? smin_(d3, smax_(d1, d2)) : smax_(d3, smin_(d1, d2));
}
// characteristicBartonDerivative1D (derivs.cpp:167-214) at ghost cell (a, b)
MOC_HD double cbd(const Ctx &c, int role, int a, int b)
{
    const int N = c.S.N;
    double lo, hi;                               // surf[a], surf[a + 1]
    if (!c.S.lower) {                            // positive_forward: cell ac writes surf[ac]; surf[N] closes the line
        lo = char_face(c, role, a, b);
        if (a == N - 1) { const double q = val(c, role, N - 1, b); hi = q + (q - lo); }
        else hi = char_face(c, role, a + 1, b);
    } else {                                     // cell ac writes surf[ac + 1]; surf[0] closes the line
        hi = char_face(c, role, a, b);
        if (a == 0) { const double q = val(c, role, 0, b); lo = q - (hi - q); }
        else lo = char_face(c, role, a - 1, b);
    }
    return (hi - lo) / dn(c, a);
}
// derivative1D along the side (derivs.cpp:223-264)
MOC_HD double d1_p(const Ctx &c, int role, int a, int b)
{
    const int b0 = wrapb(c, b - 1), b2 = wrapb(c, b + 1);
    return (interp_p(c, role, a, b, b2) - interp_p(c, role, a, b0, b)) / dp(c, b);
}
// derivative1DBackward along the side (derivs.cpp:322-359)
MOC_HD double d1_back_p(const Ctx &c, int role, int forward, int a, int b)
{
    const int b0 = wrapb(c, forward ? b - 1 : b + 1);
    const double denom = 0.5 * (dp(c, b) + dp(c, b0));
    return (forward ? 1.0 : -1.0) * (val(c, role, a, b) - val(c, role, a, b0)) / denom;
}
// the three-range sum of the alternative gradients (idealmhd.cpp:444-470): central over the padded range, one-sided in the corners
MOC_HD double d3sum(const Ctx &c, int role, int a, int b)
{
    const Side &S = c.S;
    const double u1 = (b >= S.Plo && b <= S.Phi) ? d1_p(c, role, a, b) : 0.0;
    const double u2 = (b >= S.UClo && b <= S.UChi) ? d1_back_p(c, role, 1, a, b) : 0.0;
    const double u3 = (b >= S.LClo && b <= S.LChi) ? d1_back_p(c, role, 0, a, b) : 0.0;
    return (u1 + u2) + u3;
}
MOC_HD double d3sum_acc(const Ctx &c, double acc, int role, int a, int b)
{
    const Side &S = c.S;
    const double u1 = (b >= S.Plo && b <= S.Phi) ? d1_p(c, role, a, b) : 0.0;
    const double u2 = (b >= S.UClo && b <= S.UChi) ? d1_back_p(c, role, 1, a, b) : 0.0;
    const double u3 = (b >= S.LClo && b <= S.LChi) ? d1_back_p(c, role, 0, a, b) : 0.0;
    return ((acc + u1) + u2) + u3;
}
// secondDerivative1DBackward along the normal (derivs.cpp:361-404)
MOC_HD double d2_back_n(const Ctx &c, int role, int a, int b)
{
    const int pf = !c.S.lower;
    const int a2 = a, a1 = pf ? a2 - 1 : a2 + 1, a0 = pf ? a1 - 1 : a1 + 1;
    const double d01 = 0.5 * (dn(c, a1) + dn(c, a0)), d12 = 0.5 * (dn(c, a1) + dn(c, a2));
    const int h12 = a1 > a2 ? a1 : a2, l12 = a1 < a2 ? a1 : a2, h10 = a1 > a0 ? a1 : a0, l10 = a1 < a0 ? a1 : a0;
    return (pf ? 1.0 : -1.0) * ((val(c, role, h12, b) - val(c, role, l12, b)) / d12 - (val(c, role, h10, b) - val(c, role, l10, b)) / d01) / d12;
}
// upwindSurface along the side, transport velocity v_para (derivs.cpp:10-73): face f lies between cells f-1 and f
MOC_HD double up_face_p(const Ctx &c, int role, int a, int f)
{
    const int f2 = wrapb(c, f), f1 = wrapb(c, f - 1), f0 = wrapb(c, f - 2), f3 = wrapb(c, f + 1);
    const double d2 = interp_p(c, role, a, f1, f2);
    const double vf = interp_p(c, R_VPARA, a, f1, f2);
    const double qc = val(c, role, a, f2), qm = val(c, role, a, f1);
    if (vf > 0.0) { const double d1 = extrap_p(c, role, a, f0, f1); return (qc <= qm) ? smin_(qm, smax_(d1, d2)) : smax_(qm, smin_(d1, d2)); }
    if (vf < 0.0) { const double d1 = extrap_p(c, role, a, f3, f2); return (qc <= qm) ? smax_(qc, smin_(d1, d2)) : smin_(qc, smax_(d1, d2)); }
    return d2;
}
// transportDerivative1D along the side (derivs.cpp:122-162)
MOC_HD double td_p(const Ctx &c, int role, int a, int b)
{
    const int b0 = wrapb(c, b - 1), b2 = wrapb(c, b + 1);
    return (up_face_p(c, role, a, b2) * interp_p(c, R_VPARA, a, b, b2) - up_face_p(c, role, a, b) * interp_p(c, R_VPARA, a, b0, b)) / dp(c, b);
}
// secondDerivative1D along the side (derivs.cpp:417-455)
MOC_HD double d2_p(const Ctx &c, int role, int a, int b)
{
    const int b0 = wrapb(c, b - 1), b2 = wrapb(c, b + 1);
    const double h = 0.5 * dp(c, b);
    return (interp_p(c, role, a, b, b2) - 2.0 * val(c, role, a, b) + interp_p(c, role, a, b0, b)) / (h * h);
}

// singleBoundaryTermsMOC at one ghost cell (a, b) of side S (idealmhd.cpp:334-615): adds normal + parallel terms to
// res[0..7] = d/dt of  rho, v_x, v_y, v_z, thermal_energy, b_x, b_y, b_z
MOC_HD void side_terms(const Ctx &c, int a, int b, double *res)
{
    const Field &F = *c.F;
    const Side &S = c.S;
    const size_t k = cidx(c, a, b);
    const bool xs = S.bidx == 0;
    const double PI = kPi;
    const double imask = (b >= S.Ilo && b <= S.Ihi) ? 1.0 : 0.0;
    const double rho = val(c, R_RHO, a, b), en = val(c, R_EN, a, b), press = val(c, R_PRESS, a, b);
    const double v_perp = val(c, R_VPERP, a, b), v_para = val(c, R_VPARA, a, b), v_guide = val(c, R_VGUIDE, a, b);
    const double b_perp = val(c, R_BEPERP, a, b) + val(c, R_BIPERP, a, b), b_para = val(c, R_BEPARA, a, b) + val(c, R_BIPARA, a, b),
                 b_guide = val(c, R_BEGUIDE, a, b) + val(c, R_BIGUIDE, a, b);
    const double bh = sqrt(b_para * b_para + b_guide * b_guide);
    const double grav_perp = xs ? F.gx[k] : F.gy[k], grav_para = xs ? F.gy[k] : F.gx[k];
    // b_mag of the derived variables: x, y, z order whatever the side (idealmhd.cpp:268)
    const double bx = F.bex[k] + F.bix[k], by = F.bey[k] + F.biy[k], bz = F.bez[k] + F.biz[k];
    const double b_mag = sqrt((bx * bx + by * by) + bz * bz);

    // ---- characteristic speeds :374-395
    double s_perp = b_perp > 0.0 ? 1.0 : -1.0;
    const double R_para = b_para / bh, R_guide = b_guide / bh;
    const double c_s_sq = F.gamma * press / rho;
    const double c_a_sq = b_mag * b_mag / (4.0 * PI * rho);
    const double c_perp_sq = b_perp * b_perp / (4.0 * PI * rho);
    const double t_ = c_a_sq + c_s_sq;
    const double disc = sqrt(fabs((t_ * t_) / 4.0 - c_perp_sq * c_s_sq));
    const double c_plus_sq = (c_a_sq + c_s_sq) / 2.0 + disc, c_minus_sq = (c_a_sq + c_s_sq) / 2.0 - disc;
    const double ap = (c_s_sq - c_minus_sq) / (c_plus_sq - c_minus_sq), am = (c_plus_sq - c_s_sq) / (c_plus_sq - c_minus_sq);
    const double ap_sq = (ap < 0.0) ? 0.0 : ap, am_sq = (am < 0.0) ? 0.0 : am;                 // std::max(x, 0.0)
    const double c_s = sqrt(c_s_sq), c_perp = sqrt(c_perp_sq), c_plus = sqrt(c_plus_sq), c_minus = sqrt(c_minus_sq);
    const double a_plus = sqrt(ap_sq), a_minus = sqrt(am_sq);
    if (b_perp == 0.0) s_perp = 0.0;

    // ---- characteristic derivatives along the normal :397-443
    const double b_para_grad = cbd(c, R_BEPARA, a, b) + cbd(c, R_BIPARA, a, b);
    const double b_guide_grad = cbd(c, R_BEGUIDE, a, b) + cbd(c, R_BIGUIDE, a, b);
    const double vpg = cbd(c, R_VPARA, a, b), vng = cbd(c, R_VPERP, a, b), vgg = cbd(c, R_VGUIDE, a, b), pg = cbd(c, R_PRESS, a, b);
    const double bpg = b_para_grad, bgg = b_guide_grad;
    double d1 = v_perp * cbd(c, R_BPERP, a, b);
    double d2 = v_perp * ((en + press) * cbd(c, R_RHO, a, b) - rho * cbd(c, R_EN, a, b));
    const double sgn = xs ? 1.0 : -1.0;
    const double sq = sqrt(4.0 * PI * rho);
    const double sp = s_perp, Rp = R_para, Rg = R_guide;
    double d3 = ((sp * sgn) * (v_perp + c_perp)) * (((((-sp) * Rg) * vpg + (sp * Rp) * vgg) + (Rg / sq) * bpg) - (Rp / sq) * bgg);
    double d4 = ((sp * sgn) * (v_perp - c_perp)) * (((((sp * Rg) * vpg) - (sp * Rp) * vgg) + (Rg / sq) * bpg) - (Rp / sq) * bgg);
    const double T1 = (a_plus / rho) * pg, T2 = (((sp * Rp) * c_minus) * a_minus) * vpg, T3 = (((sp * Rg) * c_minus) * a_minus) * vgg,
                 T4 = (c_plus * a_plus) * vng, T5 = (((Rp * c_s) * a_minus) / sq) * bpg, T6 = (((Rg * c_s) * a_minus) / sq) * bgg;
    double d5 = (v_perp + c_plus) * (((((T1 - T2) - T3) + T4) + T5) + T6);
    double d6 = (v_perp - c_plus) * (((((T1 + T2) + T3) - T4) + T5) + T6);
    const double A1 = (a_minus / rho) * pg, A2 = (((sp * Rp) * c_plus) * a_plus) * vpg, A3 = (((sp * Rg) * c_plus) * a_plus) * vgg,
                 A4 = (c_minus * a_minus) * vng, A5 = (((Rp * c_s) * a_plus) / sq) * bpg, A6 = (((Rg * c_s) * a_plus) / sq) * bgg;
    double d7 = (v_perp + c_minus) * (((((A1 + A2) + A3) + A4) - A5) - A6);
    double d8 = (v_perp - c_minus) * (((((A1 - A2) - A3) - A4) - A5) - A6);

    // ---- alternative (incoming-wave) amplitudes :444-500
    const double bperp_grad_alt = d3sum_acc(c, d3sum(c, R_BEPERP, a, b), R_BIPERP, a, b);
    const double bperp_sq_grad_alt = (2.0 * b_perp) * bperp_grad_alt;
    const double press_grad_alt = d3sum(c, R_PRESS, a, b);
    const double bguide_grad_alt = d3sum_acc(c, d3sum(c, R_BEGUIDE, a, b), R_BIGUIDE, a, b);
    const double pb = press_grad_alt + bperp_sq_grad_alt / (8.0 * PI);
    const double d3a = ((sp * sp) * sgn) * ((((-1.0 * imask) * Rg) * grav_para) + ((Rg * pb) + (bh * bguide_grad_alt) / (4.0 * PI)) / rho);
    const double first = (imask * grav_perp) + (b_para * bperp_grad_alt) / (4.0 * PI * rho);
    const double second = (((-imask) * Rp) * grav_para) + (Rp * pb) / rho;
    const double d5a = (c_plus * a_plus) * first + ((c_minus * a_minus) * sp) * second;
    const double d7a = (c_minus * a_minus) * first - ((c_plus * a_plus) * sp) * second;
    const double inflow = S.lower ? 1.0 : -1.0;
    if (inflow * v_perp > 0.0) { d1 = 0.0; d2 = 0.0; }
    if (inflow * (v_perp + c_perp) > 0.0) d3 = d3a;
    if (inflow * (v_perp - c_perp) > 0.0) d4 = -d3a;
    if (inflow * (v_perp + c_plus) > 0.0) d5 = d5a;
    if (inflow * (v_perp - c_plus) > 0.0) d6 = -d5a;
    if (inflow * (v_perp + c_minus) > 0.0) d7 = d7a;
    if (inflow * (v_perp - c_minus) > 0.0) d8 = -d7a;

    // output slots: rho 0, v_x 1, v_y 2, v_z 3, e 4, b_x 5, b_y 6, b_z 7
    const int s_vperp = xs ? 1 : 2, s_vpara = xs ? 2 : 1, s_bperp = xs ? 5 : 6, s_bpara = xs ? 6 : 5;
    double n0[8], p1[8];
    // ---- normal terms :534-552
    {
        const double visc = F.visc;
        const double sd_para = d2_back_n(c, R_VPARA, a, b), sd_guide = d2_back_n(c, R_VGUIDE, a, b), sd_perp = d2_back_n(c, R_VPERP, a, b);
        const double cs2 = c_s_sq;
        const double d56 = d5 + d6, d78 = d7 + d8;
        n0[0] = (-((((F.gamma / rho) * d2) + ((0.5 * rho) * a_plus) * d56) + ((0.5 * rho) * a_minus) * d78)) / cs2;
        const double cm = (c_minus * a_minus) / cs2, cp = (c_plus * a_plus) / cs2;
        n0[s_vpara] = ((-0.5 * sp) * (((sgn * Rg) * ((-d3) + d4)) + (((cm * Rp) * ((-d5) + d6)) + ((cp * Rp) * (d7 - d8))))) + visc * sd_para;
        n0[3] = ((-0.5 * sp) * (((sgn * Rp) * (d3 - d4)) + (((cm * Rg) * ((-d5) + d6)) + ((cp * Rg) * (d7 - d8))))) + visc * sd_guide;
        n0[s_vperp] = ((-0.5 / cs2) * (((c_plus * a_plus) * (d5 - d6)) + ((c_minus * a_minus) * (d7 - d8)))) + visc * sd_perp;
        const double ep = 0.5 * (en + press);
        n0[4] = (-(((ep * a_plus) * d56) + ((ep * a_minus) * d78))) / cs2;
        const double sr = -sqrt(PI * rho), d34 = d3 + d4;
        n0[s_bpara] = sr * ((((sgn * Rg) * d34) + (((a_minus / c_s) * Rp) * d56)) - (((a_plus / c_s) * Rp) * d78));
        n0[7] = sr * (((((sgn * -1.0) * Rp) * d34) + (((a_minus / c_s) * Rg) * d56)) - (((a_plus / c_s) * Rg) * d78));
        n0[s_bperp] = -d1;
    }
    // ---- parallel terms :555-604 (only where the ghost zone does not overlap another one)
    if (b >= S.Ilo && b <= S.Ihi) {
        const double visc = F.visc;
        const double td_rho = td_p(c, R_RHO, a, b);
        p1[0] = -td_rho;
        const double dpara = d1_p(c, R_BEPARA, a, b) + d1_p(c, R_BIPARA, a, b);
        const double dperp = d1_p(c, R_BEPERP, a, b) + d1_p(c, R_BIPERP, a, b);
        const double dguide = d1_p(c, R_BEGUIDE, a, b) + d1_p(c, R_BIGUIDE, a, b);
        {
            const double w1 = d1_p(c, R_PRESS, a, b), w2 = td_p(c, R_RHOVPARA, a, b), w3 = d2_p(c, R_VPARA, a, b);
            const double grad = ((w1 + ((2.0 * b_para) / (8.0 * PI)) * dpara) + ((2.0 * b_perp) / (8.0 * PI)) * dperp) + ((2.0 * b_guide) / (8.0 * PI)) * dguide;
            p1[s_vpara] = (((((-1.0 / rho) * grad) - (w2 - v_para * td_rho) / rho) + (b_para / (4.0 * PI * rho)) * dpara) + imask * grav_para) + visc * w3;
        }
        {
            const double w2 = td_p(c, R_RHOVGUIDE, a, b), w3 = d2_p(c, R_VGUIDE, a, b);
            p1[3] = (((-(w2 - v_guide * td_rho)) / rho) + (b_para / (4.0 * PI * rho)) * dguide) + visc * w3;
        }
        {
            const double w2 = td_p(c, R_RHOVPERP, a, b), w3 = d2_p(c, R_VPERP, a, b);
            p1[s_vperp] = ((((-(w2 - v_perp * td_rho)) / rho) + (b_para / (4.0 * PI * rho)) * dperp) + imask * grav_perp) + visc * w3;
        }
        p1[4] = (-td_p(c, R_EN, a, b)) - press * d1_p(c, R_VPARA, a, b);
        p1[s_bpara] = (-(td_p(c, R_BEPARA, a, b) + td_p(c, R_BIPARA, a, b))) + b_para * d1_p(c, R_VPARA, a, b);
        p1[7] = (-(td_p(c, R_BEGUIDE, a, b) + td_p(c, R_BIGUIDE, a, b))) + b_para * d1_p(c, R_VGUIDE, a, b);
        p1[s_bperp] = (-(td_p(c, R_BEPERP, a, b) + td_p(c, R_BIPERP, a, b))) + b_para * d1_p(c, R_VPERP, a, b);
    } else {
        for (int v = 0; v < 8; v++) p1[v] = 0.0;
    }
    for (int v = 0; v < 8; v++) res[v] += n0[v] + p1[v];                                        // :321-329
}

// does side s (0..3 = x1, x2, y1, y2) evolve ghost cell (i, j)?
MOC_HD bool side_owns(const Field &F, int s, int i, int j)
{
    if (F.bc[s] != MBC_OPEN_MOC) return false;
    const Side S = make_side(F, s);
    const int a = S.bidx == 0 ? i : j, b = S.bidx == 0 ? j : i;
    return a >= S.alo && a <= S.ahi && b >= S.Flo && b <= S.Fhi;
}

// ---- thread -> cell mapping of the strip kernels (moc_stage.cuh): 2 ghost layers x the slab's part of each of the four sides.
// A slab holds rows [row0, row0 + nx_local) of gnx; the x sides belong to the first / last slab (slabs are at least 2*NG rows thick).
MOC_HD int n_threads(int nx_local, int ny) { return 2 * NG * (nx_local + ny); }
MOC_HD bool thread_cell(int nx_local, int ny, int row0, int gnx, int t, int *side, int *i, int *j)
{
    const int nxs = NG * ny, nys = NG * nx_local;
    if (t < 0) return false;
    if (t < nxs) { *side = 0; *i = t / ny; *j = t % ny; return row0 == 0; }
    t -= nxs;
    if (t < nxs) { *side = 1; *i = gnx - NG + t / ny; *j = t % ny; return row0 + nx_local == gnx; }
    t -= nxs;
    if (t < nys) { *side = 2; *j = t / nx_local; *i = row0 + t % nx_local; return true; }
    t -= nys;
    if (t < nys) { *side = 3; *j = ny - NG + t / nx_local; *i = row0 + t % nx_local; return true; }
    return false;
}
// a corner cell evolved by two sides is handled by the thread of the first of them
MOC_HD bool thread_owns(const Field &F, int side, int i, int j)
{
    if (!side_owns(F, side, i, j)) return false;
    for (int s = 0; s < side; s++) if (side_owns(F, s, i, j)) return false;
    return true;
}

// computeTimeDerivativesCharacteristicBoundary at cell (i, j) (idealmhd.cpp:306-331 + :88-103): the characteristic part of
// d/dt of the EVOLVED variables rho, mom_x, mom_y, mom_z, thermal_energy, bi_x, bi_y, bi_z.  Returns false when no open_moc side
// evolves this cell (k stays untouched).
MOC_HD bool moc_cell_terms(const Field &F, int i, int j, double *k)
{
    double res[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool any = false;
    for (int s = 0; s < 4; s++) {
        if (!side_owns(F, s, i, j)) continue;
        Ctx c; c.F = &F; c.S = make_side(F, s);
        side_terms(c, c.S.bidx == 0 ? i : j, c.S.bidx == 0 ? j : i, res);
        any = true;
    }
    if (!any) return false;
    const size_t q = (size_t)i * F.pitch + j;
    const double rho = F.n[q] * F.m_i;
    const double vx = F.mx[q] / rho, vy = F.my[q] / rho, vz = F.mz[q] / rho;
    const double r0 = res[0];
    k[0] = r0;
    k[1] = rho * res[1] + vx * r0; k[2] = rho * res[2] + vy * r0; k[3] = rho * res[3] + vz * r0;
    k[4] = res[4]; k[5] = res[5]; k[6] = res[6]; k[7] = res[7];
    return true;
}

// ---- the update of one evolved ghost cell: equationset.cpp:226-228 (U += k*s), enforceMinimums (idealmhd.cpp:234-239), the
// pointwise part of updateGhostZones that can reach it (fixed / reflect sides zero the momenta of their ghost cells on the PRIMARY
// state, evolution.cpp:245-263,272-282), the rho -> n round trip of the derived step (:246-247) and recomputeDT (:279-304).
enum { MBC_FIXED = 2, MBC_REFLECT = 3 };
struct Floors { double n_min, e_min; };
struct Updated { double n, mx, my, mz, e, bx, by, bz; };

MOC_HD bool momentum_zeroed(const Field &F, int i, int j)
{
    const int xl = F.bc[0] == MBC_PERIODIC ? 0 : NG, xu = F.bc[1] == MBC_PERIODIC ? F.nx - 1 : F.nx - NG - 1;
    const int yl = F.bc[2] == MBC_PERIODIC ? 0 : NG, yu = F.bc[3] == MBC_PERIODIC ? F.ny - 1 : F.ny - NG - 1;
    const bool jin = (j >= yl && j <= yu), iin = (i >= xl && i <= xu);
    if (i <= 2        && (F.bc[0] == MBC_FIXED || (F.bc[0] == MBC_REFLECT && jin))) return true;
    if (i >= F.nx - 3 && (F.bc[1] == MBC_FIXED || (F.bc[1] == MBC_REFLECT && jin))) return true;
    if (j <= 2        && (F.bc[2] == MBC_FIXED || (F.bc[2] == MBC_REFLECT && iin))) return true;
    if (j >= F.ny - 3 && (F.bc[3] == MBC_FIXED || (F.bc[3] == MBC_REFLECT && iin))) return true;
    return false;
}
// base[0..7] = n, mom_x, mom_y, mom_z, thermal_energy, bi_x, bi_y, bi_z of the state the increment is added to
MOC_HD Updated advance_cell(const Field &F, const Floors &fl, const double *base, const double *k, double s, bool primary, int i, int j)
{
    Updated u;
    const double rho_u = (base[0] * F.m_i) + k[0] * s;
    u.mx = base[1] + k[1] * s; u.my = base[2] + k[2] * s; u.mz = base[3] + k[3] * s;
    const double e_u = base[4] + k[4] * s;
    u.bx = base[5] + k[5] * s; u.by = base[6] + k[6] * s; u.bz = base[7] + k[7] * s;
    const double n1 = smax_(rho_u / F.m_i, fl.n_min);            // enforceMinimums :237
    const double rr = n1 * F.m_i;
    u.n = smax_(rr / F.m_i, fl.n_min);                           // derived step :246
    u.e = smax_(e_u, fl.e_min);
    if (primary && momentum_zeroed(F, i, j)) { u.mx = 0.0; u.my = 0.0; u.mz = 0.0; }
    return u;
}
// recomputeDT (idealmhd.cpp:279-304) for one cell; b = be + bi
MOC_HD double cell_dt_plain(const Field &F, double n, double mx, double my, double e, double bx, double by, double bz, double dx, double dy)
{
    const double rho = n * F.m_i;
    const double vx = mx / rho, vy = my / rho;
    const double p = e * F.gm1;
    const double bm = sqrt((bx * bx + by * by) + bz * bz);
    const double cs = sqrt(F.gamma * p / rho);
    const double cs2 = cs * cs;
    const double va = bm / sqrt(rho * (4.0 * kPi));
    const double va2 = va * va;
    const double sm = cs2 + va2;
    const double delta = sqrt(1.0 - ((cs2 * 4.0) * va2) / (sm * sm));
    const double vfast = sqrt((sm * 0.5) * (1.0 + delta));
    const double vslow = sqrt((sm * 0.5) * (1.0 - delta));
    const double vmx = sqrt(vx * vx), vmy = sqrt(vy * vy);
    const double M = smax_(smax_(smax_(cs, va), vfast), vslow);
    return 1.0 / ((vmx + M) / dx + (vmy + M) / dy);
}
// is (i, j) inside the bounds of the time-step minimum (the interior widened by the ghost zone on open_moc sides, plasmadomain.cpp:155-159)?
MOC_HD bool in_dt_bounds(const Field &F, int i, int j)
{
    const int xl = (F.bc[0] == MBC_PERIODIC || F.bc[0] == MBC_OPEN_MOC) ? 0 : NG, xu = (F.bc[1] == MBC_PERIODIC || F.bc[1] == MBC_OPEN_MOC) ? F.nx - 1 : F.nx - NG - 1;
    const int yl = (F.bc[2] == MBC_PERIODIC || F.bc[2] == MBC_OPEN_MOC) ? 0 : NG, yu = (F.bc[3] == MBC_PERIODIC || F.bc[3] == MBC_OPEN_MOC) ? F.ny - 1 : F.ny - NG - 1;
    return i >= xl && i <= xu && j >= yl && j <= yu;
}
MOC_HD bool in_interior(const Field &F, int i, int j)
{
    const int xl = F.bc[0] == MBC_PERIODIC ? 0 : NG, xu = F.bc[1] == MBC_PERIODIC ? F.nx - 1 : F.nx - NG - 1;
    const int yl = F.bc[2] == MBC_PERIODIC ? 0 : NG, yu = F.bc[3] == MBC_PERIODIC ? F.ny - 1 : F.ny - NG - 1;
    return i >= xl && i <= xu && j >= yl && j <= yu;
}

// ---- applyMomThresholdingMoC / applyBThresholdingMoC (idealmhd.cpp:107-223), run at the head of every derived-variable pass when
// moc_mom_limiting / moc_b_limiting are set: on an open_moc side the two ghost layers and the first interior layer of every line are clamped
// between lower*ref and upper*ref, ref = the value in the second interior layer of that line.  The sides run one after the other (x1, x2, y1, y2):
// a later side reads cells an earlier one clamped.  Within a side the lines are independent; the cells of a line are taken outermost first because
// the reference column of the B limiter on y_bound_2 is m_xdim-2-N_GHOST (reference typo, :154) and can be one of the clamped cells itself.
struct Limits { int b_on, mom_on; double b_lo, b_hi, mom_lo, mom_hi; };
struct Mutable { double *mx, *my, *mz, *bix, *biy, *biz; };          // row-shifted like the planes of Field
MOC_HD double clamp_ref(double v, double ref, double lo, double hi) { return (ref >= 0.0) ? smin_(smax_(v, lo), hi) : smax_(smin_(v, lo), hi); }
MOC_HD void limit_line(const Field &F, const Mutable &U, const Limits &L, int s, int a)
{
    if (F.bc[s] != MBC_OPEN_MOC) return;
    const bool xside = s < 2, lower = (s % 2) == 0;
    const int ncr = xside ? F.nx : F.ny;
    const int r = lower ? NG + 1 : ncr - 2 - NG;
    if (L.mom_on) {
        double *m[3] = {U.mx, U.my, U.mz};
        for (int q = 0; q < 3; q++) for (int k = 0; k < NG + 1; k++) {
            const int e = lower ? k : ncr - 1 - k;
            const size_t c = xside ? (size_t)e * F.pitch + a : (size_t)a * F.pitch + e, cr = xside ? (size_t)r * F.pitch + a : (size_t)a * F.pitch + r;
            const double ref = m[q][cr];
            m[q][c] = clamp_ref(m[q][c], ref, L.mom_lo * ref, L.mom_hi * ref);
        }
    }
    if (L.b_on) {
        double *b[3] = {U.bix, U.biy, U.biz};
        const double *be[3] = {F.bex, F.bey, F.bez};
        const int rb = (s == 3) ? F.nx - 2 - NG : r;                              // the reference's index on y_bound_2
        if (s == 3 && (rb < 0 || rb >= F.ny)) return;                             // the reference aborts here (Grid bounds assert)
        for (int q = 0; q < 3; q++) for (int k = 0; k < NG + 1; k++) {
            const int e = lower ? k : ncr - 1 - k;
            const size_t c = xside ? (size_t)e * F.pitch + a : (size_t)a * F.pitch + e, cr = xside ? (size_t)rb * F.pitch + a : (size_t)a * F.pitch + rb;
            const double ref = be[q][cr] + b[q][cr];
            b[q][c] = clamp_ref(b[q][c], ref, L.b_lo * ref - be[q][c], L.b_hi * ref - be[q][c]);
        }
    }
}

// one term of the minimum behind global_visc_coeff (idealmhd.cpp:90) given the cell's dt
MOC_HD double visc_min_term(double dx, double dy, double dt) { return (1.0 / (1.0 / (dx * dx) + 1.0 / (dy * dy))) / dt; }

}  // namespace moc
}  // namespace spruce
