// mhd2e_cells.cuh -- the IdealMHD2E equation set (one fluid, separate ion / electron thermal energies), per-cell arithmetic.
//
// Replaces   IdealMHD2E::computeTimeDerivativesDerived                              source/equationsets/idealmhd2E.cpp:22-58
//            enforceMinimums / recomputeEvolvedVarsFromStateVars / recomputeDerivedVarsFromEvolvedVars / recomputeDT   :60-137
//            PlasmaDomain::updateGhostZones with the set's species lists {i, e}     source/mhd/evolution.cpp:126-333
//            the operators it calls                                                  source/mhd/derivs.cpp:10-73, 122-162, 223-264, 407-414, 465-474
//
// Everything here is plain C++ (no CUDA dependence): mhd2e_host.cuh wraps these functions in kernels, one thread per cell or per boundary
// index, and tests/hostcheck/mhd2e_host_check.cpp compiles THE SAME SOURCE with g++ and runs whole time steps on the host so that the
// arithmetic, the floors, the boundary passes and the integrator bookkeeping can be checked bit for bit without a GPU.  Operations are
// individually rounded in the reference's order (the CUDA unit is compiled with -fmad=false); divisions are plain IEEE divisions (first
// version: correctness first, this equation set is outside the north star's benchmark).
// The state planes hold exactly the reference's evolved variables: rho (mass density, not n), mom_x, mom_y, i_thermal_energy,
// e_thermal_energy, bi_x, bi_y; static planes be_x, be_y, grav_x, grav_y; cell sizes dx[i], dy[j].
#pragma once
#include <cmath>
#include <cstddef>

#if defined(__CUDACC__)
#define E2_HD __host__ __device__ inline
#else
#define E2_HD static inline
#endif

namespace spruce {
namespace e2 {

constexpr int NG = 2;
constexpr int NEV2 = 7;                                  // rho, mom_x, mom_y, i_thermal_energy, e_thermal_energy, bi_x, bi_y
enum { Q_RHO2 = 0, Q_MX2, Q_MY2, Q_EI2, Q_EE2, Q_BX2, Q_BY2 };
enum { BC2_PERIODIC = 0, BC2_OPEN = 1, BC2_FIXED = 2, BC2_REFLECT = 3, BC2_OPEN_MOC = 4, BC2_OPEN_UCNP = 5 };
constexpr double kPi2 = 3.14159265358979323846;          // source/constants.hpp:16
constexpr double kKB2 = 1.3807e-16;                      // K_B, source/constants.hpp:8
constexpr double kE2 = 4.80320425e-10;                   // E (not E_CHARGE), source/constants.hpp:19
constexpr double kMe2 = 9.1094e-28;                      // M_ELECTRON, source/constants.hpp:9

// Cells are addressed by their GLOBAL (i, j).  On a slab (rows [row0, row0 + nxl) of nx, plus two halo rows on each side) the caller passes plane
// pointers shifted back by row0 rows and dx shifted back by row0 entries, so that only resident rows are ever touched; with a periodic x axis the
// halo rows hold the ring neighbours' rows and x indices are not wrapped (x_halo).  A single rank has row0 = 0, nxl = nx, x_halo = 0.
struct Geo {
    const double *dx, *dy;        // cell sizes
    int nx, ny, pitch;            // GLOBAL extent
    int row0, nxl, x_halo;
    int bc[4];                    // x1, x2, y1, y2
    int xl, xu, yl, yu;           // interior bounds (computeIterationBounds, plasmadomain.cpp:138-161)
    int xper, yper;
    double m_i, gamma, n_min, T_min, e_min, open_strength;
    int eic;                      // eic_thermalization configured (source/modules/ucnp/eic_thermalization.cpp): its term joins the right-hand side
    double scale_1[4], scale_2[4];   // open boundary decay factors per side, pow(open_boundary_decay_base, dist / delta_last) evaluated on the HOST with libm
                                     // exactly as the reference evaluates them (evolution.cpp:163-167); see open_scales()
};
// evolution.cpp:163-167 for side s (0..3 = x1, x2, y1, y2); host only (std::pow of the host libm)
inline void open_scales(Geo &g, double open_decay)
{
    for (int s = 0; s < 4; s++) {
        const bool xside = s < 2, lower = (s % 2) == 0;
        const double *d = xside ? g.dx : g.dy;
        const int n = xside ? g.nx : g.ny;
        const int e1 = lower ? 0 : n - 1, e2 = lower ? 1 : n - 2, e3 = lower ? 2 : n - 3;
        const double dist23 = 0.5 * (d[e2] + d[e3]), dist12 = 0.5 * (d[e1] + d[e2]);
        g.scale_2[s] = std::pow(open_decay, dist23 / d[e3]);
        g.scale_1[s] = std::pow(open_decay, dist12 / d[e3]);
    }
}
struct Planes { double *u[NEV2]; };                      // one state
struct CPlanes { const double *u[NEV2]; };
struct Statics { const double *bex, *bey, *gx, *gy; };

E2_HD double smin2(double a, double b) { return (b < a) ? b : a; }      // std::min / std::max
E2_HD double smax2(double a, double b) { return (a < b) ? b : a; }
E2_HD int wi(const Geo &g, int i) { return (g.xper && !g.x_halo) ? (i + g.nx) % g.nx : i; }
E2_HD int wj(const Geo &g, int j) { return g.yper ? (j + g.ny) % g.ny : j; }
E2_HD size_t at(const Geo &g, int i, int j) { return (size_t)i * g.pitch + j; }
E2_HD bool interior(const Geo &g, int i, int j) { return i >= g.xl && i <= g.xu && j >= g.yl && j <= g.yu; }

// the fields the right-hand side differentiates, evaluated at an arbitrary cell from the state planes
enum Field2 { F_RHO = 0, F_MX, F_MY, F_EI, F_EE, F_BIX, F_BIY, F_PRESS, F_VX, F_VY, F_ZE, F_ZI };
E2_HD double fval(const Geo &g, const CPlanes &S, const Statics &T, int f, int i, int j)
{
    const size_t c = at(g, i, j);
    switch (f) {
    case F_RHO: return S.u[Q_RHO2][c];
    case F_MX: return S.u[Q_MX2][c];
    case F_MY: return S.u[Q_MY2][c];
    case F_EI: return S.u[Q_EI2][c];
    case F_EE: return S.u[Q_EE2][c];
    case F_BIX: return S.u[Q_BX2][c];
    case F_BIY: return S.u[Q_BY2][c];
    case F_PRESS: { const double gm1 = g.gamma - 1.0; return gm1 * S.u[Q_EI2][c] + gm1 * S.u[Q_EE2][c]; }          // press = i_press + e_press  (:93-95)
    case F_VX: return S.u[Q_MX2][c] / S.u[Q_RHO2][c];
    case F_VY: return S.u[Q_MY2][c] / S.u[Q_RHO2][c];
    case F_ZE: { const double vx = S.u[Q_MX2][c] / S.u[Q_RHO2][c], vy = S.u[Q_MY2][c] / S.u[Q_RHO2][c]; return vx * T.bey[c] - vy * T.bex[c]; }   // CrossProduct2D(v, be)
    default: { const double vx = S.u[Q_MX2][c] / S.u[Q_RHO2][c], vy = S.u[Q_MY2][c] / S.u[Q_RHO2][c]; return vx * S.u[Q_BY2][c] - vy * S.u[Q_BX2][c]; }
    }
}
// boundaryInterpolate / boundaryExtrapolate (derivs.cpp:477-499); da, db = half cell sizes of the two cells
E2_HD double interp(double qa, double qb, double da, double db) { return (qa * db + qb * da) / (db + da); }
E2_HD double extrap(double qa, double qb, double da, double db) { return qa + (qb - qa) * (da + 2.0 * db) / (da + db); }

// value of field f at the cell `k` steps along direction dir (0: x, 1: y) from (i, j), periodic wrap included; and that cell's half size
E2_HD double fa(const Geo &g, const CPlanes &S, const Statics &T, int f, int dir, int i, int j, int k)
{
    return dir == 0 ? fval(g, S, T, f, wi(g, i + k), j) : fval(g, S, T, f, i, wj(g, j + k));
}
E2_HD double ha(const Geo &g, int dir, int i, int j, int k) { return dir == 0 ? 0.5 * g.dx[wi(g, i + k)] : 0.5 * g.dy[wj(g, j + k)]; }

// derivative1D at an interior cell (derivs.cpp:223-264); zero elsewhere
E2_HD double d1(const Geo &g, const CPlanes &S, const Statics &T, int f, int dir, int i, int j)
{
    if (!interior(g, i, j)) return 0.0;
    const double q0 = fa(g, S, T, f, dir, i, j, -1), q1 = fa(g, S, T, f, dir, i, j, 0), q2 = fa(g, S, T, f, dir, i, j, 1);
    const double h0 = ha(g, dir, i, j, -1), h1 = ha(g, dir, i, j, 0), h2 = ha(g, dir, i, j, 1);
    return (interp(q1, q2, h1, h2) - interp(q0, q1, h0, h1)) / (dir == 0 ? g.dx[i] : g.dy[j]);
}
// upwindSurface (derivs.cpp:10-73): Barton face value of field f at the face between cells (k-1) and k steps from (i, j) along dir
E2_HD double up_face(const Geo &g, const CPlanes &S, const Statics &T, int f, int dir, int i, int j, int k)
{
    const int vf_field = dir == 0 ? F_VX : F_VY;
    const double qm = fa(g, S, T, f, dir, i, j, k - 1), qc = fa(g, S, T, f, dir, i, j, k);
    const double hm = ha(g, dir, i, j, k - 1), hc = ha(g, dir, i, j, k);
    const double d2 = interp(qm, qc, hm, hc);
    const double vf = interp(fa(g, S, T, vf_field, dir, i, j, k - 1), fa(g, S, T, vf_field, dir, i, j, k), hm, hc);
    if (vf > 0.0) {
        const double d1_ = extrap(fa(g, S, T, f, dir, i, j, k - 2), qm, ha(g, dir, i, j, k - 2), hm);
        return (qc <= qm) ? smin2(qm, smax2(d1_, d2)) : smax2(qm, smin2(d1_, d2));
    }
    if (vf < 0.0) {
        const double d1_ = extrap(fa(g, S, T, f, dir, i, j, k + 1), qc, ha(g, dir, i, j, k + 1), hc);
        return (qc <= qm) ? smax2(qc, smin2(d1_, d2)) : smin2(qc, smax2(d1_, d2));
    }
    return d2;
}
// transportDerivative1D at an interior cell (derivs.cpp:122-162)
E2_HD double td(const Geo &g, const CPlanes &S, const Statics &T, int f, int dir, int i, int j)
{
    const int vf_field = dir == 0 ? F_VX : F_VY;
    const double v0 = fa(g, S, T, vf_field, dir, i, j, -1), v1 = fa(g, S, T, vf_field, dir, i, j, 0), v2 = fa(g, S, T, vf_field, dir, i, j, 1);
    const double h0 = ha(g, dir, i, j, -1), h1 = ha(g, dir, i, j, 0), h2 = ha(g, dir, i, j, 1);
    return (up_face(g, S, T, f, dir, i, j, 1) * interp(v1, v2, h1, h2) - up_face(g, S, T, f, dir, i, j, 0) * interp(v0, v1, h0, h1)) / (dir == 0 ? g.dx[i] : g.dy[j]);
}
E2_HD double tdiv(const Geo &g, const CPlanes &S, const Statics &T, int f, int i, int j) { return td(g, S, T, f, 0, i, j) + td(g, S, T, f, 1, i, j); }   // derivs.cpp:216-220

// computeTimeDerivativesDerived at one cell (idealmhd2E.cpp:22-58): k[0..6] = d/dt of the evolved variables
E2_HD void rhs_cell(const Geo &g, const CPlanes &S, const Statics &T, int i, int j, double *k)
{
    if (!interior(g, i, j)) {                            // every operator is zero outside the interior and the ghost-zone mask multiplies the rest (:56-57)
        for (int v = 0; v < NEV2; v++) k[v] = 0.0;
        return;
    }
    const size_t c = at(g, i, j);
    const double gm1 = g.gamma - 1.0;
    const double rho = S.u[Q_RHO2][c];
    k[0] = -tdiv(g, S, T, F_RHO, i, j);                                                              // :32
    const double curl_db = (d1(g, S, T, F_BIY, 0, i, j) - d1(g, S, T, F_BIX, 1, i, j)) / (4.0 * kPi2);  // :34
    const double ext_x = (-curl_db) * T.bey[c], int_x = (-curl_db) * S.u[Q_BY2][c];                  // CrossProductZ2D (grid.cpp:455-460)
    const double ext_y = curl_db * T.bex[c], int_y = curl_db * S.u[Q_BX2][c];
    k[1] = ((-tdiv(g, S, T, F_MX, i, j)) - d1(g, S, T, F_PRESS, 0, i, j)) + 1.0 * ((rho * T.gx[c] + ext_x) + int_x);   // :38-40 (mask = 1 here)
    k[2] = ((-tdiv(g, S, T, F_MY, i, j)) - d1(g, S, T, F_PRESS, 1, i, j)) + 1.0 * ((rho * T.gy[c] + ext_y) + int_y);   // :41-43
    const double divv = d1(g, S, T, F_VX, 0, i, j) + d1(g, S, T, F_VY, 1, i, j);                     // divergence2D
    k[3] = (-tdiv(g, S, T, F_EI, i, j)) - (gm1 * S.u[Q_EI2][c]) * divv;                              // :45-46
    k[4] = (-tdiv(g, S, T, F_EE, i, j)) - (gm1 * S.u[Q_EE2][c]) * divv;                              // :47-48
    k[5] = d1(g, S, T, F_ZE, 1, i, j) + d1(g, S, T, F_ZI, 1, i, j);                                  // curlZ(...)[0] = d/dy   :50-52
    k[6] = (-d1(g, S, T, F_ZE, 0, i, j)) + (-d1(g, S, T, F_ZI, 0, i, j));                            // curlZ(...)[1] = -d/dx  :53
    for (int v = 0; v < NEV2; v++) k[v] = k[v] * 1.0;                                                // the final mask multiply (:56-57) is by 1 in the interior
    if (g.eic) {
        // EICThermalization::computeTimeDerivativesModule (eic_thermalization.cpp:27-44), called by EquationSet::computeTimeDerivatives after the set's own
        // right-hand side (equationset.cpp:204-210): reads n, e_temp and the two thermal energies of the grids being differentiated.  n and e_temp are what
        // recomputeDerivedVarsFromEvolvedVars left for these (settled) planes, formed as derive_cell forms them.  pow(x, 1./3.) -> cbrt, pow(G, 1.5) -> G sqrt(G)
        // as in the two-fluid kernel: the module is held to 1e-9, not to the bit.
        const double n = smax2(rho / g.m_i, g.n_min);
        const double Te = smax2((gm1 * S.u[Q_EE2][c]) / (kKB2 * n), g.T_min);
        const double a = cbrt((3. / 4. / kPi2) / n);
        const double w_pe = sqrt((4. * kPi2 * kE2 * kE2 / kMe2) * n);
        const double Gam = ((kE2 * kE2 / kKB2) / Te) / a;
        const double g15 = Gam * sqrt(Gam);
        const double Lam = (1. / sqrt(3.)) / g15;
        const double gam_ei = ((sqrt(2. / 3. / kPi2) * g15) * w_pe) * log(Lam);
        const double nu_ei = (2. * kMe2 / g.m_i) * gam_ei;
        const double dE = nu_ei * (S.u[Q_EE2][c] - S.u[Q_EI2][c]);
        k[Q_EE2] = k[Q_EE2] - dE * 1.0;                                                              // -= dEdt * mask, += dEdt * mask (:42-43)
        k[Q_EI2] = k[Q_EI2] + dE * 1.0;
    }
}

// applyTimeDerivatives + enforceMinimums at one cell (equationset.cpp:226-228, idealmhd2E.cpp:60-66): out = floor(base + step*k)
E2_HD void apply_cell(const Geo &g, const double *base, const double *k, double s, double *out)
{
    for (int v = 0; v < NEV2; v++) out[v] = base[v] + s * k[v];
    out[Q_RHO2] = g.m_i * smax2(out[Q_RHO2] / g.m_i, g.n_min);
    out[Q_EI2] = smax2(out[Q_EI2], g.e_min);
    out[Q_EE2] = smax2(out[Q_EE2], g.e_min);
}
// recomputeDerivedVarsFromEvolvedVars at one cell, the part that feeds back into the evolved planes (:78-92): rho -> n -> rho, energy floors
E2_HD void settle_cell(const Geo &g, double *u)
{
    const double n = smax2(u[Q_RHO2] / g.m_i, g.n_min);
    u[Q_RHO2] = n * g.m_i;
    u[Q_EI2] = smax2(u[Q_EI2], g.e_min);
    u[Q_EE2] = smax2(u[Q_EE2], g.e_min);
}
// recomputeDT (:111-137) from settled cell values
E2_HD double dt_cell(const Geo &g, const double *u, double bex, double bey, double dx, double dy)
{
    const double gm1 = g.gamma - 1.0;
    const double rho = u[Q_RHO2];
    const double vx = u[Q_MX2] / rho, vy = u[Q_MY2] / rho;
    const double press = gm1 * u[Q_EI2] + gm1 * u[Q_EE2];
    const double bx = bex + u[Q_BX2], by = bey + u[Q_BY2];
    const double bm = sqrt(bx * bx + by * by);
    const double cs = sqrt(g.gamma * press / rho), cs2 = cs * cs;
    const double va = bm / sqrt((4.0 * kPi2) * rho), va2 = va * va;
    const double sm = cs2 + va2;
    const double delta = sqrt(1.0 - ((4.0 * cs2) * va2) / (sm * sm));
    const double vfast = sqrt((0.5 * sm) * (1.0 + delta)), vslow = sqrt((0.5 * sm) * (1.0 - delta));
    const double M = smax2(smax2(smax2(cs, va), vfast), vslow);
    return 1. / ((fabs(vx) + M) / dx + (fabs(vy) + M) / dy);
}
// recomputeEvolvedVarsFromStateVars (:68-76): thermal energies from the temperatures of a state file
E2_HD void from_state_cell(const Geo &g, double rho, double i_temp, double e_temp, double *ei, double *ee)
{
    const double n = smax2(rho / g.m_i, g.n_min);
    const double pi_ = (n * kKB2) * smax2(i_temp, g.T_min), pe_ = (n * kKB2) * smax2(e_temp, g.T_min);
    *ei = smax2(pi_ / (g.gamma - 1.0), g.e_min);
    *ee = smax2(pe_ / (g.gamma - 1.0), g.e_min);
}

// ---- boundary passes (evolution.cpp:126-333): one call handles boundary index `a` of side `side` (0..3 = x1, x2, y1, y2).
// open / reflect / fixed write the PRIMARY state P whatever set is being propagated; open_ucnp writes the propagated set G.
// The four sides must run one after the other in this order (each reads what the earlier ones wrote).
// A slab runs the y sides for its own rows (index t -> a = row0 + t) and an x side only when it holds that side's three rows (the first / last slab).
E2_HD int side_length(const Geo &g, int side) { return side < 2 ? g.ny : g.nxl; }
E2_HD void ghost_cell(const Geo &g, const Planes &G, const Planes &P, int side, int t)
{
    const int bc = g.bc[side];
    if (bc == BC2_PERIODIC || bc == BC2_OPEN_MOC) return;
    const bool xside = side < 2, lower = (side % 2) == 0;
    const int ncross = xside ? g.nx : g.ny;
    const int e1 = lower ? 0 : ncross - 1, e2 = lower ? 1 : ncross - 2, e3 = lower ? 2 : ncross - 3;
    if (xside && (lower ? g.row0 != 0 : g.row0 + g.nxl != g.nx)) return;
    const int a = xside ? t : g.row0 + t;
    const int lo = xside ? g.yl : g.xl, hi = xside ? g.yu : g.xu;
    const size_t c1 = xside ? at(g, e1, a) : at(g, a, e1), c2 = xside ? at(g, e2, a) : at(g, a, e2), c3 = xside ? at(g, e3, a) : at(g, a, e3);
    if (bc == BC2_FIXED) {                                                                           // :268-282, the whole side
        P.u[Q_MX2][c1] = 0.0; P.u[Q_MX2][c2] = 0.0; P.u[Q_MX2][c3] = 0.0;
        P.u[Q_MY2][c1] = 0.0; P.u[Q_MY2][c2] = 0.0; P.u[Q_MY2][c3] = 0.0;
        return;
    }
    if (a < lo || a > hi) return;
    if (bc == BC2_REFLECT) {                                                                         // :231-266
        P.u[Q_EI2][c1] = P.u[Q_EI2][c3]; P.u[Q_EI2][c2] = P.u[Q_EI2][c3];
        P.u[Q_EE2][c1] = P.u[Q_EE2][c3]; P.u[Q_EE2][c2] = P.u[Q_EE2][c3];
        P.u[Q_RHO2][c1] = P.u[Q_RHO2][c3]; P.u[Q_RHO2][c2] = P.u[Q_RHO2][c3];
        P.u[Q_MX2][c1] = 0.0; P.u[Q_MX2][c2] = 0.0; P.u[Q_MX2][c3] = 0.0;
        P.u[Q_MY2][c1] = 0.0; P.u[Q_MY2][c2] = 0.0; P.u[Q_MY2][c3] = 0.0;
    } else if (bc == BC2_OPEN) {                                                                     // :158-224
        const double *d = xside ? g.dx : g.dy;
        const double d1_ = d[e1], d2_ = d[e2], d3_ = d[e3];
        const double dist23 = 0.5 * (d2_ + d3_);
        const double scale_2 = g.scale_2[side], scale_1 = g.scale_1[side];
        (void)d1_;
        P.u[Q_RHO2][c1] = scale_1 * P.u[Q_RHO2][c3]; P.u[Q_RHO2][c2] = scale_2 * P.u[Q_RHO2][c3];
        P.u[Q_EI2][c1] = scale_1 * P.u[Q_EI2][c3]; P.u[Q_EI2][c2] = scale_2 * P.u[Q_EI2][c3];
        P.u[Q_EE2][c1] = scale_1 * P.u[Q_EE2][c3]; P.u[Q_EE2][c2] = scale_2 * P.u[Q_EE2][c3];
        double c_s = 0.0;
        for (int q = 0; q < 2; q++) {                                                                // species i, e: the same rho, their own energy
            const double pr = (g.gamma - 1.0) * P.u[q == 0 ? Q_EI2 : Q_EE2][c3];
            const double c_new = sqrt(g.gamma * pr / P.u[Q_RHO2][c3]);
            if (c_new > c_s) c_s = c_new;
        }
        const double vel_x = P.u[Q_MX2][c3] / P.u[Q_RHO2][c3], vel_y = P.u[Q_MY2][c3] / P.u[Q_RHO2][c3];
        double boost = g.open_strength * c_s;
        if (lower) boost *= -1.0;
        const double vn = xside ? vel_x : vel_y, vt = xside ? vel_y : vel_x;
        const double bv = lower ? smin2(0.0, vn + boost) : smax2(0.0, vn + boost);
        const double gv = (dist23 * bv - 0.5 * d2_ * vn) / (0.5 * d3_);
        const int mn = xside ? Q_MX2 : Q_MY2, mt = xside ? Q_MY2 : Q_MX2;
        P.u[mn][c1] = P.u[Q_RHO2][c1] * gv; P.u[mn][c2] = P.u[Q_RHO2][c2] * gv;                      // (the second species pass rewrites the same values)
        P.u[mt][c1] = P.u[Q_RHO2][c1] * vt; P.u[mt][c2] = P.u[Q_RHO2][c2] * vt;
    } else if (bc == BC2_OPEN_UCNP) {                                                                // :290-333
        for (int v = 0; v < NEV2; v++) { G.u[v][c1] = G.u[v][c3]; G.u[v][c2] = G.u[v][c3]; }
    }
}

// derived variables on demand (idealmhd2E.hpp:18-22 order), from settled evolved planes
enum Var2 { V2_rho = 0, V2_i_temp, V2_e_temp, V2_mom_x, V2_mom_y, V2_bi_x, V2_bi_y, V2_grav_x, V2_grav_y, V2_n, V2_i_press, V2_e_press, V2_press, V2_i_thermal_energy,
            V2_e_thermal_energy, V2_v_x, V2_v_y, V2_kinetic_energy, V2_b_x, V2_b_y, V2_b_mag, V2_b_hat_x, V2_b_hat_y, V2_dt, V2_COUNT };
E2_HD double derive_cell(const Geo &g, const CPlanes &U, const Statics &T, int var, int i, int j)
{
    const size_t c = at(g, i, j);
    const double gm1 = g.gamma - 1.0;
    const double rho = U.u[Q_RHO2][c];
    const double n = smax2(rho / g.m_i, g.n_min);
    const double ip = gm1 * U.u[Q_EI2][c], ep = gm1 * U.u[Q_EE2][c];
    const double vx = U.u[Q_MX2][c] / rho, vy = U.u[Q_MY2][c] / rho;
    const double bx = T.bex[c] + U.u[Q_BX2][c], by = T.bey[c] + U.u[Q_BY2][c];
    const double bm = sqrt(bx * bx + by * by);
    switch (var) {
    case V2_rho: return rho;
    case V2_i_temp: return smax2(ip / (kKB2 * n), g.T_min);
    case V2_e_temp: return smax2(ep / (kKB2 * n), g.T_min);
    case V2_mom_x: return U.u[Q_MX2][c];
    case V2_mom_y: return U.u[Q_MY2][c];
    case V2_bi_x: return U.u[Q_BX2][c];
    case V2_bi_y: return U.u[Q_BY2][c];
    case V2_grav_x: return T.gx[c];
    case V2_grav_y: return T.gy[c];
    case V2_n: return n;
    case V2_i_press: return ip;
    case V2_e_press: return ep;
    case V2_press: return ip + ep;
    case V2_i_thermal_energy: return U.u[Q_EI2][c];
    case V2_e_thermal_energy: return U.u[Q_EE2][c];
    case V2_v_x: return vx;
    case V2_v_y: return vy;
    case V2_kinetic_energy: return (0.5 * rho) * (vx * vx + vy * vy);
    case V2_b_x: return bx;
    case V2_b_y: return by;
    case V2_b_mag: return bm;
    case V2_b_hat_x: return bm == 0.0 ? 0.0 : bx / bm;
    case V2_b_hat_y: return bm == 0.0 ? 0.0 : by / bm;
    default: { double u[NEV2]; for (int v = 0; v < NEV2; v++) u[v] = U.u[v][c]; return dt_cell(g, u, T.bex[c], T.bey[c], g.dx[i], g.dy[j]); }
    }
}


// ---- artificial_viscosity on this set (source/modules/viscosity.cpp:185-267; the fourth module of the UCNP set).  One term: dq = visc_coeff * laplacian(q) * scale_fac
// (+ gradient correction), q = a variable of the grid set the right-hand side is evaluated on.  timescale() = {dt, dt} (idealmhd2E.hpp:45): the timescale is the PRIMARY
// state's dt plane or its minimum, whichever species the term names.  scale_mode 1: momentum <- velocity, n m_i (the set has no i_n, :233-234);
// 2: thermal energy <- temperature, n K_B / (gamma - 1) for both species (:246-251); 0 otherwise.
struct ViscTerm2 { int opt; int var_diff; int scale_mode; double strength; const double *strength_plane; };     // opt: 0 local, 1 global, 2 boundary, 3 boundary_global
struct ViscEnv2 { const double *dt_plane; double dt_min; int gc; };
// derivative1D / secondDerivative1D (derivs.cpp:223-264, 417-456) of a per-cell functor f(i, j); zero outside the interior
template <class F> E2_HD double d1_of(const Geo &g, F f, int dir, int i, int j)
{
    if (!interior(g, i, j)) return 0.0;
    const int im = dir == 0 ? wi(g, i - 1) : i, ip = dir == 0 ? wi(g, i + 1) : i, jm = dir == 1 ? wj(g, j - 1) : j, jp = dir == 1 ? wj(g, j + 1) : j;
    const double q0 = f(im, jm), q1 = f(i, j), q2 = f(ip, jp);
    const double h0 = ha(g, dir, i, j, -1), h1 = ha(g, dir, i, j, 0), h2 = ha(g, dir, i, j, 1);
    return (interp(q1, q2, h1, h2) - interp(q0, q1, h0, h1)) / (dir == 0 ? g.dx[i] : g.dy[j]);
}
template <class F> E2_HD double d2_of(const Geo &g, F f, int dir, int i, int j)
{
    if (!interior(g, i, j)) return 0.0;
    const int im = dir == 0 ? wi(g, i - 1) : i, ip = dir == 0 ? wi(g, i + 1) : i, jm = dir == 1 ? wj(g, j - 1) : j, jp = dir == 1 ? wj(g, j + 1) : j;
    const double q0 = f(im, jm), q1 = f(i, j), q2 = f(ip, jp);
    const double h0 = ha(g, dir, i, j, -1), h1 = ha(g, dir, i, j, 0), h2 = ha(g, dir, i, j, 1);
    return ((interp(q1, q2, h1, h2) - 2.0 * q1) + interp(q0, q1, h0, h1)) / (h1 * h1);          // denominator (0.5 d)^2, derivs.cpp:423
}
E2_HD double visc_cell(const Geo &g, const CPlanes &S, const Statics &T, const ViscTerm2 &t, const ViscEnv2 &e, int i, int j)
{
    auto q = [&](int a, int b) { return derive_cell(g, S, T, t.var_diff, a, b); };
    auto coef = [&](int a, int b) {                                                                  // :213
        const size_t c = at(g, a, b);
        const double str = t.strength_plane ? t.strength_plane[c] : t.strength;
        const double dtg = (t.opt == 0 || t.opt == 2) ? e.dt_plane[c] : e.dt_min;
        const double dx = g.dx[a], dy = g.dy[b];
        return (((str * 1.0) / (1.0 / (dx * dx) + 1.0 / (dy * dy))) / 2.) / dtg;
    };
    auto scale = [&](int a, int b) {
        if (t.scale_mode == 1) return derive_cell(g, S, T, V2_n, a, b) * g.m_i;
        if (t.scale_mode == 2) return derive_cell(g, S, T, V2_n, a, b) * (kKB2 / (g.gamma - 1));
        return 1.0;
    };
    const double lap = d2_of(g, q, 0, i, j) + d2_of(g, q, 1, i, j);                                 // laplacian, derivs.cpp:458-462
    double out = (coef(i, j) * lap) * scale(i, j);                                                   // :266
    if (e.gc) {                                                                                      // :261-265
        auto cs = [&](int a, int b) { return coef(a, b) * scale(a, b); };
        out = (out + d1_of(g, cs, 0, i, j) * d1_of(g, q, 0, i, j)) + d1_of(g, cs, 1, i, j) * d1_of(g, q, 1, i, j);
    }
    return out;
}

}  // namespace e2
}  // namespace spruce
