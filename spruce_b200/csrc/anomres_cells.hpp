// anomres_cells.hpp -- the anomalous_resistivity module (localized magnetic diffusion + Joule heating around a tracked null point), per-cell
// arithmetic and the module's sequence of passes.
//
// Replaces   AnomalousResistivity::setupModule / iterateModule / computeTimeDerivatives      source/modules/solar/anomalousresistivity.cpp:18-178
//            computeNumSubcycles / computeDiffusion / computeTemplate / argminLocalized      :180-281
//            circularMask / currentThresholdMask / currentDensity                            :283-309
//            Grid::argmin, Grid::floodFill                                                   source/utils/grid.cpp:86-100, 487-509
//            derivative1D / secondDerivative1D / laplacian / boundaryInterpolate             source/mhd/derivs.cpp:223-264, 417-462, 477-487
//
// Everything here is plain C++: anomres_host.cuh runs the functors below as kernels (one thread per cell), tests/hostcheck/anomres_host_check.cpp
// compiles THE SAME SOURCE with g++ and runs the same sequence with loops, so the arithmetic and the order of the passes are checked bit for bit
// without a GPU.  An executor X provides
//     double *plane(int slot)                 a scratch plane of the domain's extent (slots are named below)
//     void cells(const F &f)                  f(i, j) for every cell, any order, no cell reads what another cell writes in the same pass
//     int reduce_min(const F &f, double *m)   minimum of f(i, j) over every cell (f may also write its own cell and neighbours monotonically)
// Reference typos are reproduced: computeDiffusion(bi_x, bi_x, bi_z) at :118, curl2D(bi_x, bi_x) at :91, and the `j < result.rows()` loops of
// circularMask / currentThresholdMask (:285, :297).  A whole domain on one rank only (the flood fill and the null-point search are global).
#pragma once
#include <cmath>
#include <cstddef>

#if defined(__CUDACC__)
#define AR_HD __host__ __device__ inline
#else
#define AR_HD inline
#endif

namespace spruce {
namespace ar {

constexpr double kPi = 3.14159265358979323846;           // PI, source/constants.hpp:16
constexpr double kC = 29979245800.0;                      // C, source/constants.hpp:18
constexpr double kECharge = 4.8032e-10;                   // E_CHARGE, source/constants.hpp:10
constexpr double kHuge = 1.7976931348623157e308;
enum { MODEL_TIME_SCALE = 0, MODEL_SYNTELIS_19 = 1, MODEL_YS_94 = 2 };
enum { TI_EULER = 0, TI_RK2 = 1, TI_RK4 = 2 };

AR_HD double smin2(double a, double b) { return (b < a) ? b : a; }
AR_HD double smax2(double a, double b) { return (a < b) ? b : a; }

struct Geom {
    int nx, ny, pitch;
    int xl, xu, yl, yu;           // interior bounds (computeIterationBounds, plasmadomain.cpp:138-161)
    int xper, yper;
    const double *dx, *dy;        // cell sizes per row / per column
    const double *px, *py;        // pos_x, pos_y planes as the state file holds them
};
AR_HD size_t at(const Geom &g, int i, int j) { return (size_t)i * g.pitch + j; }
AR_HD bool interior(const Geom &g, int i, int j) { return i >= g.xl && i <= g.xu && j >= g.yl && j <= g.yu; }
// boundaryInterpolate (derivs.cpp:477-487)
AR_HD double face(double a, double b, double da, double db) { return (a * db + b * da) / (db + da); }

// derivative1D (derivs.cpp:223-264): zero outside the interior
AR_HD double d1(const Geom &g, const double *q, int index, int i, int j)
{
    if (!interior(g, i, j)) return 0.0;
    if (index == 0) {
        int i0 = i - 1, i2 = i + 1;
        if (g.xper) { i0 = (i0 + g.nx) % g.nx; i2 = (i2 + g.nx) % g.nx; }
        const double h0 = 0.5 * g.dx[i0], h1 = 0.5 * g.dx[i], h2 = 0.5 * g.dx[i2];
        return (face(q[at(g, i, j)], q[at(g, i2, j)], h1, h2) - face(q[at(g, i0, j)], q[at(g, i, j)], h0, h1)) / g.dx[i];
    }
    int j0 = j - 1, j2 = j + 1;
    if (g.yper) { j0 = (j0 + g.ny) % g.ny; j2 = (j2 + g.ny) % g.ny; }
    const double h0 = 0.5 * g.dy[j0], h1 = 0.5 * g.dy[j], h2 = 0.5 * g.dy[j2];
    return (face(q[at(g, i, j)], q[at(g, i, j2)], h1, h2) - face(q[at(g, i, j0)], q[at(g, i, j)], h0, h1)) / g.dy[j];
}
// secondDerivative1D (derivs.cpp:417-455)
AR_HD double d2(const Geom &g, const double *q, int index, int i, int j)
{
    if (!interior(g, i, j)) return 0.0;
    if (index == 0) {
        int i0 = i - 1, i2 = i + 1;
        if (g.xper) { i0 = (i0 + g.nx) % g.nx; i2 = (i2 + g.nx) % g.nx; }
        const double h0 = 0.5 * g.dx[i0], h1 = 0.5 * g.dx[i], h2 = 0.5 * g.dx[i2];
        return (face(q[at(g, i, j)], q[at(g, i2, j)], h1, h2) - 2.0 * q[at(g, i, j)] + face(q[at(g, i0, j)], q[at(g, i, j)], h0, h1)) / (h1 * h1);
    }
    int j0 = j - 1, j2 = j + 1;
    if (g.yper) { j0 = (j0 + g.ny) % g.ny; j2 = (j2 + g.ny) % g.ny; }
    const double h0 = 0.5 * g.dy[j0], h1 = 0.5 * g.dy[j], h2 = 0.5 * g.dy[j2];
    return (face(q[at(g, i, j)], q[at(g, i, j2)], h1, h2) - 2.0 * q[at(g, i, j)] + face(q[at(g, i, j0)], q[at(g, i, j)], h0, h1)) / (h1 * h1);
}
AR_HD double lap(const Geom &g, const double *q, int i, int j) { return d2(g, q, 0, i, j) + d2(g, q, 1, i, j); }     // derivs.cpp:458-462
AR_HD double cell_diffusion_scale(const Geom &g, int i, int j) { return (1.0 / (1.0 / (g.dx[i] * g.dx[i]) + 1.0 / (g.dy[j] * g.dy[j]))) / 2.; }

struct Params {
    double time_scale, frob_coeff, sigma, safety, max_radius, argmin_radius, min_current, ramp_length, threshold, model_params[3];
    int smoothing, integrator, flood_fill, model, gradient_correction;
};

// ---------------------------------------------------------------------------------------------------------------- passes (functors)
// currentDensity (:306-309): (c/4pi) * |curl(bi)|
struct CurrentDensity {
    Geom g; const double *bix, *biy, *biz; double *J;
    AR_HD void operator()(int i, int j) const
    {
        const double c2 = d1(g, biy, 0, i, j) - d1(g, bix, 1, i, j);
        const double cx = d1(g, biz, 1, i, j), ipy = -d1(g, biz, 0, i, j);
        J[at(g, i, j)] = (kC / (4.0 * kPi)) * sqrt((c2 * c2 + cx * cx) + ipy * ipy);
    }
};
// total in-plane field and its magnitude
struct FieldMagnitude {
    Geom g; const double *bex, *bey, *bix, *biy; double *bx, *by, *bm;
    AR_HD void operator()(int i, int j) const
    {
        const size_t c = at(g, i, j);
        const double x = bex[c] + bix[c], y = bey[c] + biy[c];
        bx[c] = x; by[c] = y; bm[c] = sqrt(x * x + y * y);
    }
};
// Frobenius metric (:229-236): |grad b|_F / |b|, then frob_coeff * laplacian(.)^2 capped at 1
struct FrobeniusMetric {
    Geom g; const double *bx, *by, *bm; double *q;
    AR_HD void operator()(int i, int j) const
    {
        const double xx = d1(g, bx, 0, i, j), xy = d1(g, bx, 1, i, j), yx = d1(g, by, 0, i, j), yy = d1(g, by, 1, i, j);
        q[at(g, i, j)] = sqrt(((xx * xx + xy * xy) + yx * yx) + yy * yy) / bm[at(g, i, j)];
    }
};
struct FrobeniusTemplate {
    Geom g; const double *q; double *t; double frob_coeff;
    AR_HD void operator()(int i, int j) const { const double l = lap(g, q, i, j); t[at(g, i, j)] = smin2(frob_coeff * (l * l), 1.0); }
};
struct FrobeniusFinish {                                  // :256: pow(min(10 t, 1), 1.5)
    Geom g; double *t;
    AR_HD void operator()(int i, int j) const { t[at(g, i, j)] = pow(smin2(10.0 * t[at(g, i, j)], 1.0), 1.5); }
};
struct Fill { Geom g; double *t; double v; AR_HD void operator()(int i, int j) const { t[at(g, i, j)] = v; } };
// one value of a plane (everything else +huge): read-back through reduce_min
struct ValueAt { Geom g; const double *q; int ci, cj; AR_HD double operator()(int i, int j) const { return (i == ci && j == cj) ? q[at(g, i, j)] : kHuge; } };
// the search window of Grid::argmin (interior) or argminLocalized (index window + distance test, :261-281)
struct Window { int il, iu, jl, ju, ci, cj, localized; double radius; };
AR_HD bool in_window(const Geom &g, const Window &w, int i, int j)
{
    if (i < w.il || i > w.iu || j < w.jl || j > w.ju) return false;
    if (!w.localized) return true;
    const size_t c0 = at(g, w.ci, w.cj), c = at(g, i, j);
    return sqrt(pow(g.px[c0] - g.px[c], 2.0) + pow(g.py[c0] - g.py[c], 2.0)) <= w.radius;
}
struct WindowMin { Geom g; const double *q; Window w; AR_HD double operator()(int i, int j) const { return in_window(g, w, i, j) ? q[at(g, i, j)] : kHuge; } };
// first cell (row-major) of the window that holds the value v, as a double (exact below 2^53)
struct WindowFirst {
    Geom g; const double *q; Window w; double v;
    AR_HD double operator()(int i, int j) const { return (in_window(g, w, i, j) && q[at(g, i, j)] == v) ? (double)((long long)i * g.ny + j) : kHuge; }
};
// Grid::floodFill as label propagation: a cell below the threshold joins when a 4-neighbour has joined; returns 0 when it changed something
struct FloodSeed { Geom g; const double *bm; double *t; int ci, cj; double thr; AR_HD void operator()(int i, int j) const { t[at(g, i, j)] = (i == ci && j == cj && bm[at(g, i, j)] < thr) ? 1.0 : 0.0; } };
struct FloodStep {
    Geom g; const double *bm; double *t; double thr;
    AR_HD double operator()(int i, int j) const
    {
        const size_t c = at(g, i, j);
        if (t[c] >= 1.0 || !(bm[c] < thr)) return 1.0;
        const bool nb = (i + 1 < g.nx && t[at(g, i + 1, j)] >= 1.0) || (i > 0 && t[at(g, i - 1, j)] >= 1.0) || (j + 1 < g.ny && t[at(g, i, j + 1)] >= 1.0) || (j > 0 && t[at(g, i, j - 1)] >= 1.0);
        if (!nb) return 1.0;
        t[c] = 1.0;
        return 0.0;
    }
};
// circularMask (:283-293) and currentThresholdMask (:295-304); columns at and beyond min(xdim, ydim) stay zero (the reference's loop bound)
struct CircularMask {
    Geom g; double *t; int ci, cj; double radius;
    AR_HD void operator()(int i, int j) const
    {
        double m = 0.0;
        if (j < (g.nx < g.ny ? g.nx : g.ny)) {
            const size_t c0 = at(g, ci, cj), c = at(g, i, j);
            if (sqrt(pow(g.px[c] - g.px[c0], 2.0) + pow(g.py[c] - g.py[c0], 2.0)) <= radius) m = 1.0;
        }
        t[at(g, i, j)] = m * t[at(g, i, j)];
    }
};
struct CurrentMask {
    Geom g; double *t; const double *J; double min_current, ramp_length;
    AR_HD void operator()(int i, int j) const
    {
        double m = 0.0;
        if (j < (g.nx < g.ny ? g.nx : g.ny)) { m = (J[at(g, i, j)] - min_current) / ramp_length + 0.5; m = (m < 0.0) ? 0.0 : m; m = (1.0 < m) ? 1.0 : m; }
        t[at(g, i, j)] = m * t[at(g, i, j)];
    }
};
// Gaussian smoothing of the template (:250-254 with the kernel of :31-40), fixed summation order (k outer, l inner), cells outside the domain skipped
struct Smooth {
    Geom g; const double *t; double *out; const double *kernel; int kr;
    AR_HD void operator()(int i, int j) const
    {
        const int ks = 2 * kr + 1;
        double acc = 0.0;
        for (int k = -kr; k <= kr; k++) for (int l = -kr; l <= kr; l++) {
            if (i + k < 0 || j + l < 0 || i + k >= g.nx || j + l >= g.ny) continue;
            acc += kernel[(k + kr) * ks + (l + kr)] * t[at(g, i + k, j + l)];
        }
        out[at(g, i, j)] = acc;
    }
};
struct Copy { Geom g; double *dst; const double *src; AR_HD void operator()(int i, int j) const { dst[at(g, i, j)] = src[at(g, i, j)]; } };
// computeDiffusion (:196-224)
struct Diffusivity {
    Geom g; Params p; const double *J, *n; double *D; double time_scale;
    AR_HD void operator()(int i, int j) const
    {
        const size_t c = at(g, i, j);
        if (p.model == MODEL_TIME_SCALE) { D[c] = cell_diffusion_scale(g, i, j) / time_scale; return; }
        if (p.model == MODEL_SYNTELIS_19) {
            double s = J[c] / p.model_params[2];
            if (s < 1.0) s = 0.0;
            D[c] = p.model_params[1] * s + p.model_params[0];
            return;
        }
        const double ratio = ((J[c] / n[c]) / kECharge) / p.model_params[0];
        double v = smin2(p.model_params[1] * ((ratio - 1) * (ratio - 1)), p.model_params[2]);
        if (ratio <= 1.0) v = 0.0;
        D[c] = v;
    }
};
// computeNumSubcycles (:180-194)
struct AnyPositive { Geom g; const double *t, *D; AR_HD double operator()(int i, int j) const { return (t[at(g, i, j)] * D[at(g, i, j)] > 0.0) ? 0.0 : 1.0; } };
struct Bounds { int il, iu, jl, ju; };
struct MinTimeScale {
    Geom g; const double *t, *D; Bounds b;
    AR_HD double operator()(int i, int j) const
    {
        if (i < b.il || i > b.iu || j < b.jl || j > b.ju) return kHuge;
        return cell_diffusion_scale(g, i, j) / (t[at(g, i, j)] * D[at(g, i, j)]);
    }
};
struct MinInBounds { Geom g; const double *q; Bounds b; AR_HD double operator()(int i, int j) const { return (i < b.il || i > b.iu || j < b.jl || j > b.ju) ? kHuge : q[at(g, i, j)]; } };
// computeTimeDerivatives (:70-105): coefficient and Joule heating, then the three field components
struct Coefficient {
    Geom g; const double *t, *D, *J; double *coeff, *heat;
    AR_HD void operator()(int i, int j) const
    {
        const size_t c = at(g, i, j);
        const double cf = ((interior(g, i, j) ? 1.0 : 0.0) * t[c]) * D[c];
        coeff[c] = cf;
        const double er = (((4.0 * kPi) / kC) / kC) * cf;
        heat[c] = (er * J[c]) * J[c];
    }
};
struct FieldDerivatives {
    Geom g; const double *coeff, *bix, *biy, *biz, *lbex, *lbey, *lbez; double *kx, *ky, *kz; int gradient_correction;
    AR_HD void operator()(int i, int j) const
    {
        const size_t c = at(g, i, j);
        const double cf = coeff[c];
        double gx = 0.0, gy = 0.0, gz = 0.0;
        if (gradient_correction) {                       // grad(eta) x curl(B) with curl2D(bi_x, bi_x) (:91)
            const double ex = d1(g, coeff, 0, i, j), ey = d1(g, coeff, 1, i, j);
            const double bz = d1(g, bix, 0, i, j) - d1(g, bix, 1, i, j), az = -1.0 * bz;
            const double czx = d1(g, biz, 1, i, j), czy = d1(g, biz, 0, i, j);
            gx = (-az) * ey; gy = az * ex; gz = ex * (-czy) - ey * czx;
        }
        const double x = cf * (lbex[c] + lap(g, bix, i, j)), y = cf * (lbey[c] + lap(g, biy, i, j)), z = cf * (lbez[c] + lap(g, biz, i, j));
        kx[c] = gradient_correction ? gx + x : x;
        ky[c] = gradient_correction ? gy + y : y;
        kz[c] = gradient_correction ? gz + z : z;
    }
};
struct Laplacian { Geom g; const double *q; double *out; AR_HD void operator()(int i, int j) const { out[at(g, i, j)] = lap(g, q, i, j); } };
// out = base + s*k
struct Axpy { Geom g; double *out; const double *base, *k; double s; AR_HD void operator()(int i, int j) const { const size_t c = at(g, i, j); out[c] = base[c] + s * k[c]; } };
// rk4 (:150-170): q += (dt*(((k1 + 2 k2) + 2 k3) + k4))/6
struct Rk4Final { Geom g; double *q; const double *k1, *k2, *k3, *k4; double s; AR_HD void operator()(int i, int j) const { const size_t c = at(g, i, j); q[c] += (s * (((k1[c] + 2.0 * k2[c]) + 2.0 * k3[c]) + k4[c])) / 6.0; } };

// ---------------------------------------------------------------------------------------------------------------- the module's sequence
// scratch plane slots of the executor
enum Slot { P_BIX = 0, P_BIY, P_BIZ, P_E, P_BX, P_BY, P_BM, P_J, P_TMPL, P_DIFF, P_TMP, P_COEFF, P_LBEX, P_LBEY, P_LBEZ, P_MIDX, P_MIDY, P_MIDZ,
            P_K1, P_K2 = P_K1 + 4, P_K3 = P_K2 + 4, P_K4 = P_K3 + 4, P_COUNT = P_K4 + 4 };

struct State {                                            // what the module keeps between calls
    Params p;
    int null_i, null_j, kr, nsub;
    double time_scale;
    const double *kernel;                                 // (2 kr + 1)^2 weights where the executor's passes can read them
};
// Gaussian kernel of setupModule (:31-40), host libm like the reference; w has (2 kr + 1)^2 entries
inline int smoothing_radius(double sigma) { return (int)std::nearbyint(4.0 * sigma); }
inline void smoothing_kernel(double sigma, int kr, double *w)
{
    const int ks = 2 * kr + 1;
    const double mx = 1.0 / (2.0 * kPi * sigma * sigma);
    for (int a = 0; a < ks; a++) for (int b = 0; b < ks; b++) {
        const double ga = std::exp(-0.5 * std::pow(((double)a - (double)kr) / sigma, 2.0)), gb = std::exp(-0.5 * std::pow(((double)b - (double)kr) / sigma, 2.0));
        w[a * ks + b] = (mx - 0.0) * ga * gb + 0.0;
    }
}

// arg-min with the reference's tie rule: the start cell keeps the title unless some window cell is STRICTLY smaller; among equals the first in row-major order
template <class X>
int argmin_window(X &x, const Geom &g, const double *q, const Window &w, int *oi, int *oj)
{
    int rc;
    double start, m, first;
    if ((rc = x.reduce_min(ValueAt{g, q, w.ci, w.cj}, &start))) return rc;
    if ((rc = x.reduce_min(WindowMin{g, q, w}, &m))) return rc;
    *oi = w.ci; *oj = w.cj;
    if (!(m < start)) return 0;
    if ((rc = x.reduce_min(WindowFirst{g, q, w, m}, &first))) return rc;
    const long long idx = (long long)first;
    *oi = (int)(idx / g.ny); *oj = (int)(idx % g.ny);
    return 0;
}

// computeTemplate (:226-258) into P_TMPL; needs P_BX, P_BY, P_BM (FieldMagnitude) and P_J
// pos_row / pos_col: host copies of pos_x along the null point's column and pos_y along its row are read through `host_px(i, j)` / `host_py(i, j)`
template <class X>
int compute_template(X &x, const Geom &g, State &s)
{
    int rc;
    double *t = x.plane(P_TMPL), *tmp = x.plane(P_TMP);
    const double *bm = x.plane(P_BM);
    if (!s.p.flood_fill) {
        x.cells(FrobeniusMetric{g, x.plane(P_BX), x.plane(P_BY), bm, tmp});
        x.cells(FrobeniusTemplate{g, tmp, t, s.p.frob_coeff});
    } else {
        // argminLocalized (:261-281): index window from the positions along the start cell's row and column
        const int ci = s.null_i, cj = s.null_j;
        int il = ci, iu = ci, jl = cj, ju = cj;
        for (; il >= g.xl; il--) if (std::fabs(x.host_px(ci, cj) - x.host_px(il, cj)) > s.p.argmin_radius) break;
        for (; iu <= g.xu; iu++) if (std::fabs(x.host_px(ci, cj) - x.host_px(iu, cj)) > s.p.argmin_radius) break;
        for (; jl >= g.yl; jl--) if (std::fabs(x.host_py(ci, cj) - x.host_py(ci, jl)) > s.p.argmin_radius) break;
        for (; ju <= g.yu; ju++) if (std::fabs(x.host_py(ci, cj) - x.host_py(ci, ju)) > s.p.argmin_radius) break;
        Window w{il < 0 ? 0 : il, iu >= g.nx ? g.nx - 1 : iu, jl < 0 ? 0 : jl, ju >= g.ny ? g.ny - 1 : ju, ci, cj, 1, s.p.argmin_radius};
        if ((rc = argmin_window(x, g, bm, w, &s.null_i, &s.null_j))) return rc;
        x.cells(FloodSeed{g, bm, t, s.null_i, s.null_j, s.p.threshold});
        for (;;) {
            double unchanged;
            if ((rc = x.reduce_min(FloodStep{g, bm, t, s.p.threshold}, &unchanged))) return rc;
            if (unchanged != 0.0) break;
        }
        if (s.p.max_radius > 0.0) x.cells(CircularMask{g, t, s.null_i, s.null_j, s.p.max_radius});
        if (s.p.min_current > 0.0) x.cells(CurrentMask{g, t, x.plane(P_J), s.p.min_current, s.p.ramp_length});
    }
    if (s.p.smoothing) {
        x.cells(Smooth{g, t, tmp, s.kernel, s.kr});
        x.cells(Copy{g, t, tmp});
    }
    if (!s.p.flood_fill) x.cells(FrobeniusFinish{g, t});
    return 0;
}

// setupModule (:18-44) on the state before the first step: bi = the current perturbation field planes of the domain, be = background field
template <class X>
int setup(X &x, const Geom &g, State &s, const double *bex, const double *bey, const double *bix, const double *biy, const double *biz)
{
    int rc;
    x.cells(FieldMagnitude{g, bex, bey, bix, biy, x.plane(P_BX), x.plane(P_BY), x.plane(P_BM)});
    // Grid::argmin over the interior, started from the centre cell (grid.cpp:86-100)
    Window w{g.xl, g.xu, g.yl, g.yu, (int)(0.5 * g.nx), (int)(0.5 * g.ny), 0, 0.0};
    if ((rc = argmin_window(x, g, x.plane(P_BM), w, &s.null_i, &s.null_j))) return rc;
    x.cells(CurrentDensity{g, bix, biy, biz, x.plane(P_J)});
    return compute_template(x, g, s);
}

// computeTimeDerivatives (:70-105) of the field (b0, b1, b2) into the four planes starting at slot K
template <class X>
int time_derivatives(X &x, const Geom &g, State &s, const double *bex, const double *bey, const double *n, const double *b0, const double *b1, const double *b2, int K)
{
    int rc;
    x.cells(FieldMagnitude{g, bex, bey, b0, b1, x.plane(P_BX), x.plane(P_BY), x.plane(P_BM)});
    x.cells(CurrentDensity{g, b0, b1, b2, x.plane(P_J)});
    if ((rc = compute_template(x, g, s))) return rc;
    x.cells(Diffusivity{g, s.p, x.plane(P_J), n, x.plane(P_DIFF), s.time_scale});
    x.cells(Coefficient{g, x.plane(P_TMPL), x.plane(P_DIFF), x.plane(P_J), x.plane(P_COEFF), x.plane(K + 3)});
    x.cells(FieldDerivatives{g, x.plane(P_COEFF), b0, b1, b2, x.plane(P_LBEX), x.plane(P_LBEY), x.plane(P_LBEZ), x.plane(K), x.plane(K + 1), x.plane(K + 2), s.p.gradient_correction});
    return 0;
}

// iterateModule (:107-178) up to, not including, the write-back + propagateChanges: P_BIX..P_BIZ and P_E hold the domain's bi and thermal energy
// on entry and the module's result on exit.  dtp = the domain's dt plane; moc_ext[4] = 1 where the side is open_moc (the dt bounds grow by the ghost zone)
template <class X>
int iterate(X &x, const Geom &g, State &s, const double *bex, const double *bey, const double *bez, const double *n, const double *dtp, const int *moc_ext, double epsilon, double dt)
{
    int rc;
    double *bi[3] = {x.plane(P_BIX), x.plane(P_BIY), x.plane(P_BIZ)}, *e = x.plane(P_E);
    double *mid[3] = {x.plane(P_MIDX), x.plane(P_MIDY), x.plane(P_MIDZ)};
    x.cells(FieldMagnitude{g, bex, bey, bi[0], bi[1], x.plane(P_BX), x.plane(P_BY), x.plane(P_BM)});
    x.cells(CurrentDensity{g, bi[0], bi[1], bi[2], x.plane(P_J)});
    if ((rc = compute_template(x, g, s))) return rc;
    if (s.p.model != MODEL_TIME_SCALE) x.cells(CurrentDensity{g, bi[0], bi[0], bi[2], x.plane(P_J)});                 // computeDiffusion(bi_x, bi_x, bi_z), :118
    x.cells(Diffusivity{g, s.p, x.plane(P_J), n, x.plane(P_DIFF), s.time_scale});
    // computeNumSubcycles (:180-194)
    const Bounds b{g.xl - (moc_ext[0] ? 2 : 0), g.xu + (moc_ext[1] ? 2 : 0), g.yl - (moc_ext[2] ? 2 : 0), g.yu + (moc_ext[3] ? 2 : 0)};
    double none_positive;
    if ((rc = x.reduce_min(AnyPositive{g, x.plane(P_TMPL), x.plane(P_DIFF)}, &none_positive))) return rc;
    if (none_positive != 0.0) s.nsub = 1;
    else {
        if (s.p.model != MODEL_TIME_SCALE && (rc = x.reduce_min(MinTimeScale{g, x.plane(P_TMPL), x.plane(P_DIFF), b}, &s.time_scale))) return rc;
        double dtmin;
        if ((rc = x.reduce_min(MinInBounds{g, dtp, b}, &dtmin))) return rc;
        const double rk = epsilon * dtmin;
        s.nsub = (int)(1.0 + rk / (s.p.safety * s.time_scale));
    }
    const double dts = dt / (double)s.nsub;
    x.cells(Laplacian{g, bex, x.plane(P_LBEX)});
    x.cells(Laplacian{g, bey, x.plane(P_LBEY)});
    x.cells(Laplacian{g, bez, x.plane(P_LBEZ)});
    for (int sub = 0; sub < s.nsub; sub++) {
        if (s.p.integrator == TI_EULER) {
            if ((rc = time_derivatives(x, g, s, bex, bey, n, bi[0], bi[1], bi[2], P_K1))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, bi[q], bi[q], x.plane(P_K1 + q), dts});
            x.cells(Axpy{g, e, e, x.plane(P_K1 + 3), dts});
        } else if (s.p.integrator == TI_RK2) {
            if ((rc = time_derivatives(x, g, s, bex, bey, n, bi[0], bi[1], bi[2], P_K1))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, mid[q], bi[q], x.plane(P_K1 + q), 0.5 * dts});
            if ((rc = time_derivatives(x, g, s, bex, bey, n, mid[0], mid[1], mid[2], P_K1))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, bi[q], bi[q], x.plane(P_K1 + q), dts});
            x.cells(Axpy{g, e, e, x.plane(P_K1 + 3), dts});
        } else {
            if ((rc = time_derivatives(x, g, s, bex, bey, n, bi[0], bi[1], bi[2], P_K1))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, mid[q], bi[q], x.plane(P_K1 + q), 0.5 * dts});
            if ((rc = time_derivatives(x, g, s, bex, bey, n, mid[0], mid[1], mid[2], P_K2))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, mid[q], bi[q], x.plane(P_K2 + q), 0.5 * dts});
            if ((rc = time_derivatives(x, g, s, bex, bey, n, mid[0], mid[1], mid[2], P_K3))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Axpy{g, mid[q], bi[q], x.plane(P_K3 + q), dts});
            if ((rc = time_derivatives(x, g, s, bex, bey, n, mid[0], mid[1], mid[2], P_K4))) return rc;
            for (int q = 0; q < 3; q++) x.cells(Rk4Final{g, bi[q], x.plane(P_K1 + q), x.plane(P_K2 + q), x.plane(P_K3 + q), x.plane(P_K4 + q), dts});
            x.cells(Rk4Final{g, e, x.plane(P_K1 + 3), x.plane(P_K2 + 3), x.plane(P_K3 + 3), x.plane(P_K4 + 3), dts});
        }
    }
    return 0;
}

}  // namespace ar
}  // namespace spruce
