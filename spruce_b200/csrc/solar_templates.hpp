// solar_templates.hpp -- the static spatial templates of the pointwise solar source-term modules, built on the host exactly as the
// reference builds them (host libm: the reference binary and this library link the same glibc):
//   SolarUtils::GaussianGrid / GaussianGridRotated     source/solar/solarutils.cpp:65-98     (coordinates = GLOBAL cell indices)
//   LocalizedHeating / MassInjection templates          source/modules/solar/localizedheating.cpp:31-50, massinjection.cpp:27-46
//   MomentumInjection templates                         source/modules/solar/momentuminjection.cpp:35-67
// Plain C++ (no CUDA): capi.cu includes it, and tests/hostcheck compiles it with g++ to check it without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace spruce {
namespace solar {

constexpr double kPiS = 3.14159265358979323846;        // source/constants.hpp:16

struct Geom {
    int nx_local, ny, row0;      // the slab's rows [row0, row0 + nx_local) of ny columns, stored densely (ny doubles per row)
    int xdim, ydim;              // global extent
    bool x_periodic, y_periodic; // x_bound_1 / y_bound_1 == periodic (what the reference tests)
};

inline void gaussian_rows(const Geom &g, double mn, double mx, double sx, double sy, double cx, double cy, double angle_deg, std::vector<double> &out)
{
    const int nx = g.nx_local, ny = g.ny;
    out.assign((size_t)nx * ny, 0.0);
    if (angle_deg == 0.0) {
        std::vector<double> gx(nx), gy(ny);
        for (int r = 0; r < nx; r++) gx[r] = std::exp(-0.5 * std::pow(((double)(g.row0 + r) - cx) / sx, 2.0));
        for (int j = 0; j < ny; j++) gy[j] = std::exp(-0.5 * std::pow(((double)j - cy) / sy, 2.0));
        for (int r = 0; r < nx; r++) for (int j = 0; j < ny; j++) out[(size_t)r * ny + j] = (mx - mn) * gx[r] * gy[j] + mn;
        return;
    }
    for (int r = 0; r < nx; r++) for (int j = 0; j < ny; j++) {
        const double i = (double)(g.row0 + r);
        const double x_eff = (i - cx) * std::cos(angle_deg * kPiS / 180.0) - ((double)j - cy) * std::sin(angle_deg * kPiS / 180.0);
        const double y_eff = (i - cx) * std::sin(angle_deg * kPiS / 180.0) + ((double)j - cy) * std::cos(angle_deg * kPiS / 180.0);
        out[(size_t)r * ny + j] = (mx - mn) * std::exp(-0.5 * std::pow(x_eff / sx, 2.0)) * std::exp(-0.5 * std::pow(y_eff / sy, 2.0)) + mn;
    }
}
// centre shifts (in cells) of the periodic images a template is combined with: +xdim, -xdim, +ydim, -ydim
inline int periodic_shifts(const Geom &g, double sh[4][2])
{
    int n = 0;
    if (g.x_periodic) { sh[n][0] = (double)g.xdim; sh[n][1] = 0.0; n++; sh[n][0] = -(double)g.xdim; sh[n][1] = 0.0; n++; }
    if (g.y_periodic) { sh[n][0] = 0.0; sh[n][1] = (double)g.ydim; n++; sh[n][0] = 0.0; sh[n][1] = -(double)g.ydim; n++; }
    return n;
}
// max(0) of the Gaussian, max-combined with its images
inline void positive_template(const Geom &g, double peak, double sx, double sy, double cx, double cy, std::vector<double> &out)
{
    std::vector<double> t;
    gaussian_rows(g, -1.0e-2 * peak, peak, sx, sy, cx, cy, 0.0, out);
    for (double &q : out) q = (q < 0.0) ? 0.0 : q;                                     // Grid::max(0.0) = std::max(q, 0.0)
    double sh[4][2];
    const int ns = periodic_shifts(g, sh);
    for (int k = 0; k < ns; k++) {
        gaussian_rows(g, -1.0e-2 * peak, peak, sx, sy, cx + sh[k][0], cy + sh[k][1], 0.0, t);
        for (size_t c = 0; c < out.size(); c++) { const double im = (t[c] < 0.0) ? 0.0 : t[c]; out[c] = (out[c] < im) ? im : out[c]; }
    }
}
// one template per acceleration component; the reference min-combines the images in both of its branches (:52-53, :60-61)
inline void momentum_templates(const Geom &g, double sx, double sy, double cx, double cy, double dir_x, double dir_y, double angle_deg,
                               std::vector<double> &px, std::vector<double> &py)
{
    const double mag = std::sqrt(dir_x * dir_x + dir_y * dir_y);                        // :38-40
    const double dir[2] = {dir_x / mag, dir_y / mag};
    std::vector<double> *p[2] = {&px, &py};
    std::vector<double> t;
    double sh[4][2];
    const int ns = periodic_shifts(g, sh);
    for (int k = 0; k < 2; k++) {
        const double a = dir[k];
        gaussian_rows(g, -1.0e-2 * a, a, sx, sy, cx, cy, angle_deg, *p[k]);
        for (int q = 0; q < ns; q++) {
            gaussian_rows(g, -1.0e-2 * a, a, sx, sy, cx + sh[q][0], cy + sh[q][1], angle_deg, t);
            for (size_t c = 0; c < t.size(); c++) (*p[k])[c] = (t[c] < (*p[k])[c]) ? t[c] : (*p[k])[c];
        }
        for (double &q : *p[k]) q = (a > 0.0) ? ((q < 0.0) ? 0.0 : q) : ((0.0 < q) ? 0.0 : q);
    }
}

// ---- BoundaryOutflow (source/modules/solar/boundaryoutflow.cpp): single rank, pos_x / pos_y are full planes (ydim doubles per row)
enum { OB_X1 = 0, OB_X2 = 1, OB_Y1 = 2, OB_Y2 = 3 };            // boundary
enum { OS_EXP = 0, OS_GAUSSIAN = 1, OS_FLAT = 2 };               // falloff_shape
struct Window { int xl, xu, yl, yu; };

// interior bounds, extended into the ghost zone on the chosen side when that side is open_moc (m_*_dt, plasmadomain.cpp:138-159)
inline Window outflow_bounds(int xdim, int ydim, const int bc[4], int boundary, int bc_periodic, int bc_open_moc, int n_ghost)
{
    auto lo = [&](int b) { return b == bc_periodic ? 0 : n_ghost; };
    auto hi = [&](int b, int n) { return b == bc_periodic ? n - 1 : n - n_ghost - 1; };
    Window w{lo(bc[0]), hi(bc[1], xdim), lo(bc[2]), hi(bc[3], ydim)};
    if (boundary == OB_X1 && bc[0] == bc_open_moc) w.xl -= n_ghost;
    else if (boundary == OB_X2 && bc[1] == bc_open_moc) w.xu += n_ghost;
    else if (boundary == OB_Y2 && bc[3] == bc_open_moc) w.yu += n_ghost;
    else if (boundary == OB_Y1 && bc[2] == bc_open_moc) w.yl -= n_ghost;
    return w;
}
// constructBoundaryAccel (boundaryoutflow.cpp:76-137)
inline void outflow_template(int xdim, int ydim, const Window &w, const double *x, const double *y, double length, double feather, int boundary, int shape,
                             std::vector<double> &out)
{
    const size_t n = (size_t)xdim * ydim;
    double xmin = x[0], xmax = x[0], ymin = y[0], ymax = y[0];
    for (size_t c = 0; c < n; c++) { xmin = std::min(xmin, x[c]); xmax = std::max(xmax, x[c]); ymin = std::min(ymin, y[c]); ymax = std::max(ymax, y[c]); }
    std::vector<double> res(n, 0.0);
    for (size_t c = 0; c < n; c++) {
        if (shape == OS_EXP)
            res[c] = boundary == OB_X1 ? std::exp((-2.3 * (x[c] - xmin)) / length) : boundary == OB_X2 ? std::exp((2.3 * (x[c] - xmax)) / length)
                   : boundary == OB_Y2 ? std::exp((2.3 * (y[c] - ymax)) / length) : std::exp((-2.3 * (y[c] - ymin)) / length);
        else if (shape == OS_GAUSSIAN) {
            const double q = boundary == OB_X1 ? (x[c] - xmin) / length : boundary == OB_X2 ? ((-x[c]) + xmax) / length
                           : boundary == OB_Y2 ? ((-y[c]) + ymax) / length : (y[c] - ymin) / length;
            res[c] = std::exp(-2.3 * (q * q));
        }
    }
    if (shape == OS_FLAT) {
        const double ext = boundary == OB_X1 ? xmin : boundary == OB_X2 ? xmax : boundary == OB_Y2 ? ymax : ymin;
        const double *p = boundary < OB_Y1 ? x : y;
        for (int i = w.xl; i <= w.xu; i++) for (int j = w.yl; j <= w.yu; j++) if (std::abs(p[(size_t)i * ydim + j] - ext) <= length) res[(size_t)i * ydim + j] = 1.0;
    }
    if (feather > 0.0) {
        const double *p = boundary < OB_Y1 ? y : x;
        const double pmax = boundary < OB_Y1 ? ymax : xmax, pmin = boundary < OB_Y1 ? ymin : xmin;
        for (size_t c = 0; c < n; c++) {
            const double a = std::max(p[c] - (pmax - 2.0 * feather), 0.0) / feather, b = std::min(p[c] - (pmin + 2.0 * feather), 0.0) / feather;
            res[c] *= std::max(std::exp(-2.3 * (a * a)) * std::exp(-2.3 * (b * b)) - 0.01, 0.0);
        }
    }
    out.assign(n, 0.0);
    for (int i = w.xl; i <= w.xu; i++) for (int j = w.yl; j <= w.yu; j++) out[(size_t)i * ydim + j] = 1.0 * res[(size_t)i * ydim + j];
}
// the index window of computeMeanOutflow (boundaryoutflow.cpp:140-213): the bounds above, narrowed by the feather and falloff lengths
inline Window outflow_mean_window(int ydim, Window w, const double *x, const double *y, double falloff, double feather, int boundary)
{
    auto X = [&](int i) { return x[(size_t)i * ydim]; };           // x(i, 0)
    auto Y = [&](int j) { return y[j]; };                           // y(0, j)
    if (boundary < OB_Y1) {
        for (int j = w.yl; j <= w.yu; j++) if (Y(j) - Y(w.yl) >= feather) { w.yl = j; break; }
        for (int j = w.yu; j >= w.yl; j--) if (Y(w.yu) - Y(j) >= feather) { w.yu = j; break; }
        if (boundary == OB_X1) { for (int i = w.xl; i <= w.xu; i++) if (X(i) - X(w.xl) >= falloff) { w.xu = i; break; } }
        else { for (int i = w.xu; i >= w.xl; i--) if (X(w.xu) - X(i) >= falloff) { w.xl = i; break; } }
    } else {
        for (int i = w.xl; i <= w.xu; i++) if (X(i) - X(w.xl) >= feather) { w.xl = i; break; }
        for (int i = w.xu; i >= w.xl; i--) if (X(w.xu) - X(i) >= feather) { w.xu = i; break; }
        if (boundary == OB_Y1) { for (int j = w.yl; j <= w.yu; j++) if (Y(j) - Y(w.yl) >= falloff) { w.yu = j; break; } }
        else { for (int j = w.yu; j >= w.yl; j--) if (Y(w.yu) - Y(j) >= falloff) { w.yl = j; break; } }
    }
    return w;
}

}  // namespace solar
}  // namespace spruce
