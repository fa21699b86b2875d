// solar_templates.hpp -- the static spatial templates of the pointwise solar source-term modules, built on the host exactly as the
// reference builds them (host libm: the reference binary and this library link the same glibc):
//   SolarUtils::GaussianGrid / GaussianGridRotated     source/solar/solarutils.cpp:65-98     (coordinates = GLOBAL cell indices)
//   LocalizedHeating / MassInjection templates          source/modules/solar/localizedheating.cpp:31-50, massinjection.cpp:27-46
//   MomentumInjection templates                         source/modules/solar/momentuminjection.cpp:35-67
// Plain C++ (no CUDA): capi.cu includes it, and tests/hostcheck compiles it with g++ to check it without a GPU.
#pragma once
#include <cmath>
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

}  // namespace solar
}  // namespace spruce
