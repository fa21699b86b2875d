// viscosity_profile.hpp -- Viscosity::getBoundaryViscosity of the reference (source/modules/viscosity.cpp:278-325) on host Grids: the static strength profile of a
// `boundary` / `boundary_global` term of artificial_viscosity, for its four shapes.  Built on the host with the host libm, as the reference builds it, and handed to the
// device as a plane (spruce_module_viscosity_term).  Same expression order as the reference's Grid arithmetic; tests/test_host_viscosity_profile.py holds it bit for bit
// to the restatement of the test infrastructure, which is pinned to live runs of the reference binary for all four shapes.
#pragma once
#include "grid.hpp"
#include <algorithm>
#include <cmath>
#include <string>

// returns false for a shape the reference does not know (it asserts, :304)
inline bool boundaryViscosityProfile(const Grid &x, const Grid &y, double strength, double length, const std::string &shape, Grid &result)
{
    const size_t n = (size_t)x.size();
    const double *px = x.ptr(), *py = y.ptr();
    double x_min = px[0], x_max = px[0], y_min = py[0], y_max = py[0];
    for (size_t c = 0; c < n; c++) { x_min = std::min(x_min, px[c]); x_max = std::max(x_max, px[c]); y_min = std::min(y_min, py[c]); y_max = std::max(y_max, py[c]); }
    result = Grid((size_t)x.rows(), (size_t)x.cols());
    double *r = result.ptr();
    if (shape == "gaussian") {                                                     // :286-295: the four sides superposed
        for (size_t c = 0; c < n; c++) {
            const double a[4] = {(px[c] - x_min) / length, (px[c] - x_max) / length, (py[c] - y_max) / length, (py[c] - y_min) / length};
            double v = 0.0;
            for (double q : a) v = v + std::exp((q * q) * -2.3) * strength;
            r[c] = v;
        }
    } else if (shape == "exp") {                                                   // :296-305
        for (size_t c = 0; c < n; c++) {
            double v = 0.0;
            v = v + std::exp(((px[c] - x_min) * -2.3) / length) * strength;
            v = v + std::exp(((px[c] - x_max) * 2.3) / length) * strength;
            v = v + std::exp(((py[c] - y_max) * 2.3) / length) * strength;
            v = v + std::exp(((py[c] - y_min) * -2.3) / length) * strength;
            r[c] = v;
        }
    } else if (shape == "exp_elliptical" || shape == "gaussian_elliptical") {       // :306-319: 1 on the ellipse through the domain's edge midpoints, 0 at the centre
        const double x_center = 0.5 * (x_min + x_max), y_center = 0.5 * (y_min + y_max);
        const double ax2 = std::pow(x_max - x_center, 2.0), ay2 = std::pow(y_max - y_center, 2.0);
        const double s_length = length / std::min(x_max - x_center, y_max - y_center);          // the scale length applies to the semi-minor axis
        const bool expo = shape == "exp_elliptical";
        for (size_t c = 0; c < n; c++) {
            const double s = ((px[c] - x_center) * (px[c] - x_center)) / ax2 + ((py[c] - y_center) * (py[c] - y_center)) / ay2;
            const double m = std::min(s - 1.0, 0.0);
            if (expo) r[c] = std::exp((m * 2.3) / s_length) * strength;
            else { const double q = m / s_length; r[c] = std::exp((q * q) * -2.3) * strength; }
        }
    } else return false;
    for (size_t c = 0; c < n; c++) r[c] = (strength < r[c]) ? strength : r[c];     // result.min(strength), :324
    return true;
}
