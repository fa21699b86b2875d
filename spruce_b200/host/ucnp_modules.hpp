// ucnp_modules.hpp -- the arithmetic of the two host-resident UCNP modules of the drop-in shell, on host Grids:
//   coulombExplosionForce   CoulombExplosion::postIterateModule (source/modules/ucnp/coulomb_explosion.cpp:49-87) without its state update: radial profile of n
//                           over 101 bins, charge density of a non-neutrality that decays in time and radius, enclosed charge by a running trapezoid,
//                           E = Q r_vec / r^3, force density rho_c E
//   diffuseTemperature      GlobalTemperature::postIterateModule (source/modules/ucnp/global_temperature.cpp:66-94): `strength` midpoint sub-steps of
//                           d(temp)/dt = coeff * laplacian(temp); the laplacian is the caller's (the device operator in the shell)
// Both are global and serial by nature (a radial histogram, a profile look-up per cell) and run once per step: they stay on the host, staged through the C ABI like
// every un-ported module.  Results equal the reference's with one OpenMP thread bit for bit (its Bin1D sums into shared bins from an
// unsynchronised parallel loop, grid.cpp:306-315, so its own multi-thread result is schedule dependent): tests/test_host_ucnp_modules.py against the test
// infrastructure's restatement, which is pinned to live runs of the reference binary.  The reference's per-cell scans over all bins (Interp1D, Bin1D) are
// replaced by binary searches over the increasing bin vectors -- the selected bins and the rounded expressions are the same.
#pragma once
#include "grid.hpp"
#include <algorithm>
#include <cmath>
#include <functional>
#include <string>
#include <vector>

namespace ucnp {

constexpr double kPi = 3.14159265358979323846;      // PI, source/constants.hpp:16
constexpr double kE = 4.80320425e-10;               // E, source/constants.hpp:19
constexpr double kBoltzmann = 1.3807e-16;           // K_B, source/constants.hpp:8
constexpr int kBins = 101;                          // coulomb_explosion.cpp:63

// Grid::Linspace (grid.cpp:24-35)
inline std::vector<double> linspace(double start, double end, int num)
{
    std::vector<double> v((size_t)num);
    const double spacing = (end - start) / (num - 1);
    for (int i = 0; i < num; i++) v[(size_t)i] = start + i * spacing;
    return v;
}

// Grid::Interp1D (grid.cpp:332-375) for one query over a strictly increasing abscissa: a node returns its value (the first node is counted twice by the
// reference, :347-349: d/2 + d/2), anything else the chord between its two neighbours, rounded as  lo + (q - x_lo) * (hi - lo) / (x_hi - x_lo)
inline double interpolateProfile(const std::vector<double> &x, const std::vector<double> &y, double q)
{
    const size_t k = (size_t)(std::lower_bound(x.begin(), x.end(), q) - x.begin());          // first node >= q
    if (k < x.size() && x[k] == q) return k == 0 ? y[0] / 2 + y[0] / 2 : y[k];
    const double lo = y[k - 1], hi = y[k];
    return lo + (q - x[k - 1]) * (hi - lo) / (x[k] - x[k - 1]);
}

// Returns "" and the force density (F_x, F_y) -- or the reason the reference aborts on this grid (an empty radial bin: it then multiplies vectors of unequal
// length, coulomb_explosion.cpp:39; a radius outside the bin centres: Interp1D's assertion, grid.cpp:335).
inline std::string coulombExplosionForce(const Grid &x, const Grid &y, const Grid &n, double time, double timescale, double lengthscale, double strength, Grid &F_x, Grid &F_y)
{
    const size_t cells = (size_t)n.size();
    std::vector<double> r_sq(cells), r(cells);
    const double *px = x.ptr(), *py = y.ptr(), *pn = n.ptr();
    for (size_t c = 0; c < cells; c++) { r_sq[c] = px[c] * px[c] + py[c] * py[c]; r[c] = std::sqrt(r_sq[c]); }
    const auto [rmin_it, rmax_it] = std::minmax_element(r.begin(), r.end());
    if (!(*rmax_it > *rmin_it)) return "coulomb_explosion: all cells at one radius";
    // Bin1D(r, n, 101, r_bin): centres, then edges half a spacing outside them (grid.cpp:324-330); each cell goes to the first bin whose closed interval holds it
    const std::vector<double> r_bin = linspace(*rmin_it, *rmax_it, kBins);
    const double dr = r_bin[1] - r_bin[0];
    const std::vector<double> edges = linspace(r_bin.front() - dr / 2., r_bin.back() + dr / 2., kBins + 1);
    std::vector<double> count(kBins, 0.0), n_bin(kBins, 0.0);
    for (size_t c = 0; c < cells; c++) {                                       // cell order: the sum the reference forms with one thread
        if (!(r[c] >= r_bin.front() && r[c] <= r_bin.back())) return "coulomb_explosion: a cell radius lies outside the bin centres (the reference's Interp1D asserts)";
        const size_t j = (size_t)(std::lower_bound(edges.begin() + 1, edges.end(), r[c]) - (edges.begin() + 1));       // first bin with r <= upper edge
        if (j >= (size_t)kBins || !(r[c] >= edges[j])) continue;
        count[j]++; n_bin[j] += pn[c];
    }
    for (int j = 0; j < kBins; j++) {
        n_bin[(size_t)j] /= count[(size_t)j];
        if (!std::isfinite(n_bin[(size_t)j])) return "coulomb_explosion: radial bin " + std::to_string(j) + " of 101 is empty on this grid (the reference aborts: its profile vectors then differ in length)";
    }
    // compute_charge_density (:34-40), compute_total_charge (:42-45)
    const double amp = (strength * std::exp(-time / timescale)) * kE;
    std::vector<double> rho_c_vec(kBins), Q_vec(kBins), shell(kBins);
    for (int j = 0; j < kBins; j++) rho_c_vec[(size_t)j] = (amp * n_bin[(size_t)j]) * std::exp(-r_bin[(size_t)j] / lengthscale);
    for (int j = 0; j < kBins; j++) shell[(size_t)j] = (r_bin[(size_t)j] * r_bin[(size_t)j]) * rho_c_vec[(size_t)j];
    Q_vec[0] = 0.0;
    for (int j = 0; j + 1 < kBins; j++) Q_vec[(size_t)j + 1] = Q_vec[(size_t)j] + (shell[(size_t)j] + shell[(size_t)j + 1]) * (r_bin[(size_t)j + 1] - r_bin[(size_t)j]) / 2.;
    for (double &q : Q_vec) q = q * (4. * kPi);
    // :71-81
    F_x = Grid((size_t)n.rows(), (size_t)n.cols()); F_y = F_x;
    double *fx = F_x.ptr(), *fy = F_y.ptr();
#pragma omp parallel for schedule(static)
    for (long long cc = 0; cc < (long long)cells; cc++) {
        const size_t c = (size_t)cc;
        const double Q = interpolateProfile(r_bin, Q_vec, r[c]), rho_c = interpolateProfile(r_bin, rho_c_vec, r[c]);
        const double r_cubed = r_sq[c] * r[c];
        fx[c] = rho_c * (px[c] * Q / r_cubed);
        fy[c] = rho_c * (py[c] * Q / r_cubed);
    }
    return "";
}

// dr = (d_x^2 + d_y^2) * mask / 2 (global_temperature.cpp:42)
inline Grid diffusionLengthSquared(const Grid &d_x, const Grid &d_y, const Grid &mask)
{
    Grid dr((size_t)d_x.rows(), (size_t)d_x.cols());
    for (size_t c = 0; c < (size_t)dr.size(); c++) dr.ptr()[c] = (d_x.ptr()[c] * d_x.ptr()[c] + d_y.ptr()[c] * d_y.ptr()[c]) * mask.ptr()[c] / 2.;
    return dr;
}
// global_temperature.cpp:71-89 for one species: temp is advanced in place; returns the number of midpoint sub-steps taken
inline int diffuseTemperature(Grid &temp, const Grid &dr, double dt, double epsilon, double strength, const std::function<Grid(const Grid &)> &laplacian)
{
    const double dt_fluid = dt / epsilon;
    const double dt_diff = dt_fluid / strength;
    const double dt_rk = dt_diff * epsilon;
    const int num_rk_steps = (int)(dt / dt_rk);
    const size_t cells = (size_t)temp.size();
    std::vector<double> coeff(cells);
    for (size_t c = 0; c < cells; c++) coeff[c] = dr.ptr()[c] / dt_diff;
    Grid mid = temp;
    for (int s = 0; s < num_rk_steps; s++) {
        const Grid lap0 = laplacian(temp);
        for (size_t c = 0; c < cells; c++) mid.ptr()[c] = temp.ptr()[c] + (coeff[c] * lap0.ptr()[c]) * dt_rk / 2.;
        const Grid lap1 = laplacian(mid);
        for (size_t c = 0; c < cells; c++) temp.ptr()[c] += (coeff[c] * lap1.ptr()[c]) * dt_rk;
    }
    return num_rk_steps;
}

}  // namespace ucnp
