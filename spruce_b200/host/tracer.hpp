// tracer.hpp -- bilinearInterpolate of the reference (source/mhd/utils.cpp:55-75) for the host-resident tracer_particles module: the value of a
// host Grid at a point, bilinear in the cell-centre coordinates x (rows) / y (columns).  Same branches and the same expression order as the
// reference: std::upper_bound brackets, 0.0 past the last coordinate, the clamped lower index at the first one (where the reference divides by a
// zero width -- reproduced).  tests/test_host_tracer.py compares it with the reference's own function, compiled from its source as test infrastructure.
#pragma once
#include "grid.hpp"
#include <algorithm>
#include <vector>

inline double bilinearInterpolate(const std::vector<double> &point, const Grid &quantity, const std::vector<double> &x, const std::vector<double> &y)
{
    const auto it_x = std::upper_bound(x.cbegin(), x.cend(), point[0]);
    const auto it_y = std::upper_bound(y.cbegin(), y.cend(), point[1]);
    if (it_x == x.cend() || it_y == y.cend()) return 0.0;
    const int i_1 = (int)std::distance(x.cbegin(), it_x), i_0 = std::max(i_1 - 1, 0);
    const int j_1 = (int)std::distance(y.cbegin(), it_y), j_0 = std::max(j_1 - 1, 0);
    const double dx0 = point[0] - x[i_0], dx1 = x[i_1] - point[0];
    const double dy0 = point[1] - y[j_0], dy1 = y[j_1] - point[1];
    return (quantity(i_0, j_0) * dx1 * dy1 + quantity(i_1, j_0) * dx0 * dy1 + quantity(i_0, j_1) * dx1 * dy0 + quantity(i_1, j_1) * dx0 * dy0)
           / ((x[i_1] - x[i_0]) * (y[j_1] - y[j_0]));
}
