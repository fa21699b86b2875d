// sgfilter.hpp -- SGFilter::singleVarSavitzkyGolay of the reference (source/modules/sgfilter.cpp:46-82) on a host Grid, as the reference computes it:
// a 5 x 5 Savitzky-Golay window (3 x 3-order polynomials, coefficients of Chandra Sekhar 2015) over the iteration bounds -- with the reference's
// indexing as written: the wrapped row indices i[] are never read, every tap of the window is grid(j[v], j[v]) (sgfilter.cpp:75).  Reproduced, not
// repaired: the drop-in must write the files the reference writes.  tests/test_host_sgfilter.py checks it against the CPU restatement of the test infrastructure, which is
// pinned to live runs of the reference binary.
#pragma once
#include "grid.hpp"

inline void sgFilterPlane(Grid &grid, int xl, int xu, int yl, int yu, bool y_periodic)
{
    static const double coeff[] =
        {+7.346939E-03, -2.938776E-02, -4.163265E-02, -2.938776E-02, +7.346939E-03,
         -2.938776E-02, +1.175510E-01, +1.665306E-01, +1.175510E-01, -2.938776E-02,
         -4.163265E-02, +1.665306E-01, +2.359184E-01, +1.665306E-01, -4.163265E-02,
         -2.938776E-02, +1.175510E-01, +1.665306E-01, +1.175510E-01, -2.938776E-02,
         +7.346939E-03, -2.938776E-02, -4.163265E-02, -2.938776E-02, +7.346939E-03};
    const int ydim = grid.cols();
    Grid filtered = grid;
    for (int ci = xl; ci <= xu; ci++) {
        for (int cj = yl; cj <= yu; cj++) {
            int j[5] = {cj - 2, cj - 1, cj, cj + 1, cj + 2};
            if (y_periodic) for (int k : {0, 1, 3, 4}) j[k] = (j[k] + ydim) % ydim;
            double v = 0.0;
            for (int u = 0; u < 5; u++)
                for (int w = 0; w < 5; w++) v = v + coeff[u * 5 + w] * grid((size_t)j[w], (size_t)j[w]);
            filtered((size_t)ci, (size_t)cj) = v;
        }
    }
    grid = filtered;
}
