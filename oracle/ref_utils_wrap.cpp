// ref_utils_wrap.cpp -- TEST INFRASTRUCTURE.  C entry point over the REFERENCE's own bilinearInterpolate (source/mhd/utils.cpp:55-75), compiled together with
// the reference's utils.cpp and grid.cpp where they lie (oracle/Makefile: refutils -> oracle/_ref/libref_utils.so).  tests/test_host_tracer.py holds the host
// shell's restatement (spruce_b200/host/tracer.hpp) to it bit for bit.  No reference source is copied.
#include "utils.hpp"
#include "grid.hpp"
#include <vector>
extern "C" double ref_bilinear(double px, double py, const double *q, int nx, int ny, const double *x, const double *y)
{
    Grid g(nx, ny);
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) g(i, j) = q[(size_t)i * ny + j];
    return bilinearInterpolate({px, py}, g, std::vector<double>(x, x + nx), std::vector<double>(y, y + ny));
}
