// tracer_check.cpp -- TEST INFRASTRUCTURE.  Compiles the host shell's spruce_b200/host/tracer.hpp (bilinearInterpolate of the host-resident tracer_particles
// module) for tests/test_host_tracer.py.
#include "../../spruce_b200/host/tracer.hpp"
#include <cstring>
extern "C" double ours_bilinear(double px, double py, const double *q, int nx, int ny, const double *x, const double *y)
{
    Grid g((size_t)nx, (size_t)ny);
    std::memcpy(g.ptr(), q, sizeof(double) * (size_t)nx * ny);
    return bilinearInterpolate({px, py}, g, std::vector<double>(x, x + nx), std::vector<double>(y, y + ny));
}
