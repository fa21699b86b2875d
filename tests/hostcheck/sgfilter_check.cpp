// sgfilter_check.cpp -- TEST INFRASTRUCTURE.  Compiles the host shell's spruce_b200/host/sgfilter.hpp (the host-resident sg_filtering module's filter) for
// tests/test_host_sgfilter.py, which compares it with the oracle's restatement (pinned to live runs of the reference binary).
#include "../../spruce_b200/host/sgfilter.hpp"
#include <cstring>
extern "C" void sgfilter_apply(double *plane, int nx, int ny, int xl, int xu, int yl, int yu, int y_periodic)
{
    Grid g((size_t)nx, (size_t)ny);
    std::memcpy(g.ptr(), plane, sizeof(double) * (size_t)nx * ny);
    sgFilterPlane(g, xl, xu, yl, yu, y_periodic != 0);
    std::memcpy(plane, g.ptr(), sizeof(double) * (size_t)nx * ny);
}
