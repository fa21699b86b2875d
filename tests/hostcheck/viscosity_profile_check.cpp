// viscosity_profile_check.cpp -- TEST INFRASTRUCTURE.  Compiles the host shell's spruce_b200/host/viscosity_profile.hpp for tests/test_host_viscosity_profile.py.
#include "../../spruce_b200/host/viscosity_profile.hpp"
#include <cstring>
extern "C" int boundary_viscosity_profile(const double *x, const double *y, int nx, int ny, double strength, double length, const char *shape, double *out)
{
    Grid gx((size_t)nx, (size_t)ny), gy((size_t)nx, (size_t)ny), r;
    std::memcpy(gx.ptr(), x, sizeof(double) * (size_t)nx * ny);
    std::memcpy(gy.ptr(), y, sizeof(double) * (size_t)nx * ny);
    if (!boundaryViscosityProfile(gx, gy, strength, length, shape, r)) return 1;
    std::memcpy(out, r.ptr(), sizeof(double) * (size_t)nx * ny);
    return 0;
}
