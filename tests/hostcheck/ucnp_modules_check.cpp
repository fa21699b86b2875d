// ucnp_modules_check.cpp -- TEST INFRASTRUCTURE.  Compiles the host shell's spruce_b200/host/ucnp_modules.hpp (the arithmetic of the host-resident coulomb_explosion and
// global_temperature modules) for tests/test_host_ucnp_modules.py, which compares it with the oracle's restatements (pinned to live runs of the reference binary).
#include "../../spruce_b200/host/ucnp_modules.hpp"
#include <cstring>
static Grid to_grid(const double *p, int nx, int ny)
{
    Grid g((size_t)nx, (size_t)ny);
    std::memcpy(g.ptr(), p, sizeof(double) * (size_t)nx * ny);
    return g;
}
extern "C" int ucnp_coulomb_force(const double *x, const double *y, const double *n, int nx, int ny, double time, double timescale, double lengthscale, double strength,
                                  double *fx, double *fy)
{
    Grid F_x, F_y;
    const std::string why = ucnp::coulombExplosionForce(to_grid(x, nx, ny), to_grid(y, nx, ny), to_grid(n, nx, ny), time, timescale, lengthscale, strength, F_x, F_y);
    if (!why.empty()) return 1;
    std::memcpy(fx, F_x.ptr(), sizeof(double) * (size_t)nx * ny);
    std::memcpy(fy, F_y.ptr(), sizeof(double) * (size_t)nx * ny);
    return 0;
}
typedef void (*laplacian_fn)(const double *in, double *out);
extern "C" int ucnp_diffuse_temperature(double *temp, const double *d_x, const double *d_y, const double *mask, int nx, int ny, double dt, double epsilon, double strength, laplacian_fn lap)
{
    Grid t = to_grid(temp, nx, ny);
    const Grid dr = ucnp::diffusionLengthSquared(to_grid(d_x, nx, ny), to_grid(d_y, nx, ny), to_grid(mask, nx, ny));
    const int steps = ucnp::diffuseTemperature(t, dr, dt, epsilon, strength, [&](const Grid &q) { Grid out((size_t)nx, (size_t)ny); lap(q.ptr(), out.ptr()); return out; });
    std::memcpy(temp, t.ptr(), sizeof(double) * (size_t)nx * ny);
    return steps;
}
