// cell_math_check.cpp -- TEST INFRASTRUCTURE.  Compiles spruce_b200/csrc/cell_math.cuh (density floor, recomputeDT of one cell, the dt skip test of the
// stage kernel) for the HOST; tests/test_cell_math_host.py checks cell_dt against the reference's formula (idealmhd.cpp:279-304) bit for bit and the
// soundness of dt_can_skip (a skipped cell never has a dt below the threshold) on random states.
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __host__
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
static inline int __double2hiint(double x) { return (int)(__double_as_longlong(x) >> 32); }
static inline int __double2loint(double x) { return (int)(__double_as_longlong(x) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo)); }
#define SPRUCE_EXACT_MATH_HOST_CHECK 1
#include "../../spruce_b200/csrc/cell_math.cuh"

using namespace spruce;

static DomainParams params(double m_i, double gamma, double n_min)
{
    DomainParams P;
    std::memset(&P, 0, sizeof(P));
    P.m_i = m_i; P.rm_i = 1.0 / m_i; P.gamma = gamma; P.gm1 = gamma - 1.0; P.n_min = n_min;
    P.fourpi = 4.0 * 3.14159265358979323846; P.rfourpi = 1.0 / P.fourpi;
    return P;
}
// recomputeDT (idealmhd.cpp:279-304) with every operation written out; inputs as the kernel has them: rho, momenta, thermal energy, total field
static double ref_dt(const DomainParams &P, double rho, double mx, double my, double e, double bx, double by, double bz, double dx, double dy)
{
    const double vx = mx / rho, vy = my / rho;
    const double press = e * P.gm1;
    const double bm = std::sqrt((bx * bx + by * by) + bz * bz);
    const double cs = std::sqrt(P.gamma * press / rho), cs2 = cs * cs;
    const double va = bm / std::sqrt(rho * P.fourpi), va2 = va * va;
    const double s = cs2 + va2;
    const double delta = std::sqrt(1.0 - ((cs2 * 4.0) * va2) / (s * s));
    const double vfast = std::sqrt((s * 0.5) * (1.0 + delta)), vslow = std::sqrt((s * 0.5) * (1.0 - delta));
    const double vmx = std::sqrt(vx * vx), vmy = std::sqrt(vy * vy);
    double M = cs; M = (M < va) ? va : M; M = (M < vfast) ? vfast : M; M = (M < vslow) ? vslow : M;
    return 1.0 / ((vmx + M) / dx + (vmy + M) / dy);
}
// u[n][9] = rho, mx, my, e, bx, by, bz, dx, dy.  out[n][2] = cell_dt, reference dt; skip[n] = dt_can_skip(R)
extern "C" void cell_math_dt(int n, const double *u, double m_i, double gamma, double R, double *out, int *skip)
{
    const DomainParams P = params(m_i, gamma, 1.0);
    for (int k = 0; k < n; k++) {
        const double *c = u + 9 * k;
        out[2 * k] = cell_dt(P, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], 1.0 / c[7], c[8], 1.0 / c[8]);
        out[2 * k + 1] = ref_dt(P, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]);
        skip[k] = dt_can_skip(P, R, c[0], c[1], c[2], c[3], c[4], c[5], c[6], 1.0 / c[7], 1.0 / c[8]) ? 1 : 0;
    }
}
// enforceMinimums + the derived step for rho (idealmhd.cpp:237, 246-247): out[n][4] = n and post-floor rho of density_floor, and the reference's two values
extern "C" void cell_math_floor(int n, const double *rho_u, double m_i, double n_min, double *out)
{
    const DomainParams P = params(m_i, 5.0 / 3.0, n_min);
    for (int k = 0; k < n; k++) {
        double r1;
        out[4 * k] = density_floor(P, rho_u[k], &r1);
        out[4 * k + 1] = r1;
        const double a = rho_u[k] / m_i;
        const double rr = m_i * ((a < n_min) ? n_min : a);                  // rho = m_i*max(rho/m_i, n_min)
        const double b = rr / m_i;
        out[4 * k + 2] = (b < n_min) ? n_min : b;                          // n = max(rho/m_i, n_min)
        out[4 * k + 3] = rr;
    }
}
