// exact_math_check.cpp -- TEST INFRASTRUCTURE.  Compiles the product's device arithmetic header (spruce_b200/csrc/exact_math.cuh: the correctly
// rounded division sequence and the three forms of the Barton face value the stage kernels use) for the HOST, with one-line stand-ins for the
// handful of CUDA intrinsics it touches, so that tests/test_exact_math_host.py can hold every form to the reference's upwindSurface formula
// (source/mhd/derivs.cpp:47-68) bit for bit on random and adversarial operands -- a regression guard for kernel work that needs no GPU.
#include <cmath>
#include <cstdint>
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
#include "../../spruce_b200/csrc/exact_math.cuh"

using namespace spruce;

// the reference's formula, written out (derivs.cpp:47-68 with boundaryInterpolate / boundaryExtrapolate :477-499)
static double ref_face(double qm2, double qm1, double q0, double qp1, double vf, double hm2, double hm1, double h0, double hp1, double *d2out)
{
    const double d2 = (qm1 * h0 + q0 * hm1) / (h0 + hm1);
    *d2out = d2;
    if (vf > 0.0) {
        const double d3 = qm1, d1 = qm2 + (qm1 - qm2) * (hm2 + 2.0 * hm1) / (hm2 + hm1);
        return (q0 <= qm1) ? ((std::max(d1, d2) < d3) ? std::max(d1, d2) : d3) : ((d3 < std::min(d1, d2)) ? std::min(d1, d2) : d3);
    }
    if (vf < 0.0) {
        const double d3 = q0, d1 = qp1 + (q0 - qp1) * (hp1 + 2.0 * h0) / (hp1 + h0);
        return (q0 <= qm1) ? ((d3 < std::min(d1, d2)) ? std::min(d1, d2) : d3) : ((std::max(d1, d2) < d3) ? std::max(d1, d2) : d3);
    }
    return d2;
}

// n cases; q[4n] = qm2 qm1 q0 qp1, h[4n] = half sizes of the same cells, vf[n].  out[5n]: reference face value, d2, and the face values of
// upwind_face / upwind_face_sel / upwind_face_far (the last one is only defined for vf != 0: the caller multiplies by vf, so it is compared there).
extern "C" void exact_math_faces(int n, const double *q, const double *h, const double *vf, double *out)
{
    for (int k = 0; k < n; k++) {
        const double qm2 = q[4 * k], qm1 = q[4 * k + 1], q0 = q[4 * k + 2], qp1 = q[4 * k + 3];
        const double hm2 = h[4 * k], hm1 = h[4 * k + 1], h0 = h[4 * k + 2], hp1 = h[4 * k + 3];
        FaceGeom g;
        g.hm1 = hm1; g.h0 = h0; g.fs = h0 + hm1; g.rfs = 1.0 / g.fs;
        g.ep = hm2 + 2.0 * hm1; g.fsm = hm2 + hm1; g.rfsm = 1.0 / g.fsm;
        g.em = hp1 + 2.0 * h0; g.fsp = hp1 + h0; g.rfsp = 1.0 / g.fsp;
        double d2r, d2a, d2b, d2c;
        out[5 * k] = ref_face(qm2, qm1, q0, qp1, vf[k], hm2, hm1, h0, hp1, &d2r);
        out[5 * k + 1] = d2r;
        out[5 * k + 2] = upwind_face(qm2, qm1, q0, qp1, vf[k], g, &d2a);
        const FaceSel s = select_face(g, vf[k]);
        out[5 * k + 3] = upwind_face_sel(qm2, qm1, q0, qp1, s, &d2b);
        out[5 * k + 4] = upwind_face_far(s.pos ? qm2 : qp1, qm1, q0, s, &d2c);
        if (!(d2a == d2r || (d2a != d2a && d2r != d2r)) || !(d2b == d2r || (d2b != d2b && d2r != d2r)) || !(d2c == d2r || (d2c != d2c && d2r != d2r))) out[5 * k + 1] = std::nan("");
    }
}
extern "C" void exact_math_div(int n, const double *a, const double *b, double *out)
{
    for (int k = 0; k < n; k++) out[k] = ddiv(a[k], b[k], 1.0 / b[k]);
}
extern "C" int exact_math_all_zero8(const double *v) { return all_zero8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]) ? 1 : 0; }
