// exact_math.cuh -- individually rounded FP64 building blocks for bit-exact parity with the reference.
//
// The reference binary (g++ -O3, x86-64 baseline ISA) contains no FMA: every +,-,*,/ and sqrt is a
// separately rounded IEEE-754 binary64 operation.  This translation unit is therefore
// compiled with -fmad=false, and the only fused operations are the explicit fma() calls below, which
// implement a CORRECTLY ROUNDED division and nothing else.
//
// ddiv(a, b, rb): RN(a / b) given rb = RN(1 / b) (computed once, by a true IEEE division, for every
// divisor that depends only on the 1-D cell-size tables or on constants).  Five FP64-pipe instructions
// instead of the ~10 + slow-path check of div.rn.f64:
//     q0 = RN(a*rb)                      |q0 - a/b| <= 1.5 ulp
//     e0 = RN(a - b*q0)  (fma)           q1 = RN(q0 + e0*rb)  -> faithful (|q1 - a/b| < 1 ulp)
//     e1 = a - b*q1      (fma, exact)    q2 = RN(q1 + e1*rb)  -> = RN(a/b)   [Markstein 1990; Muller et al.,
//                                                            Handbook of Floating-Point Arithmetic, Thm "correction
//                                                            step with a correctly rounded reciprocal"]
// Valid for finite a, b with a/b in the normal range (physical CGS magnitudes here are ~1e-30..1e+30).
// a = 0 gives 0 (the sign of a zero result may differ from IEEE division; zeros compare equal).
// tests/test_exact_division.py checks the sequence exhaustively in reduced precision and on 1e8 random and
// adversarial binary64 operand pairs against true division.
#pragma once
#ifndef SPRUCE_EXACT_MATH_HOST_CHECK      // tests/hostcheck/exact_math_check.cpp compiles this header for the host with stand-ins for the intrinsics
#include <cuda_runtime.h>
#endif

namespace spruce {

#ifdef SPRUCE_RELAXED
// relaxed arithmetic (stage_relaxed.cu only): one multiplication by the correctly rounded reciprocal, |error| <= 1.5 ulp
__device__ __forceinline__ double ddiv(double a, double b, double rb) { (void)b; return a * rb; }
#else
__device__ __forceinline__ double ddiv(double a, double b, double rb)
{
    double q = a * rb;
    double e = fma(-b, q, a);
    q = fma(e, rb, q);
    e = fma(-b, q, a);
    return fma(e, rb, q);
}
#endif

// std::min / std::max semantics of the reference's Grid::min/max (source/mhd/grid.cpp:110-163):
// std::min(a,b) = (b < a) ? b : a ; std::max(a,b) = (a < b) ? b : a  -- NaN handling differs from fmin/fmax.
__device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }

// boundaryInterpolate (source/mhd/derivs.cpp:477-487): value at the face between cell A (lower index) and cell B:
// (a*dist_b + b*dist_a) / (dist_b + dist_a), dist = half cell size; fs = hb + ha, rfs = RN(1/fs).
__device__ __forceinline__ double face_interp(double a, double b, double ha, double hb, double fs, double rfs)
{
    return ddiv(a * hb + b * ha, fs, rfs);
}

// Geometry of one face f (between cells f-1 and f) along one axis, all from the 1-D cell-size table.
struct FaceGeom {
    double hm1;   // h[f-1]
    double h0;    // h[f]
    double fs;    // h[f] + h[f-1]
    double rfs;   // RN(1/fs)
    double ep;    // h[f-2] + 2*h[f-1]   numerator weight of boundaryExtrapolate for flow in +direction
    double fsm;   // h[f-2] + h[f-1]     (= fs of face f-1)
    double rfsm;
    double em;    // h[f+1] + 2*h[f]     ... for flow in -direction
    double fsp;   // h[f+1] + h[f]       (= fs of face f+1)
    double rfsp;
};

// Barton's monotone upwind face value, upwindSurface (source/mhd/derivs.cpp:47-68).
// qm2..qp1 = q[f-2], q[f-1], q[f], q[f+1]; vf = interpolated face velocity. Returns S; *d2 gets the linear interpolation.
__device__ __forceinline__ double upwind_face(double qm2, double qm1, double q0, double qp1, double vf,
                                              const FaceGeom &g, double *d2out)
{
    const double d2 = face_interp(qm1, q0, g.hm1, g.h0, g.fs, g.rfs);
    *d2out = d2;
    const bool pos = vf > 0.0, neg = vf < 0.0;
    // boundaryExtrapolate (derivs.cpp:490-499): a + (b-a)*(dist_a + 2*dist_b)/(dist_a + dist_b)
    const double a = pos ? qm2 : qp1;
    const double b = pos ? qm1 : q0;          // also d3, the upwind cell value
    const double w = pos ? g.ep : g.em;
    const double den = pos ? g.fsm : g.fsp;
    const double rden = pos ? g.rfsm : g.rfsp;
    const double d1 = a + ddiv((b - a) * w, den, rden);
    const bool le = (q0 <= qm1);
    const double rmin = smin(b, smax(d1, d2));
    const double rmax = smax(b, smin(d1, d2));
    const double r = (pos == le) ? rmin : rmax;
    return (pos || neg) ? r : d2;
}

// ---- the same face value with everything that depends only on the face (not on the transported quantity) hoisted:
// the sign of the face velocity selects the upwind side once per face, not once per quantity.
struct FaceSel {
    double hm1, h0, fs, rfs;   // interpolation weights of the face
    double w, den, rden;       // extrapolation weight / divisor of the upwind side (ep,fsm | em,fsp)
    bool pos, any;             // vf > 0 ; vf != 0 (strictly: vf > 0 || vf < 0, NaN -> false)
};
__device__ __forceinline__ FaceSel select_face(const FaceGeom &g, double vf)
{
    FaceSel s;
    s.hm1 = g.hm1; s.h0 = g.h0; s.fs = g.fs; s.rfs = g.rfs;
    s.pos = vf > 0.0;
    s.any = s.pos || (vf < 0.0);
    s.w = s.pos ? g.ep : g.em;
    s.den = s.pos ? g.fsm : g.fsp;
    s.rden = s.pos ? g.rfsm : g.rfsp;
    return s;
}
__device__ __forceinline__ double flip_sign(double x, unsigned long long m) { return __longlong_as_double(__double_as_longlong(x) ^ (long long)m); }

// upwindSurface (derivs.cpp:47-68) for one quantity.  std::max(d3, std::min(d1,d2)) == -std::min(-d3, std::max(-d1,-d2))
// selection for selection (the comparisons are mirrored exactly, NaNs and equal values included), so only ONE
// min/max pair is evaluated, on operands whose sign bit is flipped when the max-outer form is the one required.
__device__ __forceinline__ double upwind_face_sel(double qm2, double qm1, double q0, double qp1, const FaceSel &f, double *d2out)
{
    const double d2 = face_interp(qm1, q0, f.hm1, f.h0, f.fs, f.rfs);
    *d2out = d2;
    const double a = f.pos ? qm2 : qp1;
    const double b = f.pos ? qm1 : q0;          // d3, the upwind cell value
    const double d1 = a + ddiv((b - a) * f.w, f.den, f.rden);
    const bool min_outer = (f.pos == (q0 <= qm1));
    const unsigned long long m = min_outer ? 0ULL : 0x8000000000000000ULL;
    const double fb = flip_sign(b, m), f1 = flip_sign(d1, m), f2 = flip_sign(d2, m);
    const double r = flip_sign(smin(fb, smax(f1, f2)), m);
    return f.any ? r : d2;
}

// Same face value when the caller has already picked the far cell of the extrapolation (q[f-2] for flow in the + direction,
// q[f+1] for the - direction): far, qm1 = q[f-1], q0 = q[f].
__device__ __forceinline__ double upwind_face_far(double far, double qm1, double q0, const FaceSel &f, double *d2out)
{
    const double d2 = face_interp(qm1, q0, f.hm1, f.h0, f.fs, f.rfs);
    *d2out = d2;
    const double b = f.pos ? qm1 : q0;          // d3, the upwind cell value
    const double d1 = far + ddiv((b - far) * f.w, f.den, f.rden);
    const bool min_outer = (f.pos == (q0 <= qm1));
    const unsigned hm = min_outer ? 0u : 0x80000000u;       // flip the sign bit (high word only) when the max-outer form is required
    const double fb = __hiloint2double(__double2hiint(b) ^ hm, __double2loint(b));
    const double f1 = __hiloint2double(__double2hiint(d1) ^ hm, __double2loint(d1));
    const double f2 = __hiloint2double(__double2hiint(d2) ^ hm, __double2loint(d2));
    const double m = smin(fb, smax(f1, f2));
    // derivs.cpp:66 returns d2 when the face velocity is exactly zero; the caller multiplies the face value by that velocity, so
    // the flux is +-0 either way (finite operands) and the select is dropped
    return __hiloint2double(__double2hiint(m) ^ hm, __double2loint(m));
}

// exact test "all eight values are +-0" on the integer pipe
__device__ __forceinline__ bool all_zero8(double a, double b, double c, double d, double e, double f, double g, double h)
{
    const unsigned long long o = (unsigned long long)(__double_as_longlong(a) | __double_as_longlong(b) | __double_as_longlong(c) | __double_as_longlong(d)
                                                       | __double_as_longlong(e) | __double_as_longlong(f) | __double_as_longlong(g) | __double_as_longlong(h));
    return (o << 1) == 0ULL;
}

} // namespace spruce
