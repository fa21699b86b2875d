// ideal2f_sides.cuh -- the boundary passes of the two-fluid equation set in their LITERAL, ordered form (PlasmaDomain::updateGhostZones with the
// species lists of Ideal2F, source/mhd/evolution.cpp:126-152, 231-333): fixed zeroes every momentum of both species in the two ghost cells and the
// first interior cell of the whole side, reflect does the same for the interior range of the side and copies densities and thermal energies
// outwards, open_ucnp copies densities, thermal energies, E and momenta outwards.  The sides run one after the other (x1, x2, y1, y2), each reading
// what the earlier ones wrote.
//
// The default two-fluid path applies fixed / reflect POINTWISE inside the stage kernel and runs the copying passes of all sides concurrently; that
// is exact unless an open_ucnp side meets a fixed / reflect side (the ucnp pass must read the first interior cell's momentum BEFORE a later
// reflect side zeroes it).  For those mixed sets the stage runs without the pointwise part and these passes follow, in order; the dt minimum is
// then taken over the finished state.  Plain C++ (no CUDA dependence): tests/hostcheck compiles it with g++ and checks it against the pinned CPU
// restatement on random planes.  Cells are addressed by GLOBAL (i, j) (see mhd2e_cells.cuh for the slab convention).
// STATUS: written after the round-1 GPU budget was spent; not yet run on a GPU.
#pragma once
#include <cstddef>

#if defined(__CUDACC__)
#define TF2_HD __host__ __device__ inline
#else
#define TF2_HD static inline
#endif

namespace spruce {
namespace tf2 {

constexpr int NV = 14;      // i_rho, e_rho, i_mom_x, i_mom_y, e_mom_x, e_mom_y, i_thermal_energy, e_thermal_energy, E_x, E_y, E_z, bi_x, bi_y, bi_z (ideal2F.hpp:44-46)
enum { BCT_PERIODIC = 0, BCT_OPEN = 1, BCT_FIXED = 2, BCT_REFLECT = 3, BCT_OPEN_MOC = 4, BCT_OPEN_UCNP = 5 };
struct Geo { int nx, ny, pitch, row0, nxl; int bc[4]; int xl, xu, yl, yu; };      // nx: GLOBAL; the slab holds rows [row0, row0 + nxl)
struct Planes { double *u[NV]; };                                                  // row-shifted on slabs

TF2_HD int side_length(const Geo &g, int side) { return side < 2 ? g.ny : g.nxl; }
// boundary index t of side `side`; G = the set being propagated (open_ucnp writes it), P = the primary state (fixed / reflect write it)
TF2_HD void side_line(const Geo &g, const Planes &G, const Planes &P, int side, int t)
{
    const int bc = g.bc[side];
    if (bc != BCT_FIXED && bc != BCT_REFLECT && bc != BCT_OPEN_UCNP) return;
    const bool xside = side < 2, lower = (side % 2) == 0;
    const int ncross = xside ? g.nx : g.ny;
    const int e1 = lower ? 0 : ncross - 1, e2 = lower ? 1 : ncross - 2, e3 = lower ? 2 : ncross - 3;
    if (xside && (lower ? g.row0 != 0 : g.row0 + g.nxl != g.nx)) return;          // the first / last slab owns the x sides
    const int a = xside ? t : g.row0 + t;
    const int lo = xside ? g.yl : g.xl, hi = xside ? g.yu : g.xu;
    const size_t c1 = xside ? (size_t)e1 * g.pitch + a : (size_t)a * g.pitch + e1;
    const size_t c2 = xside ? (size_t)e2 * g.pitch + a : (size_t)a * g.pitch + e2;
    const size_t c3 = xside ? (size_t)e3 * g.pitch + a : (size_t)a * g.pitch + e3;
    const int moms[4] = {2, 3, 4, 5};
    if (bc == BCT_FIXED) {                                                         // evolution.cpp:268-282, whole side
        for (int m = 0; m < 4; m++) { P.u[moms[m]][c1] = 0.0; P.u[moms[m]][c2] = 0.0; P.u[moms[m]][c3] = 0.0; }
        return;
    }
    if (a < lo || a > hi) return;
    if (bc == BCT_REFLECT) {                                                       // :231-266
        const int copy[4] = {6, 7, 0, 1};                                          // thermal energies, then densities
        for (int k = 0; k < 4; k++) { P.u[copy[k]][c1] = P.u[copy[k]][c3]; P.u[copy[k]][c2] = P.u[copy[k]][c3]; }
        for (int m = 0; m < 4; m++) { P.u[moms[m]][c1] = 0.0; P.u[moms[m]][c2] = 0.0; P.u[moms[m]][c3] = 0.0; }
    } else {                                                                       // open_ucnp :290-333: densities, thermal energies, fields (E), momenta
        const int vars[11] = {0, 1, 6, 7, 8, 9, 10, 2, 3, 4, 5};
        for (int k = 0; k < 11; k++) { G.u[vars[k]][c1] = G.u[vars[k]][c3]; G.u[vars[k]][c2] = G.u[vars[k]][c3]; }
    }
}

}  // namespace tf2
}  // namespace spruce
