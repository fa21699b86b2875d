// mhd_kernels.cuh -- device kernels of the ideal-MHD per-timestep advance (sm_100a, FP64, bit-exact mode).
//
// Replaces, fused into one launch per Runge-Kutta stage:
//   IdealMHD::computeTimeDerivativesDerived           source/equationsets/idealmhd.cpp:42-105
//     (upwindSurface / transportDerivative1D / derivative1D   source/mhd/derivs.cpp:10-73,122-162,223-264)
//   EquationSet::applyTimeDerivatives                  source/equationsets/equationset.cpp:222-230
//   IdealMHD::enforceMinimums / recomputeDerived... / recomputeDT   idealmhd.cpp:234-304
//   Grid::min over the dt bounds                       source/mhd/grid.cpp:71-82
//
// Data layout: SoA planes, plane[r*pitch + j]; r = local row (x index), j = y index, j contiguous; 2 halo rows
// on each side of the slab (rows -2,-1,nx,nx+1) that are filled by the neighbour exchange when the domain is
// slab-decomposed; a single-rank periodic x axis wraps the row index instead.
//
// Plane 0 of an evolved set holds n (number density), not rho: after every propagate the reference's rho is
// exactly RN(n * m_i) (idealmhd.cpp:246-247), so n is the lossless representation and rho is one multiply away.
//
// Kernel shape ("column marching"): a CTA owns TW adjacent columns (contiguous j, one thread per column) and
// marches along x over a chunk of rows.  A 5-row ring of the 11 transported quantities + 3 velocity components
// lives in shared memory (row r+3 is being loaded while row r is computed).  x-direction face fluxes are carried
// in registers from one row to the next, so every x face is evaluated once; the x-direction cell-size tables are
// warp-uniform loads, the y-direction ones per-thread constants.
#pragma once
#include "exact_math.cuh"

namespace spruce {

constexpr int HALO = 2;          // N_GHOST, source/constants.hpp:4
constexpr int TAB_APRON = 3;     // 1-D geometry tables are valid for indices [-3, n+3)
constexpr int NEV = 8;           // evolved planes: n(rho), mom_x, mom_y, mom_z, thermal_energy, bi_x, bi_y, bi_z
constexpr int NSTATIC = 5;       // be_x, be_y, be_z, grav_x, grav_y

constexpr double kKB = 1.3807e-16;                    // K_B   source/constants.hpp:8
constexpr double kPI = 3.14159265358979323846;        // PI    source/constants.hpp:16

enum { E_N = 0, E_MX, E_MY, E_MZ, E_E, E_BX, E_BY, E_BZ };
enum { S_BEX = 0, S_BEY, S_BEZ, S_GX, S_GY };
enum { BC_PERIODIC = 0, BC_OPEN = 1, BC_FIXED = 2, BC_REFLECT = 3, BC_OPEN_MOC = 4, BC_OPEN_UCNP = 5 };
enum { KM_NONE = 0, KM_STORE_K1 = 1, KM_STORE_K2 = 2, KM_ADD_K2 = 3, KM_FINAL = 4, KM_EXPORT = 5 };

// 1-D cell-size tables of one axis; pointers are pre-offset so that index -TAB_APRON is the first element.
struct AxisTab {
    const double *h;     // 0.5*d
    const double *fs;    // fs[f] = h[f] + h[f-1]
    const double *rfs;   // RN(1/fs[f])
    const double *ep;    // ep[f] = h[f-2] + 2*h[f-1]
    const double *em;    // em[f] = h[f+1] + 2*h[f]
    const double *d;     // d
    const double *rd;    // RN(1/d)
};

struct DomainParams {
    int nx, ny, pitch;            // local rows, columns, doubles per row
    int gnx, row0;                // global xdim, global index of local row 0
    int xl, xu, yl, yu;           // GLOBAL interior bounds (computeIterationBounds, plasmadomain.cpp:138-161)
    int xper, yper;               // both sides periodic along that axis
    int xwrap;                    // single-rank periodic x: wrap the local row index
    int bc_x1, bc_x2, bc_y1, bc_y2;
    double m_i, rm_i;             // ion mass, RN(1/m_i)
    double gamma, gm1;            // adiabatic index, gamma - 1.0
    double n_min, T_min, e_min;
    double fourpi, rfourpi;       // 4.0*PI, RN(1/(4.0*PI))
    double epsilon;
    AxisTab tx, ty;               // tx indexed by LOCAL row
};

struct StageArgs {
    const double *S[NEV];         // state the right-hand side is evaluated on
    const double *B[NEV];         // state the increment is added to (own cell only)
    double *D[NEV];               // destination state (never aliases S)
    const double *st[NSTATIC];    // be_x, be_y, be_z, grav_x, grav_y
    double *K1[NEV];              // RK4: k1            (KM_STORE_K1 / KM_FINAL) ; KM_EXPORT: raw k output
    double *K2[NEV];              // RK4: k2 then k2+k3 (KM_STORE_K2 / KM_ADD_K2 / KM_FINAL)
    int kmode;
    int primary;                  // D is the primary state: pointwise boundary zeroing, r1 strips, dt minimum
    double coef;                  // 0.5 or 1.0: s = coef*step (evolution.cpp:95,100)
    const double *step_ptr;       // device scalar: step size of this iteration
    const int *done_ptr;          // device flag: run() reached max_time -> every later launch is a no-op
    unsigned long long *dtmin_bits; // device scalar: running min of dt over the interior, as ordered bits
    double *r1strip[4];           // post-floor rho of the first interior cell next to an `open` side (x1,x2,y1,y2)
    int chunk_rows;               // rows per CTA
};

constexpr int TW = 64;                       // columns (= threads) per CTA
constexpr int SW = TW + 2 * HALO;            // shared row width
constexpr int RD = 5;                        // ring depth (rows)
enum { Q_RHO = 0, Q_MX, Q_MY, Q_MZ, Q_E, Q_BIX, Q_BIY, Q_BIZ, Q_BEX, Q_BEY, Q_BEZ, Q_VX, Q_VY, Q_VZ, NARR };
constexpr int NTR = 11;                      // transported quantities Q_RHO..Q_BEZ

__device__ __forceinline__ FaceGeom load_face_geom(const AxisTab &t, int f)
{
    FaceGeom g;
    g.hm1 = t.h[f - 1]; g.h0 = t.h[f];
    g.fs = t.fs[f];     g.rfs = t.rfs[f];
    g.ep = t.ep[f];     g.fsm = t.fs[f - 1]; g.rfsm = t.rfs[f - 1];
    g.em = t.em[f];     g.fsp = t.fs[f + 1]; g.rfsp = t.rfs[f + 1];
    return g;
}

// is global row g / column j inside the domain (or reachable by periodic wrap)?
__device__ __forceinline__ bool row_exists(const DomainParams &P, int r)
{
    if (P.xwrap) return true;
    const int g = P.row0 + r;
    return (g >= 0 && g < P.gnx) || P.xper;   // xper && !xwrap: halo rows hold the neighbour's (wrapped) rows
}
__device__ __forceinline__ int phys_row(const DomainParams &P, int r)
{
    if (P.xwrap) { r = (r + P.nx) % P.nx; }
    return r;
}

// Pointwise part of updateGhostZones that lands inside the dt bounds: `fixed` and `reflect` zero every momentum
// component in the two ghost cells AND the first interior cell (evolution.cpp:245-263, 272-282).  fixed sweeps the
// whole side, reflect only the interior range of the other axis (evolution.cpp:129-151).
__device__ __forceinline__ bool momentum_zeroed(const DomainParams &P, int g, int j)
{
    const bool jin = (j >= P.yl && j <= P.yu), iin = (g >= P.xl && g <= P.xu);
    bool z = false;
    if (g <= 2)          z |= (P.bc_x1 == BC_FIXED) || (P.bc_x1 == BC_REFLECT && jin);
    if (g >= P.gnx - 3)  z |= (P.bc_x2 == BC_FIXED) || (P.bc_x2 == BC_REFLECT && jin);
    if (j <= 2)          z |= (P.bc_y1 == BC_FIXED) || (P.bc_y1 == BC_REFLECT && iin);
    if (j >= P.ny - 3)   z |= (P.bc_y2 == BC_FIXED) || (P.bc_y2 == BC_REFLECT && iin);
    return z;
}

// enforceMinimums + recomputeDerivedVarsFromEvolvedVars for rho (idealmhd.cpp:237,246-247). Returns n; *r1 = post-floor rho.
__device__ __forceinline__ double density_floor(const DomainParams &P, double rho_u, double *r1)
{
    const double n1 = smax(ddiv(rho_u, P.m_i, P.rm_i), P.n_min);
    const double rr = n1 * P.m_i;
    *r1 = rr;
    return smax(ddiv(rr, P.m_i, P.rm_i), P.n_min);
}

// recomputeDT for one cell (idealmhd.cpp:279-304)
__device__ __forceinline__ double cell_dt(const DomainParams &P, double rho, double mx, double my, double e,
                                          double bx, double by, double bz, double dx, double rdx, double dy, double rdy)
{
    const double vx = mx / rho, vy = my / rho;
    const double p = e * P.gm1;
    const double bm = sqrt((bx * bx + by * by) + bz * bz);
    const double cs = sqrt((p * P.gamma) / rho);
    const double cs2 = cs * cs;
    const double va = bm / sqrt(rho * P.fourpi);
    const double va2 = va * va;
    const double s = cs2 + va2;
    const double delta = sqrt(1.0 - ((cs2 * 4.0) * va2) / (s * s));
    const double vfast = sqrt((s * 0.5) * (1.0 + delta));
    const double vslow = sqrt((s * 0.5) * (1.0 - delta));
    const double vmx = sqrt(vx * vx), vmy = sqrt(vy * vy);
    const double M = smax(smax(smax(cs, va), vfast), vslow);
    return 1.0 / (ddiv(vmx + M, dx, rdx) + ddiv(vmy + M, dy, rdy));
}

// block-wide NaN-ignoring minimum of positive doubles -> atomicMin on the ordered bit pattern
__device__ __forceinline__ void block_min_to_global(double v, unsigned long long *target)
{
    unsigned long long b = (v == v) ? (unsigned long long)__double_as_longlong(v) : 0x7FF0000000000000ULL;
    if (v < 0.0) b = 0ULL;   // cannot happen for a valid dt; keeps ordering total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = (t < b) ? t : b;
    }
    __shared__ unsigned long long wmin[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wmin[w] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int k = 1; k < nw; k++) b = (wmin[k] < b) ? wmin[k] : b;
        atomicMin(target, b);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused Runge-Kutta stage: D = B + (coef*step) * f(S), floors, pointwise boundary zeroing, dt minimum.
// grid = (ceil(ny/TW), ceil(nx/chunk_rows)), block = TW.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TW) k_mhd_stage(const DomainParams P, const StageArgs A)
{
    __shared__ double ring[RD][NARR][SW];
    if (*A.done_ptr) return;

    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * TW;
    const int j = j0 + tid;
    const int c = tid + HALO;
    const int r0 = blockIdx.y * A.chunk_rows;
    const int r1 = min(r0 + A.chunk_rows, P.nx);
    const bool col_out = (j < P.ny);

    // column this thread loads: its own, or the periodic image when the strip overhangs the domain
    int jl = j;
    bool jl_ok = col_out;
    if (!col_out && P.yper && j - P.ny < P.ny) { jl = j - P.ny; jl_ok = true; }
    // halo column handled by threads 0..3 : c = 0,1,TW+2,TW+3
    const bool halo_thread = tid < 2 * HALO;
    const int hc = (tid < HALO) ? tid : (TW + tid);
    int jh = j0 - HALO + hc;
    bool jh_ok = (jh >= 0 && jh < P.ny);
    if (!jh_ok && P.yper) { jh = (jh + 2 * P.ny) % P.ny; jh_ok = true; }

    const double step = *A.step_ptr;
    const double s = A.coef * step;

    // ---- row loader: global -> ring slot, with rho = n*m_i and v = mom/rho computed once per cell
    auto load_cell = [&](int slot, int cc, size_t off, bool ok) {
        double n_ = 1.0, mx = 0.0, my = 0.0, mz = 0.0, e = 0.0, bx = 0.0, by = 0.0, bz = 0.0, ex = 0.0, ey = 0.0, ez = 0.0;
        if (ok) {
            n_ = A.S[E_N][off]; mx = A.S[E_MX][off]; my = A.S[E_MY][off]; mz = A.S[E_MZ][off];
            e = A.S[E_E][off]; bx = A.S[E_BX][off]; by = A.S[E_BY][off]; bz = A.S[E_BZ][off];
            ex = A.st[S_BEX][off]; ey = A.st[S_BEY][off]; ez = A.st[S_BEZ][off];
        }
        const double rho = n_ * P.m_i;
        ring[slot][Q_RHO][cc] = rho;
        ring[slot][Q_MX][cc] = mx; ring[slot][Q_MY][cc] = my; ring[slot][Q_MZ][cc] = mz;
        ring[slot][Q_E][cc] = e;
        ring[slot][Q_BIX][cc] = bx; ring[slot][Q_BIY][cc] = by; ring[slot][Q_BIZ][cc] = bz;
        ring[slot][Q_BEX][cc] = ex; ring[slot][Q_BEY][cc] = ey; ring[slot][Q_BEZ][cc] = ez;
        ring[slot][Q_VX][cc] = mx / rho; ring[slot][Q_VY][cc] = my / rho; ring[slot][Q_VZ][cc] = mz / rho;
    };
    auto slot_of = [&](int r) { return (r - r0 + HALO + RD) % RD; };
    auto load_row = [&](int r) {
        const int slot = slot_of(r);
        const bool rok = row_exists(P, r);
        const size_t rowoff = (size_t)phys_row(P, r) * P.pitch;
        load_cell(slot, c, rowoff + jl, rok && jl_ok);
        if (halo_thread) load_cell(slot, hc, rowoff + jh, rok && jh_ok);
    };

    // ---- per-thread y geometry (faces j and j+1) and cell sizes
    const FaceGeom gyL = load_face_geom(P.ty, col_out ? j : 0);
    const FaceGeom gyR = load_face_geom(P.ty, col_out ? j + 1 : 1);
    const double dy = P.ty.d[col_out ? j : 0], rdy = P.ty.rd[col_out ? j : 0];

    // ---- prologue: rows r0-2 .. r0+2
    for (int r = r0 - HALO; r <= r0 + HALO; r++) load_row(r);
    __syncthreads();

    // x-face carry (face r0, between rows r0-1 and r0)
    double Fx[NTR];                   // S*vf per transported quantity
    double cIx_biy, cIx_biz, cIx_p, cVfx, cIx_vy, cIx_vz;
    {
        const FaceGeom g = load_face_geom(P.tx, r0);
        const int sm2 = slot_of(r0 - 2), sm1 = slot_of(r0 - 1), s0 = slot_of(r0), sp1 = slot_of(r0 + 1);
        cVfx = face_interp(ring[sm1][Q_VX][c], ring[s0][Q_VX][c], g.hm1, g.h0, g.fs, g.rfs);
        cIx_vy = face_interp(ring[sm1][Q_VY][c], ring[s0][Q_VY][c], g.hm1, g.h0, g.fs, g.rfs);
        cIx_vz = face_interp(ring[sm1][Q_VZ][c], ring[s0][Q_VZ][c], g.hm1, g.h0, g.fs, g.rfs);
        cIx_p = face_interp(ring[sm1][Q_E][c] * P.gm1, ring[s0][Q_E][c] * P.gm1, g.hm1, g.h0, g.fs, g.rfs);
        double d2;
#pragma unroll
        for (int q = 0; q < NTR; q++) {
            const double S = upwind_face(ring[sm2][q][c], ring[sm1][q][c], ring[s0][q][c], ring[sp1][q][c], cVfx, g, &d2);
            Fx[q] = S * cVfx;
            if (q == Q_BIY) cIx_biy = d2;
            if (q == Q_BIZ) cIx_biz = d2;
        }
    }

    double dtmin_local = 1.7976931348623157e308;

    for (int r = r0; r < r1; r++) {
        // prefetch row r+3 into the free slot (the slot of row r-2)
        if (r + 3 <= r1 + HALO - 1) load_row(r + 3);

        const int sm1 = slot_of(r - 1), s0 = slot_of(r), sp1 = slot_of(r + 1), sp2 = slot_of(r + 2);
        const int g = P.row0 + r;                                   // global row
        const bool interior = col_out && g >= P.xl && g <= P.xu && j >= P.yl && j <= P.yu;
        const double dx = P.tx.d[r], rdx = P.tx.rd[r];

        // ---------------- x face r+1 (between rows r and r+1)
        const FaceGeom gx = load_face_geom(P.tx, r + 1);
        const double vfx1 = face_interp(ring[s0][Q_VX][c], ring[sp1][Q_VX][c], gx.hm1, gx.h0, gx.fs, gx.rfs);
        const double Ix1_vy = face_interp(ring[s0][Q_VY][c], ring[sp1][Q_VY][c], gx.hm1, gx.h0, gx.fs, gx.rfs);
        const double Ix1_vz = face_interp(ring[s0][Q_VZ][c], ring[sp1][Q_VZ][c], gx.hm1, gx.h0, gx.fs, gx.rfs);
        const double pc = ring[s0][Q_E][c] * P.gm1;                 // press = (gamma-1)*thermal_energy  idealmhd.cpp:253
        const double Ix1_p = face_interp(pc, ring[sp1][Q_E][c] * P.gm1, gx.hm1, gx.h0, gx.fs, gx.rfs);
        // ---------------- y faces j (L) and j+1 (R) of row r
        const double vyc = ring[s0][Q_VY][c];
        const double vfyL = face_interp(ring[s0][Q_VY][c - 1], vyc, gyL.hm1, gyL.h0, gyL.fs, gyL.rfs);
        const double vfyR = face_interp(vyc, ring[s0][Q_VY][c + 1], gyR.hm1, gyR.h0, gyR.fs, gyR.rfs);

        // transportDivergence2D of the 11 transported quantities (derivs.cpp:216-220) and the face interpolations
        // that the central derivatives reuse
        double T[NTR];
        double Ix1_biy = 0.0, Ix1_biz = 0.0, IyL_bix = 0.0, IyR_bix = 0.0, IyL_biz = 0.0, IyR_biz = 0.0;
#pragma unroll
        for (int q = 0; q < NTR; q++) {
            double d2x, d2L, d2R;
            const double qc = ring[s0][q][c];
            const double Sx = upwind_face(ring[sm1][q][c], qc, ring[sp1][q][c], ring[sp2][q][c], vfx1, gx, &d2x);
            const double fx1 = Sx * vfx1;
            const double SL = upwind_face(ring[s0][q][c - 2], ring[s0][q][c - 1], qc, ring[s0][q][c + 1], vfyL, gyL, &d2L);
            const double SR = upwind_face(ring[s0][q][c - 1], qc, ring[s0][q][c + 1], ring[s0][q][c + 2], vfyR, gyR, &d2R);
            const double tx_ = ddiv(fx1 - Fx[q], dx, rdx);                      // derivs.cpp:155-156
            const double ty_ = ddiv(SR * vfyR - SL * vfyL, dy, rdy);
            T[q] = tx_ + ty_;
            Fx[q] = fx1;
            if (q == Q_BIY) Ix1_biy = d2x;
            if (q == Q_BIZ) { Ix1_biz = d2x; IyL_biz = d2L; IyR_biz = d2R; }
            if (q == Q_BIX) { IyL_bix = d2L; IyR_bix = d2R; }
        }

        // central derivatives, derivative1D (derivs.cpp:259)
        const double dbiy_dx = ddiv(Ix1_biy - cIx_biy, dx, rdx);
        const double dbix_dy = ddiv(IyR_bix - IyL_bix, dy, rdy);
        const double dbiz_dy = ddiv(IyR_biz - IyL_biz, dy, rdy);
        const double dbiz_dx = ddiv(Ix1_biz - cIx_biz, dx, rdx);
        const double dp_dx = ddiv(Ix1_p - cIx_p, dx, rdx);
        const double pL = ring[s0][Q_E][c - 1] * P.gm1, pR = ring[s0][Q_E][c + 1] * P.gm1;
        const double dp_dy = ddiv(face_interp(pc, pR, gyR.hm1, gyR.h0, gyR.fs, gyR.rfs)
                                  - face_interp(pL, pc, gyL.hm1, gyL.h0, gyL.fs, gyL.rfs), dy, rdy);
        const double dvx_dx = ddiv(vfx1 - cVfx, dx, rdx);
        const double dvy_dy = ddiv(vfyR - vfyL, dy, rdy);
        const double vxc = ring[s0][Q_VX][c], vzc = ring[s0][Q_VZ][c];
        const double dvx_dy = ddiv(face_interp(vxc, ring[s0][Q_VX][c + 1], gyR.hm1, gyR.h0, gyR.fs, gyR.rfs)
                                   - face_interp(ring[s0][Q_VX][c - 1], vxc, gyL.hm1, gyL.h0, gyL.fs, gyL.rfs), dy, rdy);
        const double dvz_dy = ddiv(face_interp(vzc, ring[s0][Q_VZ][c + 1], gyR.hm1, gyR.h0, gyR.fs, gyR.rfs)
                                   - face_interp(ring[s0][Q_VZ][c - 1], vzc, gyL.hm1, gyL.h0, gyL.fs, gyL.rfs), dy, rdy);
        const double dvy_dx = ddiv(Ix1_vy - cIx_vy, dx, rdx);
        const double dvz_dx = ddiv(Ix1_vz - cIx_vz, dx, rdx);
        // roll the x carries
        cIx_biy = Ix1_biy; cIx_biz = Ix1_biz; cIx_p = Ix1_p; cVfx = vfx1; cIx_vy = Ix1_vy; cIx_vz = Ix1_vz;

        // ---------------- own-cell values
        const size_t off = (size_t)r * P.pitch + j;    // destination / base offset (local, unwrapped)
        const double rho = ring[s0][Q_RHO][c];
        const double bix = ring[s0][Q_BIX][c], biy = ring[s0][Q_BIY][c], biz = ring[s0][Q_BIZ][c];
        const double bex = ring[s0][Q_BEX][c], bey = ring[s0][Q_BEY][c], bez = ring[s0][Q_BEZ][c];
        double gxv = 0.0, gyv = 0.0;
        if (col_out) { gxv = A.st[S_GX][off]; gyv = A.st[S_GY][off]; }

        // ---------------- right-hand side, idealmhd.cpp:51-103 (expression order is load-bearing)
        double k[NEV];
        k[E_N] = T[Q_RHO] * -1.0;                                                       // :52
        const double cdb = ddiv(dbiy_dx - dbix_dy, P.fourpi, P.rfourpi);                // :54  curl2D/(4 pi)
        const double ncdb = cdb * -1.0;
        const double czx = dbiz_dy, czy = dbiz_dx * -1.0;                               // curlZ, derivs.cpp:465-469
        const double bzi = ddiv(biz, P.fourpi, P.rfourpi), bze = ddiv(bez, P.fourpi, P.rfourpi);   // :57-58
        // CrossProduct2DZ(a,bz) = CrossProductZ2D(-1.0*bz, a) = { -(-bz)*a_y , (-bz)*a_x }   grid.cpp:455-468
        k[E_MX] = ((((((T[Q_MX] * -1.0) - dp_dx) + rho * gxv) + ncdb * bey) + ncdb * biy) + bzi * czy) + bze * czy;      // :62-66
        k[E_MY] = ((((((T[Q_MY] * -1.0) - dp_dy) + rho * gyv) + cdb * bex) + cdb * bix) + (bzi * -1.0) * czx) + (bze * -1.0) * czx;  // :67-71
        const double fze = ddiv(czx * bey - czy * bex, P.fourpi, P.rfourpi);            // :59
        const double fzi = ddiv(czx * biy - czy * bix, P.fourpi, P.rfourpi);            // :60
        k[E_MZ] = ((T[Q_MZ] * -1.0) + fze) + fzi;                                       // :72-73
        k[E_E] = (T[Q_E] * -1.0) - pc * (dvx_dx + dvy_dy);                              // :75-76
        const double bxs = bix + bex, bys = biy + bey;
        k[E_BX] = (((T[Q_BIX] * -1.0) - T[Q_BEX]) + bxs * dvx_dx) + bys * dvx_dy;       // :78-80
        k[E_BY] = (((T[Q_BIY] * -1.0) - T[Q_BEY]) + bxs * dvy_dx) + bys * dvy_dy;       // :81-83
        k[E_BZ] = (((T[Q_BIZ] * -1.0) - T[Q_BEZ]) + bxs * dvz_dx) + bys * dvz_dy;       // :84-86
        // ghost mask (:99-103): operators return 0 outside [xl..xu]x[yl..yu] and the mask zeroes the rest
#pragma unroll
        for (int v = 0; v < NEV; v++) k[v] = interior ? k[v] : 0.0;

        if (col_out) {
            // ---------------- RK4 bookkeeping (evolution.cpp:103-124)
            if (A.kmode == KM_STORE_K1 || A.kmode == KM_EXPORT) {
#pragma unroll
                for (int v = 0; v < NEV; v++) A.K1[v][off] = k[v];
            } else if (A.kmode == KM_STORE_K2) {
#pragma unroll
                for (int v = 0; v < NEV; v++) A.K2[v][off] = k[v];
            } else if (A.kmode == KM_ADD_K2) {
#pragma unroll
                for (int v = 0; v < NEV; v++) A.K2[v][off] = A.K2[v][off] + k[v];
            } else if (A.kmode == KM_FINAL) {
#pragma unroll
                for (int v = 0; v < NEV; v++) k[v] = (A.K1[v][off] + k[v]) / 6.0 + A.K2[v][off] / 3.0;   // :121
            }
            if (A.kmode != KM_EXPORT) {
                // ---------------- applyTimeDerivatives: U += step*k  (two roundings)  equationset.cpp:226-228
                double U[NEV];
                U[E_N] = (A.B[E_N][off] * P.m_i) + k[E_N] * s;      // rho = n*m_i
#pragma unroll
                for (int v = 1; v < NEV; v++) U[v] = A.B[v][off] + k[v] * s;
                // ---------------- propagateChanges, pointwise part (equationset.cpp:212-220)
                double rfl;
                const double nn = density_floor(P, U[E_N], &rfl);
                const double e1 = smax(U[E_E], P.e_min);
                if (A.primary) {
                    if (momentum_zeroed(P, g, j)) { U[E_MX] = 0.0; U[E_MY] = 0.0; U[E_MZ] = 0.0; }
                    if (P.bc_x1 == BC_OPEN && g == 2 && A.r1strip[0]) A.r1strip[0][j] = rfl;
                    if (P.bc_x2 == BC_OPEN && g == P.gnx - 3 && A.r1strip[1]) A.r1strip[1][j] = rfl;
                    if (P.bc_y1 == BC_OPEN && j == 2 && A.r1strip[2]) A.r1strip[2][r] = rfl;
                    if (P.bc_y2 == BC_OPEN && j == P.ny - 3 && A.r1strip[3]) A.r1strip[3][r] = rfl;
                }
                A.D[E_N][off] = nn;
                A.D[E_MX][off] = U[E_MX]; A.D[E_MY][off] = U[E_MY]; A.D[E_MZ][off] = U[E_MZ];
                A.D[E_E][off] = e1;
                A.D[E_BX][off] = U[E_BX]; A.D[E_BY][off] = U[E_BY]; A.D[E_BZ][off] = U[E_BZ];
                if (A.primary && interior) {
                    const double dtc = cell_dt(P, nn * P.m_i, U[E_MX], U[E_MY], e1, bex + U[E_BX], bey + U[E_BY], bez + U[E_BZ],
                                               dx, rdx, dy, rdy);
                    dtmin_local = smin(dtmin_local, dtc);
                }
            }
        }
        __syncthreads();
    }
    if (A.primary && A.kmode != KM_EXPORT) block_min_to_global(dtmin_local, A.dtmin_bits);
}

// ---------------------------------------------------------------------------------------------------------
// Pointwise propagateChanges on the primary state (module hooks, setup): floors, boundary zeroing, dt minimum.
// raw_rho: plane 0 currently holds rho as uploaded (not n).  from_state: also derive thermal_energy from temp
// (recomputeEvolvedVarsFromStateVars, idealmhd.cpp:226-232).
// ---------------------------------------------------------------------------------------------------------
struct PropArgs {
    double *U[NEV];
    const double *st[NSTATIC];
    const double *temp;           // only when from_state
    int raw_rho, from_state;
    unsigned long long *dtmin_bits;
    double *r1strip[4];
};

__global__ void __launch_bounds__(256) k_mhd_propagate(const DomainParams P, const PropArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double dtc = 1.7976931348623157e308;
    if (j < P.ny) {
        const size_t off = (size_t)r * P.pitch + j;
        const int g = P.row0 + r;
        double rho_u = A.raw_rho ? A.U[E_N][off] : A.U[E_N][off] * P.m_i;
        double e = A.U[E_E][off];
        if (A.from_state) {
            const double n0 = smax(ddiv(rho_u, P.m_i, P.rm_i), P.n_min);
            const double press = ((n0 * 2.0) * kKB) * smax(A.temp[off], P.T_min);      // idealmhd.cpp:230
            e = press / P.gm1;                                                          // :231
        }
        double rfl;
        const double nn = density_floor(P, rho_u, &rfl);
        const double e1 = smax(e, P.e_min);
        double mx = A.U[E_MX][off], my = A.U[E_MY][off], mz = A.U[E_MZ][off];
        if (momentum_zeroed(P, g, j)) { mx = 0.0; my = 0.0; mz = 0.0; A.U[E_MX][off] = 0.0; A.U[E_MY][off] = 0.0; A.U[E_MZ][off] = 0.0; }
        if (P.bc_x1 == BC_OPEN && g == 2) A.r1strip[0][j] = rfl;
        if (P.bc_x2 == BC_OPEN && g == P.gnx - 3) A.r1strip[1][j] = rfl;
        if (P.bc_y1 == BC_OPEN && j == 2) A.r1strip[2][r] = rfl;
        if (P.bc_y2 == BC_OPEN && j == P.ny - 3) A.r1strip[3][r] = rfl;
        A.U[E_N][off] = nn;
        A.U[E_E][off] = e1;
        if (g >= P.xl && g <= P.xu && j >= P.yl && j <= P.yu)
            dtc = cell_dt(P, nn * P.m_i, mx, my, e1, A.st[S_BEX][off] + A.U[E_BX][off], A.st[S_BEY][off] + A.U[E_BY][off],
                          A.st[S_BEZ][off] + A.U[E_BZ][off], P.tx.d[r], P.tx.rd[r], P.ty.d[j], P.ty.rd[j]);
    }
    block_min_to_global(dtc, A.dtmin_bits);
}

// ---------------------------------------------------------------------------------------------------------
// Ghost cells of the non-periodic sides: reflect / open / open_ucnp (evolution.cpp:126-333).  `fixed` and the
// momentum zeroing of `reflect` are pointwise and already applied.  One thread per boundary index and side.
// The four sides touch disjoint cells and read only first-interior cells, so they run concurrently.
// ---------------------------------------------------------------------------------------------------------
struct GhostArgs {
    double *U[NEV];
    const double *r1strip[4];
    int primary;                       // reflect/open act on the primary state only
    // open boundary scalars per side (x1,x2,y1,y2), evaluated on the host with libm pow (evolution.cpp:163-167)
    double scale_1[4], scale_2[4], dist23[4], h2[4], h3[4], rh3[4];   // h2 = 0.5*d(i2), h3 = 0.5*d(i3), rh3 = RN(1/h3)
    double open_strength;
};

__device__ __forceinline__ void ghost_one(const DomainParams &P, const GhostArgs &A, int side, int idx)
{
    const int bc = side == 0 ? P.bc_x1 : side == 1 ? P.bc_x2 : side == 2 ? P.bc_y1 : P.bc_y2;
    if (bc == BC_PERIODIC || bc == BC_FIXED || bc == BC_OPEN_MOC) return;
    const bool xside = side < 2;
    // sweep range of the other axis: m_yl..m_yu / m_xl..m_xu (evolution.cpp:129-150)
    if (xside) { if (idx < P.yl || idx > P.yu) return; }
    else {
        const int g = P.row0 + idx;
        if (g < P.xl || g > P.xu) return;
    }
    // local (row, col) of the edge cell i1, next cell i2, first interior cell i3
    int r1_, r2_, r3_, c1_, c2_, c3_;
    if (side == 0) { r1_ = 0 - P.row0; r2_ = 1 - P.row0; r3_ = 2 - P.row0; c1_ = c2_ = c3_ = idx; }
    else if (side == 1) { r1_ = P.gnx - 1 - P.row0; r2_ = P.gnx - 2 - P.row0; r3_ = P.gnx - 3 - P.row0; c1_ = c2_ = c3_ = idx; }
    else if (side == 2) { r1_ = r2_ = r3_ = idx; c1_ = 0; c2_ = 1; c3_ = 2; }
    else { r1_ = r2_ = r3_ = idx; c1_ = P.ny - 1; c2_ = P.ny - 2; c3_ = P.ny - 3; }
    if (xside && (r3_ < 0 || r3_ >= P.nx)) return;     // that side belongs to another rank's slab
    const size_t o1 = (size_t)r1_ * P.pitch + c1_, o2 = (size_t)r2_ * P.pitch + c2_, o3 = (size_t)r3_ * P.pitch + c3_;

    if (bc == BC_OPEN_UCNP) {
        // copies the nearest interior cell into both ghost cells for densities, thermal energies, fields, momenta
        // (evolution.cpp:321-331); acts on whichever set is being propagated.
#pragma unroll
        for (int v = 0; v < NEV; v++) { const double x = A.U[v][o3]; A.U[v][o1] = x; A.U[v][o2] = x; }
        return;
    }
    if (!A.primary) return;
    if (bc == BC_REFLECT) {
        // thermal energy and density of the nearest interior cell (evolution.cpp:237-244); n(i3) is exactly what the
        // reference's derived step computes for the copied rho.  Momenta are already zero.
        const double e3 = A.U[E_E][o3], n3 = A.U[E_N][o3];
        A.U[E_E][o1] = e3; A.U[E_E][o2] = e3; A.U[E_N][o1] = n3; A.U[E_N][o2] = n3;
        return;
    }
    // BC_OPEN (evolution.cpp:158-224)
    const double rho3 = A.r1strip[side][idx];                 // post-floor rho of i3 (before the n round trip)
    const double e3 = A.U[E_E][o3];
    const double rho_1 = A.scale_1[side] * rho3, rho_2 = A.scale_2[side] * rho3;
    A.U[E_E][o1] = smax(A.scale_1[side] * e3, P.e_min);       // derived step re-applies the energy floor (idealmhd.cpp:252)
    A.U[E_E][o2] = smax(A.scale_2[side] * e3, P.e_min);
    const double press = e3 * P.gm1;
    double c_s = 0.0;
    const double c_new = sqrt(P.gamma * press / rho3);
    if (c_new > c_s) c_s = c_new;
    const double mx3 = A.U[E_MX][o3], my3 = A.U[E_MY][o3];
    const double vel_x = mx3 / rho3, vel_y = my3 / rho3;
    double boost = A.open_strength * c_s;
    const bool lower = (side == 0 || side == 2);              // i2 > i1 || j2 > j1
    if (lower) boost *= -1.0;
    const double vn = xside ? vel_x : vel_y, vt = xside ? vel_y : vel_x;
    const double bv = lower ? smin(0.0, vn + boost) : smax(0.0, vn + boost);
    const double gv = ddiv(A.dist23[side] * bv - A.h2[side] * vn, A.h3[side], A.rh3[side]);
    const double mn1 = rho_1 * gv, mn2 = rho_2 * gv, mt1 = rho_1 * vt, mt2 = rho_2 * vt;
    if (xside) { A.U[E_MX][o1] = mn1; A.U[E_MX][o2] = mn2; A.U[E_MY][o1] = mt1; A.U[E_MY][o2] = mt2; }
    else       { A.U[E_MY][o1] = mn1; A.U[E_MY][o2] = mn2; A.U[E_MX][o1] = mt1; A.U[E_MX][o2] = mt2; }
    // derived step: n = max(rho/m_i, n_min) (idealmhd.cpp:246)
    A.U[E_N][o1] = smax(ddiv(rho_1, P.m_i, P.rm_i), P.n_min);
    A.U[E_N][o2] = smax(ddiv(rho_2, P.m_i, P.rm_i), P.n_min);
}

__global__ void k_mhd_ghosts(const DomainParams P, const GhostArgs A)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int side = blockIdx.y;
    const int n = side < 2 ? P.ny : P.nx;
    if (idx < n) ghost_one(P, A, side, idx);
}

// ---------------------------------------------------------------------------------------------------------
// Derived variables on demand (download / output / host-side modules): idealmhd.cpp:241-304 evaluated from the
// evolved planes.  which: index into the reference's variable list (idealmhd.hpp:19-23).
// ---------------------------------------------------------------------------------------------------------
enum { V_rho = 0, V_temp, V_mom_x, V_mom_y, V_mom_z, V_bi_x, V_bi_y, V_bi_z, V_grav_x, V_grav_y,
       V_n, V_press, V_thermal_energy, V_v_x, V_v_y, V_v_z, V_kinetic_energy,
       V_b_x, V_b_y, V_b_z, V_b_mag, V_b_hat_x, V_b_hat_y, V_b_hat_z, V_dt, V_COUNT };

struct DeriveArgs { const double *U[NEV]; const double *st[NSTATIC]; double *out; int which; };

__global__ void __launch_bounds__(256) k_mhd_derive(const DomainParams P, const DeriveArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    const double n_ = A.U[E_N][off];
    const double rho = n_ * P.m_i;
    const double e = A.U[E_E][off];
    const double p = e * P.gm1;
    double out = 0.0;
    switch (A.which) {
    case V_rho: out = rho; break;
    case V_n: out = n_; break;
    case V_press: out = p; break;
    case V_temp: out = smax(p / (n_ * (2 * kKB)), P.T_min); break;                 // idealmhd.cpp:254
    case V_v_x: out = A.U[E_MX][off] / rho; break;
    case V_v_y: out = A.U[E_MY][off] / rho; break;
    case V_v_z: out = A.U[E_MZ][off] / rho; break;
    case V_kinetic_energy: { const double vx = A.U[E_MX][off] / rho, vy = A.U[E_MY][off] / rho; out = (rho * 0.5) * (vx * vx + vy * vy); } break;
    case V_b_x: out = A.st[S_BEX][off] + A.U[E_BX][off]; break;
    case V_b_y: out = A.st[S_BEY][off] + A.U[E_BY][off]; break;
    case V_b_z: out = A.st[S_BEZ][off] + A.U[E_BZ][off]; break;
    case V_b_mag: case V_b_hat_x: case V_b_hat_y: case V_b_hat_z: {
        const double bx = A.st[S_BEX][off] + A.U[E_BX][off], by = A.st[S_BEY][off] + A.U[E_BY][off], bz = A.st[S_BEZ][off] + A.U[E_BZ][off];
        const double bm = sqrt((bx * bx + by * by) + bz * bz);
        if (A.which == V_b_mag) out = bm;
        else if (bm == 0.0) out = 0.0;                                              // catchNullFieldDirection :265-277
        else out = (A.which == V_b_hat_x ? bx : A.which == V_b_hat_y ? by : bz) / bm;
    } break;
    case V_dt:
        out = cell_dt(P, rho, A.U[E_MX][off], A.U[E_MY][off], e, A.st[S_BEX][off] + A.U[E_BX][off], A.st[S_BEY][off] + A.U[E_BY][off],
                      A.st[S_BEZ][off] + A.U[E_BZ][off], P.tx.d[r], P.tx.rd[r], P.ty.d[j], P.ty.rd[j]);
        break;
    default: break;
    }
    A.out[off] = out;
}

// ---------------------------------------------------------------------------------------------------------
// Scalar bookkeeping of advanceTime (evolution.cpp:62,80-81), one thread.
// ctl[0] = step (double), ctl[1] = time, ctl[2] = max_time (<=0: none); ictl[0] = iter, ictl[1] = done flag
// ---------------------------------------------------------------------------------------------------------
struct StepCtl { double step, time, max_time, epsilon; long long iter; int done; int pad; unsigned long long dtmin_bits; };

// begin: step = epsilon * min(dt) ; reset the running minimum for the propagate at the end of this step
__global__ void k_step_begin(StepCtl *c, double *dt_hist, int slot)
{
    if (c->max_time > 0.0 && !(c->time < c->max_time)) c->done = 1;
    if (c->done) { return; }
    c->step = c->epsilon * __longlong_as_double((long long)c->dtmin_bits);   // evolution.cpp:62
    if (dt_hist) dt_hist[slot] = c->step;
}
__global__ void k_dtmin_reset(StepCtl *c) { if (!c->done) c->dtmin_bits = 0x7FEFFFFFFFFFFFFFULL; }
__global__ void k_step_end(StepCtl *c) { if (c->done) return; c->time += c->step; c->iter += 1; }

} // namespace spruce
