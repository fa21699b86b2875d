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
#include "cell_math.cuh"

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

struct StageArgs {
    const double *S[NEV];         // state the right-hand side is evaluated on
    const double *B[NEV];         // state the increment is added to (own cell only)
    double *D[NEV];               // destination state (never aliases S)
    const double *st[NSTATIC];    // be_x, be_y, be_z, grav_x, grav_y
    double *K1[NEV];              // RK4: k1            (KM_STORE_K1 / KM_FINAL) ; KM_EXPORT: raw k output
    double *K2[NEV];              // RK4: k2 then k2+k3 (KM_STORE_K2 / KM_ADD_K2 / KM_FINAL)
    int kmode;
    int b_is_s;                   // B aliases S (first stage): own-cell base values come from the shared ring
    int primary;                  // D is the primary state: pointwise boundary zeroing, r1 strips, dt minimum
    double coef;                  // 0.5 or 1.0: s = coef*step (evolution.cpp:95,100)
    const double *step_ptr;       // device scalar: step size of this iteration
    const int *done_ptr;          // device flag: run() reached max_time -> every later launch is a no-op
    unsigned long long *dtmin_bits; // device scalar: running min of dt over the interior, as ordered bits
    const double *inv_thr_ptr;    // device scalar R = 1/(F * previous global min dt), 0 = evaluate every cell (see dt_can_skip)
    double *strip[4];             // per side (x1,x2,y1,y2): what that side's ghost pass reads from its first interior cell
    int strip_pitch;              //   strip[s][c*strip_pitch + idx], c = 0: post-floor rho, 1..3: momentum as that pass sees it
    // rows of a launch: CTA row b covers [row_begin + b*chunk_rows, min(.. + chunk_rows, row_end)); in the edge launch of a slab (edge2_begin >= 0)
    // CTA row 0 covers [row_begin, row_end) and CTA row 1 covers [edge2_begin, edge2_end): the first and the last rows of the slab
    int chunk_rows, row_begin, row_end, edge2_begin, edge2_end;
    // module contributions to the right-hand side (Module::computeTimeDerivativesModule): already masked planes that are
    // added to k[target] in module order, after the ghost mask (equationset.cpp:208, viscosity.cpp:117-118)
    const double *xterm[4]; int xtarget[4]; int n_xterm;
    int vec16;                    // k_mhd_stage_xy: ring rows may be copied in 16-byte chunks (every plane 16-byte aligned, even pitch)
    int walls;                    // some side is not periodic: the primary stage evaluates zero_zones / record_strips
    int grav;                     // a gravity plane of this slab is non-zero (otherwise k_mhd_stage_xy adds rho * 0.0 without reading the planes)
};

// ---- tiling of the fused stage kernel ("column marching", see the header comment)
constexpr int STAGE_WARPS = 2;
constexpr int NT = 32 * STAGE_WARPS;         // threads per CTA
constexpr int CW = 31 * STAGE_WARPS;         // output columns per CTA: a warp evaluates 32 y-faces = 31 cells (+ the face it hands on)
constexpr int SW = CW + 2 * HALO;            // shared row width (66 doubles = 528 B, 16-byte multiple)
constexpr int RD = 5;                        // ring depth: rows r-1..r+2 in use, row r+3 in flight
enum { Q_RHO = 0, Q_MX, Q_MY, Q_MZ, Q_E, Q_BIX, Q_BIY, Q_BIZ, Q_BEX, Q_BEY, Q_BEZ, Q_VX, Q_VY, Q_VZ, NARR };
constexpr int NTR = 11;                      // transported quantities Q_RHO..Q_BEZ
constexpr int NLOAD = 11;                    // arrays filled from global memory (Q_RHO holds n until converted)

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

// updateGhostZones runs its four sides in the fixed order x1, x2, y1, y2 (evolution.cpp:126-152), each side reading
// cells the earlier sides may have changed.  The part that lands inside the dt bounds is pointwise: `fixed` and `reflect`
// zero every momentum component in the two ghost cells AND the first interior cell (evolution.cpp:245-263, 272-282);
// fixed sweeps the whole side, reflect only the interior range of the other axis (evolution.cpp:129-151).
// zone bit s is set when side s zeroes the momentum of cell (g, j).
__device__ __forceinline__ unsigned zero_zones(const DomainParams &P, int g, int j)
{
    const bool jin = (j >= P.yl && j <= P.yu), iin = (g >= P.xl && g <= P.xu);
    unsigned z = 0;
    if (g <= 2         && ((P.bc_x1 == BC_FIXED) || (P.bc_x1 == BC_REFLECT && jin))) z |= 1u;
    if (g >= P.gnx - 3 && ((P.bc_x2 == BC_FIXED) || (P.bc_x2 == BC_REFLECT && jin))) z |= 2u;
    if (j <= 2         && ((P.bc_y1 == BC_FIXED) || (P.bc_y1 == BC_REFLECT && iin))) z |= 4u;
    if (j >= P.ny - 3  && ((P.bc_y2 == BC_FIXED) || (P.bc_y2 == BC_REFLECT && iin))) z |= 8u;
    return z;
}
__device__ __forceinline__ bool reads_interior(int bc) { return bc == BC_OPEN || bc == BC_OPEN_UCNP; }

// The open / open_ucnp pass of side s reads rho and the momenta of its first interior cell at the moment that pass
// runs: after the floors, after the zeroing of the sides that ran BEFORE it, before the later ones.  The full-plane
// kernels record exactly that in the strips (rho additionally before its n round trip, idealmhd.cpp:237 vs :246-247).
__device__ __forceinline__ void record_strips(const DomainParams &P, double *const *strip, int sp, int g, int r, int j, unsigned z,
                                              double rfl, double mx, double my, double mz)
{
    if (reads_interior(P.bc_x1) && g == 2) {
        double *s = strip[0]; s[j] = rfl; s[sp + j] = mx; s[2 * sp + j] = my; s[3 * sp + j] = mz;
    }
    if (reads_interior(P.bc_x2) && g == P.gnx - 3) {
        const bool zz = z & 1u; double *s = strip[1];
        s[j] = rfl; s[sp + j] = zz ? 0.0 : mx; s[2 * sp + j] = zz ? 0.0 : my; s[3 * sp + j] = zz ? 0.0 : mz;
    }
    if (reads_interior(P.bc_y1) && j == 2) {
        const bool zz = z & 3u; double *s = strip[2];
        s[r] = rfl; s[sp + r] = zz ? 0.0 : mx; s[2 * sp + r] = zz ? 0.0 : my; s[3 * sp + r] = zz ? 0.0 : mz;
    }
    if (reads_interior(P.bc_y2) && j == P.ny - 3) {
        const bool zz = z & 7u; double *s = strip[3];
        s[r] = rfl; s[sp + r] = zz ? 0.0 : mx; s[2 * sp + r] = zz ? 0.0 : my; s[3 * sp + r] = zz ? 0.0 : mz;
    }
}

// block-wide NaN-ignoring minimum of positive doubles -> atomicMin on the ordered bit pattern
__device__ __forceinline__ void block_min_impl(double v, unsigned long long *target, unsigned long long *wmin)
{
    unsigned long long b = (v == v) ? (unsigned long long)__double_as_longlong(v) : 0x7FF0000000000000ULL;
    if (v < 0.0) b = 0ULL;   // cannot happen for a valid dt; keeps ordering total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = (t < b) ? t : b;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wmin[w] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int k = 1; k < nw; k++) b = (wmin[k] < b) ? wmin[k] : b;
        atomicMin(target, b);
    }
}
__device__ __forceinline__ void block_min_to_global(double v, unsigned long long *target)
{
    __shared__ unsigned long long wmin[32];
    block_min_impl(v, target, wmin);
}

// ---------------------------------------------------------------------------------------------------------
// Helpers of the fused Runge-Kutta stage kernel k_mhd_stage_xy (mhd_stage_xy.cuh): D = B + (coef*step) * f(S), floors, pointwise boundary
// zeroing, dt minimum.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ double shfl_next(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

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
    double *strip[4];
    int strip_pitch;
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
        const unsigned z = zero_zones(P, g, j);
        record_strips(P, A.strip, A.strip_pitch, g, r, j, z, rfl, mx, my, mz);
        if (z) { mx = 0.0; my = 0.0; mz = 0.0; A.U[E_MX][off] = 0.0; A.U[E_MY][off] = 0.0; A.U[E_MZ][off] = 0.0; }
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
    const double *strip[4];
    int strip_pitch;
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
    // A y side that is `fixed` runs AFTER the x sides and sweeps every i, so it zeroes the momenta an x side has just
    // written into its ghost cells at j <= 2 / j >= ny-3 (evolution.cpp:143,149).  (reflect only sweeps i in [xl,xu].)
    const bool later_zero = A.primary && xside && ((P.bc_y1 == BC_FIXED && idx <= 2) || (P.bc_y2 == BC_FIXED && idx >= P.ny - 3));
    const double *sp = A.strip[side];
    const int spitch = A.strip_pitch;

    if (bc == BC_OPEN_UCNP) {
        // copies the nearest interior cell into both ghost cells for densities, thermal energies, fields, momenta
        // (evolution.cpp:321-331); acts on whichever set is being propagated.
        const int evs[5] = {E_N, E_E, E_BX, E_BY, E_BZ};
#pragma unroll
        for (int k = 0; k < 5; k++) { const double x = A.U[evs[k]][o3]; A.U[evs[k]][o1] = x; A.U[evs[k]][o2] = x; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double m = A.primary ? sp[(k + 1) * spitch + idx] : A.U[E_MX + k][o3];
            if (later_zero) m = 0.0;
            A.U[E_MX + k][o1] = m; A.U[E_MX + k][o2] = m;
        }
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
    const double rho3 = sp[idx];                               // post-floor rho of i3 (before the n round trip)
    const double e3 = A.U[E_E][o3];
    const double rho_1 = A.scale_1[side] * rho3, rho_2 = A.scale_2[side] * rho3;
    A.U[E_E][o1] = smax(A.scale_1[side] * e3, P.e_min);       // derived step re-applies the energy floor (idealmhd.cpp:252)
    A.U[E_E][o2] = smax(A.scale_2[side] * e3, P.e_min);
    const double press = e3 * P.gm1;
    double c_s = 0.0;
    const double c_new = sqrt(P.gamma * press / rho3);
    if (c_new > c_s) c_s = c_new;
    const double mx3 = sp[spitch + idx], my3 = sp[2 * spitch + idx];
    const double vel_x = mx3 / rho3, vel_y = my3 / rho3;
    double boost = A.open_strength * c_s;
    const bool lower = (side == 0 || side == 2);              // i2 > i1 || j2 > j1
    if (lower) boost *= -1.0;
    const double vn = xside ? vel_x : vel_y, vt = xside ? vel_y : vel_x;
    const double bv = lower ? smin(0.0, vn + boost) : smax(0.0, vn + boost);
    const double gv = ddiv(A.dist23[side] * bv - A.h2[side] * vn, A.h3[side], A.rh3[side]);
    double mn1 = rho_1 * gv, mn2 = rho_2 * gv, mt1 = rho_1 * vt, mt2 = rho_2 * vt;
    if (later_zero) { mn1 = mn2 = mt1 = mt2 = 0.0; A.U[E_MZ][o1] = 0.0; A.U[E_MZ][o2] = 0.0; }
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

struct DeriveArgs { const double *U[NEV]; const double *st[NSTATIC]; double *out; int which; int row_off; };   // row_off = -HALO: also the halo rows of a slab

__global__ void __launch_bounds__(256) k_mhd_derive(const DomainParams P, const DeriveArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)blockIdx.y + A.row_off;
    if (j >= P.ny) return;
    const long long off = (long long)r * P.pitch + j;
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
// Slab decomposition: pack the first/last HALO rows of the 8 evolved planes into contiguous staging buffers
// (send_lo = rows [0,2), send_hi = rows [nx-2,nx)), and unpack the neighbours' rows into the halo rows
// (recv_lo -> rows [-2,0), recv_hi -> rows [nx,nx+2)).  Buffer layout: [plane][halo row][pitch].
// ---------------------------------------------------------------------------------------------------------
struct HaloArgs { double *U[NEV]; double *lo, *hi; int unpack; };
__global__ void __launch_bounds__(256) k_halo_copy(const DomainParams P, const HaloArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y / HALO, h = blockIdx.y % HALO;
    if (j >= P.pitch) return;
    const size_t b = ((size_t)v * HALO + h) * P.pitch + j;
    if (!A.unpack) {
        A.lo[b] = A.U[v][(size_t)h * P.pitch + j];
        A.hi[b] = A.U[v][(size_t)(P.nx - HALO + h) * P.pitch + j];
    } else {
        A.U[v][(size_t)(h - HALO) * P.pitch + j] = A.lo[b];          // rows -2,-1 (the allocation starts 2 rows earlier)
        A.U[v][(size_t)(P.nx + h) * P.pitch + j] = A.hi[b];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Peer-store halo exchange over NVLink (slab decomposition, one process per GPU).  Every rank exports ONE device
// segment through CUDA IPC (PeerSegment below); its ring neighbours map it and WRITE their edge rows straight into it
// from the pack kernel (st.global to peer memory travels over NVLink / NVSwitch), then publish a sequence number with a
// system-scope release.  The receiver's unpack kernel acquires that number before it copies the rows into its halo.
// No host synchronisation, no collective library call on the data path: a whole spruce_advance(n) is enqueued at once.
//
// Buffer reuse: exchange number q uses buffer q & 1.  A rank can be at most one exchange ahead of a neighbour (it
// needs that neighbour's rows of exchange q+1 before it can produce exchange q+2, and the neighbour only sends those
// after unpacking exchange q), so two buffers are enough.  The same argument covers the dt slots (one global
// all-gather per step, which every rank completes before it starts the next step).
// ---------------------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 16;
struct PeerFlags {                                            // one 128-byte line per independently written word
    unsigned long long halo_seq[2][16];                       // [side][0]: last exchange written by my lower / upper neighbour
    unsigned long long dt_seq[MAX_RANKS][16];                 // [r][0]: last step whose local dt minimum rank r has stored
    unsigned long long dt_bits[2][MAX_RANKS];                 // [parity][r]
    unsigned long long red_seq[MAX_RANKS][16];                // [r][0]: last module reduction rank r has stored
    unsigned long long red_bits[2][MAX_RANKS][4];             // [parity][r][min, max, min, max] (bit patterns of non-negative doubles)
    unsigned long long full_seq[MAX_RANKS][16];               // [r][0]: last re-evaluated dt minimum (skip window missed) rank r has stored
    unsigned long long full_bits[2][MAX_RANKS];               // [parity][r]
    int error; int pad[31];                                   // set when a wait timed out (peer died): every later launch drains
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= seq; gives up after ~20 s (a dead peer must not hang the box) and raises the segment's error word
__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq, int *error)
{
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(64);
        if ((++spins & 0x3FFu) == 0) {
            if (*(volatile int *)error) return false;
            if (global_timer_ns() - t0 > 20000000000ULL) { atomicExch(error, 1); return false; }
        }
    }
    return true;
}

struct PushArgs {
    const double *U[NEV];
    double *peer_lo_buf, *peer_hi_buf;          // lower neighbour's "from above" buffer / upper neighbour's "from below" buffer (null: no neighbour)
    unsigned long long *peer_lo_flag, *peer_hi_flag;
    unsigned long long seq;
    unsigned int *counter;                       // local: blocks that have finished their stores
    const int *done_ptr;
};
__global__ void __launch_bounds__(256) k_halo_push(const DomainParams P, const PushArgs A)
{
    if (*A.done_ptr) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y / HALO, h = blockIdx.y % HALO;
    if (j < P.pitch) {
        const size_t b = ((size_t)v * HALO + h) * P.pitch + j;
        if (A.peer_lo_buf) A.peer_lo_buf[b] = A.U[v][(size_t)h * P.pitch + j];
        if (A.peer_hi_buf) A.peer_hi_buf[b] = A.U[v][(size_t)(P.nx - HALO + h) * P.pitch + j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(A.counter, 1u) == total - 1) {        // last block: every block's rows are visible system-wide
            *A.counter = 0;
            __threadfence_system();
            if (A.peer_lo_flag) st_release_sys(A.peer_lo_flag, A.seq);
            if (A.peer_hi_flag) st_release_sys(A.peer_hi_flag, A.seq);
        }
    }
}

struct PullArgs {
    double *U[NEV];
    const double *lo_buf, *hi_buf;               // my segment: rows from the lower / upper neighbour (null: physical boundary)
    const unsigned long long *lo_flag, *hi_flag;
    unsigned long long seq;
    int *error;
    const int *done_ptr;
};
__global__ void __launch_bounds__(256) k_halo_pull(const DomainParams P, const PullArgs A)
{
    if (*A.done_ptr) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        bool good = true;
        if (A.lo_buf) good = wait_seq(A.lo_flag, A.seq, A.error);
        if (good && A.hi_buf) good = wait_seq(A.hi_flag, A.seq, A.error);
        ok = good ? 1 : 0;
    }
    __syncthreads();
    if (!ok) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y / HALO, h = blockIdx.y % HALO;
    if (j >= P.pitch) return;
    const size_t b = ((size_t)v * HALO + h) * P.pitch + j;
    if (A.lo_buf) A.U[v][(size_t)(h - HALO) * P.pitch + j] = __ldcv(A.lo_buf + b);
    if (A.hi_buf) A.U[v][(size_t)(P.nx + h) * P.pitch + j] = __ldcv(A.hi_buf + b);
}

// ---------------------------------------------------------------------------------------------------------
// Scalar bookkeeping of advanceTime (evolution.cpp:62,80-81), one thread.
// ctl[0] = step (double), ctl[1] = time, ctl[2] = max_time (<=0: none); ictl[0] = iter, ictl[1] = done flag
// ---------------------------------------------------------------------------------------------------------
struct StepCtl { double step, time, max_time, epsilon; long long iter; int done; int need_full; unsigned long long dtmin_bits;
                 double inv_thr; unsigned long long thr_bits; double prune_factor; unsigned long long full_seq; };

// begin: step = epsilon * min(dt) ; reset the running minimum for the propagate at the end of this step
__global__ void k_step_begin(StepCtl *c, double *dt_hist, int slot)
{
    if (c->max_time > 0.0 && !(c->time < c->max_time)) c->done = 1;
    if (c->done) { return; }
    c->step = c->epsilon * __longlong_as_double((long long)c->dtmin_bits);   // evolution.cpp:62
    if (dt_hist) dt_hist[slot] = c->step;
}
// reset the running minimum; prune != 0: the stage kernel that follows may skip cells whose dt is certainly above F * (current minimum)
__global__ void k_dtmin_reset(StepCtl *c, int prune)
{
    if (c->done) return;
    const double prev = __longlong_as_double((long long)c->dtmin_bits);
    c->need_full = 0;
    if (prune && c->prune_factor > 0.0 && prev > 0.0 && prev < 1.0e300) {
        const double thr = c->prune_factor * prev;
        c->thr_bits = (unsigned long long)__double_as_longlong(thr);
        c->inv_thr = 1.0 / thr;
    } else { c->inv_thr = 0.0; c->thr_bits = 0x7FF0000000000000ULL; }
    c->dtmin_bits = 0x7FEFFFFFFFFFFFFFULL;
}
// after the primary stage (and, on slabs, the all-gather): was the new minimum inside the window the skip test assumed?
__global__ void k_dt_validate(StepCtl *c) { if (!c->done) c->need_full = (c->inv_thr > 0.0 && c->dtmin_bits > c->thr_bits) ? 1 : 0; }
__global__ void k_step_end(StepCtl *c) { if (c->done) return; c->time += c->step; c->iter += 1; }

// the rare fallback of the skip test: dt of every interior cell of the primary state, exact (recomputeDT, idealmhd.cpp:279-304)
struct DtFullArgs { const double *U[NEV]; const double *st[NSTATIC]; StepCtl *ctl; };
__global__ void __launch_bounds__(256) k_dt_full(const DomainParams P, const DtFullArgs A)
{
    if (A.ctl->done || !A.ctl->need_full) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double dtc = 1.7976931348623157e308;
    for (int r = blockIdx.y; r < P.nx; r += gridDim.y) {       // a few row-strided CTAs: the launch that returns at once stays cheap
        const int g = P.row0 + r;
        if (j < P.ny && g >= P.xl && g <= P.xu && j >= P.yl && j <= P.yu) {
            const size_t off = (size_t)r * P.pitch + j;
            const double c = cell_dt(P, A.U[E_N][off] * P.m_i, A.U[E_MX][off], A.U[E_MY][off], A.U[E_E][off], A.st[S_BEX][off] + A.U[E_BX][off],
                                     A.st[S_BEY][off] + A.U[E_BY][off], A.st[S_BEZ][off] + A.U[E_BZ][off], P.tx.d[r], P.tx.rd[r], P.ty.d[j], P.ty.rd[j]);
            dtc = smin(dtc, c);
        }
    }
    block_min_to_global(dtc, &A.ctl->dtmin_bits);
}

// dt all-gather over peer memory: thread r stores my local minimum into rank r's slot, then publishes the step number;
// the collect kernel waits for every rank's number and takes the minimum (min of positive doubles == min of their bit patterns)
struct DtGatherArgs { StepCtl *ctl; PeerFlags *peer[MAX_RANKS]; PeerFlags *mine; int rank, world; unsigned long long seq; };
__global__ void k_dt_publish(const DtGatherArgs A)
{
    if (A.ctl->done) return;
    const int r = threadIdx.x;
    if (r >= A.world) return;
    const unsigned long long bits = A.ctl->dtmin_bits;
    PeerFlags *f = A.peer[r];
    f->dt_bits[A.seq & 1][A.rank] = bits;
    __threadfence_system();
    st_release_sys(&f->dt_seq[A.rank][0], A.seq);
}
__global__ void k_dt_collect(const DtGatherArgs A)
{
    if (A.ctl->done) return;
    __shared__ unsigned long long m[MAX_RANKS];
    const int r = threadIdx.x;
    if (r < A.world) {
        const bool ok = wait_seq(&A.mine->dt_seq[r][0], A.seq, &A.mine->error);
        m[r] = ok ? *(volatile unsigned long long *)&A.mine->dt_bits[A.seq & 1][r] : 0x7FEFFFFFFFFFFFFFULL;
    }
    __syncthreads();
    if (r == 0) {
        unsigned long long best = m[0];
        for (int k = 1; k < A.world; k++) best = m[k] < best ? m[k] : best;
        A.ctl->dtmin_bits = best;
        if (A.mine->error) A.ctl->done = 2;               // a peer stopped answering: drain the remaining launches
    }
}

// ---- the fused step control of plain runs (no modules, no open_moc): three one-block kernels per step instead of eight.
//   k_step_open   (first step of an advance call)   begin + reset
//   k_step_mid    after the primary stage           [slabs: all-gather of the dt minimum] + window check
//   k_dt_full     (returns at once unless the window was missed)
//   k_step_close                                    [slabs, window missed: all-gather again] + end of this step + begin / reset of the next one
// One regular all-gather per step with sequence number = step count (buffer parity alternates per step: when a rank publishes gather k+1
// every rank has finished reading gather k-1, see the halo argument above); the rare second gather has its own words and a device-side counter.
__device__ __forceinline__ void step_begin_reset(StepCtl *c, double *dt_hist, int slot, int prune)
{
    if (c->max_time > 0.0 && !(c->time < c->max_time)) c->done = 1;
    if (c->done) return;
    const double prev = __longlong_as_double((long long)c->dtmin_bits);
    c->step = c->epsilon * prev;                                             // evolution.cpp:62
    if (dt_hist) dt_hist[slot] = c->step;
    c->need_full = 0;
    if (prune && c->prune_factor > 0.0 && prev > 0.0 && prev < 1.0e300) {
        const double thr = c->prune_factor * prev;
        c->thr_bits = (unsigned long long)__double_as_longlong(thr);
        c->inv_thr = 1.0 / thr;
    } else { c->inv_thr = 0.0; c->thr_bits = 0x7FF0000000000000ULL; }
    c->dtmin_bits = 0x7FEFFFFFFFFFFFFFULL;
}
__global__ void k_step_open(StepCtl *c, double *dt_hist, int slot, int prune) { step_begin_reset(c, dt_hist, slot, prune); }

// all MAX_RANKS threads of the block: exchange `mine_bits` with every rank through the given words, return the minimum in thread 0
__device__ __forceinline__ unsigned long long block_allgather_min(const DtGatherArgs &A, unsigned long long mine_bits, unsigned long long seq, int full, unsigned long long *m)
{
    const int r = threadIdx.x;
    if (r < A.world) {
        PeerFlags *f = A.peer[r];
        if (full) f->full_bits[seq & 1][A.rank] = mine_bits; else f->dt_bits[seq & 1][A.rank] = mine_bits;
        __threadfence_system();
        st_release_sys(full ? &f->full_seq[A.rank][0] : &f->dt_seq[A.rank][0], seq);
        const bool ok = wait_seq(full ? &A.mine->full_seq[r][0] : &A.mine->dt_seq[r][0], seq, &A.mine->error);
        m[r] = ok ? *(volatile unsigned long long *)(full ? &A.mine->full_bits[seq & 1][r] : &A.mine->dt_bits[seq & 1][r]) : 0x7FEFFFFFFFFFFFFFULL;
    }
    __syncthreads();
    unsigned long long best = m[0];
    if (r == 0) for (int k = 1; k < A.world; k++) best = m[k] < best ? m[k] : best;
    return best;
}
__global__ void k_step_mid(const DtGatherArgs A)
{
    __shared__ unsigned long long m[MAX_RANKS];
    StepCtl *c = A.ctl;
    if (c->done) return;
    if (A.world > 1) {
        const unsigned long long best = block_allgather_min(A, c->dtmin_bits, A.seq, 0, m);
        if (threadIdx.x == 0) { c->dtmin_bits = best; if (A.mine->error) c->done = 2; }
    }
    if (threadIdx.x == 0) c->need_full = (c->inv_thr > 0.0 && c->dtmin_bits > c->thr_bits) ? 1 : 0;
}
__global__ void k_step_close(const DtGatherArgs A, double *dt_hist, int next_slot, int prune)
{
    __shared__ unsigned long long m[MAX_RANKS];
    StepCtl *c = A.ctl;
    if (c->done) return;
    if (A.world > 1 && c->need_full) {                                       // block-uniform: every thread reads the same words
        const unsigned long long fs = c->full_seq + 1;
        const unsigned long long best = block_allgather_min(A, c->dtmin_bits, fs, 1, m);
        if (threadIdx.x == 0) { c->dtmin_bits = best; c->full_seq = fs; if (A.mine->error) c->done = 2; }
    }
    if (threadIdx.x == 0 && !c->done) {
        c->time += c->step; c->iter += 1;                                    // evolution.cpp:80-81
        if (next_slot >= 0) step_begin_reset(c, dt_hist, next_slot, prune);
    }
}

// all-gather + reduce of the four module reduction words (sub-cycle counts): slots 0, 2 are minima, 1, 3 maxima
struct RedGatherArgs { unsigned long long *red; PeerFlags *peer[MAX_RANKS]; PeerFlags *mine; int rank, world; unsigned long long seq; };
__global__ void k_red_publish(const RedGatherArgs A)
{
    const int r = threadIdx.x;
    if (r >= A.world) return;
    PeerFlags *f = A.peer[r];
#pragma unroll
    for (int k = 0; k < 4; k++) f->red_bits[A.seq & 1][A.rank][k] = A.red[k];
    __threadfence_system();
    st_release_sys(&f->red_seq[A.rank][0], A.seq);
}
__global__ void k_red_collect(const RedGatherArgs A)
{
    __shared__ unsigned long long m[MAX_RANKS][4];
    __shared__ int okf[MAX_RANKS];
    const int r = threadIdx.x;
    if (r < A.world) {
        okf[r] = wait_seq(&A.mine->red_seq[r][0], A.seq, &A.mine->error) ? 1 : 0;
        for (int k = 0; k < 4; k++) m[r][k] = *(volatile unsigned long long *)&A.mine->red_bits[A.seq & 1][r][k];
    }
    __syncthreads();
    if (r < 4) {
        unsigned long long best = m[0][r];
        for (int k = 1; k < A.world; k++) { const unsigned long long v = m[k][r]; best = (r & 1) ? (v > best ? v : best) : (v < best ? v : best); }
        A.red[r] = best;
    }
}

} // namespace spruce
