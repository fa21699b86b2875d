// mhd_stage_xy.cuh -- the fused Runge-Kutta stage kernel, direction-specialised warps ("v5").
//
// A plain column-marching kernel (one thread per column doing both directions, round 1's first version) is bound by dependency latency at 8 warps/SM: the shared-memory ring (5 rows x 14 arrays)
// and ~200 registers per thread cap the residency.  Here a CTA still owns 62 columns and marches along x, but FOUR
// warps share the ring:
//     warps 0,1 ("X")  evaluate the x-direction faces of the transported quantities (flux carried row to row),
//                      the x central derivatives, and finish the outputs  n, mom_x, mom_y, mom_z;
//     warps 2,3 ("Y")  evaluate the y-direction faces (left face per lane, right face by shuffle), the y central
//                      derivatives, finish  thermal_energy, bi_x, bi_y, bi_z  and the cell's dt.
// Partial results cross through small shared-memory arrays at one barrier per row (two in the final stage, for dt).
// Per-thread state halves (<= 128 registers), 16 warps/SM are resident on the same shared-memory budget.
//
// Zero planes: the z system {mom_z, bi_z} stays identically zero when mom_z, bi_z and be_z start as zero planes
// (idealmhd.cpp:72-73,84-86: every term of d(mom_z)/dt and d(bi_z)/dt carries one of them), and a zero external-field
// plane is static.  The host tracks this (capi.cu: active_quantities) and hands the kernel the list of transported
// quantities that can be non-zero; the others contribute exact zeros in the reference too and are not evaluated.
#pragma once
#include "mhd_kernels.cuh"
#include "chunk_plan.hpp"

namespace spruce {

constexpr int XY_NT = 128;                   // 2 X warps + 2 Y warps
constexpr int XY_RV = 3;                     // velocity ring depth (rows r, r+1 in use, r+2 being formed)
constexpr int XY_CHUNK = 56;                 // rows per CTA (upper bound: sizes the x tables in shared memory)
constexpr int XY_EDGE_DELTA = 8;             // a slab's first / last CTA rows are this much shorter than the interior ones when the halo exchange overlaps the interior
constexpr int XY_XT = XY_CHUNK + 6;          // x-table entries: local rows -3 .. chunk+2
constexpr int XW = 64;                       // width of the per-column exchange arrays: one private slot per (warp column, lane)
// shared memory (doubles): transported ring, velocity ring, x tables, x-flux carry, the x parts of the transports the Y warps finish (TX), the y parts
// the X warps finish (TY), derivative exchange, dt exchange (read by the Y warps one iteration later, before the X warps overwrite it), the own-cell
// values (gravity, base state) and, for the 2-D instance, the y geometry tables.
// nq = quantity rows kept in the ring and in the flux carry: all 11, or only the 6 of the 2-D instance.
__host__ __device__ constexpr int xy_ntx(int nq) { return nq == 6 ? 3 : 7; }         // thermal_energy, bi_x, bi_y (+ bi_z, be_x, be_y, be_z)
__host__ __device__ constexpr int xy_nty(int nq) { return nq == 6 ? 3 : 4; }         // rho, mom_x, mom_y (+ mom_z)
__host__ __device__ constexpr int xy_off_vel(int nq) { return RD * nq * SW; }
__host__ __device__ constexpr int xy_off_xt(int nq) { return xy_off_vel(nq) + XY_RV * 3 * SW; }
__host__ __device__ constexpr int xy_off_fx(int nq) { return xy_off_xt(nq) + 7 * XY_XT; }
__host__ __device__ constexpr int xy_off_tx(int nq) { return xy_off_fx(nq) + nq * XW; }
__host__ __device__ constexpr int xy_off_ty(int nq) { return xy_off_tx(nq) + xy_ntx(nq) * XW; }
__host__ __device__ constexpr int xy_off_dc(int nq) { return xy_off_ty(nq) + xy_nty(nq) * XW; }
__host__ __device__ constexpr int xy_off_dt(int nq) { return xy_off_dc(nq) + 3 * XW; }             // aliases Dc_s[3..5]: a thread reads its Dc/Dt slot before it overwrites it
__host__ __device__ constexpr int xy_off_own(int nq) { return xy_off_dc(nq) + 6 * XW; }
// own-cell rows: grav_x, grav_y, and -- unless the stage's base state is the state it differentiates (first rk stage, euler) -- the base state B
__host__ __device__ constexpr int xy_nown(int nq, int var) { return ((var & 3) == 1 || (var & 3) == 3) ? 2 : 2 + (nq == 6 ? 6 : 8); }
// y geometry: the 2-D instance (96 registers) keeps the seven y tables of its 66 columns in shared memory instead of ten doubles per thread in registers
__host__ __device__ constexpr int xy_nyt(int nq) { return nq == 6 ? 7 * SW : 0; }
__host__ __device__ constexpr int xy_off_yt(int nq, int var) { return xy_off_own(nq) + xy_nown(nq, var) * XW; }
__host__ __device__ constexpr int xy_doubles(int nq, int var) { return xy_off_yt(nq, var) + xy_nyt(nq); }
__host__ __device__ constexpr size_t xy_smem_bytes(int nq, int var) { return (size_t)xy_doubles(nq, var) * sizeof(double); }
__host__ __device__ constexpr int xy_rows(int ln) { return ln == 6 ? 6 : NTR; }
__host__ __device__ constexpr int xy_ctas_per_sm(int ln) { return ln == 6 ? 5 : 4; }               // the 2-D instance: <= 45 KB of shared memory and <= 96 registers -> 20 warps / SM
static_assert(XY_CHUNK == PLAN_MAX_ROWS && XY_EDGE_DELTA == PLAN_EDGE_DELTA && HALO == PLAN_HALO, "chunk_plan.hpp mirrors these constants");
static_assert(xy_smem_bytes(NTR, 0) <= 57344, "four CTAs per SM need <= 56 KB each");
static_assert(xy_smem_bytes(6, 0) <= 45000, "five CTAs per SM need <= 45 KB each");

struct ActiveList { int n; unsigned long long q; };   // transported quantities that can be non-zero (one nibble each), padded to an even count

// LN > 0: the active list is the compile-time constant (LN, LQ) -- the quantity loops unroll completely, every shared-memory
// offset and every "is this bi_y?" test folds away.  LN == 0: the list comes from the launch argument (rolled loops).
constexpr unsigned long long xy_list(int a = -1, int b = -1, int c = -1, int d = -1, int e = -1, int f = -1, int g = -1, int h = -1, int i = -1, int j = -1, int k = -1, int l = -1)
{
    const int v[12] = {a, b, c, d, e, f, g, h, i, j, k, l};
    unsigned long long q = 0;
    for (int t = 0; t < 12; t++) if (v[t] >= 0) q |= (unsigned long long)v[t] << (4 * t);
    return q;
}
constexpr unsigned long long XY_LIST_2D = xy_list(Q_RHO, Q_E, Q_MX, Q_MY, Q_BIX, Q_BIY);                                              // no z system, no external field
constexpr unsigned long long XY_LIST_FULL = xy_list(Q_RHO, Q_E, Q_MX, Q_MY, Q_BIX, Q_BIY, Q_MZ, Q_BIZ, Q_BEX, Q_BEY, Q_BEZ, Q_BEZ);   // everything (padded to 12)

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}

// VAR > 0: the integrator stage is a compile-time constant too (no K planes, no module right-hand-side terms).  VAR & 3 =
//   1: B == S, secondary (first stage of rk2)   2: B != S, primary (last stage of rk2)   3: B == S, primary (euler).
// VAR == 0 takes kmode / b_is_s / primary / n_xterm from the launch arguments (every integrator, every module).
//
// How rows reach shared memory.  A CTA whose 66 staged columns lie inside the row (every strip except the first and the last one or two) takes the
// VECTOR path: the two Y warps copy ring row r+3 with 16-byte cp.async (a 528-byte row is 33 chunks: one per lane, lane 0 takes the last one too;
// warp 2 the first half of the quantities, warp 3 the second half), 512 contiguous bytes per instruction.  Strips that touch the y boundary (periodic
// wrap or out-of-range columns, where a chunk could straddle the seam) keep the per-column path (8-byte cp.async by the 66 loader threads).  The
// own-cell values of row r (gravity, base state) are per-thread cp.async copies into private shared-memory slots, issued at the top of iteration r
// and waited for before phase 2: they occupy no register during phase 1.
// (Measured and not kept: cp.async.bulk + mbarrier for the same rows -- every bulk copy costs ~13 uniform-datapath instructions around the
//  UBLKCP, issued by single lanes; 4096^2 exact: 2.104 ms/step against 2.072 with the per-column path)
template <int LN, unsigned long long LQ, int VAR = 0>
__global__ void __launch_bounds__(XY_NT, xy_ctas_per_sm(LN)) k_mhd_stage_xy(const DomainParams P, const StageArgs A, const ActiveList Larg)
{
    constexpr int UNR = LN > 0 ? LN : 1;
#define kmode (VAR ? (int)KM_NONE : A.kmode)
#define b_is_s (VAR ? ((VAR & 3) != 2 ? 1 : 0) : A.b_is_s)
#define primary (VAR ? ((VAR & 3) != 1 ? 1 : 0) : A.primary)
#define n_xterm (VAR ? 0 : A.n_xterm)
    // Z: the z system (mom_z, bi_z, v_z) and the external field can be non-zero.  In the 2-D instance (LN == 6) they are exact zeros in the
    // reference as well, so every term that only adds +-0 is left out (the results can differ in the sign of a zero, nothing else) and
    // the all-zero planes mom_z / bi_z are neither computed nor stored (their planes are zero-initialised and stay zero).
    constexpr bool Z = (LN != 6);
    ActiveList L;
    L.n = LN > 0 ? LN : Larg.n;
    L.q = LN > 0 ? LQ : Larg.q;
    extern __shared__ __align__(16) double smem[];
    if (*A.done_ptr) return;
    constexpr int NQ = xy_rows(LN);
    constexpr int NOWN = xy_nown(NQ, VAR);
    constexpr bool HAVE_B = NOWN > 2;                                            // the own-cell buffer has rows for the base state
    constexpr int NB = Z ? 4 : 3;                                                // base-state rows per role
    // row of quantity q inside a ring slot / the flux carry: identity, or the compact order rho, mom_x, mom_y, e, bi_x, bi_y of the 2-D instance
    auto QI = [](int q) { return LN == 6 ? (q < Q_MZ ? q : q - 1) : q; };
    double *ringp = smem;
    double (*vel)[3][SW] = reinterpret_cast<double (*)[3][SW]>(smem + xy_off_vel(NQ));
    double (*xt)[XY_XT] = reinterpret_cast<double (*)[XY_XT]>(smem + xy_off_xt(NQ));
    double *Fx_p = smem + xy_off_fx(NQ);                                          // [quantity row][XW]
    double *TX_p = smem + xy_off_tx(NQ);                                          // x parts of transportDivergence2D of the quantities the Y warps finish
    double *TY_p = smem + xy_off_ty(NQ);                                          // y parts ... the X warps finish
    double (*Dc_s)[XW] = reinterpret_cast<double (*)[XW]>(smem + xy_off_dc(NQ));
    double (*Dt_s)[XW] = reinterpret_cast<double (*)[XW]>(smem + xy_off_dt(NQ));
    double (*own)[XW] = reinterpret_cast<double (*)[XW]>(smem + xy_off_own(NQ));  // [0] grav_x [1] grav_y [2..] B: n, mom_x, mom_y, (mom_z), e, bi_x, bi_y, (bi_z)
    constexpr bool YTAB = xy_nyt(NQ) > 0;
    double (*yt)[SW] = reinterpret_cast<double (*)[SW]>(smem + xy_off_yt(NQ, VAR));   // [h, fs, rfs, ep, em, d, rd][shared column]  (YTAB)
#define RG(sl, q, cc) ringp[((sl) * NQ + QI(q)) * SW + (cc)]
#define FXS(q, cc) Fx_p[QI(q) * XW + (cc)]
#define TXS(q, cc) TX_p[((q) - Q_E) * XW + (cc)]
#define TYS(q, cc) TY_p[(q) * XW + (cc)]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool isX = warp < 2;
    const int wcol = warp & 1;
    const int ccol = 31 * wcol + lane;            // column inside the CTA; lane 31 duplicates the next warp's lane 0
    const int col = 32 * wcol + lane;             // private slot in the per-column exchange arrays (the X and the Y thread of a
                                                  //  column share it; lane 31 duplicates a column but owns its own slot)
    const int fcol = col;
    const int j0 = blockIdx.x * CW;
    const int j = j0 + ccol;
    const int c = ccol + HALO;
    const bool edge2 = A.edge2_begin >= 0 && blockIdx.y == 1;
    const int r0 = edge2 ? A.edge2_begin : A.row_begin + (int)blockIdx.y * A.chunk_rows;
    const int r1 = edge2 ? A.edge2_end : min(r0 + A.chunk_rows, A.row_end);
    const bool col_out = (lane < 31) && (j < P.ny);
    const bool vec = A.vec16 && j0 >= HALO && j0 + CW + HALO <= P.ny;             // CTA-uniform: all 66 staged columns are inside the row -> 16-byte copies

    // ---- per-thread loader (strips at the y boundary, rows beyond a physical x boundary): thread t < SW owns shared column t
    const bool loader = tid < SW;
    int jl = j0 - HALO + tid;
    bool jl_ok = loader && (jl >= 0 && jl < P.ny);
    if (loader && !jl_ok && P.yper) { jl = (jl + 2 * P.ny) % P.ny; jl_ok = true; }
    auto slot_of = [&](int r) { return (r - r0 + HALO + RD) % RD; };
    auto vslot_of = [&](int r) { return (r - r0 + HALO + 3 * XY_RV) % XY_RV; };
    auto next_slot = [](int s_) { return s_ == RD - 1 ? 0 : s_ + 1; };
    auto next_vslot = [](int s_) { return s_ == XY_RV - 1 ? 0 : s_ + 1; };
    auto wrap_row = [&](int r) { return P.xwrap ? (r < 0 ? r + P.nx : (r >= P.nx ? r - P.nx : r)) : r; };   // single-rank periodic x: one add, no modulo
    auto fill_row = [&](int slot) {                              // a row beyond a physical x boundary: never used by an in-range operator
        if (!loader) return;
        if (LN == 6) {
            RG(slot, Q_RHO, tid) = 1.0; RG(slot, Q_MX, tid) = 0.0; RG(slot, Q_MY, tid) = 0.0;
            RG(slot, Q_E, tid) = 0.0; RG(slot, Q_BIX, tid) = 0.0; RG(slot, Q_BIY, tid) = 0.0;
        } else {
#pragma unroll
            for (int v = 0; v < NTR; v++) RG(slot, v, tid) = (v == Q_RHO) ? 1.0 : 0.0;
        }
    };
    auto issue_row = [&](int r, int slot) {                      // per-thread path
        if (!loader) return;
        if (row_exists(P, r) && jl_ok) {
            const size_t off = (size_t)wrap_row(r) * P.pitch + jl;
            if (LN == 6) {
                // 2-D list: mom_z, bi_z and the external field are identically zero planes (host-tracked) and have no ring rows
                cp_async8(&RG(slot, Q_RHO, tid), A.S[E_N] + off); cp_async8(&RG(slot, Q_MX, tid), A.S[E_MX] + off);
                cp_async8(&RG(slot, Q_MY, tid), A.S[E_MY] + off); cp_async8(&RG(slot, Q_E, tid), A.S[E_E] + off);
                cp_async8(&RG(slot, Q_BIX, tid), A.S[E_BX] + off); cp_async8(&RG(slot, Q_BIY, tid), A.S[E_BY] + off);
            } else {
#pragma unroll
                for (int v = 0; v < NEV; v++) cp_async8(&RG(slot, v, tid), A.S[v] + off);
                cp_async8(&RG(slot, Q_BEX, tid), A.st[S_BEX] + off);
                cp_async8(&RG(slot, Q_BEY, tid), A.st[S_BEY] + off);
                cp_async8(&RG(slot, Q_BEZ, tid), A.st[S_BEZ] + off);
            }
        } else fill_row(slot);
    };
    constexpr int NLD = (LN == 6 ? 6 : NLOAD), NLD_A = NLD / 2;                  // planes per ring row; warp 2 copies the first NLD_A, warp 3 the others
    auto vec_row = [&](int r, int slot) {                        // vector path, warps 2 and 3: lane l owns columns 2l, 2l+1 (lane 0 also 64, 65)
        if (warp < 2) return;
        const bool ex = row_exists(P, r);
        const size_t off = (size_t)wrap_row(r) * P.pitch + (j0 - HALO) + 2 * lane;
        auto one = [&](int q, const double *plane) {
            double *dst = &RG(slot, q, 2 * lane);
            if (ex) { cp_async16(dst, plane + off); if (lane == 0) cp_async16(dst + 64, plane + off + 64); }
            else {                                               // a row beyond a physical x boundary
                const double f = (q == Q_RHO) ? 1.0 : 0.0;
                *reinterpret_cast<double2 *>(dst) = make_double2(f, f);
                if (lane == 0) *reinterpret_cast<double2 *>(dst + 64) = make_double2(f, f);
            }
        };
        if (LN == 6) {
            if (warp == 2) { one(Q_RHO, A.S[E_N]); one(Q_MX, A.S[E_MX]); one(Q_MY, A.S[E_MY]); }
            else { one(Q_E, A.S[E_E]); one(Q_BIX, A.S[E_BX]); one(Q_BIY, A.S[E_BY]); }
        } else if (warp == 2) {
#pragma unroll
            for (int v = 0; v < NLD_A; v++) one(v, A.S[v]);
        } else {
#pragma unroll
            for (int v = NLD_A; v < NEV; v++) one(v, A.S[v]);
            one(Q_BEX, A.st[S_BEX]); one(Q_BEY, A.st[S_BEY]); one(Q_BEZ, A.st[S_BEZ]);
        }
    };
    // own-cell values of row r (a row of the slab itself: it always exists): gravity for the X warps, the base state for both roles.  Every thread
    // copies its own (cp.async into its private slots): no register is held across phase 1 and nothing has to be visible to another thread.
    const bool want_b = HAVE_B && !b_is_s;
    const bool grav = A.grav != 0;                               // a gravity plane is non-zero (host-tracked); otherwise rho * 0.0 is added as the reference does
    auto own_row = [&](int r) {
        if (!col_out) return;
        const size_t off = (size_t)r * P.pitch + j;
        if (isX) {
            if (grav) { cp_async8(&own[0][ccol], A.st[S_GX] + off); cp_async8(&own[1][ccol], A.st[S_GY] + off); }
            if (HAVE_B && want_b) {
                cp_async8(&own[2][ccol], A.B[E_N] + off); cp_async8(&own[3][ccol], A.B[E_MX] + off); cp_async8(&own[4][ccol], A.B[E_MY] + off);
                if (Z) cp_async8(&own[HAVE_B ? 5 : 0][ccol], A.B[E_MZ] + off);
            }
        } else if (HAVE_B && want_b) {
            cp_async8(&own[HAVE_B ? 2 + NB : 0][ccol], A.B[E_E] + off); cp_async8(&own[HAVE_B ? 3 + NB : 0][ccol], A.B[E_BX] + off);
            cp_async8(&own[HAVE_B ? 4 + NB : 0][ccol], A.B[E_BY] + off);
            if (Z) cp_async8(&own[HAVE_B ? 5 + NB : 0][ccol], A.B[E_BZ] + off);
        }
    };
    // rho = n * m_i (idealmhd.cpp:247) for the 66 columns of a ring row: warp 2, the columns each lane copied itself on the vector path
    auto convert_rho = [&](int s_) {
        if (warp != 2) return;
        double2 *p2 = reinterpret_cast<double2 *>(&RG(s_, Q_RHO, 0));
        double2 v = p2[lane]; v.x = v.x * P.m_i; v.y = v.y * P.m_i; p2[lane] = v;
        if (lane == 0) { double2 w = p2[32]; w.x = w.x * P.m_i; w.y = w.y * P.m_i; p2[32] = w; }
    };
    // v = mom / rho (idealmhd.cpp:248-250): one IEEE reciprocal, then the exact-division correction per component (exact_math.cuh).
    // Velocities are read at shared columns 1 .. 64 only (own column and the left neighbour of a y face): warps 0 and 1, one column per thread.
    auto form_vel = [&](int s_, int v_) {
        if (warp >= 2) return;
        const int cc = tid + 1;
        const double rho = RG(s_, Q_RHO, cc);
        const double rr = 1.0 / rho;
        vel[v_][0][cc] = ddiv(RG(s_, Q_MX, cc), rho, rr);
        vel[v_][1][cc] = ddiv(RG(s_, Q_MY, cc), rho, rr);
        vel[v_][2][cc] = (LN == 6) ? 0.0 : ddiv(RG(s_, Q_MZ, cc), rho, rr);      // the 2-D list is only chosen when mom_z is identically zero
    };

    const double step = *A.step_ptr;
    const double s = A.coef * step;
    const double inv_thr = primary ? *A.inv_thr_ptr : 0.0;

    // ---- x tables of this chunk and zeroed exchange arrays (inactive quantities read as exact zeros)
    {
        const double *src[7] = {P.tx.h, P.tx.fs, P.tx.rfs, P.tx.ep, P.tx.em, P.tx.d, P.tx.rd};
        const int nent = (r1 - r0) + 6;     // local rows -3 .. chunk+2: everything x_geom() touches, nothing past the table apron
        for (int e = tid; e < 7 * XY_XT; e += XY_NT) {
            const int t = e / XY_XT, i = e - t * XY_XT;
            if (i < nent) xt[t][i] = src[t][r0 - 3 + i];
        }
        for (int e = tid; e < xy_off_own(NQ) - xy_off_fx(NQ); e += XY_NT) Fx_p[e] = 0.0;              // Fx, TX, TY, Dc are contiguous
        if (YTAB) {
            const double *srcy[7] = {P.ty.h, P.ty.fs, P.ty.rfs, P.ty.ep, P.ty.em, P.ty.d, P.ty.rd};
            for (int e = tid; e < 7 * SW; e += XY_NT) {
                const int t = e / SW, i = e - t * SW;
                yt[t][i] = srcy[t][min(j0 - HALO + i, P.ny + 2)];          // shared column i = global column j0 - 2 + i; the tables reach ny + 2
            }
        }
    }
    auto x_geom = [&](int f) {
        const int i = f - r0 + 3;
        FaceGeom g;
        g.hm1 = xt[0][i - 1]; g.h0 = xt[0][i];
        g.fs = xt[1][i];      g.rfs = xt[2][i];
        g.ep = xt[3][i];      g.fsm = xt[1][i - 1]; g.rfsm = xt[2][i - 1];
        g.em = xt[4][i];      g.fsp = xt[1][i + 1]; g.rfsp = xt[2][i + 1];
        return g;
    };

    // ---- per-thread y geometry (Y warps: left face of column j; both roles: cell size)
    const int jt = min(j, P.ny + 1);
    FaceGeom gy_reg{};
    double dy_reg = 1.0, rdy_reg = 1.0;
    if (!YTAB) { gy_reg = load_face_geom(P.ty, jt); dy_reg = P.ty.d[min(j, P.ny)]; rdy_reg = P.ty.rd[min(j, P.ny)]; }
    auto y_geom = [&]() {
        if (!YTAB) return gy_reg;
        FaceGeom g;
        g.hm1 = yt[0][c - 1]; g.h0 = yt[0][c];
        g.fs = yt[1][c];      g.rfs = yt[2][c];
        g.ep = yt[3][c];      g.fsm = yt[1][c - 1]; g.rfsm = yt[2][c - 1];
        g.em = yt[4][c];      g.fsp = yt[1][c + 1]; g.rfsp = yt[2][c + 1];
        return g;
    };
#define dy (YTAB ? yt[5][c] : dy_reg)
#define rdy (YTAB ? yt[6][c] : rdy_reg)

    // ---- prologue: rows r0-2 .. r0+2 land, rho is formed for all of them, velocity for rows r0-1, r0, r0+1
    if (vec) for (int r = r0 - HALO; r <= r0 + HALO; r++) vec_row(r, slot_of(r));
    else for (int r = r0 - HALO; r <= r0 + HALO; r++) issue_row(r, slot_of(r));
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();                             // per-thread fills / copies of other threads
    for (int r = r0 - HALO; r <= r0 + HALO; r++) convert_rho(slot_of(r));
    __syncthreads();
    for (int r = r0 - 1; r <= r0 + 1; r++) form_vel(slot_of(r), vslot_of(r));
    __syncthreads();

    // X warps: carries of face r0 (between rows r0-1 and r0)
    double cIx_biy = 0.0, cIx_biz = 0.0, cIx_p = 0.0, cVfx = 0.0, cIx_vy = 0.0, cIx_vz = 0.0;
    if (isX) {
        const FaceGeom g = x_geom(r0);
        const int sm2 = slot_of(r0 - 2), sm1 = slot_of(r0 - 1), s0 = slot_of(r0), sp1 = slot_of(r0 + 1);
        const int vm1 = vslot_of(r0 - 1), v0 = vslot_of(r0);
        cVfx = face_interp(vel[vm1][0][c], vel[v0][0][c], g.hm1, g.h0, g.fs, g.rfs);
        cIx_vy = face_interp(vel[vm1][1][c], vel[v0][1][c], g.hm1, g.h0, g.fs, g.rfs);
        if (Z) cIx_vz = face_interp(vel[vm1][2][c], vel[v0][2][c], g.hm1, g.h0, g.fs, g.rfs);
        cIx_p = face_interp(RG(sm1, Q_E, c) * P.gm1, RG(s0, Q_E, c) * P.gm1, g.hm1, g.h0, g.fs, g.rfs);
        const FaceSel fs0 = select_face(g, cVfx);
#pragma unroll UNR
        for (int k = 0; k < L.n; k++) {
            const int q = (int)((L.q >> (4 * k)) & 15ULL);
            double d2;
            const double S = upwind_face_sel(RG(sm2, q, c), RG(sm1, q, c), RG(s0, q, c), RG(sp1, q, c), fs0, &d2);
            FXS(q, fcol) = S * cVfx;
            if (q == Q_BIY) cIx_biy = d2;
            if (q == Q_BIZ) cIx_biz = d2;
        }
    }
    __syncthreads();      // row r0-2 was read above; its ring slot is the prefetch target of the first iteration

    double dtmin_local = 1.7976931348623157e308;
    // Y warps: the cell whose dt is still to be evaluated (its rho / momenta arrive from the X warps one barrier later)
    bool dt_pending = false;
    double dt_e = 0.0, dt_bx = 0.0, dt_by = 0.0, dt_bz = 0.0, dt_dx = 1.0, dt_rdx = 1.0;

    int s0 = slot_of(r0), v0 = vslot_of(r0);       // ring slots rotate by one per row: no modulo in the loop
    for (int r = r0; r < r1; r++) {
        const int sm1 = s0 == 0 ? RD - 1 : s0 - 1, sp1 = next_slot(s0), sp2 = next_slot(sp1), sp3 = next_slot(sp2);
        const int v1 = next_vslot(v0), v2 = next_vslot(v1);
        const bool pre = (r + 3 <= r1 + HALO - 1);
        // own-cell values of this row (consumed in phase 2), then the ring row three rows ahead: two cp.async groups
        own_row(r);
        cp_async_commit();
        if (pre) { if (vec) vec_row(r + 3, sp3); else issue_row(r + 3, sp3); }
        cp_async_commit();

        const int g = P.row0 + r;
        const bool interior = col_out && g >= P.xl && g <= P.xu && j >= P.yl && j <= P.yu;
        const double dx = xt[5][r - r0 + 3], rdx = xt[6][r - r0 + 3];
        const size_t off = (size_t)r * P.pitch + (col_out ? j : 0);
        const double pc = RG(s0, Q_E, c) * P.gm1;                 // press = (gamma-1)*thermal_energy  idealmhd.cpp:253

        // ================================================================ phase 1: one direction per warp
        double d_a = 0.0, d_b = 0.0, d_c = 0.0, d_d = 0.0, d_e = 0.0, d_f = 0.0;       // central derivatives kept by this role
        // the transport parts this role finishes itself stay in registers: X: rho, mom_x, mom_y, mom_z ; Y: e, bi_x, bi_y, bi_z, be_x, be_y, be_z
        double t_0 = 0.0, t_1 = 0.0, t_2 = 0.0, t_3 = 0.0, t_4 = 0.0, t_5 = 0.0, t_6 = 0.0;
        if (isX) {
            // ---- x face r+1 (between rows r and r+1)
            const FaceGeom gx = x_geom(r + 1);
            const double vfx1 = face_interp(vel[v0][0][c], vel[v1][0][c], gx.hm1, gx.h0, gx.fs, gx.rfs);
            const double Ix1_vy = face_interp(vel[v0][1][c], vel[v1][1][c], gx.hm1, gx.h0, gx.fs, gx.rfs);
            const double Ix1_vz = Z ? face_interp(vel[v0][2][c], vel[v1][2][c], gx.hm1, gx.h0, gx.fs, gx.rfs) : 0.0;
            const double Ix1_p = face_interp(pc, RG(sp1, Q_E, c) * P.gm1, gx.hm1, gx.h0, gx.fs, gx.rfs);
            const FaceSel fsx = select_face(gx, vfx1);
            // the far cell of the extrapolation: row r-1 for flow in +x, row r+2 for flow in -x -- one load from a selected row
            const double *far_row = &RG(fsx.pos ? sm1 : sp2, 0, c);
            double Ix1_biy = 0.0, Ix1_biz = 0.0;
            auto keep_x = [&](int q, double t) {                  // folds away for compile-time lists
                if (q == Q_RHO) t_0 = t; else if (q == Q_MX) t_1 = t; else if (q == Q_MY) t_2 = t; else if (q == Q_MZ) t_3 = t; else TXS(q, col) = t;
            };
#pragma unroll UNR
            for (int k = 0; k < L.n; k += 2) {
                const int qa = (int)((L.q >> (4 * k)) & 15ULL), qb = (int)((L.q >> (4 * k + 4)) & 15ULL);
                double ad2, bd2;
                const double aS = upwind_face_far(far_row[QI(qa) * SW], RG(s0, qa, c), RG(sp1, qa, c), fsx, &ad2);
                const double bS = upwind_face_far(far_row[QI(qb) * SW], RG(s0, qb, c), RG(sp1, qb, c), fsx, &bd2);
                const double af1 = aS * vfx1, bf1 = bS * vfx1;
                const double at = ddiv(af1 - FXS(qa, fcol), dx, rdx), bt = ddiv(bf1 - FXS(qb, fcol), dx, rdx);   // derivs.cpp:155-156
                FXS(qa, fcol) = af1; FXS(qb, fcol) = bf1;
                keep_x(qa, at); keep_x(qb, bt);
                if (qa == Q_BIY) Ix1_biy = ad2; if (qb == Q_BIY) Ix1_biy = bd2;
                if (qa == Q_BIZ) Ix1_biz = ad2; if (qb == Q_BIZ) Ix1_biz = bd2;
            }
            d_a = ddiv(Ix1_biy - cIx_biy, dx, rdx);                 // d(bi_y)/dx
            if (Z) d_b = ddiv(Ix1_biz - cIx_biz, dx, rdx);          // d(bi_z)/dx
            d_c = ddiv(Ix1_p - cIx_p, dx, rdx);                     // d(p)/dx
            Dc_s[0][col] = ddiv(vfx1 - cVfx, dx, rdx);              // d(v_x)/dx  -> Y
            Dc_s[1][col] = ddiv(Ix1_vy - cIx_vy, dx, rdx);          // d(v_y)/dx  -> Y
            if (Z) Dc_s[2][col] = ddiv(Ix1_vz - cIx_vz, dx, rdx);   // d(v_z)/dx  -> Y
            cIx_biy = Ix1_biy; cIx_biz = Ix1_biz; cIx_p = Ix1_p; cVfx = vfx1; cIx_vy = Ix1_vy; cIx_vz = Ix1_vz;
        } else {
            // ---- dt of the previous row: its rho and momenta were published by the X warps before the last barrier
            if (dt_pending) {
                const double rho_ = Dt_s[0][col], mx_ = Dt_s[1][col], my_ = Dt_s[2][col];
                if (!dt_can_skip(P, inv_thr, rho_, mx_, my_, dt_e, dt_bx, dt_by, dt_bz, dt_rdx, rdy)) {
                    const double dtc = cell_dt(P, rho_, mx_, my_, dt_e, dt_bx, dt_by, dt_bz, dt_dx, dt_rdx, dy, rdy);
                    dtmin_local = smin(dtmin_local, dtc);
                }
            }
            __syncwarp();
            // ---- y face j (left face of this column); the right face comes from lane+1
            const FaceGeom gy = y_geom();
            const double vxc = vel[v0][0][c], vyc = vel[v0][1][c], vzc = vel[v0][2][c];
            const double vfyL = face_interp(vel[v0][1][c - 1], vyc, gy.hm1, gy.h0, gy.fs, gy.rfs);
            const double IyL_vx = face_interp(vel[v0][0][c - 1], vxc, gy.hm1, gy.h0, gy.fs, gy.rfs);
            const double IyL_vz = Z ? face_interp(vel[v0][2][c - 1], vzc, gy.hm1, gy.h0, gy.fs, gy.rfs) : 0.0;
            const double IyL_p = face_interp(RG(s0, Q_E, c - 1) * P.gm1, pc, gy.hm1, gy.h0, gy.fs, gy.rfs);
            const double vfyR = shfl_next(vfyL), IyR_vx = shfl_next(IyL_vx), IyR_vz = Z ? shfl_next(IyL_vz) : 0.0, IyR_p = shfl_next(IyL_p);
            const FaceSel fsy = select_face(gy, vfyL);
            const double *far_col = &RG(s0, 0, fsy.pos ? c - 2 : c + 1);
            double IyL_bix = 0.0, IyR_bix = 0.0, IyL_biz = 0.0, IyR_biz = 0.0;
            auto keep_y = [&](int q, double t) {
                if (q == Q_E) t_0 = t; else if (q == Q_BIX) t_1 = t; else if (q == Q_BIY) t_2 = t; else if (q == Q_BIZ) t_3 = t;
                else if (q == Q_BEX) t_4 = t; else if (q == Q_BEY) t_5 = t; else if (q == Q_BEZ) t_6 = t; else TYS(q, col) = t;
            };
#pragma unroll UNR
            for (int k = 0; k < L.n; k += 2) {
                const int qa = (int)((L.q >> (4 * k)) & 15ULL), qb = (int)((L.q >> (4 * k + 4)) & 15ULL);
                double ad2, bd2;
                const double aS = upwind_face_far(far_col[QI(qa) * SW], RG(s0, qa, c - 1), RG(s0, qa, c), fsy, &ad2);
                const double bS = upwind_face_far(far_col[QI(qb) * SW], RG(s0, qb, c - 1), RG(s0, qb, c), fsy, &bd2);
                const double afL = aS * vfyL, bfL = bS * vfyL;
                const double afR = shfl_next(afL), bfR = shfl_next(bfL), ad2R = shfl_next(ad2), bd2R = shfl_next(bd2);
                keep_y(qa, ddiv(afR - afL, dy, rdy));
                keep_y(qb, ddiv(bfR - bfL, dy, rdy));
                if (qa == Q_BIX) { IyL_bix = ad2; IyR_bix = ad2R; } if (qb == Q_BIX) { IyL_bix = bd2; IyR_bix = bd2R; }
                if (qa == Q_BIZ) { IyL_biz = ad2; IyR_biz = ad2R; } if (qb == Q_BIZ) { IyL_biz = bd2; IyR_biz = bd2R; }
            }
            Dc_s[3][col] = ddiv(IyR_bix - IyL_bix, dy, rdy);        // d(bi_x)/dy -> X
            if (Z) Dc_s[4][col] = ddiv(IyR_biz - IyL_biz, dy, rdy); // d(bi_z)/dy -> X
            Dc_s[5][col] = ddiv(IyR_p - IyL_p, dy, rdy);            // d(p)/dy    -> X
            d_d = ddiv(vfyR - vfyL, dy, rdy);                       // d(v_y)/dy
            d_e = ddiv(IyR_vx - IyL_vx, dy, rdy);                   // d(v_x)/dy
            if (Z) d_f = ddiv(IyR_vz - IyL_vz, dy, rdy);            // d(v_z)/dy
        }
        // the own-cell rows have landed (they had the whole of phase 1), then the partial results become visible to the other role
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");     // the own-cell values have landed (the younger group is ring row r+3)
        // The exchange slots are private to a warp column: X warp w only has to meet Y warp w + 2 here.  The full instance takes the pair-wise form
        // (named barrier 1 + wcol, 64 threads); the ring itself is published by the CTA-wide barrier that ends the iteration.  Measured at 4096^2,
        // exact arithmetic: full instance 3.61 -> 3.48 ms/step, 2-D instance 2.06 -> 2.19 ms/step (so it keeps the CTA-wide barrier).
        if (LN != 6) asm volatile("bar.sync %0, 64;" ::"r"(1 + wcol) : "memory");
        else __syncthreads();

        // ================================================================ phase 2: finish the outputs
        dt_pending = false;
        if (col_out) {
            const double bix = RG(s0, Q_BIX, c), biy = RG(s0, Q_BIY, c), biz = Z ? RG(s0, Q_BIZ, c) : 0.0;
            const double bex = Z ? RG(s0, Q_BEX, c) : 0.0, bey = Z ? RG(s0, Q_BEY, c) : 0.0, bez = Z ? RG(s0, Q_BEZ, c) : 0.0;
            if (isX) {
                const double rho = RG(s0, Q_RHO, c);
                const double g0 = grav ? own[0][ccol] : 0.0, g1 = grav ? own[1][ccol] : 0.0;     // gravity
                const double dbix_dy = Dc_s[3][col], dbiz_dy = Dc_s[4][col], dp_dy = Dc_s[5][col];
                const double T_rho = t_0 + TYS(Q_RHO, col), T_mx = t_1 + TYS(Q_MX, col);
                const double T_my = t_2 + TYS(Q_MY, col), T_mz = Z ? t_3 + TYS(Q_MZ, col) : 0.0;
                double k0 = T_rho * -1.0;                                                        // idealmhd.cpp:52
                const double cdb = ddiv(d_a - dbix_dy, P.fourpi, P.rfourpi);                    // :54
                const double ncdb = cdb * -1.0;
                const double czx = dbiz_dy, czy = d_b * -1.0;                                   // curlZ, derivs.cpp:465-469
                double k1, k2, k3 = 0.0;
                if (Z) {
                    const double bzi = ddiv(biz, P.fourpi, P.rfourpi), bze = ddiv(bez, P.fourpi, P.rfourpi);   // :57-58
                    k1 = ((((((T_mx * -1.0) - d_c) + rho * g0) + ncdb * bey) + ncdb * biy) + bzi * czy) + bze * czy;                       // :62-66
                    k2 = ((((((T_my * -1.0) - dp_dy) + rho * g1) + cdb * bex) + cdb * bix) + (bzi * -1.0) * czx) + (bze * -1.0) * czx;   // :67-71
                    const double fze = ddiv(czx * bey - czy * bex, P.fourpi, P.rfourpi);        // :59
                    const double fzi = ddiv(czx * biy - czy * bix, P.fourpi, P.rfourpi);        // :60
                    k3 = ((T_mz * -1.0) + fze) + fzi;                                           // :72-73
                } else {                               // be_* = bi_z = 0: the omitted products are exact zeros
                    k1 = (((T_mx * -1.0) - d_c) + rho * g0) + ncdb * biy;
                    k2 = (((T_my * -1.0) - dp_dy) + rho * g1) + cdb * bix;
                }
                if (!interior) { k0 = 0.0; k1 = 0.0; k2 = 0.0; k3 = 0.0; }                      // ghost mask :99-103
                for (int t = 0; t < n_xterm; t++) {                                           // module RHS terms, in module order
                    const int tg = A.xtarget[t];
                    if (tg <= E_MZ) { const double x = A.xterm[t][off]; if (tg == E_N) k0 = k0 + x; if (tg == E_MX) k1 = k1 + x; if (tg == E_MY) k2 = k2 + x; if (tg == E_MZ) k3 = k3 + x; }
                }
                if (kmode == KM_STORE_K1 || kmode == KM_EXPORT) { A.K1[E_N][off] = k0; A.K1[E_MX][off] = k1; A.K1[E_MY][off] = k2; if (Z || kmode == KM_EXPORT) A.K1[E_MZ][off] = k3; }
                else if (kmode == KM_STORE_K2) { A.K2[E_N][off] = k0; A.K2[E_MX][off] = k1; A.K2[E_MY][off] = k2; if (Z) A.K2[E_MZ][off] = k3; }
                else if (kmode == KM_ADD_K2) { A.K2[E_N][off] = A.K2[E_N][off] + k0; A.K2[E_MX][off] = A.K2[E_MX][off] + k1; A.K2[E_MY][off] = A.K2[E_MY][off] + k2; if (Z) A.K2[E_MZ][off] = A.K2[E_MZ][off] + k3; }
                else if (kmode == KM_FINAL) {                                                 // evolution.cpp:121
                    k0 = (A.K1[E_N][off] + k0) / 6.0 + A.K2[E_N][off] / 3.0;   k1 = (A.K1[E_MX][off] + k1) / 6.0 + A.K2[E_MX][off] / 3.0;
                    k2 = (A.K1[E_MY][off] + k2) / 6.0 + A.K2[E_MY][off] / 3.0; if (Z) k3 = (A.K1[E_MZ][off] + k3) / 6.0 + A.K2[E_MZ][off] / 3.0;
                }
                if (kmode != KM_EXPORT) {
                    double Un, Umx, Umy, Umz;                                                   // equationset.cpp:226-228
                    if (!HAVE_B || b_is_s) { Un = rho + k0 * s; Umx = RG(s0, Q_MX, c) + k1 * s; Umy = RG(s0, Q_MY, c) + k2 * s; Umz = Z ? RG(s0, Q_MZ, c) + k3 * s : 0.0; }
                    else { Un = (own[2][ccol] * P.m_i) + k0 * s; Umx = own[3][ccol] + k1 * s; Umy = own[4][ccol] + k2 * s; Umz = Z ? own[HAVE_B ? 5 : 0][ccol] + k3 * s : 0.0; }
                    double rfl;
                    const double nn = density_floor(P, Un, &rfl);
                    if (primary && A.walls) {
                        const unsigned z = zero_zones(P, g, j);
                        record_strips(P, A.strip, A.strip_pitch, g, r, j, z, rfl, Umx, Umy, Umz);
                        if (z) { Umx = 0.0; Umy = 0.0; Umz = 0.0; }
                    }
                    A.D[E_N][off] = nn; A.D[E_MX][off] = Umx; A.D[E_MY][off] = Umy; if (Z) A.D[E_MZ][off] = Umz;
                    if (primary) { Dt_s[0][col] = nn * P.m_i; Dt_s[1][col] = Umx; Dt_s[2][col] = Umy; }
                }
            } else {
                const double dvx_dx = Dc_s[0][col], dvy_dx = Dc_s[1][col], dvz_dx = Dc_s[2][col];
                const double T_e = TXS(Q_E, col) + t_0;
                const double T_bix = TXS(Q_BIX, col) + t_1, T_biy = TXS(Q_BIY, col) + t_2;
                double k4 = (T_e * -1.0) - pc * (dvx_dx + d_d);                                 // :75-76
                double k5, k6, k7 = 0.0;
                if (Z) {
                    const double T_biz = TXS(Q_BIZ, col) + t_3;
                    const double T_bex = TXS(Q_BEX, col) + t_4, T_bey = TXS(Q_BEY, col) + t_5, T_bez = TXS(Q_BEZ, col) + t_6;
                    const double bxs = bix + bex, bys = biy + bey;
                    k5 = (((T_bix * -1.0) - T_bex) + bxs * dvx_dx) + bys * d_e;                 // :78-80
                    k6 = (((T_biy * -1.0) - T_bey) + bxs * dvy_dx) + bys * d_d;                 // :81-83
                    k7 = (((T_biz * -1.0) - T_bez) + bxs * dvz_dx) + bys * d_f;                 // :84-86
                } else {                               // be_* = 0, v_z = 0: T(be_k) and the bi_z equation are exact zeros
                    k5 = ((T_bix * -1.0) + bix * dvx_dx) + biy * d_e;
                    k6 = ((T_biy * -1.0) + bix * dvy_dx) + biy * d_d;
                }
                if (!interior) { k4 = 0.0; k5 = 0.0; k6 = 0.0; k7 = 0.0; }
                for (int t = 0; t < n_xterm; t++) {
                    const int tg = A.xtarget[t];
                    if (tg > E_MZ) { const double x = A.xterm[t][off]; if (tg == E_E) k4 = k4 + x; if (tg == E_BX) k5 = k5 + x; if (tg == E_BY) k6 = k6 + x; if (tg == E_BZ) k7 = k7 + x; }
                }
                if (kmode == KM_STORE_K1 || kmode == KM_EXPORT) { A.K1[E_E][off] = k4; A.K1[E_BX][off] = k5; A.K1[E_BY][off] = k6; if (Z || kmode == KM_EXPORT) A.K1[E_BZ][off] = k7; }
                else if (kmode == KM_STORE_K2) { A.K2[E_E][off] = k4; A.K2[E_BX][off] = k5; A.K2[E_BY][off] = k6; if (Z) A.K2[E_BZ][off] = k7; }
                else if (kmode == KM_ADD_K2) { A.K2[E_E][off] = A.K2[E_E][off] + k4; A.K2[E_BX][off] = A.K2[E_BX][off] + k5; A.K2[E_BY][off] = A.K2[E_BY][off] + k6; if (Z) A.K2[E_BZ][off] = A.K2[E_BZ][off] + k7; }
                else if (kmode == KM_FINAL) {
                    k4 = (A.K1[E_E][off] + k4) / 6.0 + A.K2[E_E][off] / 3.0;   k5 = (A.K1[E_BX][off] + k5) / 6.0 + A.K2[E_BX][off] / 3.0;
                    k6 = (A.K1[E_BY][off] + k6) / 6.0 + A.K2[E_BY][off] / 3.0; if (Z) k7 = (A.K1[E_BZ][off] + k7) / 6.0 + A.K2[E_BZ][off] / 3.0;
                }
                if (kmode != KM_EXPORT) {
                    double Ue, Ubx, Uby, Ubz;
                    if (!HAVE_B || b_is_s) { Ue = RG(s0, Q_E, c) + k4 * s; Ubx = bix + k5 * s; Uby = biy + k6 * s; Ubz = Z ? biz + k7 * s : 0.0; }
                    else { Ue = own[HAVE_B ? 2 + NB : 0][ccol] + k4 * s; Ubx = own[HAVE_B ? 3 + NB : 0][ccol] + k5 * s; Uby = own[HAVE_B ? 4 + NB : 0][ccol] + k6 * s; Ubz = Z ? own[HAVE_B ? 5 + NB : 0][ccol] + k7 * s : 0.0; }
                    const double e1 = smax(Ue, P.e_min);
                    A.D[E_E][off] = e1; A.D[E_BX][off] = Ubx; A.D[E_BY][off] = Uby; if (Z) A.D[E_BZ][off] = Ubz;
                    if (primary && interior) {                                                 // dt of this cell: evaluated after the next barrier
                        dt_pending = true; dt_e = e1; dt_bx = bex + Ubx; dt_by = bey + Uby; dt_bz = bez + Ubz; dt_dx = dx; dt_rdx = rdx;
                    }
                }
            }
        }
        cp_async_wait_all();
        if (pre && !vec) __syncthreads();                           // per-column path: the row came through the 66 loader threads, warp 2 converts it below
        if (pre) convert_rho(sp3);                                  // the row that just landed
        if (r + 2 <= r1) form_vel(sp2, v2);                         // into the velocity slot of row r-1 (dead since the last barrier)
        __syncthreads();
        s0 = sp1; v0 = v1;
    }
    if (primary && kmode != KM_EXPORT) {
        if (!isX && dt_pending && !dt_can_skip(P, inv_thr, Dt_s[0][col], Dt_s[1][col], Dt_s[2][col], dt_e, dt_bx, dt_by, dt_bz, dt_rdx, rdy)) {
            const double dtc = cell_dt(P, Dt_s[0][col], Dt_s[1][col], Dt_s[2][col], dt_e, dt_bx, dt_by, dt_bz, dt_dx, dt_rdx, dy, rdy);
            dtmin_local = smin(dtmin_local, dtc);
        }
        block_min_impl(dtmin_local, A.dtmin_bits, reinterpret_cast<unsigned long long *>(Fx_p));   // the flux carry is dead after the last barrier
    }
}
#undef dy
#undef rdy
#undef kmode
#undef b_is_s
#undef primary
#undef n_xterm
#undef RG
#undef FXS
#undef TXS
#undef TYS

}  // namespace spruce
