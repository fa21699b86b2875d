// mhd2e_host_check.cpp -- TEST INFRASTRUCTURE.  Runs whole IdealMHD2E time steps on the HOST with the product's own per-cell functions
// (spruce_b200/csrc/mhd2e_cells.cuh) in the product's own stage order (spruce_b200/csrc/mhd2e_step.hpp: the template the device executor
// instantiates too); the only thing replaced is "launch a kernel over the cells" by "loop over the cells".  tests/test_mhd2e_host_check.py
// compares the result bit for bit with the CPU restatement that is pinned to the reference.  Nothing in the product links this.
#include "../../spruce_b200/csrc/mhd2e_cells.cuh"
#include "../../spruce_b200/csrc/mhd2e_step.hpp"
#include <cmath>
#include <cstring>
#include <vector>

using namespace spruce::e2;

static int g_eic = 0;                                      // eic_thermalization configured for the runs that follow
extern "C" void mhd2e_host_set_eic(int on) { g_eic = on; }

struct HostExec {
    Geo g;
    Statics T;
    std::vector<double> sets[3][NEV2], K1[NEV2], K23[NEV2];
    double step = 0.0, dtmin = 0.0;
    int order[3] = {0, 1, 2};                              // logical set -> storage
    Planes planes(int s) { Planes p; for (int v = 0; v < NEV2; v++) p.u[v] = sets[order[s]][v].data(); return p; }
    CPlanes cplanes(int s) { CPlanes p; for (int v = 0; v < NEV2; v++) p.u[v] = sets[order[s]][v].data(); return p; }
    void swap_sets(int a, int b) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    int stage(int S, int B, int D, double coef, int kmode, int ghost_primary, int final_stage)
    {
        const CPlanes s = cplanes(S), bb = cplanes(B);
        const Planes d = planes(D), p = planes(ghost_primary);
        const double sc = coef * step;
        for (int i = 0; i < g.nx; i++) for (int j = 0; j < g.ny; j++) {                   // "kernel" 1: right-hand side, K rule, apply, floors
            const size_t c = at(g, i, j);
            double k[NEV2], k1[NEV2], k23[NEV2], base[NEV2], out[NEV2];
            rhs_cell(g, s, T, i, j, k);
            for (int v = 0; v < NEV2; v++) { k1[v] = K1[v][c]; k23[v] = K23[v][c]; }
            k_rule(kmode, k, k1, k23, NEV2);
            for (int v = 0; v < NEV2; v++) { K1[v][c] = k1[v]; K23[v][c] = k23[v]; }
            if (kmode == KM2_EXPORT) continue;
            for (int v = 0; v < NEV2; v++) base[v] = bb.u[v][c];
            apply_cell(g, base, k, sc, out);
            for (int v = 0; v < NEV2; v++) d.u[v][c] = out[v];
        }
        if (kmode == KM2_EXPORT) return 0;
        for (int side = 0; side < 4; side++)                                                // "kernels" 2-5: the four sides, one after the other
            for (int a = 0; a < side_length(g, side); a++) ghost_cell(g, d, p, side, a);
        double m = 1.7976931348623157e308;
        for (int i = 0; i < g.nx; i++) for (int j = 0; j < g.ny; j++) {                   // "kernel" 6: settle, dt
            const size_t c = at(g, i, j);
            double u[NEV2];
            for (int v = 0; v < NEV2; v++) u[v] = d.u[v][c];
            settle_cell(g, u);
            for (int v = 0; v < NEV2; v++) d.u[v][c] = u[v];
            if (final_stage && interior(g, i, j)) m = smin2(m, dt_cell(g, u, T.bex[c], T.bey[c], g.dx[i], g.dy[j]));
        }
        if (final_stage) dtmin = m;
        return 0;
    }
};

// planes_in: rho, i_temp, e_temp, mom_x, mom_y, bi_x, bi_y, be_x, be_y, grav_x, grav_y.  Runs setup (recomputeEvolvedVarsFromStateVars + propagateChanges)
// and n_steps steps; out[7][nx*ny] = evolved planes, dt_out = dt plane, steps_out = step sizes; rhs_out (optional) = right-hand side of the final state.
extern "C" int mhd2e_host_run(const double *const *planes_in, const double *dx, const double *dy, int nx, int ny, const int *bc, int integrator, double m_i, double gamma,
                              double epsilon, double n_min, double T_min, double e_min, double open_strength, double open_decay, int n_steps,
                              double *out, double *dt_out, double *steps_out, double *rhs_out)
{
    HostExec x;
    Geo &g = x.g;
    g = Geo{};
    g.eic = g_eic;
    g.dx = dx; g.dy = dy; g.nx = nx; g.ny = ny; g.pitch = ny; g.row0 = 0; g.nxl = nx; g.x_halo = 0;
    for (int s = 0; s < 4; s++) g.bc[s] = bc[s];
    g.xl = bc[0] == BC2_PERIODIC ? 0 : NG; g.xu = bc[1] == BC2_PERIODIC ? nx - 1 : nx - NG - 1;
    g.yl = bc[2] == BC2_PERIODIC ? 0 : NG; g.yu = bc[3] == BC2_PERIODIC ? ny - 1 : ny - NG - 1;
    g.xper = bc[0] == BC2_PERIODIC && bc[1] == BC2_PERIODIC; g.yper = bc[2] == BC2_PERIODIC && bc[3] == BC2_PERIODIC;
    g.m_i = m_i; g.gamma = gamma; g.n_min = n_min; g.T_min = T_min; g.e_min = e_min; g.open_strength = open_strength;
    open_scales(g, open_decay);
    const size_t n = (size_t)nx * ny;
    x.T.bex = planes_in[7]; x.T.bey = planes_in[8]; x.T.gx = planes_in[9]; x.T.gy = planes_in[10];
    for (int s = 0; s < 3; s++) for (int v = 0; v < NEV2; v++) x.sets[s][v].assign(n, 0.0);
    for (int v = 0; v < NEV2; v++) { x.K1[v].assign(n, 0.0); x.K23[v].assign(n, 0.0); }
    // setup: state -> evolved (equationset.cpp:96-104), then propagateChanges = a stage with k = 0 is NOT the same thing (no increment): floors, ghosts, settle
    {
        Planes P = x.planes(0);
        for (size_t c = 0; c < n; c++) {
            P.u[Q_RHO2][c] = planes_in[0][c]; P.u[Q_MX2][c] = planes_in[3][c]; P.u[Q_MY2][c] = planes_in[4][c]; P.u[Q_BX2][c] = planes_in[5][c]; P.u[Q_BY2][c] = planes_in[6][c];
            from_state_cell(g, planes_in[0][c], planes_in[1][c], planes_in[2][c], &P.u[Q_EI2][c], &P.u[Q_EE2][c]);
            double u[NEV2], zero[NEV2] = {0, 0, 0, 0, 0, 0, 0}, o2[NEV2];
            for (int v = 0; v < NEV2; v++) u[v] = P.u[v][c];
            apply_cell(g, u, zero, 0.0, o2);                                               // enforceMinimums (base + 0*0 = base)
            for (int v = 0; v < NEV2; v++) P.u[v][c] = o2[v];
        }
        for (int side = 0; side < 4; side++) for (int a = 0; a < side_length(g, side); a++) ghost_cell(g, P, P, side, a);
        double m = 1.7976931348623157e308;
        for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
            const size_t c = at(g, i, j);
            double u[NEV2];
            for (int v = 0; v < NEV2; v++) u[v] = P.u[v][c];
            settle_cell(g, u);
            for (int v = 0; v < NEV2; v++) P.u[v][c] = u[v];
            if (interior(g, i, j)) m = smin2(m, dt_cell(g, u, x.T.bex[c], x.T.bey[c], dx[i], dy[j]));
        }
        x.dtmin = m;
    }
    for (int it = 0; it < n_steps; it++) {
        x.step = epsilon * x.dtmin;                                                         // evolution.cpp:62
        steps_out[it] = x.step;
        if (advance(x, integrator)) return 1;
    }
    const CPlanes P = x.cplanes(0);
    for (int v = 0; v < NEV2; v++) std::memcpy(out + v * n, P.u[v], n * sizeof(double));
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) dt_out[at(g, i, j)] = derive_cell(g, P, x.T, V2_dt, i, j);
    if (rhs_out) {
        x.stage(0, 0, 1, 0.0, KM2_EXPORT, 0, 0);
        for (int v = 0; v < NEV2; v++) std::memcpy(rhs_out + v * n, x.K1[v].data(), n * sizeof(double));
    }
    return 0;
}


// ---- the slab form: n_ranks executors, each holding its rows plus two halo rows per side (NaN where nothing is resident), addressed by global row
// through shifted pointers exactly as mhd2e_host.cuh does; every phase ("kernel") runs on all ranks, then the halo exchange the device performs
// after a stage copies the edge rows to the neighbours.  The integrator sequence is the same advance() template.
struct SlabExec {
    int n_ranks, nx, ny;
    bool xper;
    std::vector<Geo> g;                                                     // per rank
    std::vector<std::vector<double>> dxl;                                   // per rank: cell sizes with an apron of 3
    std::vector<std::vector<double>> st[4];                                 // per rank statics (with halo rows)
    std::vector<std::vector<double>> sets[3][NEV2], K1[NEV2], K23[NEV2];    // [..][rank]
    double step = 0.0, dtmin = 0.0;
    int order[3] = {0, 1, 2};
    bool any_wall = false;
    static constexpr int H = NG;
    const double *shifted(const std::vector<double> &v, int r) const { return v.data() + (size_t)H * ny - (long long)g[r].row0 * ny; }
    double *shifted(std::vector<double> &v, int r) { return v.data() + (size_t)H * ny - (long long)g[r].row0 * ny; }
    Planes planes(int s, int r) { Planes p; for (int v = 0; v < NEV2; v++) p.u[v] = shifted(sets[order[s]][v][r], r); return p; }
    CPlanes cplanes(int s, int r) { CPlanes p; for (int v = 0; v < NEV2; v++) p.u[v] = shifted(sets[order[s]][v][r], r); return p; }
    Statics statics(int r) { Statics T; T.bex = shifted(st[0][r], r); T.bey = shifted(st[1][r], r); T.gx = shifted(st[2][r], r); T.gy = shifted(st[3][r], r); return T; }
    void swap_sets(int a, int b) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    // halo rows of one logical set from the ring neighbours (what peer_exchange does)
    void exchange(int s)
    {
        for (int v = 0; v < NEV2; v++) for (int r = 0; r < n_ranks; r++) {
            const int lo = r > 0 ? r - 1 : (xper ? n_ranks - 1 : -1), hi = r < n_ranks - 1 ? r + 1 : (xper ? 0 : -1);
            std::vector<double> &mine = sets[order[s]][v][r];
            for (int k = 0; k < H; k++) {
                if (lo >= 0) { const std::vector<double> &o = sets[order[s]][v][lo]; std::memcpy(&mine[(size_t)k * ny], &o[(size_t)(g[lo].nxl + k) * ny], ny * sizeof(double)); }      // its last H rows
                if (hi >= 0) { const std::vector<double> &o = sets[order[s]][v][hi]; std::memcpy(&mine[(size_t)(H + g[r].nxl + k) * ny], &o[(size_t)(H + k) * ny], ny * sizeof(double)); }  // its first H rows
            }
        }
    }
    int stage(int S, int B, int D, double coef, int kmode, int ghost_primary, int final_stage)
    {
        const double sc = coef * step;
        for (int r = 0; r < n_ranks; r++) {
            const Geo &gg = g[r]; const CPlanes s = cplanes(S, r), bb = cplanes(B, r); const Planes d = planes(D, r); const Statics T = statics(r);
            double *k1p[NEV2], *k23p[NEV2];
            for (int v = 0; v < NEV2; v++) { k1p[v] = shifted(K1[v][r], r); k23p[v] = shifted(K23[v][r], r); }
            for (int i = gg.row0; i < gg.row0 + gg.nxl; i++) for (int j = 0; j < ny; j++) {
                const size_t c = at(gg, i, j);
                double k[NEV2], k1[NEV2], k23[NEV2], base[NEV2], out[NEV2];
                rhs_cell(gg, s, T, i, j, k);
                for (int v = 0; v < NEV2; v++) { k1[v] = k1p[v][c]; k23[v] = k23p[v][c]; }
                k_rule(kmode, k, k1, k23, NEV2);
                for (int v = 0; v < NEV2; v++) { k1p[v][c] = k1[v]; k23p[v][c] = k23[v]; }
                if (kmode == KM2_EXPORT) continue;
                for (int v = 0; v < NEV2; v++) base[v] = bb.u[v][c];
                apply_cell(gg, base, k, sc, out);
                for (int v = 0; v < NEV2; v++) d.u[v][c] = out[v];
            }
        }
        if (kmode == KM2_EXPORT) return 0;
        for (int side = 0; side < 4; side++) for (int r = 0; r < n_ranks; r++) {
            const Planes d = planes(D, r), p = planes(ghost_primary, r);
            for (int t = 0; t < side_length(g[r], side); t++) ghost_cell(g[r], d, p, side, t);
        }
        double m = 1.7976931348623157e308;
        for (int r = 0; r < n_ranks; r++) {
            const Geo &gg = g[r]; const Planes d = planes(D, r); const Statics T = statics(r);
            for (int i = gg.row0; i < gg.row0 + gg.nxl; i++) for (int j = 0; j < ny; j++) {
                const size_t c = at(gg, i, j);
                double u[NEV2];
                for (int v = 0; v < NEV2; v++) u[v] = d.u[v][c];
                settle_cell(gg, u);
                for (int v = 0; v < NEV2; v++) d.u[v][c] = u[v];
                if (final_stage && interior(gg, i, j)) m = smin2(m, dt_cell(gg, u, T.bex[c], T.bey[c], gg.dx[i], gg.dy[j]));
            }
        }
        if (final_stage) dtmin = m;                                         // all-gathered minimum
        exchange(D);
        if (ghost_primary != D && any_wall) exchange(ghost_primary);        // the wall-type passes wrote the primary state
        return 0;
    }
};

extern "C" int mhd2e_host_run_slabs(const double *const *planes_in, const double *dx, const double *dy, int nx, int ny, const int *bc, int integrator, double m_i, double gamma,
                                    double epsilon, double n_min, double T_min, double e_min, double open_strength, double open_decay, int n_steps, int n_ranks,
                                    double *out, double *steps_out)
{
    SlabExec x;
    x.n_ranks = n_ranks; x.nx = nx; x.ny = ny; x.xper = bc[0] == BC2_PERIODIC && bc[1] == BC2_PERIODIC;
    for (int s = 0; s < 4; s++) if (bc[s] == BC2_OPEN || bc[s] == BC2_REFLECT || bc[s] == BC2_FIXED) x.any_wall = true;
    x.g.resize(n_ranks); x.dxl.resize(n_ranks);
    const int H = SlabExec::H, APR = 3;
    Geo g0{};
    g0.eic = g_eic;
    g0.dx = dx; g0.dy = dy; g0.nx = nx; g0.ny = ny; g0.pitch = ny; g0.row0 = 0; g0.nxl = nx; g0.x_halo = 0;
    for (int s = 0; s < 4; s++) g0.bc[s] = bc[s];
    g0.xl = bc[0] == BC2_PERIODIC ? 0 : NG; g0.xu = bc[1] == BC2_PERIODIC ? nx - 1 : nx - NG - 1;
    g0.yl = bc[2] == BC2_PERIODIC ? 0 : NG; g0.yu = bc[3] == BC2_PERIODIC ? ny - 1 : ny - NG - 1;
    g0.xper = x.xper; g0.yper = bc[2] == BC2_PERIODIC && bc[3] == BC2_PERIODIC;
    g0.m_i = m_i; g0.gamma = gamma; g0.n_min = n_min; g0.T_min = T_min; g0.e_min = e_min; g0.open_strength = open_strength;
    open_scales(g0, open_decay);
    const double nan = std::nan("");
    for (int k = 0; k < 4; k++) x.st[k].resize(n_ranks);
    for (int s = 0; s < 3; s++) for (int v = 0; v < NEV2; v++) x.sets[s][v].resize(n_ranks);
    for (int v = 0; v < NEV2; v++) { x.K1[v].resize(n_ranks); x.K23[v].resize(n_ranks); }
    std::vector<std::vector<double>> it(n_ranks), et(n_ranks);
    for (int r = 0; r < n_ranks; r++) {
        Geo &g = x.g[r];
        g = g0;
        g.row0 = (int)((long long)nx * r / n_ranks); g.nxl = (int)((long long)nx * (r + 1) / n_ranks) - g.row0;
        g.x_halo = (x.xper && n_ranks > 1) ? 1 : 0;
        x.dxl[r].assign(g.nxl + 2 * APR, 1.0);
        for (int k = 0; k < g.nxl + 2 * APR; k++) { int gi = g.row0 + k - APR; if (gi < 0 || gi >= nx) { if (!x.xper) continue; gi = (gi + nx) % nx; } x.dxl[r][k] = dx[gi]; }
        g.dx = x.dxl[r].data() + APR - g.row0;
        const size_t nl = (size_t)(g.nxl + 2 * H) * ny;
        auto fill = [&](std::vector<double> &dst, const double *src) {               // own rows + halo rows from the global plane (ring wrap / NaN)
            dst.assign(nl, nan);
            for (int rr = -H; rr < g.nxl + H; rr++) { int gi = g.row0 + rr; if (gi < 0 || gi >= nx) { if (!x.xper) continue; gi = (gi + nx) % nx; }
                std::memcpy(&dst[(size_t)(rr + H) * ny], &src[(size_t)gi * ny], ny * sizeof(double)); }
        };
        for (int k = 0; k < 4; k++) fill(x.st[k][r], planes_in[7 + k]);
        for (int s = 0; s < 3; s++) for (int v = 0; v < NEV2; v++) x.sets[s][v][r].assign(nl, nan);
        for (int v = 0; v < NEV2; v++) { x.K1[v][r].assign(nl, 0.0); x.K23[v][r].assign(nl, 0.0); }
        fill(x.sets[0][Q_RHO2][r], planes_in[0]); fill(x.sets[0][Q_MX2][r], planes_in[3]); fill(x.sets[0][Q_MY2][r], planes_in[4]);
        fill(x.sets[0][Q_BX2][r], planes_in[5]); fill(x.sets[0][Q_BY2][r], planes_in[6]);
        fill(it[r], planes_in[1]); fill(et[r], planes_in[2]);
    }
    // setup on every rank: state -> evolved, floors, boundary passes, settle + dt, then the initial exchange
    for (int r = 0; r < n_ranks; r++) {
        const Geo &g = x.g[r]; Planes P = x.planes(0, r);
        const double *itp = x.shifted(it[r], r), *etp = x.shifted(et[r], r);
        for (int i = g.row0; i < g.row0 + g.nxl; i++) for (int j = 0; j < ny; j++) {
            const size_t c = at(g, i, j);
            double u[NEV2], zero[NEV2] = {0, 0, 0, 0, 0, 0, 0}, o2[NEV2];
            for (int v = 0; v < NEV2; v++) u[v] = P.u[v][c];
            from_state_cell(g, u[Q_RHO2], itp[c], etp[c], &u[Q_EI2], &u[Q_EE2]);
            apply_cell(g, u, zero, 0.0, o2);
            for (int v = 0; v < NEV2; v++) P.u[v][c] = o2[v];
        }
    }
    for (int side = 0; side < 4; side++) for (int r = 0; r < n_ranks; r++) { Planes P = x.planes(0, r); for (int t = 0; t < side_length(x.g[r], side); t++) ghost_cell(x.g[r], P, P, side, t); }
    double m = 1.7976931348623157e308;
    for (int r = 0; r < n_ranks; r++) {
        const Geo &g = x.g[r]; Planes P = x.planes(0, r); const Statics T = x.statics(r);
        for (int i = g.row0; i < g.row0 + g.nxl; i++) for (int j = 0; j < ny; j++) {
            const size_t c = at(g, i, j);
            double u[NEV2];
            for (int v = 0; v < NEV2; v++) u[v] = P.u[v][c];
            settle_cell(g, u);
            for (int v = 0; v < NEV2; v++) P.u[v][c] = u[v];
            if (interior(g, i, j)) m = smin2(m, dt_cell(g, u, T.bex[c], T.bey[c], g.dx[i], g.dy[j]));
        }
    }
    x.dtmin = m;
    x.exchange(0);
    for (int s = 0; s < n_steps; s++) {
        x.step = epsilon * x.dtmin;
        steps_out[s] = x.step;
        if (advance(x, integrator)) return 1;
    }
    const size_t n = (size_t)nx * ny;
    for (int r = 0; r < n_ranks; r++) for (int v = 0; v < NEV2; v++)
        std::memcpy(out + v * n + (size_t)x.g[r].row0 * ny, &x.sets[x.order[0]][v][r][(size_t)H * ny], (size_t)x.g[r].nxl * ny * sizeof(double));
    return 0;
}
