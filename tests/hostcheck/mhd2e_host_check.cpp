// mhd2e_host_check.cpp -- TEST INFRASTRUCTURE.  Runs whole IdealMHD2E time steps on the HOST with the product's own per-cell functions
// (spruce_b200/csrc/mhd2e_cells.cuh) in the product's own stage order (spruce_b200/csrc/mhd2e_step.hpp: the template the device executor
// instantiates too); the only thing replaced is "launch a kernel over the cells" by "loop over the cells".  tests/test_mhd2e_host_check.py
// compares the result bit for bit with the CPU restatement that is pinned to the reference.  Nothing in the product links this.
#include "../../spruce_b200/csrc/mhd2e_cells.cuh"
#include "../../spruce_b200/csrc/mhd2e_step.hpp"
#include <cstring>
#include <vector>

using namespace spruce::e2;

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
    g.dx = dx; g.dy = dy; g.nx = nx; g.ny = ny; g.pitch = ny;
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
