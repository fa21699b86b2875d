// mhd2e_host.cuh -- kernels and launch sequences of the IdealMHD2E equation set (included by capi.cu inside its anonymous namespace, after
// the plane helpers).  The arithmetic is mhd2e_cells.cuh, the stage order mhd2e_step.hpp -- both proven on the host against the CPU
// restatement (tests/test_mhd2e_host_check.py); this file only maps "every cell" / "every boundary index" onto threads.
// First version: one thread per cell, operands straight from global memory (the pattern the two-fluid kernel started from).  Slabs: cells are
// addressed by global row through row-shifted plane pointers, the halo rows of a stage's result (and of the primary state when a wall-type
// boundary pass wrote it) travel over the peer transport after the stage.
// STATUS: written after the round-1 GPU budget was spent; compiled, not yet run on a GPU (tests/test_zz_gpu_unvalidated.py).
#pragma once
#include "mhd2e_cells.cuh"
#include "mhd2e_step.hpp"

struct E2Args {
    e2::Geo g;
    e2::CPlanes S, B;
    e2::Planes D, Pg;             // destination; the set the open / reflect / fixed passes write
    e2::Statics T;
    double *K1[e2::NEV2], *K23[e2::NEV2];
    const double *i_temp, *e_temp;
    double *out;
    int kmode, final_stage, side, var, from_state;
    double coef;
    const double *step_ptr;
    const int *done_ptr;
    unsigned long long *dtmin_bits;
    // artificial_viscosity (mhd2e_cells.cuh: visc_cell): right-hand-side terms of this stage as planes (k_2e_visc_term wrote them from set S), and one term's description
    const double *xterm[4]; int xtarget[4]; int n_xterm;
    e2::ViscTerm2 vt; const double *dt_plane; int visc_gc;
    // hyper-viscous sub-step (k_2e_visc_apply): out = base + (mask * (frac * step / n_sub)) * term, or the rk4 combination of four terms
    const double *base, *t1, *t2, *t3, *t4; double frac; int n_sub;
};

// right-hand side, K rule, increment, floors: one thread per cell
__global__ void __launch_bounds__(128) k_2e_cells(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;      // GLOBAL row
    if (j >= A.g.ny || *A.done_ptr) return;
    const size_t c = e2::at(A.g, i, j);
    double k[e2::NEV2], k1[e2::NEV2], k23[e2::NEV2];
    e2::rhs_cell(A.g, A.S, A.T, i, j, k);
    if (A.n_xterm) {                                                                        // Viscosity::computeTimeDerivativesModule: grids_dt[evol] += dqdt * mask (viscosity.cpp:116-118)
        const double mask = e2::interior(A.g, i, j) ? 1.0 : 0.0;
        for (int x = 0; x < A.n_xterm; x++) k[A.xtarget[x]] = k[A.xtarget[x]] + A.xterm[x][c] * mask;
    }
    const bool need_k = A.kmode != e2::KM2_NONE;
    if (need_k) {
        for (int v = 0; v < e2::NEV2; v++) { k1[v] = A.K1[v][c]; k23[v] = (A.kmode == e2::KM2_STORE_K1 || A.kmode == e2::KM2_EXPORT) ? 0.0 : A.K23[v][c]; }
        e2::k_rule(A.kmode, k, k1, k23, e2::NEV2);
        if (A.kmode == e2::KM2_STORE_K1 || A.kmode == e2::KM2_EXPORT) { for (int v = 0; v < e2::NEV2; v++) A.K1[v][c] = k1[v]; }
        else if (A.kmode != e2::KM2_FINAL) { for (int v = 0; v < e2::NEV2; v++) A.K23[v][c] = k23[v]; }
    }
    if (A.kmode == e2::KM2_EXPORT) return;
    double base[e2::NEV2], out[e2::NEV2];
    for (int v = 0; v < e2::NEV2; v++) base[v] = A.B.u[v][c];
    e2::apply_cell(A.g, base, k, A.coef * *A.step_ptr, out);
    for (int v = 0; v < e2::NEV2; v++) A.D.u[v][c] = out[v];
}
// one boundary pass (launched four times, side = 0..3 in the reference's order)
__global__ void __launch_bounds__(128) k_2e_ghost(const E2Args A)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (*A.done_ptr || a >= e2::side_length(A.g, A.side)) return;
    e2::ghost_cell(A.g, A.D, A.Pg, A.side, a);
}
// recomputeDerivedVarsFromEvolvedVars feedback (rho round trip, energy floors) and, in the step's last stage, the dt minimum;
// from_state: recomputeEvolvedVarsFromStateVars + enforceMinimums first (setup)
__global__ void __launch_bounds__(128) k_2e_settle(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;
    double dtc = 1.7976931348623157e308;
    if (j < A.g.ny && !*A.done_ptr) {
        const size_t c = e2::at(A.g, i, j);
        double u[e2::NEV2];
        for (int v = 0; v < e2::NEV2; v++) u[v] = A.D.u[v][c];
        e2::settle_cell(A.g, u);
        for (int v = 0; v < e2::NEV2; v++) A.D.u[v][c] = u[v];
        if (A.final_stage && e2::interior(A.g, i, j)) dtc = e2::dt_cell(A.g, u, A.T.bex[c], A.T.bey[c], A.g.dx[i], A.g.dy[j]);
    }
    if (A.final_stage) block_min_to_global(dtc, A.dtmin_bits);
}
__global__ void __launch_bounds__(128) k_2e_from_state(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;
    if (j >= A.g.ny) return;
    const size_t c = e2::at(A.g, i, j);
    double u[e2::NEV2], zero[e2::NEV2] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, out[e2::NEV2];
    for (int v = 0; v < e2::NEV2; v++) u[v] = A.D.u[v][c];
    if (A.from_state) e2::from_state_cell(A.g, u[e2::Q_RHO2], A.i_temp[c], A.e_temp[c], &u[e2::Q_EI2], &u[e2::Q_EE2]);
    e2::apply_cell(A.g, u, zero, 0.0, out);                                  // enforceMinimums
    for (int v = 0; v < e2::NEV2; v++) A.D.u[v][c] = out[v];
}
__global__ void __launch_bounds__(128) k_2e_derive(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;
    if (j >= A.g.ny) return;
    A.out[e2::at(A.g, i, j)] = e2::derive_cell(A.g, A.S, A.T, A.var, i, j);
}

// one viscosity term on grid set S -> out (constructSingleViscosityGrid, viscosity.cpp:185-267); the timescale comes from the PRIMARY state's dt plane / minimum
__global__ void __launch_bounds__(128) k_2e_visc_term(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;
    if (j >= A.g.ny || *A.done_ptr) return;
    e2::ViscEnv2 env;
    env.dt_plane = A.dt_plane; env.gc = A.visc_gc;
    env.dt_min = __longlong_as_double((long long)*A.dtmin_bits);
    A.out[e2::at(A.g, i, j)] = e2::visc_cell(A.g, A.S, A.T, A.vt, env, i, j);
}
// grid_to_evol = base + (mask * c) * term, c = frac * (step / n_sub) (viscosity.cpp:135, 145, 150, ...); the rk4 combination when t2..t4 are given (:172-174)
__global__ void __launch_bounds__(128) k_2e_visc_apply(const E2Args A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = A.g.row0 + (int)blockIdx.y;
    if (j >= A.g.ny || *A.done_ptr) return;
    const size_t c = e2::at(A.g, i, j);
    const double mask = e2::interior(A.g, i, j) ? 1.0 : 0.0;
    const double dts = *A.step_ptr / (double)A.n_sub;
    const double cc = A.frac == 1.0 ? dts : A.frac * dts;
    double t = A.t1[c];
    if (A.t4) t = (((A.t1[c] + A.t2[c] * 2.0) + A.t3[c] * 2.0) + A.t4[c]) / 6.0;
    A.out[c] = A.base[c] + (mask * cc) * t;
}

struct OneFluid2E {
    double *set[3][e2::NEV2] = {{nullptr}};          // storage of the three state sets
    double *K1[e2::NEV2] = {nullptr}, *K23[e2::NEV2] = {nullptr};
    double *i_temp = nullptr, *e_temp = nullptr;     // uploaded state temperatures, consumed by setup
    int order[3] = {0, 1, 2};                        // logical set (0 = primary) -> storage
    bool rk4_alloc = false;
    e2::Geo g{};
    // artificial_viscosity (source/modules/viscosity.cpp), terms in config order
    struct VTerm { int opt; int var_diff; int evol; int scale_mode; double strength; double *strength_plane; };
    std::vector<VTerm> visc;
    int visc_hv_integrator = 0, visc_gc = 0; double visc_hv_epsilon = 1.0;
    double *dt_plane = nullptr, *vplane[6] = {nullptr};     // primary dt plane; vplane[0..3]: right-hand-side terms / the four rk4 evaluations, [4]: the sub-step's initial plane, [5]: spare
};

const char *const kE2EvolvedNames[e2::NEV2] = {"rho", "mom_x", "mom_y", "i_thermal_energy", "e_thermal_energy", "bi_x", "bi_y"};
const char *const kE2VarNames[e2::V2_COUNT] = {"rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y", "n", "i_press", "e_press", "press",
                                               "i_thermal_energy", "e_thermal_energy", "v_x", "v_y", "kinetic_energy", "b_x", "b_y", "b_mag", "b_hat_x", "b_hat_y", "dt"};
int e2_var_index(const char *name) { for (int v = 0; v < e2::V2_COUNT; v++) if (!strcmp(kE2VarNames[v], name)) return v; return -1; }
int e2_evolved_slot(const char *name) { for (int v = 0; v < e2::NEV2; v++) if (!strcmp(kE2EvolvedNames[v], name)) return v; return -1; }

int e2_check(const spruce_config &c)
{
    const int b[4] = {c.x_bound_1, c.x_bound_2, c.y_bound_1, c.y_bound_2};
    for (int s = 0; s < 4; s++) if (b[s] == SPRUCE_BC_OPEN_MOC) return fail(SPRUCE_ERR_UNSUPPORTED, "open_moc boundaries exist for ideal_mhd only (idealmhd.cpp:306)");
    return SPRUCE_OK;
}
int e2_create(spruce_domain *d)
{
    OneFluid2E *t = new OneFluid2E();
    d->e2 = t;
    int rc;
    for (int s = 0; s < 2; s++) for (int v = 0; v < e2::NEV2; v++) if ((rc = alloc_plane(d, &t->set[s][v]))) return rc;
    if ((rc = alloc_plane(d, &t->i_temp)) || (rc = alloc_plane(d, &t->e_temp))) return rc;
    return SPRUCE_OK;
}
int e2_ensure_rk4(spruce_domain *d)
{
    OneFluid2E *t = d->e2;
    if (t->rk4_alloc) return SPRUCE_OK;
    int rc;
    for (int v = 0; v < e2::NEV2; v++) if ((rc = alloc_plane(d, &t->set[2][v])) || (rc = alloc_plane(d, &t->K1[v])) || (rc = alloc_plane(d, &t->K23[v]))) return rc;
    t->rk4_alloc = true;
    return SPRUCE_OK;
}
// geometry and parameters of the per-cell functions; the open-boundary decay factors come from the host libm as in the reference
void e2_geometry(spruce_domain *d)
{
    e2::Geo &g = d->e2->g;                                // (g.eic stays as spruce_module_eic_thermalization left it)
    const spruce_config &c = d->cfg;
    g.nx = d->P.gnx; g.ny = d->P.ny; g.pitch = d->P.pitch;
    g.row0 = d->P.row0; g.nxl = d->P.nx; g.x_halo = (d->P.xper && !d->P.xwrap) ? 1 : 0;
    g.bc[0] = c.x_bound_1; g.bc[1] = c.x_bound_2; g.bc[2] = c.y_bound_1; g.bc[3] = c.y_bound_2;
    g.xl = d->P.xl; g.xu = d->P.xu; g.yl = d->P.yl; g.yu = d->P.yu; g.xper = d->P.xper; g.yper = d->P.yper;
    g.m_i = c.ion_mass; g.gamma = c.adiabatic_index; g.n_min = c.density_min; g.T_min = c.temp_min; g.e_min = c.thermal_energy_min;
    g.open_strength = c.open_boundary_strength;
    g.dx = d->dxg.data(); g.dy = d->dyg.data();          // host copies for open_scales() ...
    e2::open_scales(g, c.open_boundary_decay_base);
    g.dx = d->P.tx.d - d->P.row0; g.dy = d->P.ty.d;      // ... device tables for the kernels (x table indexed by global row)
}
// every plane pointer handed to the kernels is shifted back by the slab's first row: cells are addressed by GLOBAL row
template <class T> T *e2_shift(const spruce_domain *d, T *p) { return p ? p - (long long)d->P.row0 * d->P.pitch : p; }
void e2_base(spruce_domain *d, E2Args &A)
{
    OneFluid2E *t = d->e2;
    A.g = t->g;
    A.T.bex = e2_shift(d, d->stat[S_BEX]); A.T.bey = e2_shift(d, d->stat[S_BEY]); A.T.gx = e2_shift(d, d->stat[S_GX]); A.T.gy = e2_shift(d, d->stat[S_GY]);
    for (int v = 0; v < e2::NEV2; v++) { A.K1[v] = e2_shift(d, t->K1[v]); A.K23[v] = e2_shift(d, t->K23[v]); }
    A.step_ptr = &d->ctl->step; A.done_ptr = &d->ctl->done; A.dtmin_bits = &d->ctl->dtmin_bits;
    A.i_temp = e2_shift(d, t->i_temp); A.e_temp = e2_shift(d, t->e_temp);
}
void e2_set(const spruce_domain *d, int logical, e2::Planes &p) { const OneFluid2E *t = d->e2; for (int v = 0; v < e2::NEV2; v++) p.u[v] = e2_shift(d, t->set[t->order[logical]][v]); }
void e2_cset(const spruce_domain *d, int logical, e2::CPlanes &p) { const OneFluid2E *t = d->e2; for (int v = 0; v < e2::NEV2; v++) p.u[v] = e2_shift(d, t->set[t->order[logical]][v]); }
// slab decomposition: halo rows of the 7 planes of one logical set from the ring neighbours (one packed exchange of 8 plane slots)
int e2_exchange(spruce_domain *d, int logical)
{
    if (d->cfg.n_ranks == 1) return SPRUCE_OK;
    OneFluid2E *t = d->e2;
    double *v[NEV];
    for (int k = 0; k < NEV; k++) v[k] = t->set[t->order[logical]][k < e2::NEV2 ? k : 0];
    return peer_exchange(d, v, nullptr);
}
bool e2_any_wall(const spruce_domain *d)
{
    const int b[4] = {d->cfg.x_bound_1, d->cfg.x_bound_2, d->cfg.y_bound_1, d->cfg.y_bound_2};
    for (int s = 0; s < 4; s++) if (b[s] == SPRUCE_BC_OPEN || b[s] == SPRUCE_BC_REFLECT || b[s] == SPRUCE_BC_FIXED) return true;
    return false;
}

// the boundary passes (four launches, in order) and the settle / dt pass of set D
int e2_finish(spruce_domain *d, E2Args &A, int final_stage)
{
    const int n = d->P.ny > d->P.nx ? d->P.ny : d->P.nx;
    for (int side = 0; side < 4; side++) {
        const int bc = A.g.bc[side];
        if (bc == SPRUCE_BC_PERIODIC || bc == SPRUCE_BC_OPEN_MOC) continue;
        A.side = side;
        k_2e_ghost<<<(n + 127) / 128, 128, 0, d->stream>>>(A);
        d->launches++;
    }
    A.final_stage = final_stage;
    if (final_stage) { k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, 0); d->launches++; }
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2e_settle<<<grid, 128, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

// ---- artificial_viscosity on this set: the primary dt plane, one term as a plane, the right-hand-side terms of a stage
int e2_visc_refresh_dt(spruce_domain *d)                 // Viscosity reads the PRIMARY state's dt plane: materialised at the step's start and after every hyper-viscous sub-step
{
    OneFluid2E *t = d->e2;
    E2Args A{};
    e2_base(d, A);
    e2_cset(d, 0, A.S);
    A.out = e2_shift(d, t->dt_plane); A.var = e2::V2_dt;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2e_derive<<<grid, 128, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int e2_visc_term(spruce_domain *d, int S, int i, double *out)
{
    OneFluid2E *t = d->e2;
    const OneFluid2E::VTerm &v = t->visc[i];
    E2Args A{};
    e2_base(d, A);
    e2_cset(d, S, A.S);
    A.vt.opt = v.opt; A.vt.var_diff = v.var_diff; A.vt.scale_mode = v.scale_mode; A.vt.strength = v.strength;
    A.vt.strength_plane = (v.opt == 2 || v.opt == 3) ? e2_shift(d, v.strength_plane) : nullptr;
    A.dt_plane = e2_shift(d, t->dt_plane); A.visc_gc = t->visc_gc; A.out = e2_shift(d, out);
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2e_visc_term<<<grid, 128, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// Viscosity::computeTimeDerivativesModule (viscosity.cpp:112-123): every right-hand-side-form term (strength <= 1) evaluated on set S, handed to k_2e_cells as planes
int e2_visc_rhs_terms(spruce_domain *d, int S, E2Args &A)
{
    OneFluid2E *t = d->e2;
    A.n_xterm = 0;
    for (size_t i = 0; i < t->visc.size(); i++) {
        if (t->visc[i].strength > 1.0) continue;
        if (A.n_xterm >= 4) return fail(SPRUCE_ERR_UNSUPPORTED, "at most 4 right-hand-side viscosity terms");
        int rc = e2_visc_term(d, S, (int)i, t->vplane[A.n_xterm]);
        if (rc) return rc;
        A.xterm[A.n_xterm] = e2_shift(d, t->vplane[A.n_xterm]); A.xtarget[A.n_xterm] = t->visc[i].evol;
        A.n_xterm++;
    }
    return SPRUCE_OK;
}

// the executor mhd2e_step.hpp's advance() drives (the host check drives the same template with loops)
struct E2DeviceExec {
    spruce_domain *d;
    int stage(int S, int B, int D, double coef, int kmode, int ghost_primary, int final_stage)
    {
        OneFluid2E *t = d->e2;
        E2Args A{};
        e2_base(d, A);
        e2_cset(d, S, A.S); e2_cset(d, B, A.B); e2_set(d, D, A.D); e2_set(d, ghost_primary, A.Pg);
        A.coef = coef; A.kmode = kmode;
        if (!t->visc.empty()) { int rv = e2_visc_rhs_terms(d, S, A); if (rv) return rv; }
        dim3 grid((d->P.ny + 127) / 128, d->P.nx);
        k_2e_cells<<<grid, 128, 0, d->stream>>>(A);
        d->launches++;
        CUDA_TRY(cudaGetLastError());
        if (kmode == e2::KM2_EXPORT) return SPRUCE_OK;
        int rc = e2_finish(d, A, final_stage);
        if (rc || d->cfg.n_ranks == 1) return rc;
        if ((rc = e2_exchange(d, D))) return rc;
        if (ghost_primary != D && e2_any_wall(d) && (rc = e2_exchange(d, ghost_primary))) return rc;     // the wall-type passes wrote the primary state
        return final_stage ? peer_dt_allgather(d) : SPRUCE_OK;
    }
    void swap_sets(int a, int b) { std::swap(d->e2->order[a], d->e2->order[b]); }
};

// setupEquationSet / propagateChanges on the primary state (equationset.cpp:96-104, 212-220)
int e2_launch_propagate(spruce_domain *d, int from_state)
{
    OneFluid2E *t = d->e2;
    if (from_state) e2_geometry(d);
    E2Args A{};
    e2_base(d, A);
    e2_set(d, 0, A.D); e2_set(d, 0, A.Pg); e2_cset(d, 0, A.S);
    A.from_state = from_state;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2e_from_state<<<grid, 128, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return e2_finish(d, A, 1);
}
// Viscosity::iterateModule (viscosity.cpp:125-180): the hyper-viscous terms (strength > 1) sub-cycled on the primary state, a propagate after every sub-step
int e2_av_iterate(spruce_domain *d)
{
    OneFluid2E *t = d->e2;
    int rc;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    for (size_t i = 0; i < t->visc.size(); i++) {
        const OneFluid2E::VTerm &v = t->visc[i];
        if (v.strength <= 1.0) continue;
        const int ns = (int)(std::ceil(v.strength / t->visc_hv_epsilon) + 0.1);                // :130
        double *T[4] = {t->vplane[0], t->vplane[1], t->vplane[2], t->vplane[3]}, *init = t->vplane[4];
        auto evol = [&]() { return t->set[t->order[0]][v.evol]; };                              // (the primary set's storage does not move inside a module hook)
        auto apply = [&](const double *base, double frac, int n_terms) -> int {
            E2Args A{};
            e2_base(d, A);
            A.base = e2_shift(d, base); A.t1 = e2_shift(d, T[0]); A.t2 = n_terms == 4 ? e2_shift(d, T[1]) : nullptr; A.t3 = n_terms == 4 ? e2_shift(d, T[2]) : nullptr;
            A.t4 = n_terms == 4 ? e2_shift(d, T[3]) : nullptr; A.frac = frac; A.n_sub = ns; A.out = e2_shift(d, evol());
            k_2e_visc_apply<<<grid, 128, 0, d->stream>>>(A);
            d->launches++;
            CUDA_TRY(cudaGetLastError());
            int rp = e2_launch_propagate(d, 0);
            return rp ? rp : e2_visc_refresh_dt(d);
        };
        auto term = [&](double *out) -> int { return e2_visc_term(d, 0, (int)i, out); };
        const size_t plane_bytes = (size_t)d->P.nx * d->P.pitch * sizeof(double);
        for (int sc = 0; sc < ns; sc++) {
            if (t->visc_hv_integrator == SPRUCE_TI_EULER) {
                if ((rc = term(T[0])) || (rc = apply(evol(), 1.0, 1))) return rc;
            } else {
                CUDA_TRY(cudaMemcpyAsync(init, evol(), plane_bytes, cudaMemcpyDeviceToDevice, d->stream));
                if (t->visc_hv_integrator == SPRUCE_TI_RK2) {
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5, 1))) return rc;
                    if ((rc = term(T[0])) || (rc = apply(init, 1.0, 1))) return rc;
                } else {
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5, 1))) return rc;
                    std::swap(T[0], T[1]);                                                      // dqdt1 waits in T[1] while T[0] is the working term
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5, 1))) return rc;
                    std::swap(T[0], T[2]);
                    if ((rc = term(T[0])) || (rc = apply(init, 1.0, 1))) return rc;
                    std::swap(T[0], T[3]);
                    if ((rc = term(T[0]))) return rc;
                    double *d1 = T[1], *d2 = T[2], *d3 = T[3], *d4 = T[0];                     // -> (d1 + 2 d2 + 2 d3 + d4) / 6
                    T[0] = d1; T[1] = d2; T[2] = d3; T[3] = d4;
                    if ((rc = apply(init, 1.0, 4))) return rc;
                }
            }
        }
        if ((rc = e2_launch_propagate(d, 0)) || (rc = e2_visc_refresh_dt(d))) return rc;        // :178
    }
    return SPRUCE_OK;
}
int e2_enqueue_step(spruce_domain *d, int hist_slot)
{
    k_step_begin<<<1, 1, 0, d->stream>>>(d->ctl, d->dt_hist, hist_slot);
    d->launches++;
    int rc;
    if (!d->e2->visc.empty()) {                                                                 // iterateModules (evolution.cpp:66), then the stages read the primary dt plane
        if ((rc = e2_visc_refresh_dt(d)) || (rc = e2_av_iterate(d))) return rc;
    }
    if (d->cfg.time_integrator == SPRUCE_TI_RK4 && (rc = e2_ensure_rk4(d))) return rc;
    E2DeviceExec x{d};
    if ((rc = e2::advance(x, d->cfg.time_integrator))) return rc;                  // SPRUCE_TI_* = e2::TI2_* = 0, 1, 2
    k_step_end<<<1, 1, 0, d->stream>>>(d->ctl);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// spruce_module_viscosity / spruce_module_viscosity_term on this set (Viscosity::setupModule, viscosity.cpp:37-110)
int e2_visc_config(spruce_domain *d, int hv_time_integrator, double hv_epsilon, int gradient_correction)
{
    OneFluid2E *t = d->e2;
    if (d->cfg.n_ranks > 1) return fail(SPRUCE_ERR_UNSUPPORTED, "artificial_viscosity on ideal_mhd_2E is single-rank (its terms differentiate planes that have no halo exchange yet)");
    t->visc_hv_integrator = hv_time_integrator; t->visc_hv_epsilon = hv_epsilon; t->visc_gc = gradient_correction ? 1 : 0;
    int rc;
    if (!t->dt_plane && (rc = alloc_plane(d, &t->dt_plane))) return rc;
    for (int k = 0; k < 6; k++) if (!t->vplane[k] && (rc = alloc_plane(d, &t->vplane[k]))) return rc;
    return SPRUCE_OK;
}
int e2_visc_add_term(spruce_domain *d, const char *visc_opt, double strength, const char *var_to_diff, const char *var_to_evol, const char *species, const double *strength_plane, size_t count)
{
    OneFluid2E *t = d->e2;
    if (!t->dt_plane) return fail(SPRUCE_ERR_STATE, "spruce_module_viscosity must precede its terms");
    OneFluid2E::VTerm v{};
    v.opt = !strcmp(visc_opt, "local") ? 0 : !strcmp(visc_opt, "global") ? 1 : !strcmp(visc_opt, "boundary") ? 2 : !strcmp(visc_opt, "boundary_global") ? 3 : -1;
    if (v.opt < 0) return fail(SPRUCE_ERR_ARG, "Viscosity option must be global, local, or boundary.");
    v.strength = strength;
    v.var_diff = e2_var_index(var_to_diff);
    v.evol = e2_evolved_slot(var_to_evol);
    if (v.var_diff < 0) return fail(SPRUCE_ERR_ARG, "Each variable to differentiate must be a valid variable within the chosen equation set.");
    if (v.evol < 0) return fail(SPRUCE_ERR_ARG, "Each variable to evolve must be a valid evolved variable within the chosen equation set.");
    const char sp = (species && species[0]) ? species[0] : 'i';
    const bool evol_mom = v.evol == e2::Q_MX2 || v.evol == e2::Q_MY2, diff_vel = v.var_diff == e2::V2_v_x || v.var_diff == e2::V2_v_y;
    const bool evol_e = v.evol == e2::Q_EI2 || v.evol == e2::Q_EE2, diff_T = v.var_diff == e2::V2_i_temp || v.var_diff == e2::V2_e_temp;
    if (evol_mom && diff_vel && sp != 'i') return fail(SPRUCE_ERR_ARG, "Grid <e_n> was not found within the EquationSet.");      // viscosity.cpp:236: the set has no e_n
    v.scale_mode = (evol_mom && diff_vel) ? 1 : (evol_e && diff_T) ? 2 : 0;
    if (v.opt >= 2) {
        if (!strength_plane || count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "boundary viscosity needs its strength profile (%zu values)", (size_t)d->P.nx * d->P.ny);
        int rc = alloc_plane(d, &v.strength_plane);
        if (rc) return rc;
        if ((rc = h2d_plane(d, v.strength_plane, strength_plane))) return rc;
    }
    t->visc.push_back(v);
    return SPRUCE_OK;
}
int e2_upload(spruce_domain *d, const char *name, const double *host)
{
    OneFluid2E *t = d->e2;
    if (!strcmp(name, "be_z")) return SPRUCE_OK;                                     // not a variable of this set (2-D field); accepted and ignored
    const int s = static_slot(name);
    if (s >= 0) return h2d_plane(d, d->stat[s], host);
    if (!strcmp(name, "i_temp")) return h2d_plane(d, t->i_temp, host);
    if (!strcmp(name, "e_temp")) return h2d_plane(d, t->e_temp, host);
    const int ev = e2_evolved_slot(name);
    if (ev >= 0) return h2d_plane(d, t->set[t->order[0]][ev], host);
    if (e2_var_index(name) < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);
    return fail(SPRUCE_ERR_ARG, "<%s> is a derived variable and cannot be uploaded", name);
}
int e2_download(spruce_domain *d, const char *name, double *host)
{
    OneFluid2E *t = d->e2;
    const int s = static_slot(name);
    if (s >= 0) return d2h_plane(d, host, d->stat[s]);
    const int var = e2_var_index(name);
    if (var < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "download of <%s> before spruce_eqs_setup", name);
    const int ev = e2_evolved_slot(name);
    if (ev >= 0) return d2h_plane(d, host, t->set[t->order[0]][ev]);
    E2Args A{};
    e2_base(d, A);
    e2_cset(d, 0, A.S);
    A.out = e2_shift(d, d->scratch_out); A.var = var;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2e_derive<<<grid, 128, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return d2h_plane(d, host, d->scratch_out);
}
int e2_time_derivatives(spruce_domain *d, double *k_out, size_t count)
{
    OneFluid2E *t = d->e2;
    const size_t np = (size_t)d->P.nx * d->P.ny;
    if (!k_out || count != e2::NEV2 * np) return fail(SPRUCE_ERR_ARG, "k_out needs %zu values", e2::NEV2 * np);
    int rc = e2_ensure_rk4(d);
    if (rc) return rc;
    E2DeviceExec x{d};
    if ((rc = x.stage(0, 0, 1, 0.0, e2::KM2_EXPORT, 0, 0))) return rc;
    for (int v = 0; v < e2::NEV2; v++) if ((rc = d2h_plane(d, k_out + v * np, t->K1[v]))) return rc;
    return SPRUCE_OK;
}

// after spruce_eqs_setup on every rank of a decomposed run: halo rows of the static planes and the primary state, global dt minimum
int e2_initial_exchange(spruce_domain *d)
{
    double *stat_view[NEV];
    for (int v = 0; v < NEV; v++) stat_view[v] = d->stat[v < NSTATIC ? v : 0];
    int rc;
    if ((rc = peer_exchange(d, stat_view, nullptr))) return rc;
    if ((rc = e2_exchange(d, 0))) return rc;
    return peer_dt_allgather(d);
}
