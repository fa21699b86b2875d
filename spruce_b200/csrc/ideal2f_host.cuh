// ideal2f_host.cuh -- launch sequences of the two-fluid equation set (included by capi.cu inside its anonymous namespace,
// after the plane helpers).  Same step structure as the ideal-MHD path: PlasmaDomain::advanceTime, evolution.cpp:59-124.
#pragma once
#include "ideal2f_kernels.cuh"
#include "ideal2f_sides.cuh"

// ---- ordered boundary passes (ideal2f_sides.cuh): boundary sets in which an open_ucnp side meets a fixed / reflect side.
// STATUS: written after the round-1 GPU budget was spent; the passes are proven on the host (tests/test_ideal2f_sides_host_check.py), the launch side
// has not run on a GPU yet (tests/test_zz_gpu_unvalidated.py).  The default path (every other boundary set) is untouched.
struct TfSideArgs { tf2::Geo g; tf2::Planes G, P; int side; const int *done_ptr; };
__global__ void __launch_bounds__(128) k_2f_side(const TfSideArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (*A.done_ptr || t >= tf2::side_length(A.g, A.side)) return;
    tf2::side_line(A.g, A.G, A.P, A.side, t);
}
// k_2f_propagate without the pointwise zeroing and without dt: recomputeEvolvedVarsFromStateVars (setup) and enforceMinimums only
__global__ void __launch_bounds__(256) k_2f_floors(const DomainParams P, const TfPropArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= P.ny) return;
    const size_t off = (size_t)r * P.pitch + j;
    double i_rho = A.U[F_IRHO][off], e_rho = A.U[F_ERHO][off], i_e = A.U[F_IE][off], e_e = A.U[F_EE][off];
    if (A.from_state) {                                                                                 // ideal2F.cpp:107-117
        const double i_n = ddiv(i_rho, P.m_i, P.rm_i), e_n = ddiv(e_rho, A.base.m_e, A.base.rm_e);
        i_e = ((i_n * kKB) * A.i_temp[off]) / P.gm1;
        e_e = ((e_n * kKB) * A.e_temp[off]) / P.gm1;
    }
    A.U[F_IRHO][off] = smax(ddiv(i_rho, P.m_i, P.rm_i), P.n_min) * P.m_i;                               // :98-105
    A.U[F_ERHO][off] = smax(ddiv(e_rho, A.base.m_e, A.base.rm_e), P.n_min) * A.base.m_e;
    A.U[F_IE][off] = smax(i_e, P.e_min); A.U[F_EE][off] = smax(e_e, P.e_min);
}
// recomputeDT over the interior of a finished state (ideal2F.cpp:169-198) -> running minimum
__global__ void __launch_bounds__(256) k_2f_dt_full(const DomainParams P, const TfPropArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    double dtc = 1.7976931348623157e308;
    if (!*A.base.done_ptr && j < P.ny && is_interior(P, r, j)) {
        const size_t off = (size_t)r * P.pitch + j;
        dtc = tf_cell_dt(P, A.base, A.U[F_ERHO][off], A.U[F_EMX][off], A.U[F_EMY][off], A.U[F_EE][off], P.tx.d[r], P.ty.d[j]);
    }
    block_min_to_global(dtc, A.dtmin_bits);
}

struct PlaneSet2 { double *p[NEV2] = {nullptr}; };

struct TwoFluid {
    PlaneSet2 P, M, M2, K1, K2;
    double *i_temp = nullptr, *e_temp = nullptr;     // uploaded state temperatures, consumed by setup
    double *vel[4] = {nullptr};                        // species velocities of the state a stage evaluates (k_2f_velocity)
    bool rk4_alloc = false;
    int use_sub_cycling = 1;                           // Ideal2F default (ideal2F.hpp:62)
    int remove_curl_terms = 0;
    int eic = 0;                                       // eic_thermalization module configured
    bool ordered_sides = false;                        // an open_ucnp side meets a fixed / reflect side: literal, ordered boundary passes (ideal2f_sides.cuh)
};

const char *const kTfEvolvedNames[NEV2] = {"i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy",
                                           "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"};
const char *const kTfVarNames[W_COUNT] = {"i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_temp", "e_temp", "bi_x", "bi_y", "bi_z", "E_x", "E_y", "E_z",
                                          "grav_x", "grav_y", "i_n", "e_n", "i_v_x", "i_v_y", "e_v_x", "e_v_y", "j_x", "j_y", "i_press", "e_press", "press",
                                          "i_thermal_energy", "e_thermal_energy", "rho", "rho_c", "n", "dn", "dt", "dt_i", "b_x", "b_y", "b_z", "b_mag", "b_mag_xy",
                                          "b_hat_x", "b_hat_y", "curlE_z", "divE", "divB", "i_dPdx", "e_dPdx"};

int tf_var_index(const char *name) { for (int v = 0; v < W_COUNT; v++) if (!strcmp(kTfVarNames[v], name)) return v; return -1; }
int tf_evolved_slot(const char *name) { for (int v = 0; v < NEV2; v++) if (!strcmp(kTfEvolvedNames[v], name)) return v; return -1; }

int tf_alloc_set(spruce_domain *d, PlaneSet2 &s)
{
    for (int v = 0; v < NEV2; v++) { int rc = alloc_plane(d, &s.p[v]); if (rc) return rc; }
    return SPRUCE_OK;
}
int tf_create(spruce_domain *d)
{
    TwoFluid *t = new TwoFluid();
    d->tf = t;
    {
        const int b[4] = {d->cfg.x_bound_1, d->cfg.x_bound_2, d->cfg.y_bound_1, d->cfg.y_bound_2};
        bool ucnp = false, wall = false;
        for (int s = 0; s < 4; s++) { ucnp |= (b[s] == SPRUCE_BC_OPEN_UCNP); wall |= (b[s] == SPRUCE_BC_FIXED || b[s] == SPRUCE_BC_REFLECT); }
        t->ordered_sides = ucnp && wall;
    }
    int rc;
    if ((rc = tf_alloc_set(d, t->P))) return rc;
    if ((rc = tf_alloc_set(d, t->M))) return rc;
    if ((rc = alloc_plane(d, &t->i_temp))) return rc;
    if ((rc = alloc_plane(d, &t->e_temp))) return rc;
    for (int k = 0; k < 4; k++) if ((rc = alloc_plane(d, &t->vel[k]))) return rc;
    return SPRUCE_OK;
}
int tf_ensure_rk4(spruce_domain *d)
{
    TwoFluid *t = d->tf;
    if (t->rk4_alloc) return SPRUCE_OK;
    int rc;
    if ((rc = tf_alloc_set(d, t->M2))) return rc;
    if ((rc = tf_alloc_set(d, t->K1))) return rc;
    if ((rc = tf_alloc_set(d, t->K2))) return rc;
    t->rk4_alloc = true;
    return SPRUCE_OK;
}

void tf_base(const spruce_domain *d, TfArgs &A)
{
    const TwoFluid *t = d->tf;
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    for (int v = 0; v < NEV2; v++) { A.K1[v] = t->K1.p[v]; A.K2[v] = t->K2.p[v]; }
    A.step_ptr = &d->ctl->step; A.done_ptr = &d->ctl->done; A.dtmin_bits = &d->ctl->dtmin_bits;
    A.m_e = kMElectron; A.rm_e = 1.0 / kMElectron;
    A.curl_terms = t->remove_curl_terms ? 0 : 1;
    A.fast = d->fast_interior ? 1 : 0;
    A.eic = t->eic;
}

// The supported boundary sets: any mix of periodic and open_ucnp, or any mix of periodic / fixed / reflect.  A ucnp pass next to a
// fixed / reflect side reads momenta the later side zeroes afterwards (evolution.cpp:126-152 runs x1, x2, y1, y2 in that order),
// which this first version does not track; `open` needs a one-fluid rho / thermal_energy (evolution.cpp:163-229 aborts on name2index).
int tf_check_boundaries(const spruce_config &c)
{
    const int b[4] = {c.x_bound_1, c.x_bound_2, c.y_bound_1, c.y_bound_2};
    for (int s = 0; s < 4; s++) {
        if (b[s] == SPRUCE_BC_OPEN) return fail(SPRUCE_ERR_UNSUPPORTED, "open boundaries need a single-fluid equation set (name2index(\"rho\") aborts in the reference)");
        if (b[s] == SPRUCE_BC_OPEN_MOC) return fail(SPRUCE_ERR_UNSUPPORTED, "open_moc boundaries exist for ideal_mhd only (idealmhd.cpp:306)");
    }
    return SPRUCE_OK;
}

int tf_launch_ghosts(spruce_domain *d, const PlaneSet2 &U, int primary)
{
    if (!(d->any_ucnp || (primary && d->any_primary_ghost))) return SPRUCE_OK;
    TfGhostArgs G{};
    for (int v = 0; v < NEV2; v++) G.U[v] = U.p[v];
    G.primary = primary;
    const int n = d->P.ny > d->P.nx ? d->P.ny : d->P.nx;
    dim3 grid((n + 127) / 128, 4);
    k_2f_ghosts<<<grid, 128, 0, d->stream>>>(d->P, G);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// slab decomposition: halo rows of the 14 evolved planes from the ring neighbours (two packed exchanges of 8 planes)
int tf_exchange(spruce_domain *d, const PlaneSet2 &U)
{
    if (d->cfg.n_ranks == 1) return SPRUCE_OK;
    double *a[NEV], *b[NEV];
    for (int k = 0; k < NEV; k++) { a[k] = U.p[k]; b[k] = U.p[NEV + (k % (NEV2 - NEV))]; }
    int rc = peer_exchange(d, a, nullptr);
    if (rc) return rc;
    return peer_exchange(d, b, nullptr);
}
// the four boundary passes of the primary state in the reference's order, then the dt minimum over the finished state
int tf_ordered_sides_and_dt(spruce_domain *d, const PlaneSet2 &U)
{
    TfSideArgs A{};
    A.g.nx = d->P.gnx; A.g.ny = d->P.ny; A.g.pitch = d->P.pitch; A.g.row0 = d->P.row0; A.g.nxl = d->P.nx;
    A.g.bc[0] = d->cfg.x_bound_1; A.g.bc[1] = d->cfg.x_bound_2; A.g.bc[2] = d->cfg.y_bound_1; A.g.bc[3] = d->cfg.y_bound_2;
    A.g.xl = d->P.xl; A.g.xu = d->P.xu; A.g.yl = d->P.yl; A.g.yu = d->P.yu;
    const long long sh = (long long)d->P.row0 * d->P.pitch;
    for (int v = 0; v < NEV2; v++) { A.G.u[v] = U.p[v] - sh; A.P.u[v] = U.p[v] - sh; }
    A.done_ptr = &d->ctl->done;
    const int n = d->P.ny > d->P.nx ? d->P.ny : d->P.nx;
    for (int side = 0; side < 4; side++) {
        const int bc = A.g.bc[side];
        if (bc != SPRUCE_BC_FIXED && bc != SPRUCE_BC_REFLECT && bc != SPRUCE_BC_OPEN_UCNP) continue;
        A.side = side;
        k_2f_side<<<(n + 127) / 128, 128, 0, d->stream>>>(A);
        d->launches++;
    }
    k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, 0);
    TfPropArgs Q{};
    tf_base(d, Q.base);
    for (int v = 0; v < NEV2; v++) Q.U[v] = U.p[v];
    Q.dtmin_bits = &d->ctl->dtmin_bits;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_2f_dt_full<<<grid, 256, 0, d->stream>>>(d->P, Q);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int tf_finish_stage(spruce_domain *d, const PlaneSet2 &U, int primary)
{
    // ordered mode: an intermediate set only needs the open_ucnp copies (fixed / reflect act on the primary state, where they change nothing
    // between two of its own propagates), and those never overlap -> the concurrent kernel; the primary state gets the literal ordered passes
    int rc = (d->tf->ordered_sides && primary) ? tf_ordered_sides_and_dt(d, U) : tf_launch_ghosts(d, U, primary);
    if (rc) return rc;
    return tf_exchange(d, U);
}

int tf_launch_stage(spruce_domain *d, const PlaneSet2 &S, const PlaneSet2 &B, const PlaneSet2 &D, double coef, int primary, int kmode)
{
    TfArgs A{};
    tf_base(d, A);
    for (int v = 0; v < NEV2; v++) { A.S[v] = S.p[v]; A.B[v] = B.p[v]; A.D[v] = D.p[v]; }
    A.coef = coef; A.primary = primary; A.kmode = kmode;
    if (d->tf->ordered_sides) A.primary = 0;          // no pointwise zeroing, no dt in the kernel: tf_ordered_sides_and_dt follows for the primary state
    if (primary && kmode != KM_EXPORT) { k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, 0); d->launches++; }
    TfVelArgs V{};
    for (int v = 0; v < NEV2; v++) V.U[v] = S.p[v];
    for (int k = 0; k < 4; k++) { V.vel[k] = d->tf->vel[k]; A.vel[k] = d->tf->vel[k]; }
    V.done_ptr = &d->ctl->done;
    const int halo = d->cfg.n_ranks > 1 ? HALO : 0;                   // the neighbours' cells are resident in the halo rows
    V.row_off = -halo;
    dim3 vgrid((d->P.ny + 255) / 256, d->P.nx + 2 * halo);
    k_2f_velocity<<<vgrid, 256, 0, d->stream>>>(d->P, V);
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2f_stage<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

int tf_launch_propagate(spruce_domain *d, int from_state)
{
    TwoFluid *t = d->tf;
    k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, 0);
    TfPropArgs A{};
    tf_base(d, A.base);
    for (int v = 0; v < NEV2; v++) A.U[v] = t->P.p[v];
    A.i_temp = t->i_temp; A.e_temp = t->e_temp; A.from_state = from_state;
    A.dtmin_bits = &d->ctl->dtmin_bits;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    if (t->ordered_sides) {
        k_2f_floors<<<grid, 256, 0, d->stream>>>(d->P, A);
        d->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return tf_ordered_sides_and_dt(d, t->P);
    }
    k_2f_propagate<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return tf_launch_ghosts(d, t->P, 1);
}
// after spruce_eqs_setup on every rank of a decomposed two-fluid run: halo rows of the static planes and the primary state, global dt minimum
int tf_initial_exchange(spruce_domain *d)
{
    double *stat_view[NEV];
    for (int v = 0; v < NEV; v++) stat_view[v] = d->stat[v < NSTATIC ? v : 0];
    int rc;
    if ((rc = peer_exchange(d, stat_view, nullptr))) return rc;
    if ((rc = tf_exchange(d, d->tf->P))) return rc;
    return peer_dt_allgather(d);
}

int tf_enqueue_step(spruce_domain *d, int hist_slot)
{
    TwoFluid *t = d->tf;
    int rc;
    k_step_begin<<<1, 1, 0, d->stream>>>(d->ctl, d->dt_hist, hist_slot);
    d->launches++;
    const int ti = d->cfg.time_integrator;
    if (ti == SPRUCE_TI_EULER) {
        if ((rc = tf_launch_stage(d, t->P, t->P, t->M, 1.0, 1, KM_NONE))) return rc;
        std::swap(t->P, t->M);
        if ((rc = tf_finish_stage(d, t->P, 1))) return rc;
    } else if (ti == SPRUCE_TI_RK2) {
        if ((rc = tf_launch_stage(d, t->P, t->P, t->M, 0.5, 0, KM_NONE))) return rc;
        if ((rc = tf_finish_stage(d, t->M, 0))) return rc;
        if ((rc = tf_launch_stage(d, t->M, t->P, t->P, 1.0, 1, KM_NONE))) return rc;
        if ((rc = tf_finish_stage(d, t->P, 1))) return rc;
    } else {
        if ((rc = tf_ensure_rk4(d))) return rc;
        if ((rc = tf_launch_stage(d, t->P, t->P, t->M, 0.5, 0, KM_STORE_K1))) return rc;
        if ((rc = tf_finish_stage(d, t->M, 0))) return rc;
        if ((rc = tf_launch_stage(d, t->M, t->P, t->M2, 0.5, 0, KM_STORE_K2))) return rc;
        if ((rc = tf_finish_stage(d, t->M2, 0))) return rc;
        if ((rc = tf_launch_stage(d, t->M2, t->P, t->M, 1.0, 0, KM_ADD_K2))) return rc;
        if ((rc = tf_finish_stage(d, t->M, 0))) return rc;
        if ((rc = tf_launch_stage(d, t->M, t->P, t->P, 1.0, 1, KM_FINAL))) return rc;
        if ((rc = tf_finish_stage(d, t->P, 1))) return rc;
    }
    if (d->cfg.n_ranks > 1 && (rc = peer_dt_allgather(d))) return rc;
    k_step_end<<<1, 1, 0, d->stream>>>(d->ctl);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

int tf_upload(spruce_domain *d, const char *name, const double *host)
{
    TwoFluid *t = d->tf;
    const int s = static_slot(name);
    if (s >= 0) return h2d_plane(d, d->stat[s], host);
    if (!strcmp(name, "i_temp")) return h2d_plane(d, t->i_temp, host);
    if (!strcmp(name, "e_temp")) return h2d_plane(d, t->e_temp, host);
    const int ev = tf_evolved_slot(name);
    if (ev >= 0) return h2d_plane(d, t->P.p[ev], host);
    if (tf_var_index(name) < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);
    return fail(SPRUCE_ERR_ARG, "<%s> is a derived variable and cannot be uploaded", name);
}

int tf_download(spruce_domain *d, const char *name, double *host)
{
    TwoFluid *t = d->tf;
    const int s = static_slot(name);
    if (s >= 0) return d2h_plane(d, host, d->stat[s]);
    const int var = tf_var_index(name);
    if (var < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "download of <%s> before spruce_eqs_setup", name);
    const int ev = tf_evolved_slot(name);
    if (ev >= 0) return d2h_plane(d, host, t->P.p[ev]);
    TfDeriveArgs A{};
    tf_base(d, A.base);
    for (int v = 0; v < NEV2; v++) A.U[v] = t->P.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.out = d->scratch_out; A.which = var;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_2f_derive<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return d2h_plane(d, host, d->scratch_out);
}

int tf_time_derivatives(spruce_domain *d, double *k_out, size_t count)
{
    TwoFluid *t = d->tf;
    const size_t np = (size_t)d->P.nx * d->P.ny;
    if (!k_out || count != NEV2 * np) return fail(SPRUCE_ERR_ARG, "k_out needs %zu values", NEV2 * np);
    int rc = tf_ensure_rk4(d);
    if (rc) return rc;
    if ((rc = tf_launch_stage(d, t->P, t->P, t->M, 0.0, 0, KM_EXPORT))) return rc;
    for (int v = 0; v < NEV2; v++) if ((rc = d2h_plane(d, k_out + v * np, t->K1.p[v]))) return rc;
    return SPRUCE_OK;
}
