/* capi_stub.c -- TEST INFRASTRUCTURE.  A recording stand-in for libspruce_b200.so, preloaded under the host shell (spruce_b200/bin/run) by
 * tests/test_host_shell_calls.py: every entry point the shell calls logs its name and arguments to $SPRUCE_STUB_LOG and returns SPRUCE_OK, planes
 * that were uploaded come back on download, time advances by a fixed step.  It checks -- without a GPU -- that the shell turns a .config / .state
 * into the right C-ABI calls in the right order (the reference's ModuleHandler order, parameter meaning and defaults).  Including the public
 * header makes every signature here a compile-time check against include/spruce_b200.h.  Nothing in the product links this. */
#include "spruce_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXP 64
struct spruce_domain { spruce_config cfg; char *names[MAXP]; double *planes[MAXP]; size_t counts[MAXP]; int n; double time; long long iter; };
static int g_rank = 0, g_n_ranks = 1;   /* set by spruce_domain_create before the first line is logged: one log per rank (`run -g N` forks before any call) */
static FILE *lg(void)
{
    static FILE *f = NULL;
    if (!f) {
        const char *p = getenv("SPRUCE_STUB_LOG");
        char path[4096];
        if (p && g_n_ranks > 1) { snprintf(path, sizeof(path), "%s.r%d", p, g_rank); p = path; }
        f = p ? fopen(p, "a") : stderr;
    }
    return f;
}
#define LOG(...) do { fprintf(lg(), __VA_ARGS__); fputc('\n', lg()); fflush(lg()); } while (0)
static void log_vec(const char *what, const double *p, size_t n)
{
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += p[i];
    LOG("  %s count=%zu first=%.17g last=%.17g sum=%.17g", what, n, n ? p[0] : 0.0, n ? p[n - 1] : 0.0, s);
}
const char *spruce_last_error(void) { return ""; }
int spruce_abi_version(void) { return SPRUCE_ABI_VERSION; }
int spruce_domain_create(const spruce_config *c, spruce_domain **out)
{
    spruce_domain *d = (spruce_domain *)calloc(1, sizeof(*d));
    d->cfg = *c; d->time = c->time;
    g_rank = c->rank; g_n_ranks = c->n_ranks;
    LOG("spruce_domain_create eqs=%d xdim=%d ydim=%d bc=%d,%d,%d,%d ti=%d row0=%d nx_local=%d n_ranks=%d ion_mass=%.17g gamma=%.17g epsilon=%.17g floors=%.17g,%.17g,%.17g open=%.17g,%.17g time=%.17g",
        c->equation_set, c->xdim, c->ydim, c->x_bound_1, c->x_bound_2, c->y_bound_1, c->y_bound_2, c->time_integrator, c->row0, c->nx_local, c->n_ranks, c->ion_mass, c->adiabatic_index,
        c->epsilon, c->density_min, c->temp_min, c->thermal_energy_min, c->open_boundary_strength, c->open_boundary_decay_base, c->time);
    *out = d;
    return SPRUCE_OK;
}
void spruce_domain_destroy(spruce_domain *d) { LOG("spruce_domain_destroy"); (void)d; }
int spruce_set_cell_sizes(spruce_domain *d, const double *dx, size_t nx, const double *dy, size_t ny) { (void)d; LOG("spruce_set_cell_sizes"); log_vec("d_x", dx, nx); log_vec("d_y", dy, ny); return SPRUCE_OK; }
int spruce_grid_upload(spruce_domain *d, const char *name, const double *h, size_t count)
{
    LOG("spruce_grid_upload %s", name);
    int k = 0;
    for (; k < d->n; k++) if (!strcmp(d->names[k], name)) break;
    if (k == d->n) { if (d->n == MAXP) return SPRUCE_ERR_ARG; d->names[k] = strdup(name); d->planes[k] = (double *)malloc(count * sizeof(double)); d->counts[k] = count; d->n++; }
    memcpy(d->planes[k], h, count * sizeof(double));
    return SPRUCE_OK;
}
int spruce_grid_download(spruce_domain *d, const char *name, double *h, size_t count)
{
    for (int k = 0; k < d->n; k++) if (!strcmp(d->names[k], name) && d->counts[k] == count) { memcpy(h, d->planes[k], count * sizeof(double)); return SPRUCE_OK; }
    for (size_t i = 0; i < count; i++) h[i] = 1.0 + d->cfg.rank;    /* derived variables: a recognisable constant, per rank */
    return SPRUCE_OK;
}
int spruce_plane_activity(spruce_domain *d, int *local_mask, int set_global_mask)
{
    if (local_mask) *local_mask = 1 << d->cfg.rank;                  /* the OR over the ranks must come back */
    LOG("spruce_plane_activity set=%d", set_global_mask);
    return SPRUCE_OK;
}
int spruce_mgpu_ipc_export(spruce_domain *d, void *handle64) { memset(handle64, 0xA0 + d->cfg.rank, 64); LOG("spruce_mgpu_ipc_export"); return SPRUCE_OK; }
int spruce_mgpu_ipc_connect(spruce_domain *d, const void *handles, int n)
{
    const unsigned char *h = (const unsigned char *)handles;
    int ok = n == d->cfg.n_ranks;
    for (int r = 0; r < n && ok; r++) for (int b = 0; b < 64; b++) if (h[r * 64 + b] != (unsigned char)(0xA0 + r)) ok = 0;
    LOG("spruce_mgpu_ipc_connect n=%d handles_in_rank_order=%d", n, ok);
    return SPRUCE_OK;
}
int spruce_mgpu_initial_exchange(spruce_domain *d) { (void)d; LOG("spruce_mgpu_initial_exchange"); return SPRUCE_OK; }
int spruce_eqs_setup(spruce_domain *d) { (void)d; LOG("spruce_eqs_setup"); return SPRUCE_OK; }
int spruce_eqs_propagate_changes(spruce_domain *d) { (void)d; LOG("spruce_eqs_propagate_changes"); return SPRUCE_OK; }
int spruce_next_step_size(spruce_domain *d, double *step) { (void)d; *step = 0.5; return SPRUCE_OK; }
int spruce_advance(spruce_domain *d, int n, double max_time, double *dt_used, int *done)
{
    const char *fr = getenv("SPRUCE_STUB_FAIL_RANK");                /* a rank whose device call fails mid-run (tests/test_host_slab_ranks.py) */
    if (fr && atoi(fr) == d->cfg.rank && d->iter >= 2) { LOG("spruce_advance FAILS on rank %d", d->cfg.rank); return SPRUCE_ERR_CUDA; }
    int k = 0;
    for (; k < n && d->time < max_time; k++) { double s = 0.5; if (d->time + s > max_time) s = max_time - d->time; d->time += s; d->iter++; if (dt_used) dt_used[k] = s; }
    if (done) *done = k;
    LOG("spruce_advance n=%d max_time=%.17g done=%d", n, max_time, k);
    return SPRUCE_OK;
}
int spruce_get_time(spruce_domain *d, double *t, int64_t *it) { if (t) *t = d->time; if (it) *it = d->iter; return SPRUCE_OK; }
int spruce_operator(spruce_domain *d, const char *op, int index, const double *q, const double *vel, double *out, size_t count) { (void)d; (void)q; (void)vel; LOG("spruce_operator %s %d", op, index); memset(out, 0, count * sizeof(double)); return SPRUCE_OK; }
int spruce_operator2(spruce_domain *d, const char *op, const double *a, const double *b, const double *c, double *out, size_t count) { (void)d; (void)a; (void)b; (void)c; LOG("spruce_operator2 %s", op); memset(out, 0, count * sizeof(double)); return SPRUCE_OK; }
int spruce_eqs_time_derivatives(spruce_domain *d, double *k, size_t count) { (void)d; memset(k, 0, count * sizeof(double)); LOG("spruce_eqs_time_derivatives"); return SPRUCE_OK; }
int spruce_module_thermal_conduction(spruce_domain *d, int fs, int ti, double eps, double dtmin, double weak)
{ (void)d; LOG("spruce_module_thermal_conduction flux_saturation=%d ti=%d epsilon=%.17g dt_subcycle_min=%.17g weakening_factor=%.17g", fs, ti, eps, dtmin, weak); return SPRUCE_OK; }
int spruce_module_radiative_losses(spruce_domain *d, int ti, double ramp, double temp, double eps, int prevent)
{ (void)d; LOG("spruce_module_radiative_losses ti=%d cutoff_ramp=%.17g cutoff_temp=%.17g epsilon=%.17g prevent_subcycling=%d", ti, ramp, temp, eps, prevent); return SPRUCE_OK; }
int spruce_module_ambient_heating(spruce_domain *d, const double *h, size_t count) { (void)d; LOG("spruce_module_ambient_heating"); log_vec("heating", h, count); return SPRUCE_OK; }
int spruce_module_viscosity(spruce_domain *d, int ti, double eps, int gc) { (void)d; LOG("spruce_module_viscosity hv_ti=%d hv_epsilon=%.17g gradient_correction=%d", ti, eps, gc); return SPRUCE_OK; }
int spruce_module_viscosity_term(spruce_domain *d, const char *opt, double strength, const char *vd, const char *ve, const char *sp, const double *plane, size_t count)
{ (void)d; LOG("spruce_module_viscosity_term opt=%s strength=%.17g diff=%s evol=%s species=%s plane=%d", opt, strength, vd, ve, sp, plane != NULL); if (plane) log_vec("strength_plane", plane, count); return SPRUCE_OK; }
int spruce_module_physical_viscosity(spruce_domain *d, double coeff, const double *plane, size_t count, double eps, int heat, int force, int gc, int ti, int inactive)
{ (void)d; LOG("spruce_module_physical_viscosity coeff=%.17g epsilon=%.17g heating_on=%d force_on=%d gradient_correction=%d ti=%d inactive=%d", coeff, eps, heat, force, gc, ti, inactive); log_vec("coeff_plane", plane, count); return SPRUCE_OK; }
int spruce_module_output_to_file(spruce_domain *d, const char *m, int on) { (void)d; LOG("spruce_module_output_to_file %s %d", m, on); return SPRUCE_OK; }
int spruce_module_inactive_mode(spruce_domain *d, const char *m, int on) { (void)d; LOG("spruce_module_inactive_mode %s %d", m, on); return SPRUCE_OK; }
int spruce_multispecies_mode(spruce_domain *d, int on) { (void)d; LOG("spruce_multispecies_mode %d", on); return SPRUCE_OK; }
int spruce_multispecies_reset(spruce_domain *d) { (void)d; LOG("spruce_multispecies_reset"); return SPRUCE_OK; }
int spruce_module_ms_fraction(spruce_domain *d, const char *m, double f) { (void)d; LOG("spruce_module_ms_fraction %s fraction=%.17g", m, f); return SPRUCE_OK; }
int spruce_module_output(spruce_domain *d, const char *name, double *h, size_t count) { (void)d; LOG("spruce_module_output %s", name); for (size_t i = 0; i < count; i++) h[i] = 7.0; return SPRUCE_OK; }
int spruce_module_ambient_heating_sink(spruce_domain *d, const double *r, size_t count) { (void)d; LOG("spruce_module_ambient_heating_sink"); log_vec("reduction", r, count); return SPRUCE_OK; }
int spruce_module_localized_heating(spruce_domain *d, double t0, double dur, double rate, double sx, double sy, double cx, double cy, double ramp)
{ (void)d; LOG("spruce_module_localized_heating start=%.17g duration=%.17g max=%.17g stddev=%.17g,%.17g center=%.17g,%.17g ramp_time=%.17g", t0, dur, rate, sx, sy, cx, cy, ramp); return SPRUCE_OK; }
int spruce_module_mass_injection(spruce_domain *d, double t0, double dur, double rate, double sx, double sy, double cx, double cy)
{ (void)d; LOG("spruce_module_mass_injection start=%.17g duration=%.17g max=%.17g stddev=%.17g,%.17g center=%.17g,%.17g", t0, dur, rate, sx, sy, cx, cy); return SPRUCE_OK; }
int spruce_module_momentum_injection(spruce_domain *d, double t0, double dur, double acc, double sx, double sy, double cx, double cy, double dirx, double diry, double ang, int osc, double per)
{ (void)d; LOG("spruce_module_momentum_injection start=%.17g duration=%.17g max_accel=%.17g stddev=%.17g,%.17g center=%.17g,%.17g dir=%.17g,%.17g angle=%.17g oscillatory=%d period=%.17g", t0, dur, acc, sx, sy, cx, cy, dirx, diry, ang, osc, per); return SPRUCE_OK; }
int spruce_module_div_cleaning(spruce_domain *d, double eps, double ts) { (void)d; LOG("spruce_module_div_cleaning epsilon=%.17g time_scale=%.17g", eps, ts); return SPRUCE_OK; }
int spruce_module_field_heating(spruce_domain *d, double c, double cp, double bp, double np, double rp, int inactive)
{ (void)d; LOG("spruce_module_field_heating coeff=%.17g current_pow=%.17g b_pow=%.17g n_pow=%.17g roc_pow=%.17g inactive=%d", c, cp, bp, np, rp, inactive); return SPRUCE_OK; }
int spruce_module_boundary_outflow(spruce_domain *d, const double *px, const double *py, size_t count, double acc, double len, int bnd, int shape, double feather, int fa, int dyn, double dt, double target)
{ (void)d; LOG("spruce_module_boundary_outflow max_accel=%.17g falloff_length=%.17g boundary=%d shape=%d feather=%.17g field_aligned=%d dynamic=%d dynamic_time=%.17g target=%.17g", acc, len, bnd, shape, feather, fa, dyn, dt, target);
  log_vec("pos_x", px, count); log_vec("pos_y", py, count); return SPRUCE_OK; }
int spruce_module_boundary_outflow_state(spruce_domain *d, double *m, double *a) { (void)d; if (m) *m = 1.5; if (a) *a = 2.5; return SPRUCE_OK; }
int spruce_module_anomalous_resistivity(spruce_domain *d, const double *px, const double *py, size_t count, const double *p, int np)
{
    (void)d;
    LOG("spruce_module_anomalous_resistivity n_params=%d", np);
    for (int k = 0; k < np; k++) LOG("  p[%d]=%.17g", k, p[k]);
    log_vec("pos_x", px, count); log_vec("pos_y", py, count);
    return SPRUCE_OK;
}
int spruce_module_anomalous_resistivity_state(spruce_domain *d, int *i, int *j, int *n) { (void)d; if (i) *i = 3; if (j) *j = 4; if (n) *n = 5; return SPRUCE_OK; }
int spruce_eqs_ideal_mhd_options(spruce_domain *d, double gv) { (void)d; LOG("spruce_eqs_ideal_mhd_options global_viscosity=%.17g", gv); return SPRUCE_OK; }
int spruce_eqs_ideal_mhd_moc_limiting(spruce_domain *d, int b, double bl, double bu, int m, double ml, double mu)
{ (void)d; LOG("spruce_eqs_ideal_mhd_moc_limiting b=%d %.17g %.17g mom=%d %.17g %.17g", b, bl, bu, m, ml, mu); return SPRUCE_OK; }
int spruce_eqs_ideal2f_options(spruce_domain *d, int sub, int curl) { (void)d; LOG("spruce_eqs_ideal2f_options use_sub_cycling=%d remove_curl_terms=%d", sub, curl); return SPRUCE_OK; }
int spruce_module_eic_thermalization(spruce_domain *d) { (void)d; LOG("spruce_module_eic_thermalization"); return SPRUCE_OK; }
int spruce_module_subcycles(spruce_domain *d, const char *w, int *c) { (void)d; (void)w; *c = 9; return SPRUCE_OK; }
