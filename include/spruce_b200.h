/* spruce_b200.h -- C ABI of the B200-native per-timestep advance for SPRUCE (gszypko/spruce).
 *
 * This is the drop-in boundary: plain C, opaque handle, plain pointers and sizes, int status codes
 * (0 = ok; spruce_last_error() gives the message).  The reference has NO FFI of its own -- its
 * plug-in surface is C++ inheritance inside one binary -- so each entry point below names the
 * reference member function(s) whose arithmetic it replaces (file:line in the reference tree).
 * The host-side C++ mirror of the reference classes (spruce_b200/host/) and the Python binding
 * (spruce_b200/capi.py) call only these functions.
 *
 * Data layout at the boundary = the reference's Grid: row-major plane[i*ydim + j], i = x index,
 * j = y index, j contiguous (source/mhd/grid.cpp:516-526), FP64.  With a slab decomposition a rank
 * passes only its own rows [row0, row0+nx_local) of every plane.
 *
 * Threading: one host thread drives one handle; all device work is ordered on the handle's stream.
 * There is no CPU fallback: every call fails with SPRUCE_ERR_CUDA when no CUDA device is usable.
 */
#ifndef SPRUCE_B200_H
#define SPRUCE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPRUCE_ABI_VERSION 1

/* status codes */
enum {
    SPRUCE_OK = 0,
    SPRUCE_ERR_ARG = 1,       /* bad argument / unknown plane name (reference: assert + abort)            */
    SPRUCE_ERR_CUDA = 2,      /* CUDA runtime failure or no device                                         */
    SPRUCE_ERR_STATE = 3,     /* call order violated (e.g. advance before setup)                           */
    SPRUCE_ERR_UNSUPPORTED = 4 /* feature outside the hot-path scope (open_moc, non rank-1 d_x/d_y, ...)   */
};

/* PlasmaDomain::BoundaryCondition, source/mhd/plasmadomain.hpp:22 (same order => same integer values) */
enum { SPRUCE_BC_PERIODIC = 0, SPRUCE_BC_OPEN = 1, SPRUCE_BC_FIXED = 2, SPRUCE_BC_REFLECT = 3,
       SPRUCE_BC_OPEN_MOC = 4, SPRUCE_BC_OPEN_UCNP = 5 };
/* PlasmaDomain::TimeIntegrator, source/mhd/plasmadomain.hpp:29 */
enum { SPRUCE_TI_EULER = 0, SPRUCE_TI_RK2 = 1, SPRUCE_TI_RK4 = 2 };
/* EquationSet::m_sets, source/equationsets/equationset.hpp:22 (index in that list) */
enum { SPRUCE_EQS_IDEAL_MHD = 0, SPRUCE_EQS_IDEAL_MHD_2E = 2, SPRUCE_EQS_IDEAL_2F = 3 };   /* index in EquationSet::m_sets (equationset.hpp:22); ideal_mhd_2E: one fluid with
                                                                                            separate ion / electron thermal energies (idealmhd2E.cpp) -- written after the round-1
                                                                                            GPU budget was spent, host-checked, not yet run on a GPU; no modules */

/* Everything PlasmaDomain reads from the .config / .state headers that the device needs
 * (source/mhd/plasmadomain.hpp:93-153, source/mhd/fileio.cpp:297-325). */
typedef struct spruce_config {
    int32_t abi_version;        /* = SPRUCE_ABI_VERSION */
    int32_t equation_set;       /* SPRUCE_EQS_* */
    int32_t xdim, ydim;         /* GLOBAL grid size (m_xdim, m_ydim) */
    int32_t x_bound_1, x_bound_2, y_bound_1, y_bound_2;  /* SPRUCE_BC_* */
    int32_t time_integrator;    /* SPRUCE_TI_* */
    int32_t device;             /* CUDA device ordinal; -1 = current device */
    /* slab decomposition along x (i): this handle owns global rows [row0, row0 + nx_local).
     * Single GPU: row0 = 0, nx_local = xdim, n_ranks = 1. */
    int32_t row0, nx_local, rank, n_ranks;
    double ion_mass;            /* m_ion_mass   (.state header)  */
    double adiabatic_index;     /* m_adiabatic_index             */
    double epsilon;             /* CFL safety factor             */
    double density_min, temp_min, thermal_energy_min;
    double open_boundary_strength, open_boundary_decay_base;
    double time;                /* m_time at start (continue mode) */
} spruce_config;

typedef struct spruce_domain spruce_domain;

const char *spruce_last_error(void);
int spruce_abi_version(void);

/* PlasmaDomain::PlasmaDomain + computeIterationBounds (source/mhd/plasmadomain.cpp:13-75,138-161):
 * allocates the SoA FP64 device arena for the equation set's variables and the domain grids. */
int spruce_domain_create(const spruce_config *cfg, spruce_domain **out);
void spruce_domain_destroy(spruce_domain *dom);

/* Cell sizes.  The reference stores d_x, d_y as full planes (plasmadomain.hpp:35) but requires d_x to vary
 * with i only and d_y with j only (README.md:41); the arena keeps them as 1-D tables.  d_x: xdim GLOBAL
 * entries (every rank passes the whole array), d_y: ydim entries.  Also derives the exact-division tables
 * used by the stencil kernels (DESIGN.md "exact division"). */
int spruce_set_cell_sizes(spruce_domain *dom, const double *d_x, size_t n_x, const double *d_y, size_t n_y);

/* Grid upload / download by the reference's variable names: domain grids "be_x","be_y","be_z","pos_x","pos_y"
 * (plasmadomain.hpp:36) and every EquationSet variable (idealmhd.hpp:19-23 / ideal2F.hpp:23-38).
 * host: nx_local*ydim doubles, reference layout.  Derived variables are materialised on download from the
 * evolved state exactly as IdealMHD::recomputeDerivedVarsFromEvolvedVars / recomputeDT would have left them
 * (source/equationsets/idealmhd.cpp:241-304).  Replaces EquationSet::grid(name) (equationset.cpp:113-121). */
int spruce_grid_upload(spruce_domain *dom, const char *name, const double *host, size_t count);
int spruce_grid_download(spruce_domain *dom, const char *name, double *host, size_t count);

/* EquationSet::setupEquationSet -> populateVariablesFromState (source/equationsets/equationset.cpp:87-104):
 * state variables -> evolved variables, floors, ghost zones, derived variables, dt. */
int spruce_eqs_setup(spruce_domain *dom);

/* EquationSet::propagateChanges on the primary state (source/equationsets/equationset.cpp:212-220);
 * called by host-side modules after they edited an evolved plane through spruce_grid_upload. */
int spruce_eqs_propagate_changes(spruce_domain *dom);

/* epsilon * getDT().min(m_xl_dt, m_yl_dt, m_xu_dt, m_yu_dt) (source/mhd/evolution.cpp:62,
 * source/mhd/grid.cpp:71-82): the step the NEXT advance will use.  With n_ranks > 1 this is the LOCAL slab
 * minimum until spruce_set_global_dt_min() has been fed the all-reduced value. */
int spruce_next_step_size(spruce_domain *dom, double *step);

/* PlasmaDomain::advanceTime x n_steps (source/mhd/evolution.cpp:59-82) with the configured integrator
 * (integrateEuler/RK2/RK4, :84-124) and the device-resident modules, entirely on the device: no host
 * round trip between steps.  dt_used (may be NULL) receives the n_steps step sizes.  Stops early, like
 * PlasmaDomain::run (:26), once time >= max_time when max_time > 0; *steps_done (may be NULL) gets the count. */
int spruce_advance(spruce_domain *dom, int n_steps, double max_time, double *dt_used, int *steps_done);
int spruce_get_time(spruce_domain *dom, double *time, int64_t *iter);

/* One evaluation of EquationSet::computeTimeDerivatives on the primary state
 * (source/equationsets/equationset.cpp:204-210 -> idealmhd.cpp:42-105); k_out: n_evolved planes of
 * nx_local*ydim doubles in evolved_variables() order (idealmhd.hpp:32-34).  Test / module hook. */
int spruce_eqs_time_derivatives(spruce_domain *dom, double *k_out, size_t count);

/* PlasmaDomain differential operators on a host plane (source/mhd/derivs.cpp), so that host-side modules
 * that are not ported keep working: op = "derivative1D" (:223), "secondDerivative1D" (:417), "laplacian" (:458),
 * "transportDerivative1D" (:122; vel required).  index: 0 = x, 1 = y. Single-rank only. */
int spruce_operator(spruce_domain *dom, const char *op, int index, const double *q, const double *vel, double *out, size_t count);
/* The two-operand operators (source/mhd/plasmadomain.hpp:201-242): op = "divergence2D" (a = a_x, b = a_y; derivs.cpp:407-409),
 * "curl2D" (a = x, b = y; derivs.cpp:472-474: d(y)/dx - d(x)/dy), "transportDivergence2D" (a = quantity, b = vel_x, c = vel_y;
 * derivs.cpp:216-220).  c may be null for the first two.  Each is two passes of spruce_operator combined per cell on the host. */
int spruce_operator2(spruce_domain *dom, const char *op, const double *a, const double *b, const double *c, double *out, size_t count);

/* ---- physics modules on the device (source/modules/...) ; call order = execution order (modulehandler.cpp:27-65) */
/* ThermalConduction::parseModuleConfigs (solar/thermalconduction.cpp:16-30) */
int spruce_module_thermal_conduction(spruce_domain *dom, int flux_saturation, int time_integrator, double epsilon,
                                     double dt_subcycle_min, double weakening_factor);
/* RadiativeLosses::parseModuleConfigs (solar/radiativelosses.cpp:17-31) */
int spruce_module_radiative_losses(spruce_domain *dom, int time_integrator, double cutoff_ramp, double cutoff_temp,
                                   double epsilon, int prevent_subcycling);
/* AmbientHeating::setupModule (solar/ambientheating.cpp:28-40): heating = nx_local*ydim plane built on the host
 * (the exp() profile is evaluated there with the host libm, as the reference does once at setup). */
int spruce_module_ambient_heating(spruce_domain *dom, const double *heating, size_t count);
/* Viscosity (source/modules/viscosity.cpp): module-level options (parseModuleConfigs :6-24), then one call per term of the
 * comma-separated lists visc_opt / visc_strength / visc_vars_to_diff / visc_vars_to_evol / visc_species (setupModule :37-110).
 * Terms with strength <= 1 are added to the right-hand side (computeTimeDerivativesModule :112-123), terms with strength > 1
 * are sub-cycled (iterateModule :125-180).  visc_opt "boundary"/"boundary_global" need the strength profile of
 * getBoundaryViscosity (:278-325), built on the host (libm exp) like the reference does.  ideal_mhd domains, and ideal_mhd_2E domains on one rank (variable names
 * of idealmhd2E.hpp:18-22; the timescale is the primary dt plane for either species, :45); ideal_2F domains are refused. */
int spruce_module_viscosity(spruce_domain *dom, int hv_time_integrator, double hv_epsilon, int gradient_correction);
int spruce_module_viscosity_term(spruce_domain *dom, const char *visc_opt, double strength, const char *var_to_diff, const char *var_to_evol,
                                 const char *species, const double *strength_plane, size_t count);
/* PhysicalViscosity (source/modules/solar/physicalviscosity.cpp:47-245): Braginskii eta_0 viscous heating and force, sub-cycled
 * (euler | rk2) inside iterateModule.  coeff_plane = constructCoefficientGrid(coeff, ramp_length, buffer_length) (:247-267), built by
 * the host exactly as the reference does.  count of sub-cycles: spruce_module_subcycles(dom, "physical_viscosity", &n). */
int spruce_module_physical_viscosity(spruce_domain *dom, double coeff, const double *coeff_plane, size_t count, double epsilon, int heating_on,
                                     int force_on, int gradient_correction, int time_integrator, int inactive_mode);
/* output_to_file = true of ThermalConduction / RadiativeLosses (thermalconduction.cpp:226-237, radiativelosses.cpp:172-179): the planes Module::fileOutput
 * appends to mhd.out -- "thermal_conduction" and "rad" = the module's (e_after - e_before)/dt of the last step, "flux_saturation" = the saturation coefficient
 * of that step's first temperature field (zero planes before the first step, as in the reference).  Enable per module, then download by plane name.
 * Also: module "anomalous_resistivity" -> planes "anomalous_diffusivity", "anomalous_template", "joule_heating" (anomalousresistivity.cpp:320-329), and the
 * plane "field_heating" (fieldheating.cpp:73-80: mask*(dt*heating) of the last step), which needs no enabling; module "physical_viscosity" -> planes
 * "viscous_heating", "viscous_force_x" / "_y" / "_z" (physicalviscosity.cpp:292-308: the averages over the last step's sub-cycles, also in inactive_mode); module
 * "artificial_viscosity" (enable AFTER its terms are configured) -> planes "visc_dqdt:<i>", "visc_lap:<i>", "visc_str:<i>", "visc_dt:<i>" of term i in config order
 * (viscosity.cpp:351-376: m_grids_dqdt / _lap / _strength / _dt, each what the term's last evaluation left; zero planes before the first one). */
int spruce_module_output_to_file(spruce_domain *dom, const char *module, int on);
/* multispecies_mode = true of PlasmaDomain (source/mhd/plasmadomain.hpp:134-135, fileio.cpp:164-183, 322): three cumulative planes -- "cumulative_electron_heating",
 * "cumulative_ion_heating", "cumulative_joule_heating", downloaded by name with spruce_module_output -- to which thermal_conduction, radiative_losses, ambient_heating,
 * ambient_heating_sink, localized_heating and physical_viscosity add their energy input, split by the module's ms_electron_heating_fraction (defaults as in the
 * reference: 1, 1, 0.5, 0.5, 0, 0), and anomalous_resistivity its joule heating.  The host resets them after every stored frame (evolution.cpp:36-41).
 * ideal_mhd domains only. */
int spruce_multispecies_mode(spruce_domain *dom, int on);
/* inactive_mode = true of "thermal_conduction" / "radiative_losses" (thermalconduction.cpp:25,109, radiativelosses.cpp:26,98): the module is evaluated every step -- sub-cycle
 * count, output_to_file planes, cumulative planes -- and nothing is applied to the state. */
int spruce_module_inactive_mode(spruce_domain *dom, const char *module, int on);
int spruce_multispecies_reset(spruce_domain *dom);
int spruce_module_ms_fraction(spruce_domain *dom, const char *module, double ms_electron_heating_fraction);
int spruce_module_output(spruce_domain *dom, const char *plane_name, double *host, size_t count);
/* Pointwise solar source terms applied in postIterateModule (evolution.cpp:74), each followed by propagateChanges.  Gaussian templates
 * (SolarUtils::GaussianGrid / GaussianGridRotated, source/solar/solarutils.cpp:65-98; centres and widths in grid cells, combined with
 * their periodic images as the reference does) are built by the library from the reference's config keys:
 *   AmbientHeatingSink  source/modules/solar/ambientheatingsink.cpp:27-42   reduction = the plane setupModule builds (needs pos_x / pos_y: host)
 *   LocalizedHeating    source/modules/solar/localizedheating.cpp:31-66     active for start_time <= t <= start_time + duration, linear ramps
 *   MassInjection       source/modules/solar/massinjection.cpp:27-52
 *   MomentumInjection   source/modules/solar/momentuminjection.cpp:35-74    incl. the reference's min-combination with the periodic images
 * Written after the round-1 GPU budget was spent: compiled, not yet run on a GPU. */
int spruce_module_ambient_heating_sink(spruce_domain *dom, const double *reduction, size_t count);
int spruce_module_localized_heating(spruce_domain *dom, double start_time, double duration, double max_heating_rate, double stddev_x, double stddev_y,
                                    double center_x, double center_y, double ramp_time);
int spruce_module_mass_injection(spruce_domain *dom, double start_time, double duration, double max_injection_rate, double stddev_x, double stddev_y,
                                 double center_x, double center_y);
int spruce_module_momentum_injection(spruce_domain *dom, double start_time, double duration, double max_accel, double stddev_x, double stddev_y,
                                     double center_x, double center_y, double dir_x, double dir_y, double template_angle, int oscillatory,
                                     double oscillation_period);
/* DivCleaning (source/modules/solar/divcleaning.cpp:21-47): post-iterate, sub-cycled diffusion of div(b) out of bi_x / bi_y; sub-cycle count through
 * spruce_module_subcycles(dom, "div_cleaning", &n).  FieldHeating (source/modules/solar/fieldheating.cpp:30-58): heating from |curl b|, |b|, n and the field
 * line radius of curvature, evaluated before the modules iterate and applied in iterateModule; inactive_mode computes it without applying it.
 * Written after the round-1 GPU budget was spent: compiled, not yet run on a GPU. */
int spruce_module_div_cleaning(spruce_domain *dom, double epsilon, double time_scale);
int spruce_module_field_heating(spruce_domain *dom, double coeff, double current_pow, double b_pow, double n_pow, double roc_pow, int inactive_mode);
/* BoundaryOutflow (source/modules/solar/boundaryoutflow.cpp:33-236): boundary 0..3 = x_bound_1, x_bound_2, y_bound_1, y_bound_2; falloff_shape 0 exp,
 * 1 gaussian, 2 flat.  pos_x / pos_y = the PlasmaDomain position grids of the WHOLE domain (xdim*ydim values, also on a slab): the library builds accel_template (constructBoundaryAccel, :76-137)
 * and the window of computeMeanOutflow (:140-213) from them.  spruce_module_boundary_outflow_state returns mean_outflow / curr_accel of the last step for
 * commandLineMessage (:65-74).  Written after the round-1 GPU budget was spent: compiled, not yet run on a GPU. */
int spruce_module_boundary_outflow(spruce_domain *dom, const double *pos_x, const double *pos_y, size_t count, double max_accel, double falloff_length,
                                   int boundary, int falloff_shape, double feather_length, int field_aligned_mode, int dynamic_mode, double dynamic_time,
                                   double dynamic_target_speed);
int spruce_module_boundary_outflow_state(spruce_domain *dom, double *mean_outflow, double *curr_accel);
/* AnomalousResistivity (source/modules/solar/anomalousresistivity.cpp:18-309): localized magnetic diffusion with Joule heating around the tracked null
 * point, sub-cycled euler / rk2 / rk4 inside iterateModule.  pos_x / pos_y = the PlasmaDomain position grids (xdim*ydim values).  params[SPRUCE_AR_N_PARAMS], the
 * config keys of parseModuleConfigs (:46-68) with the defaults of anomalousresistivity.hpp:16-39:
 *   [0] time_scale  [1] frobenius_metric_coeff  [2] smoothing_sigma  [3] safety_factor  [4] metric_smoothing (0/1)  [5] time_integrator (SPRUCE_TI_*)
 *   [6] template_mode (1 flood_fill, 0 frobenius)  [7] flood_fill_max_radius  [8] flood_fill_argmin_radius  [9] flood_fill_min_current
 *   [10] flood_fill_current_ramp_length  [11] flood_fill_threshold  [12] resistivity_model (0 time_scale, 1 syntelis_19, 2 ys_94)  [13] gradient_correction (0/1)
 *   [14..16] resistivity_model_params
 * setupModule (:18-44: null point = arg-min of the in-plane field over the interior, Gaussian kernel, first template) runs on the state before the first step.
 * The reference's typos are kept (computeDiffusion(bi_x, bi_x, bi_z) :118, curl2D(bi_x, bi_x) :91, the row-count loop bound of the masks :285/:297).
 * One rank only: refused on a slab.  spruce_module_anomalous_resistivity_state returns the tracked null point and the last sub-cycle count.
 * Written after the round-1 GPU budget was spent: the passes are proven on the host (tests/test_anomres_host_check.py), the launches have not run on a GPU. */
#define SPRUCE_AR_N_PARAMS 17
int spruce_module_anomalous_resistivity(spruce_domain *dom, const double *pos_x, const double *pos_y, size_t count, const double *params, int n_params);
int spruce_module_anomalous_resistivity_state(spruce_domain *dom, int *null_i, int *null_j, int *num_subcycles);
/* IdealMHD::parseEquationSetConfigs (source/equationsets/idealmhd.cpp:12-40): global_viscosity (idealmhd.hpp:48, default 0), read only
 * by the characteristic open boundary (global_visc_coeff, idealmhd.cpp:90).  open_moc sides themselves (idealmhd.cpp:306-615) are
 * selected through spruce_config.x_bound_* / y_bound_* = SPRUCE_BC_OPEN_MOC (ideal_mhd only). */
int spruce_eqs_ideal_mhd_options(spruce_domain *dom, double global_viscosity);
/* moc_b_limiting / moc_mom_limiting with moc_{b,mom}_{lower,upper}_lim (idealmhd.cpp:17-36; applyBThresholdingMoC / applyMomThresholdingMoC :107-223): clamps of
 * the ghost layers and the first interior layer of every open_moc side at the head of each derived-variable pass, the reference's y_bound_2 index quirk
 * (:154) included.  Call before spruce_eqs_setup (the setup's own derived pass applies them).  Not available through spruce_mgpu_stage. */
int spruce_eqs_ideal_mhd_moc_limiting(spruce_domain *dom, int b_limiting, double b_lower_lim, double b_upper_lim, int mom_limiting, double mom_lower_lim,
                                      double mom_upper_lim);
/* Ideal2F::parseEquationSetConfigs (source/equationsets/ideal2F.cpp:5-28): use_sub_cycling (reference default true, which aborts
 * in computeTimeDerivatives -- only false can run, and spruce_eqs_setup refuses true) and remove_curl_terms. */
int spruce_eqs_ideal2f_options(spruce_domain *dom, int use_sub_cycling, int remove_curl_terms);
/* EICThermalization (source/modules/ucnp/eic_thermalization.cpp:27-44): electron-ion collisional energy exchange added to the
 * right-hand side of e_thermal_energy / i_thermal_energy.  ideal_2F and ideal_mhd_2E domains (the equation sets that hold the four grids the module
 * looks up by name, :12-25: n, e_temp, e_thermal_energy, i_thermal_energy); on ideal_mhd it fails with the reference's message. */
int spruce_module_eic_thermalization(spruce_domain *dom);
/* curr_num_subcycles of the last advance: which = "thermal_conduction" | "radiative_losses" | "physical_viscosity" | "div_cleaning" | "anomalous_resistivity".
 * With SPRUCE_DEVICE_SUBCYCLES=1 in the environment of spruce_domain_create, steps whose modules are thermal_conduction / radiative_losses / ambient_heating are planned on
 * the device (numberSubcycles, thermalconduction.cpp:135-149, radiativelosses.cpp:161-166, evaluated by one device thread) and spruce_advance(n) waits for the host once,
 * not per step; which = "device_plan" (1 when the configured module set runs that way), "device_plan_budget" (conduction sub-cycles the next advance enqueues per step),
 * "device_plan_replans" (how often an advance had to raise that number and enqueue again). */
int spruce_module_subcycles(spruce_domain *dom, const char *which, int *count);

/* ---- multi-GPU (slab decomposition along x, one process per GPU; DESIGN.md "multi-GPU") */
/* Device pointers + byte counts of the contiguous halo staging buffers of the state that the next
 * right-hand-side evaluation reads: send_lo/send_hi = my first/last 2 rows (all evolved planes packed),
 * recv_lo/recv_hi = where the neighbours' rows must land.  The caller moves them (NCCL send/recv). */
int spruce_halo_buffers(spruce_domain *dom, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes);
/* Stepping split at the exchange points, for callers that own the communicator (bench.py, torch.distributed):
 * phase 0: pack halos of the primary state; then per RK stage s: stage_begin(s) computes stage s and packs the halos
 * of its output; the caller exchanges; stage_end(s) unpacks.  The dt minimum is exchanged with
 * spruce_local_dt_min / spruce_set_global_dt_min (ncclAllReduce(min)). */
/* Peer-store transport (the default): every rank exports one device segment through CUDA IPC (64-byte cudaIpcMemHandle_t),
 * the caller gathers the handles of all ranks (any host-side channel; 64 bytes per rank, rank order) and hands them to
 * spruce_mgpu_ipc_connect.  From then on spruce_advance works on a slab exactly as on a whole domain: after every stage the
 * pack kernel stores the edge rows straight into the ring neighbours' segments over NVLink and publishes a sequence number,
 * the unpack kernel acquires the neighbours' numbers; the dt minimum is all-gathered the same way.  No host synchronisation
 * and no collective-library call inside the time loop.  Call order: create, upload, plane_activity, ipc_export / gather /
 * ipc_connect, eqs_setup, initial_exchange (needs every rank to have connected: barrier before it), advance. */
int spruce_mgpu_ipc_export(spruce_domain *dom, void *handle64);
int spruce_mgpu_ipc_connect(spruce_domain *dom, const void *handles, int n_handles);
int spruce_mgpu_initial_exchange(spruce_domain *dom);
int spruce_mgpu_pack(spruce_domain *dom, int which_state);     /* 0 primary, 1/2 stage copies, 3 static planes (be_*, grav_*) */
int spruce_mgpu_unpack(spruce_domain *dom, int which_state);
int spruce_mgpu_stage(spruce_domain *dom, int stage);
int spruce_mgpu_n_stages(spruce_domain *dom, int *n);
int spruce_mgpu_stage_output(spruce_domain *dom, int stage, int *which_state);   /* set the stage wrote: 0 primary, 1/2 stage copies */
int spruce_mgpu_dt_min_ptr(spruce_domain *dom, void **device_double);   /* 1 double on the device: all-reduce(min) it in place */
int spruce_mgpu_begin_step(spruce_domain *dom);   /* fixes `step` from the (all-reduced) dt minimum */
int spruce_mgpu_end_step(spruce_domain *dom);     /* t += step; iter++ */

/* Planes uploaded as identically zero (mom_z, bi_z, be_x, be_y, be_z: bits 0..4) let the stage kernel skip transports that
 * are exactly zero in the reference as well.  With n_ranks > 1 the knowledge must be global: read the local mask, OR it over
 * the ranks, set it (set_global_mask < 0: only read). */
int spruce_plane_activity(spruce_domain *dom, int *local_mask, int set_global_mask);

/* stream on which every kernel of this handle is launched (cudaStream_t as void*), for event timing */
int spruce_stream(spruce_domain *dom, void **stream);
int spruce_synchronize(spruce_domain *dom);
/* number of kernels this library has launched on the handle since creation (bench.py's gpu_launches) */
int spruce_launch_count(spruce_domain *dom, int64_t *count);
/* Stage kernel only: launches it `reps` times on a scratch copy and returns the mean duration in ms measured with
 * CUDA events on the handle's stream (bench.py roofline leg). */
int spruce_time_stage_kernel(spruce_domain *dom, int reps, float *ms_mean);

#ifdef __cplusplus
}
#endif
#endif /* SPRUCE_B200_H */
