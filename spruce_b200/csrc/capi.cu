// capi.cu -- host side of the C ABI declared in include/spruce_b200.h: device arena, geometry tables,
// plane transfer, and the launch sequences that replace PlasmaDomain::advanceTime (source/mhd/evolution.cpp:59-124).
// Compiled with -fmad=false (see exact_math.cuh).  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/spruce_b200.h"
#include "mhd_kernels.cuh"
#include "module_kernels.cuh"
#include "mhd_stage_xy.cuh"
#include "moc_stage.cuh"
#include "solar_templates.hpp"
#include "mhd2e_cells.cuh"
#include "ideal2f_sides.cuh"
#include "mhd2e_step.hpp"
#include "anomres_cells.hpp"
#include "ideal2f_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>
#include "chunk_plan.hpp"
#include "host_scan.hpp"

using namespace spruce;

// NVTX ranges around what the host enqueues: the advance call, every step, every Runge-Kutta stage with its exchange, every
// module hook.  Header-only NVTX v3: without a profiler attached the calls return at once.  SPRUCE_NVTX=0 leaves them out altogether.
namespace {
const bool g_nvtx_on = [] { const char *e = getenv("SPRUCE_NVTX"); return !(e && atoi(e) == 0); }();
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char *name) : on(g_nvtx_on) { if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
}  // namespace

// stage_relaxed.cu: the stage kernel in relaxed arithmetic (its own translation unit, compiled with FMA contraction)
extern "C" __attribute__((visibility("hidden"))) int spruce_relaxed_launch_stage(unsigned gx, unsigned gy, void *stream, const void *P, const void *A, const void *L, int list, int var);

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(SPRUCE_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CUDA_TRY_D(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { delete d; return fail(SPRUCE_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
#define NOT_2F(d, what) do { if ((d)->tf) return fail(SPRUCE_ERR_UNSUPPORTED, "%s is not available with the ideal_2F equation set", what); \
                             if ((d)->e2) return fail(SPRUCE_ERR_UNSUPPORTED, "%s is not available with the ideal_mhd_2E equation set", what); } while (0)
#define CHECK_DOM(d) do { if (!(d)) return fail(SPRUCE_ERR_ARG, "null domain handle"); } while (0)

namespace {

struct PlaneSet { double *p[NEV] = {nullptr}; };

struct HostAxis {          // 1-D tables of one axis on the host, index -TAB_APRON .. n+TAB_APRON-1 stored at [k+TAB_APRON]
    std::vector<double> h, fs, rfs, ep, em, d, rd;
};

struct TwoFluid;
struct OneFluid2E;

}  // namespace

struct spruce_domain {
    spruce_config cfg{};
    TwoFluid *tf = nullptr;        // ideal_2F state (ideal2f_host.cuh); null for ideal_mhd
    OneFluid2E *e2 = nullptr;      // ideal_mhd_2E state (mhd2e_host.cuh)
    DomainParams P{};
    cudaStream_t stream = nullptr;
    size_t plane_doubles = 0;      // allocation size of one plane incl. halo rows
    size_t row_off = 0;            // offset (doubles) of local row 0 inside an allocation
    std::vector<double *> allocs;  // every cudaMalloc, for destroy
    double *pool = nullptr; size_t pool_planes = 0, pool_used = 0;   // pooled planes of the base arena
    PlaneSet Pset, Mset, M2set, K1set, K2set;
    double *stat[NSTATIC] = {nullptr};
    double *scratch_temp = nullptr, *scratch_out = nullptr;
    double *strip[4] = {nullptr, nullptr, nullptr, nullptr}; int strip_pitch = 0;
    double *tab_dev = nullptr;     // all 1-D tables in one allocation
    StepCtl *ctl = nullptr;
    double *dt_hist = nullptr; size_t dt_hist_cap = 0;
    bool have_geom = false, is_setup = false, raw_rho = true, rk4_alloc = false;
    std::vector<double> dxg, dyg;  // global cell sizes (host copy)
    GhostArgs ghost_proto{};
    int64_t launches = 0;
    bool any_ucnp = false, any_primary_ghost = false;
    // physics modules, in config order (ModuleHandler::instantiateModule, modulehandler.cpp:92-111)
    enum { MOD_TC = 1, MOD_RL = 2, MOD_AH = 3, MOD_AV = 4, MOD_PV = 5 };
    // physical_viscosity (source/modules/solar/physicalviscosity.cpp)
    struct { double coeff = 0.0, epsilon = 1.0; int heating_on = 1, force_on = 1, gc = 0, integrator = 0, inactive = 0, nsub = 1;
             double *cg = nullptr, *v[2][3] = {{nullptr}}, *T[2] = {nullptr}, *bh[3] = {nullptr}; bool cg_halo_done = false;
             double *avg[4] = {nullptr}; bool output = false; } pv;   // avg: output_to_file planes viscous_heating, viscous_force_x/y/z
    std::vector<int> module_order;                 // MOD_* ; MOD_SRC0 + k = sources[k]
    enum { MOD_SRC0 = 100, MOD_DC = 6, MOD_FH = 7, MOD_BO = 8, MOD_AR = 9 };
    struct { double max_accel = 0.0, dynamic_time = 1.0, target = 0.0, mean = 0.0, accel = 0.0; int boundary = 3, field_aligned = 0, dynamic = 0;
             int win[4] = {0, 0, 0, 0}; double *tmpl = nullptr; } bo;                // boundary_outflow
    struct { double epsilon = 0.1, time_scale = 1.0; int nsub = 0; } dc;                            // div_cleaning (divcleaning.hpp:22-23)
    struct { double coeff = 0.0, current_pow = 0.0, b_pow = 0.0, n_pow = 0.0, roc_pow = 0.0; int inactive = 0; double *H = nullptr; } fh;   // field_heating
    // anomalous_resistivity (anomres_cells.hpp / anomres_host.cuh)
    struct { ar::State s{}; bool on = false, ready = false, output = false; double *planes[ar::P_COUNT] = {nullptr}; double *px = nullptr, *py = nullptr, *dtp = nullptr, *kernel = nullptr, *avg = nullptr, *prod = nullptr;
             std::vector<double> hpx, hpy; } ar;
    // pointwise solar source terms (module_kernels.cuh: k_source_term), in config order
    struct SourceTerm { int kind = 0; double start = 0.0, duration = 0.0, ramp_time = 0.0, max_accel = 0.0, period = 1.0; int oscillatory = 0; double *plane[2] = {nullptr, nullptr};
                        double ms_fraction = 0.0; };      // ms_electron_heating_fraction: 0.5 for the sink, 0.0 for localized_heating (ambientheatingsink.hpp:28, localizedheating.hpp:28)
    // multispecies_mode (plasmadomain.hpp:134-135): cumulative electron / ion / joule heating between outputs; the modules' electron fractions with the reference's defaults
    bool ms_on = false; double *ms_cum[3] = {nullptr, nullptr, nullptr};
    double ms_frac_tc = 1.0, ms_frac_rl = 1.0, ms_frac_ah = 0.5, ms_frac_pv = 0.0;
    std::vector<SourceTerm> sources;
    TcParams tc{}; int tc_integrator = 0; double tc_epsilon = 0.0; int tc_nsub = 0; bool tc_planes_valid = false;
    bool tc_two_pass = true; double *tc_pass[3] = {nullptr, nullptr, nullptr};      // SPRUCE_TC_TWO_PASS=0: saturated conduction evaluates the coefficient at five points per cell (k_tc_coef off)
    RlParams rl{}; int rl_nsub = 0;
    double *heating = nullptr;
    // diagnostic planes of output_to_file = true (thermalconduction.cpp:226-237, radiativelosses.cpp:172-179)
    bool tc_output = false, rl_output = false;
    bool tc_inactive = false, rl_inactive = false;      // inactive_mode (thermalconduction.cpp:109, radiativelosses.cpp:98): evaluated for the output / cumulative planes, not applied
    double *tc_avg = nullptr, *tc_sat = nullptr, *rl_avg = nullptr, *old_e = nullptr;
    // artificial_viscosity (source/modules/viscosity.cpp): terms in config order
    struct ViscTerm { int opt; double strength; int var_diff; int var_evol; int species; double *strength_plane; bool halo_done = false;
                      double *o_dqdt = nullptr, *o_lap = nullptr, *o_dt = nullptr; };      // output planes (spruce_module_output_to_file "artificial_viscosity")
    std::vector<ViscTerm> visc;
    bool visc_output = false; bool visc_evaluated = false;   // visc_evaluated: some term has been evaluated (the reference's strength planes are zero until then)
    int visc_hv_integrator = 0, visc_gradient_correction = 0; double visc_hv_epsilon = 1.0;
    double *vscratch[8] = {nullptr}; double *dt_plane = nullptr;
    const double *cur_xterm[4] = {nullptr}; int cur_xtarget[4] = {0}; int cur_nx = 0;
    unsigned long long *red = nullptr;     // 4 reduction scalars for the sub-cycle counts
    double *halo[4] = {nullptr, nullptr, nullptr, nullptr};   // send_lo, send_hi, recv_lo, recv_hi
    // planes that were uploaded as identically zero: bit 0 mom_z, 1 bi_z, 2 be_x, 3 be_y, 4 be_z (global knowledge; see spruce_plane_activity);
    // bits 5, 6: grav_x, grav_y (own-cell values only: local knowledge is enough)
    unsigned nonzero_mask = 0x7F;
    bool in_mgpu_stage_api = false;        // inside spruce_mgpu_stage (caller-owned exchange and dt reduction)
    bool static_lists = true;              // use the fully unrolled instances of k_mhd_stage_xy when the active list matches one
    // open_moc (moc_stage.cuh): evolved ghost cells
    bool moc_any = false; double global_viscosity = 0.0; double *moc_base = nullptr;
    moc::Limits moc_lim{0, 0, 0.1, 10.0, 0.1, 10.0};                      // moc_b_limiting / moc_mom_limiting and their bounds (idealmhd.hpp:59-64)
    int chunk_rows_override = 0;           // SPRUCE_CHUNK_ROWS (8 .. XY_CHUNK): rows per CTA of the stage kernel, for tuning sweeps; 0 = plan_chunks' own choice
    bool relaxed = false;                  // SPRUCE_ARITH=relaxed: stage kernel from stage_relaxed.cu (FMA contraction, one-multiplication table divisions)
    int stage_variants = 1;                // compile-time integrator-stage instances of k_mhd_stage_xy (SPRUCE_STAGE_VARIANTS=0: only the run-time-stage instances)
    bool vec_rows = true;                 // k_mhd_stage_xy copies ring rows in 16-byte chunks where a strip allows it (SPRUCE_VEC_ROWS=0: 8-byte per-column copies everywhere)
    // SPRUCE_TIMELINE=n: CUDA events around the launches of the first n plain steps of an advance call, printed (ms since the first one) when the
    // call returns -- the per-step timeline of a slab rank without a system profiler
    struct Mark { const char *label; cudaEvent_t ev; };
    std::vector<Mark> marks; int timeline_steps = 0; bool timeline_on = false;
    bool fused_ctl = false;                // inside a plain step (no modules, no open_moc): the step control runs in k_step_open / k_step_mid / k_step_close
    bool fuse_ctl_enabled = true;          // SPRUCE_FUSED_CTL=0: always the separate one-thread control kernels
    // SPRUCE_DEVICE_SUBCYCLES=1: the sub-cycle counts of thermal_conduction / radiative_losses stay on the device (module_kernels.cuh: SubPlan); the host enqueues
    // tc_budget conduction sub-cycles per step (SPRUCE_TC_BUDGET; adapted after every advance) and no step waits for the host
    bool dev_sub = false; int tc_budget = 8; SubPlan *plan = nullptr; int replans = 0;
    bool fast_interior = true;             // SPRUCE_FAST_INTERIOR=0: the module stencil kernels use their general (wrapping, range-testing) instance for every cell
    size_t halo_doubles = 0;
    // peer-store transport (CUDA IPC segment: PeerFlags + 2 sides x 2 parities of packed halo rows; mhd_kernels.cuh)
    void *seg = nullptr; size_t seg_bytes = 0;
    void *peer_seg[MAX_RANKS] = {nullptr};      // mapped segments of the other ranks (own rank: seg)
    bool peers_connected = false;
    unsigned long long halo_seq = 0, dt_seq = 0, red_seq = 0;
    unsigned int *push_counter = nullptr;
    cudaStream_t comm_stream = nullptr;          // high priority: edge chunks + push + pull, concurrent with the interior chunks
    cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
    bool overlap = true;
};

namespace {

void mark(spruce_domain *d, const char *label, cudaStream_t st)
{
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    d->marks.push_back({label, e});
}
void dump_marks(spruce_domain *d)
{
    if (d->marks.empty()) return;
    cudaDeviceSynchronize();
    for (size_t k = 0; k < d->marks.size(); k++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, d->marks[0].ev, d->marks[k].ev);
        fprintf(stderr, "[spruce timeline rank %d] %9.4f ms  %s\n", d->cfg.rank, ms, d->marks[k].label);
    }
    for (auto &m : d->marks) cudaEventDestroy(m.ev);
    d->marks.clear();
}
#define MARK(d, label, st) do { if ((d)->timeline_on) mark((d), (label), (st)); } while (0)

// planes are carved from one pooled allocation made at create time (one cudaMalloc instead of ~25); later requests beyond the
// pool (RK4 sets, module work planes) get their own allocation
int alloc_plane(spruce_domain *d, double **out)
{
    double *base = nullptr;
    if (d->pool && d->pool_used < d->pool_planes) {
        base = d->pool + d->pool_used * d->plane_doubles;
        d->pool_used++;
    } else {
        CUDA_TRY(cudaMalloc(&base, d->plane_doubles * sizeof(double)));
        d->allocs.push_back(base);
    }
    CUDA_TRY(cudaMemsetAsync(base, 0, d->plane_doubles * sizeof(double), d->stream));
    *out = base + d->row_off;
    return SPRUCE_OK;
}
int alloc_set(spruce_domain *d, PlaneSet &s)
{
    for (int v = 0; v < NEV; v++) { int rc = alloc_plane(d, &s.p[v]); if (rc) return rc; }
    return SPRUCE_OK;
}

// computeIterationBounds, source/mhd/plasmadomain.cpp:138-161
void iteration_bounds(const spruce_config &c, DomainParams &P)
{
    P.xl = (c.x_bound_1 == SPRUCE_BC_PERIODIC) ? 0 : HALO;
    P.xu = (c.x_bound_2 == SPRUCE_BC_PERIODIC) ? c.xdim - 1 : c.xdim - HALO - 1;
    P.yl = (c.y_bound_1 == SPRUCE_BC_PERIODIC) ? 0 : HALO;
    P.yu = (c.y_bound_2 == SPRUCE_BC_PERIODIC) ? c.ydim - 1 : c.ydim - HALO - 1;
}

// 1-D tables for indices [lo-TAB_APRON, lo+n+TAB_APRON) of a global axis of length gn (periodic wrap when per)
void build_axis(const std::vector<double> &dg, int gn, bool per, int lo, int n, HostAxis &A)
{
    const int m = n + 2 * TAB_APRON;
    auto dval = [&](int g) -> double {          // cell size at global index g
        if (per) { g = ((g % gn) + gn) % gn; return dg[g]; }
        if (g < 0 || g >= gn) return 1.0;       // never used by an in-range operator
        return dg[g];
    };
    auto hval = [&](int g) { return 0.5 * dval(g); };
    A.h.resize(m); A.fs.resize(m); A.rfs.resize(m); A.ep.resize(m); A.em.resize(m); A.d.resize(m); A.rd.resize(m);
    for (int k = 0; k < m; k++) {
        const int g = lo + k - TAB_APRON;
        A.d[k] = dval(g);
        A.rd[k] = 1.0 / A.d[k];
        A.h[k] = hval(g);
        A.fs[k] = hval(g) + hval(g - 1);
        A.rfs[k] = 1.0 / A.fs[k];
        A.ep[k] = hval(g - 2) + 2.0 * hval(g - 1);
        A.em[k] = hval(g + 1) + 2.0 * hval(g);
    }
}

int upload_tables(spruce_domain *d)
{
    const spruce_config &c = d->cfg;
    HostAxis X, Y;
    build_axis(d->dxg, c.xdim, d->P.xper, c.row0, c.nx_local, X);
    build_axis(d->dyg, c.ydim, d->P.yper, 0, c.ydim, Y);
    const size_t mx = X.h.size(), my = Y.h.size();
    std::vector<double> all;
    all.reserve(7 * (mx + my));
    const std::vector<double> *xs[7] = {&X.h, &X.fs, &X.rfs, &X.ep, &X.em, &X.d, &X.rd};
    const std::vector<double> *ys[7] = {&Y.h, &Y.fs, &Y.rfs, &Y.ep, &Y.em, &Y.d, &Y.rd};
    for (auto v : xs) all.insert(all.end(), v->begin(), v->end());
    for (auto v : ys) all.insert(all.end(), v->begin(), v->end());
    if (!d->tab_dev) { CUDA_TRY(cudaMalloc(&d->tab_dev, all.size() * sizeof(double))); }
    CUDA_TRY(cudaMemcpyAsync(d->tab_dev, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    const double *b = d->tab_dev + TAB_APRON;
    AxisTab &tx = d->P.tx, &ty = d->P.ty;
    tx.h = b; tx.fs = b + mx; tx.rfs = b + 2 * mx; tx.ep = b + 3 * mx; tx.em = b + 4 * mx; tx.d = b + 5 * mx; tx.rd = b + 6 * mx;
    b += 7 * mx;
    ty.h = b; ty.fs = b + my; ty.rfs = b + 2 * my; ty.ep = b + 3 * my; ty.em = b + 4 * my; ty.d = b + 5 * my; ty.rd = b + 6 * my;
    return SPRUCE_OK;
}

// open-boundary scalars, evolution.cpp:163-167 (host libm pow, exactly as the reference evaluates them)
void build_ghost_proto(spruce_domain *d)
{
    const spruce_config &c = d->cfg;
    GhostArgs &G = d->ghost_proto;
    memset(&G, 0, sizeof(G));
    G.open_strength = c.open_boundary_strength;
    const int bcs[4] = {c.x_bound_1, c.x_bound_2, c.y_bound_1, c.y_bound_2};
    for (int side = 0; side < 4; side++) {
        const std::vector<double> &dd = side < 2 ? d->dxg : d->dyg;
        const int n = side < 2 ? c.xdim : c.ydim;
        const bool lower = (side == 0 || side == 2);
        const int i1 = lower ? 0 : n - 1, i2 = lower ? 1 : n - 2, i3 = lower ? 2 : n - 3;
        const double delta_last = dd[i3];
        const double dist23 = 0.5 * (dd[i2] + dd[i3]);
        const double dist12 = 0.5 * (dd[i1] + dd[i2]);
        G.scale_2[side] = std::pow(c.open_boundary_decay_base, dist23 / delta_last);
        G.scale_1[side] = std::pow(c.open_boundary_decay_base, dist12 / delta_last);
        G.dist23[side] = dist23;
        G.h2[side] = 0.5 * dd[i2];
        G.h3[side] = 0.5 * dd[i3];
        G.rh3[side] = 1.0 / G.h3[side];
        if (bcs[side] == SPRUCE_BC_OPEN_UCNP) d->any_ucnp = true;
        if (bcs[side] == SPRUCE_BC_OPEN || bcs[side] == SPRUCE_BC_REFLECT) d->any_primary_ghost = true;
    }
}

void fill_sets(const spruce_domain *d, StageArgs &A, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D)
{
    for (int v = 0; v < NEV; v++) { A.S[v] = S.p[v]; A.B[v] = B.p[v]; A.D[v] = D.p[v]; A.K1[v] = d->K1set.p[v]; A.K2[v] = d->K2set.p[v]; }
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    for (int s = 0; s < 4; s++) A.strip[s] = d->strip[s];
    A.strip_pitch = d->strip_pitch;
    A.step_ptr = &d->ctl->step;
    A.done_ptr = &d->ctl->done;
    A.dtmin_bits = &d->ctl->dtmin_bits;
    A.inv_thr_ptr = &d->ctl->inv_thr;
}

// the dt skip test (dt_can_skip) lives in k_mhd_stage_xy only, and needs k_dt_validate / k_dt_full after the step's last stage:
// spruce_advance provides that; the caller-driven spruce_mgpu_stage path (NCCL transport) evaluates every cell
// open_moc: the minimum also runs over evolved ghost cells, which k_dt_full does not visit -> every cell is evaluated
int dt_prune_enabled(const spruce_domain *d) { return (!d->in_mgpu_stage_api && !d->moc_any) ? 1 : 0; }

ActiveList active_quantities(const spruce_domain *d);

// Row chunks of one stage launch.  A CTA marches over `rows` rows (plus 4 warm-up rows and a prologue), so few, long chunks are cheap, but the grid
// should fill whole waves of resident CTAs (148 SMs x 5 for the 2-D instance, x 4 otherwise).  The plan takes the smallest number of waves the upper
// bound XY_CHUNK allows and spreads the rows evenly over the CTA rows that fit in them.  On a slab whose halo exchange overlaps the interior
// (split), the first and the last `edge` rows form their own short launch on the communication stream: it finishes early, its rows travel while
// the long interior chunks still run.
ChunkPlan plan_chunks(const spruce_domain *d, bool split)
{
    const int strips = (d->P.ny + CW - 1) / CW;
    const ActiveList L = active_quantities(d);
    const int cap = 148 * ((L.n == 6 && L.q == XY_LIST_2D && d->static_lists) ? xy_ctas_per_sm(6) : xy_ctas_per_sm(0));
    return plan_chunk_rows(d->P.nx, strips, cap, split, d->chunk_rows_override);      // chunk_plan.hpp
}
bool can_split(const spruce_domain *d, int primary)
{
    const bool ghosts = d->any_ucnp || (primary && d->any_primary_ghost);
    return d->cfg.n_ranks > 1 && d->peers_connected && d->overlap && !ghosts && d->visc.empty() && !d->moc_any      // the strip kernel follows the whole stage kernel and writes edge rows
           && d->P.nx >= 3 * XY_CHUNK;
}

// Transported quantities that can be non-zero.  The z system {mom_z, bi_z} stays identically zero when mom_z, bi_z and be_z
// are zero planes (idealmhd.cpp:72-73,84-86), and a zero external-field plane is static.
ActiveList active_quantities(const spruce_domain *d)
{
    ActiveList L{};
    const unsigned m = d->nonzero_mask;
    const bool zsys = (m & 0x13u) != 0;           // mom_z | bi_z | be_z
    int n = 0, last = 0;
    auto add = [&](int q) { L.q |= (unsigned long long)q << (4 * n); n++; last = q; };
    add(Q_RHO); add(Q_E); add(Q_MX); add(Q_MY); add(Q_BIX); add(Q_BIY);
    if (zsys) { add(Q_MZ); add(Q_BIZ); }
    if (m & 0x04u) add(Q_BEX);
    if (m & 0x08u) add(Q_BEY);
    if (m & 0x10u) add(Q_BEZ);
    if (n & 1) add(last);                         // pad to an even count: the duplicate recomputes the same values
    L.n = n;
    return L;
}

int prepare_rhs_modules(spruce_domain *d, const PlaneSet &S);
int launch_moc(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode, int dt_only);
int reset_reductions(spruce_domain *d);
int peer_red_allgather(spruce_domain *d);
int launch_moc_save(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D);
int launch_stage(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode, int part = 0)
{
    // part 0: every row chunk on the main stream; 1: the first and last chunk on the communication stream; 2: the chunks between on the main stream
    if (part == 0 && !d->visc.empty()) { int rcv = prepare_rhs_modules(d, S); if (rcv) return rcv; }
    if (d->moc_any && kmode != KM_EXPORT) { int rcm = launch_moc_save(d, S, B, D); if (rcm) return rcm; }
    StageArgs A{};
    fill_sets(d, A, S, B, D);
    A.n_xterm = d->visc.empty() ? 0 : d->cur_nx;
    for (int t = 0; t < A.n_xterm; t++) { A.xterm[t] = d->cur_xterm[t]; A.xtarget[t] = d->cur_xtarget[t]; }
    A.coef = coef; A.primary = primary; A.kmode = kmode;
    A.b_is_s = (S.p[0] == B.p[0]) ? 1 : 0;
    A.vec16 = d->vec_rows ? 1 : 0;
    A.grav = (d->nonzero_mask & 0x60u) ? 1 : 0;
    A.walls = (d->cfg.x_bound_1 != SPRUCE_BC_PERIODIC || d->cfg.x_bound_2 != SPRUCE_BC_PERIODIC || d->cfg.y_bound_1 != SPRUCE_BC_PERIODIC || d->cfg.y_bound_2 != SPRUCE_BC_PERIODIC) ? 1 : 0;
    if (part == 0 && primary && kmode != KM_EXPORT && !d->fused_ctl) { k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, dt_prune_enabled(d)); d->launches++; }
    const ChunkPlan cp = plan_chunks(d, part != 0);
    const StageRows sr = stage_rows(cp, d->P.nx, part);
    A.chunk_rows = sr.chunk_rows; A.row_begin = sr.row_begin; A.row_end = sr.row_end; A.edge2_begin = sr.edge2_begin; A.edge2_end = sr.edge2_end;
    const int gy = sr.grid_y;
    cudaStream_t st = part == 1 ? d->comm_stream : d->stream;
    dim3 grid((d->P.ny + CW - 1) / CW, gy);
    {
        const ActiveList L = active_quantities(d);
        // compile-time integrator stage (SPRUCE_STAGE_VARIANTS=0 turns it off): plain euler / rk2 stages without module terms
        int var = 0;
        if (d->stage_variants && kmode == KM_NONE && A.n_xterm == 0) var = A.b_is_s ? (primary ? 3 : 1) : (primary ? 2 : 0);
        const bool list2d = L.n == 6 && L.q == XY_LIST_2D, listfull = L.n == 12 && L.q == XY_LIST_FULL;
        if (d->relaxed && d->static_lists && (list2d || listfull)) {          // any other list runs the exact run-time-list kernel below
            const int e = spruce_relaxed_launch_stage(grid.x, grid.y, (void *)st, &d->P, &A, &L, list2d ? 6 : 12, var);
            if (e != 0) return fail(SPRUCE_ERR_CUDA, "relaxed stage kernel: %s", cudaGetErrorString((cudaError_t)e));
        } else
        if (d->static_lists && list2d) {
            if (var == 1) k_mhd_stage_xy<6, XY_LIST_2D, 1><<<grid, XY_NT, xy_smem_bytes(6, 1), st>>>(d->P, A, L);
            else if (var == 2) k_mhd_stage_xy<6, XY_LIST_2D, 2><<<grid, XY_NT, xy_smem_bytes(6, 2), st>>>(d->P, A, L);
            else if (var == 3) k_mhd_stage_xy<6, XY_LIST_2D, 3><<<grid, XY_NT, xy_smem_bytes(6, 3), st>>>(d->P, A, L);
            else k_mhd_stage_xy<6, XY_LIST_2D><<<grid, XY_NT, xy_smem_bytes(6, 0), st>>>(d->P, A, L);
        } else if (d->static_lists && listfull) {
            if (var == 1) k_mhd_stage_xy<12, XY_LIST_FULL, 1><<<grid, XY_NT, xy_smem_bytes(NTR, 1), st>>>(d->P, A, L);
            else if (var == 2) k_mhd_stage_xy<12, XY_LIST_FULL, 2><<<grid, XY_NT, xy_smem_bytes(NTR, 2), st>>>(d->P, A, L);
            else if (var == 3) k_mhd_stage_xy<12, XY_LIST_FULL, 3><<<grid, XY_NT, xy_smem_bytes(NTR, 3), st>>>(d->P, A, L);
            else k_mhd_stage_xy<12, XY_LIST_FULL><<<grid, XY_NT, xy_smem_bytes(NTR, 0), st>>>(d->P, A, L);
        }
        else k_mhd_stage_xy<0, 0ULL><<<grid, XY_NT, xy_smem_bytes(NTR, 0), st>>>(d->P, A, L);
    }
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    if (d->moc_any) return launch_moc(d, S, B, D, coef, primary, kmode, 0);      // single rank: part == 0, same stream
    return SPRUCE_OK;
}

// open_moc sides: the characteristic update of their ghost cells, after the stage kernel of the same stage (moc_stage.cuh);
// dt_only: after a propagateChanges, the dt of those cells of the primary state
void fill_moc(spruce_domain *d, MocArgs &A, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode, int dt_only)
{
    for (int v = 0; v < NEV; v++) { A.S[v] = S.p[v]; A.B[v] = B.p[v]; A.D[v] = D.p[v]; A.K1[v] = d->K1set.p[v]; A.K2[v] = d->K2set.p[v]; }
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.kmode = kmode; A.primary = primary; A.dt_only = dt_only; A.coef = coef;
    A.step_ptr = &d->ctl->step; A.done_ptr = &d->ctl->done; A.dtmin_bits = &d->ctl->dtmin_bits;
    A.global_viscosity = d->global_viscosity;
    A.base_buf = d->moc_base;
}
// before the stage kernel of a stage: the base state of the evolved ghost cells (see MocArgs::base_buf)
int launch_moc_save(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D)
{
    if (!d->moc_base) CUDA_TRY(cudaMalloc(&d->moc_base, (size_t)NEV * moc_threads(d->P) * sizeof(double)));
    MocArgs A{};
    fill_moc(d, A, S, B, D, 0.0, 0, KM_NONE, 0);
    k_moc_save<<<(moc_threads(d->P) + 127) / 128, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int launch_moc(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode, int dt_only)
{
    if (!d->moc_any) return SPRUCE_OK;
    MocArgs A{};
    fill_moc(d, A, S, B, D, coef, primary, kmode, dt_only);
    if (!dt_only && d->global_viscosity != 0.0) {             // global_visc_coeff of this right-hand-side evaluation (idealmhd.cpp:90)
        // the minimum goes through the module reduction scalars (slot 0 = a minimum): on slabs they are all-gathered over the peer segments
        if (d->cfg.n_ranks > 1 && !d->peers_connected) return fail(SPRUCE_ERR_UNSUPPORTED, "open_moc with global_viscosity != 0 on slabs needs the peer transport (spruce_mgpu_ipc_connect)");
        int rc = reset_reductions(d);
        if (rc) return rc;
        dim3 grid((d->P.ny + 127) / 128, d->P.nx);
        k_moc_visc_min<<<grid, 128, 0, d->stream>>>(d->P, A, d->red);
        d->launches++;
        if (d->cfg.n_ranks > 1 && (rc = peer_red_allgather(d))) return rc;
        A.visc_min_bits = d->red;
    }
    k_moc_stage<<<(moc_threads(d->P) + 127) / 128, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

// moc_b_limiting / moc_mom_limiting (idealmhd.cpp:107-223): the head of the derived-variable pass, i.e. after the boundary passes of a propagate.
// The clamps reach the first interior layer, whose dt the stage kernel has already folded into the minimum: in the step's last stage the minimum
// is rebuilt from every cell of U (interior by k_dt_full, evolved ghost cells by k_moc_stage in dt-only mode).
int moc_limit(spruce_domain *d, const PlaneSet &U, int primary)
{
    if (!d->moc_any || !(d->moc_lim.b_on || d->moc_lim.mom_on)) return SPRUCE_OK;
    MocLimitArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = U.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.L = d->moc_lim; A.done_ptr = &d->ctl->done;
    const int bcs[4] = {d->cfg.x_bound_1, d->cfg.x_bound_2, d->cfg.y_bound_1, d->cfg.y_bound_2};
    for (int s = 0; s < 4; s++) {
        if (bcs[s] != SPRUCE_BC_OPEN_MOC) continue;
        A.side = s;
        const int n = s < 2 ? d->P.ny : d->P.nx;
        k_moc_limit<<<(n + 127) / 128, 128, 0, d->stream>>>(d->P, A);
        d->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    if (!primary) return SPRUCE_OK;
    k_moc_force_full_dt<<<1, 1, 0, d->stream>>>(d->ctl);
    DtFullArgs F{};
    for (int v = 0; v < NEV; v++) F.U[v] = U.p[v];
    for (int v = 0; v < NSTATIC; v++) F.st[v] = d->stat[v];
    F.ctl = d->ctl;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_dt_full<<<grid, 256, 0, d->stream>>>(d->P, F);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return launch_moc(d, U, U, U, 0.0, 1, KM_NONE, 1);
}

int launch_ghosts(spruce_domain *d, const PlaneSet &U, int primary)
{
    if (!(d->any_ucnp || (primary && d->any_primary_ghost))) return SPRUCE_OK;
    GhostArgs G = d->ghost_proto;
    for (int v = 0; v < NEV; v++) G.U[v] = U.p[v];
    for (int s = 0; s < 4; s++) G.strip[s] = d->strip[s];
    G.strip_pitch = d->strip_pitch;
    G.primary = primary;
    const int n = d->P.ny > d->P.nx ? d->P.ny : d->P.nx;
    dim3 grid((n + 127) / 128, 4);
    k_mhd_ghosts<<<grid, 128, 0, d->stream>>>(d->P, G);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

int launch_propagate(spruce_domain *d, int from_state)
{
    k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, 0);
    PropArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    for (int s = 0; s < 4; s++) A.strip[s] = d->strip[s];
    A.strip_pitch = d->strip_pitch;
    A.temp = d->scratch_temp; A.raw_rho = d->raw_rho ? 1 : 0; A.from_state = from_state;
    A.dtmin_bits = &d->ctl->dtmin_bits;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_mhd_propagate<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    d->raw_rho = false;
    int rc = launch_ghosts(d, d->Pset, 1);
    if (rc || !d->moc_any) return rc;
    if (d->moc_lim.b_on || d->moc_lim.mom_on) return moc_limit(d, d->Pset, 1);       // clamps, then the whole dt minimum again
    return launch_moc(d, d->Pset, d->Pset, d->Pset, 0.0, 1, KM_NONE, 1);
}

int ensure_rk4(spruce_domain *d)
{
    if (d->rk4_alloc) return SPRUCE_OK;
    int rc;
    if ((rc = alloc_set(d, d->M2set))) return rc;
    if ((rc = alloc_set(d, d->K1set))) return rc;
    if ((rc = alloc_set(d, d->K2set))) return rc;
    d->rk4_alloc = true;
    return SPRUCE_OK;
}

int derive_to(spruce_domain *d, int var, double *out, const PlaneSet *set = nullptr)
{
    DeriveArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = (set ? set : &d->Pset)->p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.out = out; A.which = var;
    // a slab also fills its halo rows (the neighbours' cells are resident there), so that stencils on derived planes need no exchange
    const int halo = d->cfg.n_ranks > 1 ? HALO : 0;
    A.row_off = -halo;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx + 2 * halo);
    k_mhd_derive<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

int peer_red_allgather(spruce_domain *d);
int read_reductions(spruce_domain *d, unsigned long long h[4])
{
    if (d->cfg.n_ranks > 1) { int rc = peer_red_allgather(d); if (rc) return rc; }     // min / max over all slabs (exact: same counts as one GPU)
    CUDA_TRY(cudaMemcpyAsync(h, d->red, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return SPRUCE_OK;
}
int reset_reductions(spruce_domain *d)
{
    const unsigned long long init[4] = {0x7FEFFFFFFFFFFFFFULL, 0ULL, 0x7FEFFFFFFFFFFFFFULL, 0ULL};
    CUDA_TRY(cudaMemcpyAsync(d->red, init, sizeof(init), cudaMemcpyHostToDevice, d->stream));
    return SPRUCE_OK;
}
double bits_to_double(unsigned long long b) { double v; memcpy(&v, &b, sizeof(v)); return v; }

int peer_exchange(spruce_domain *d, double *const *U, cudaStream_t st);
int peer_dt_allgather(spruce_domain *d);
// slab decomposition: halo rows of one plane (a module's working plane) from the ring neighbours
int exchange_plane(spruce_domain *d, double *plane)
{
    if (d->cfg.n_ranks == 1) return SPRUCE_OK;
    double *v[NEV];
    for (int k = 0; k < NEV; k++) v[k] = plane;
    return peer_exchange(d, v, nullptr);
}
// after a module's closing propagateChanges: halo rows of the primary state and the global dt minimum
int after_module_propagate(spruce_domain *d)
{
    if (d->cfg.n_ranks == 1) return SPRUCE_OK;
    int rc = peer_exchange(d, d->Pset.p, nullptr);
    if (rc) return rc;
    return peer_dt_allgather(d);
}

// ThermalConduction::numberSubcycles (thermalconduction.cpp:135-149); scratch: Mset planes 0..2 (temp, b_hat_x, b_hat_y)
// one feed of the cumulative planes of multispecies_mode (k_ms_feed); dt comes from the device step control when the form needs it
int ms_feed(spruce_domain *d, int mode, const double *a, const double *b, double f, double sign = 1.0)
{
    if (!d->ms_on) return SPRUCE_OK;
    MsArgs A{};
    A.cum_i = mode == MS_JOULE ? d->ms_cum[2] : d->ms_cum[1]; A.cum_e = d->ms_cum[0];
    A.a = a; A.b = b; A.f = f; A.sign = sign; A.mode = mode; A.dt_ptr = &d->ctl->step; A.done_ptr = &d->ctl->done;
    const dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_ms_feed<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// temp, b_hat_x, b_hat_y of the primary state into Mset planes 0..2, one pass (k_tc_derive)
int tc_derive(spruce_domain *d)
{
    TcDeriveArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.T = d->Mset.p[0]; A.bhx = d->Mset.p[1]; A.bhy = d->Mset.p[2];
    const int halo = d->cfg.n_ranks > 1 ? HALO : 0;                     // as derive_to: a slab also fills its halo rows
    A.row_off = -halo;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx + 2 * halo);
    k_tc_derive<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// the planes tc_derive left are still those of the state tc_iterate starts from: thermal_conduction is the first module of the run (its iterate hook runs first) and no
// field_heating pre-iterate hook has used the same scratch planes in between (the other pre-iterate hooks only reduce)
bool tc_planes_reusable(const spruce_domain *d)
{
    if (d->module_order.empty() || d->module_order.front() != spruce_domain::MOD_TC) return false;
    for (int m : d->module_order) if (m == spruce_domain::MOD_FH) return false;
    return true;
}
int tc_count_launch(spruce_domain *d)
{
    int rc;
    if ((rc = tc_derive(d))) return rc;
    d->tc_planes_valid = tc_planes_reusable(d);
    TcFields F{d->Mset.p[0], d->Pset.p[E_N], d->Mset.p[1], d->Mset.p[2]};
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_tc_count<<<grid, 128, 0, d->stream>>>(d->P, d->tc, F, d->red);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int tc_count(spruce_domain *d, double dt, int *nsub)
{
    int rc;
    if ((rc = reset_reductions(d))) return rc;
    if ((rc = tc_count_launch(d))) return rc;
    unsigned long long h[4];
    if ((rc = read_reductions(d, h))) return rc;
    if (d->tc.flux_saturation && bits_to_double(h[1]) == 0.0) { *nsub = 0; return SPRUCE_OK; }   // :143
    const double a = d->tc_epsilon * bits_to_double(h[0]), b = d->tc.dt_subcycle_min;
    const double md = (a < b) ? b : a;                                                              // std::max :147
    *nsub = (int)(dt / md) + 1;                                                                     // :148
    return SPRUCE_OK;
}

// ThermalConduction::iterateModule (thermalconduction.cpp:47-112); scratch = the 8 planes of Mset (free between RK steps)
// dev: the device-resident plan decides (SubPlan): tc_budget sub-cycles are enqueued, the stage kernels take count and sub-cycle step from the plan, dt is not used
int tc_iterate(spruce_domain *d, double dt, bool dev = false)
{
    int rc;
    double *Ta = d->Mset.p[0], *bhx = d->Mset.p[1], *bhy = d->Mset.p[2], *Tb = d->Mset.p[3], *Tc = d->Mset.p[4];
    double *K1 = d->Mset.p[5], *K2 = d->Mset.p[6], *K3 = d->Mset.p[7];
    if (!d->tc_planes_valid && (rc = tc_derive(d))) return rc;          // Ta, bhx, bhy (unless the count kernel's planes still stand)
    d->tc_planes_valid = false;
    const int ns = dev ? d->tc_budget : d->tc_nsub;
    const double dts = dev ? 0.0 : dt / (double)ns;                                                 // :60
    // inactive_mode: the sub-cycles run on a copy of the thermal energy (old_e) and only the output / cumulative planes keep their result (:101-109)
    const bool inactive = d->tc_inactive;
    if (inactive && !d->tc_output && !d->ms_on) return SPRUCE_OK;                                   // nothing would be kept
    double *const e_primary = d->Pset.p[E_E];
    double *e = inactive ? d->old_e : e_primary;
    const double *e_before = inactive ? e_primary : d->old_e;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    const dim3 grid256((d->P.ny + 255) / 256, d->P.nx);
    if (d->tc_output || d->ms_on) {                                                                 // :53-59
        k_plane_copy<<<grid256, 256, 0, d->stream>>>(d->P, d->old_e, e_primary);
        d->launches++;
    }
    if (d->tc_output) {
        if (d->tc.flux_saturation) {
            k_tc_saturation_plane<<<grid, 128, 0, d->stream>>>(d->P, d->tc, TcFields{Ta, d->Pset.p[E_N], bhx, bhy}, d->tc_sat, dev ? &d->ctl->done : nullptr);
            d->launches++;
        }
    }
    int sub = 0;
    const bool two_pass = d->tc.flux_saturation && d->tc_two_pass;
    if (two_pass) for (int k = 0; k < 3; k++) if (!d->tc_pass[k] && (rc = alloc_plane(d, &d->tc_pass[k]))) return rc;
    auto stage = [&](const double *Tin, double *Tout, int mode, int half, double *Kst) -> int {
        TcStageArgs A{};
        A.F = TcFields{Tin, d->Pset.p[E_N], bhx, bhy};
        if (two_pass) {                                  // saturation coefficient and raw flux of this temperature plane, once per cell (k_tc_coef)
            TcCoefArgs K{};
            K.F = A.F; K.C = d->tc; K.coef = d->tc_pass[0]; K.rx = d->tc_pass[1]; K.ry = d->tc_pass[2]; K.fast = d->fast_interior ? 1 : 0;
            const int halo1 = d->cfg.n_ranks > 1 ? 1 : 0;                // a slab differentiates the coefficient across its edges: one halo row per side
            K.row_off = -halo1;
            if (dev) { K.plan = d->plan; K.sub = sub; }
            const dim3 gridk((d->P.ny + 127) / 128, d->P.nx + 2 * halo1);
            k_tc_coef<<<gridk, 128, 0, d->stream>>>(d->P, K);
            d->launches++;
            A.S.coef = K.coef; A.S.rx = K.rx; A.S.ry = K.ry;
        }
        A.C = d->tc; A.e_base = e; A.e_out = e; A.T_out = Tout; A.K_store = Kst; A.K1 = K1; A.K2 = K2; A.K3 = K3; A.mode = mode; A.c = half ? 0.5 * dts : dts;
        A.fast = d->fast_interior ? 1 : 0;
        if (dev) { A.plan = d->plan; A.sub = sub; A.half = half; }
        k_tc_stage<<<grid, 128, 0, d->stream>>>(d->P, A);
        d->launches++;
        CUDA_TRY(cudaGetLastError());
        return exchange_plane(d, Tout);                  // the next stage differentiates this temperature plane across the slab edge
    };
    for (sub = 0; sub < ns; sub++) {
        if (d->tc_integrator == SPRUCE_TI_EULER) {
            if ((rc = stage(Ta, Tb, TC_FINAL, 0, nullptr))) return rc;
            std::swap(Ta, Tb);
        } else if (d->tc_integrator == SPRUCE_TI_RK2) {
            if ((rc = stage(Ta, Tb, TC_INTERMEDIATE, 1, nullptr))) return rc;
            if ((rc = stage(Tb, Ta, TC_FINAL, 0, nullptr))) return rc;
        } else {
            if ((rc = stage(Ta, Tb, TC_INTERMEDIATE, 1, K1))) return rc;
            if ((rc = stage(Tb, Tc, TC_INTERMEDIATE, 1, K2))) return rc;
            if ((rc = stage(Tc, Tb, TC_INTERMEDIATE, 0, K3))) return rc;
            if ((rc = stage(Tb, Ta, TC_RK4_FINAL, 0, nullptr))) return rc;
        }
    }
    if (d->tc_output) {                                                                             // :101-104
        if (dev) k_avg_change_dev<<<grid256, 256, 0, d->stream>>>(d->P, d->tc_avg, e, e_before, &d->ctl->step, &d->ctl->done);
        else k_avg_change<<<grid256, 256, 0, d->stream>>>(d->P, d->tc_avg, e, e_before, dt);
        d->launches++;
    }
    if ((rc = ms_feed(d, MS_DIFF, e, e_before, d->ms_frac_tc))) return rc;                          // :105-108
    if (inactive) return SPRUCE_OK;                                                                 // :109
    if ((rc = launch_propagate(d, 0))) return rc;                                                   // :110-111
    return after_module_propagate(d);
}

int rl_launch(spruce_domain *d, int count_mode, double dt, double *e_out = nullptr, bool dev = false)
{
    RlArgs A{};
    A.R = d->rl;
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.e_out = e_out ? e_out : d->Pset.p[E_E]; A.n_sub = d->rl_nsub; A.dt = dt; A.red = d->red + 2; A.count_mode = count_mode;
    if (dev) { A.plan = d->plan; A.dt_ptr = &d->ctl->step; A.done_ptr = &d->ctl->done; }
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_rl<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// RadiativeLosses::numberSubcycles (radiativelosses.cpp:161-166)
int rl_count(spruce_domain *d, double dt, int *nsub)
{
    int rc;
    if ((rc = reset_reductions(d))) return rc;
    if ((rc = rl_launch(d, 1, dt))) return rc;
    unsigned long long h[4];
    if ((rc = read_reductions(d, h))) return rc;
    if (bits_to_double(h[3]) == 0.0) { *nsub = 0; return SPRUCE_OK; }
    const double sdt = d->rl.epsilon * bits_to_double(h[2]);
    *nsub = (int)(dt / sdt) + 1;
    return SPRUCE_OK;
}
int rl_iterate(spruce_domain *d, double dt, bool dev = false)
{
    const dim3 grid256((d->P.ny + 255) / 256, d->P.nx);
    // inactive_mode: the kernel writes its result into the copy (old_e) instead of the primary plane; only the output / cumulative planes keep it (:93-98)
    const bool inactive = d->rl_inactive;
    if (inactive && !d->rl_output && !d->ms_on) return SPRUCE_OK;
    if (d->rl_output || d->ms_on) { k_plane_copy<<<grid256, 256, 0, d->stream>>>(d->P, d->old_e, d->Pset.p[E_E]); d->launches++; }     // radiativelosses.cpp:50
    double *const e_after = inactive ? d->old_e : d->Pset.p[E_E];
    const double *const e_before = inactive ? d->Pset.p[E_E] : d->old_e;
    int rc = rl_launch(d, 0, dt, inactive ? d->old_e : nullptr, dev);
    if (rc) return rc;
    if (d->rl_output) {                                                                             // :93
        if (dev) k_avg_change_dev<<<grid256, 256, 0, d->stream>>>(d->P, d->rl_avg, e_after, e_before, &d->ctl->step, &d->ctl->done);
        else k_avg_change<<<grid256, 256, 0, d->stream>>>(d->P, d->rl_avg, e_after, e_before, dt);
        d->launches++;
    }
    if ((rc = ms_feed(d, MS_DIFF, e_after, e_before, d->ms_frac_rl))) return rc;                    // :94-97
    if (inactive) return SPRUCE_OK;                                                                 // :98
    if ((rc = launch_propagate(d, 0))) return rc;                                                   // radiativelosses.cpp:99-100
    return after_module_propagate(d);
}
// ---- device-resident sub-cycle plan (SPRUCE_DEVICE_SUBCYCLES=1; module_kernels.cuh: SubPlan / k_sub_plan)
// which module sets can run without the host knowing the step size: thermal_conduction, radiative_losses, ambient_heating (the set of the reference's solar runs);
// every other hook computes launch counts or time windows on the host and keeps the per-step readback
bool dev_subcycles(const spruce_domain *d)
{
    if (!d->dev_sub || !d->plan || d->module_order.empty() || d->ar.on || !d->visc.empty()) return false;
    for (int m : d->module_order) if (m != spruce_domain::MOD_TC && m != spruce_domain::MOD_RL && m != spruce_domain::MOD_AH) return false;
    return true;
}
// preIterateModules (evolution.cpp:65) of the two sub-cycling modules: both count kernels on the unchanged state, one all-gather over the slabs, one planning thread
int plan_subcycles(spruce_domain *d)
{
    int rc;
    bool tc_on = false, rl_on = false;
    for (int m : d->module_order) { tc_on = tc_on || m == spruce_domain::MOD_TC; rl_on = rl_on || m == spruce_domain::MOD_RL; }
    if ((rc = reset_reductions(d))) return rc;
    if (tc_on && (rc = tc_count_launch(d))) return rc;
    if (rl_on && (rc = rl_launch(d, 1, 0.0))) return rc;                 // count mode does not read the step size
    if (d->cfg.n_ranks > 1 && (rc = peer_red_allgather(d))) return rc;
    SubPlanArgs A{};
    A.ctl = d->ctl; A.red = d->red; A.plan = d->plan;
    A.tc_on = tc_on ? 1 : 0; A.tc_sat = d->tc.flux_saturation; A.tc_eps = d->tc_epsilon; A.tc_dtmin = d->tc.dt_subcycle_min;
    A.rl_on = rl_on ? 1 : 0; A.rl_eps = d->rl.epsilon; A.tc_budget = d->tc_budget;
    k_sub_plan<<<1, 1, 0, d->stream>>>(A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// postIterateModule of a pointwise source term; time = m_time at the start of the step (evolution.cpp:74 runs before the time update)
int src_post(spruce_domain *d, const spruce_domain::SourceTerm &m, double time, double step)
{
    double f = step;
    if (m.kind != SRC_SINK && (time < m.start || time > m.start + m.duration)) return SPRUCE_OK;        // localizedheating.cpp:52 etc.: no propagate either
    if (m.kind == SRC_HEATING) {                                                                         // localizedheating.cpp:53-59
        const double t = time - m.start;
        double ramp = 1.0;
        if (t < m.ramp_time && t <= 0.5 * m.duration) ramp = t / m.ramp_time;
        else if (t > m.duration - m.ramp_time && t > 0.5 * m.duration) ramp = (m.duration - t) / m.ramp_time;
        f = step * ramp;
    } else if (m.kind == SRC_MOMENTUM) {                                                                 // momentuminjection.cpp:70
        const double osc = m.oscillatory ? std::sin(2.0 * kPI * (time - m.start) / m.period) : 1.0;
        f = step * (osc * m.max_accel);
    }
    SrcArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    A.p0 = m.plane[0]; A.p1 = m.plane[1]; A.kind = m.kind; A.f = f; A.done_ptr = &d->ctl->done;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_source_term<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    if (m.kind == SRC_MASS) d->raw_rho = true;                     // E_N holds rho until the propagate's floor / n round trip
    int rc = launch_propagate(d, 0);
    if (rc) return rc;
    if ((rc = after_module_propagate(d))) return rc;
    if (m.kind == SRC_SINK) return ms_feed(d, MS_RATE, m.plane[0], nullptr, m.ms_fraction, -1.0);        // ambientheatingsink.cpp:38-41
    if (m.kind == SRC_HEATING) return ms_feed(d, MS_PULSE, m.plane[0], nullptr, m.ms_fraction);           // localizedheating.cpp:63-66
    return SPRUCE_OK;
}
int launch_op(spruce_domain *d, int code, int index, const double *q, double *out)
{
    OpArgs A{};
    A.q = q; A.vel = nullptr; A.out = out; A.op = code; A.index = index;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_operator<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// DivCleaning::postIterateModule (divcleaning.cpp:26-47); scratch: Mset planes 0..4 (b_x, b_y, first derivative, mixed derivative, second derivative)
int dc_post(spruce_domain *d, double dt)
{
    int rc;
    double *bx = d->Mset.p[0], *by = d->Mset.p[1], *a = d->Mset.p[2], *b = d->Mset.p[3], *s2 = d->Mset.p[4];
    double *bix = d->Pset.p[E_BX], *biy = d->Pset.p[E_BY];
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    const int ns = (int)(dt / (d->dc.epsilon * d->dc.time_scale)) + 1;                              // :30
    const double dts = dt / ((double)ns);
    d->dc.nsub = ns;
    k_plane_sum<<<grid, 256, 0, d->stream>>>(d->P, bx, d->stat[S_BEX], bix);                        // b_x = be_x + bi_x (idealmhd.cpp:262)
    k_plane_sum<<<grid, 256, 0, d->stream>>>(d->P, by, d->stat[S_BEY], biy);
    d->launches += 2;
    // slabs: a plane that is differentiated along x needs its halo rows (b_x, and the y-derivative of b_y before its x-derivative)
    if ((rc = exchange_plane(d, bx))) return rc;
    for (int s = 0; s < ns; s++) {
        DcArgs A{};
        A.dts = dts; A.time_scale = d->dc.time_scale;
        if ((rc = launch_op(d, 0, 1, by, a)) || (rc = exchange_plane(d, a)) || (rc = launch_op(d, 0, 0, a, b)) || (rc = launch_op(d, 1, 0, bx, s2))) return rc;    // d/dx(d/dy b_y) + d2/dx2 b_x  :36-37
        A.bi = bix; A.mixed = b; A.second = s2;
        k_dc_update<<<grid, 256, 0, d->stream>>>(d->P, A);
        if ((rc = launch_op(d, 0, 0, bx, a)) || (rc = launch_op(d, 0, 1, a, b)) || (rc = launch_op(d, 1, 1, by, s2))) return rc;    // b_x is still the sub-cycle's starting value  :38-40
        A.bi = biy;
        k_dc_update<<<grid, 256, 0, d->stream>>>(d->P, A);
        k_plane_sum<<<grid, 256, 0, d->stream>>>(d->P, bx, bix, d->stat[S_BEX]);                    // :41-42
        k_plane_sum<<<grid, 256, 0, d->stream>>>(d->P, by, biy, d->stat[S_BEY]);
        d->launches += 4;
        if ((rc = exchange_plane(d, bx))) return rc;
    }
    CUDA_TRY(cudaGetLastError());
    if ((rc = launch_propagate(d, 0))) return rc;                                                   // :46
    return after_module_propagate(d);
}
// FieldHeating::preIterateModule (fieldheating.cpp:30-46); scratch: Mset planes 0..4
int fh_pre(spruce_domain *d)
{
    int rc;
    const int vars[5] = {V_b_x, V_b_y, V_b_hat_x, V_b_hat_y, V_b_mag};
    for (int k = 0; k < 5; k++) if ((rc = derive_to(d, vars[k], d->Mset.p[k]))) return rc;
    FhArgs A{};
    A.bx = d->Mset.p[0]; A.by = d->Mset.p[1]; A.bhx = d->Mset.p[2]; A.bhy = d->Mset.p[3]; A.bmag = d->Mset.p[4]; A.n = d->Pset.p[E_N];
    A.H = d->fh.H; A.coeff = d->fh.coeff; A.current_pow = d->fh.current_pow; A.b_pow = d->fh.b_pow; A.n_pow = d->fh.n_pow; A.roc_pow = d->fh.roc_pow;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_fh_compute<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// FieldHeating::iterateModule (fieldheating.cpp:48-58)
int fh_iterate(spruce_domain *d, double dt)
{
    FhArgs A{};
    A.H = d->fh.H; A.e = d->Pset.p[E_E]; A.dt = dt; A.inactive = d->fh.inactive;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_fh_apply<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    if (d->fh.inactive) return SPRUCE_OK;                                                           // :53
    int rc = launch_propagate(d, 0);
    if (rc) return rc;
    return after_module_propagate(d);
}
// BoundaryOutflow::postIterateModule (boundaryoutflow.cpp:39-63)
int bo_post(spruce_domain *d, double dt)
{
    auto &bo = d->bo;
    BoArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.tmpl = bo.tmpl; A.xl = bo.win[0]; A.xu = bo.win[1]; A.yl = bo.win[2]; A.yu = bo.win[3];
    A.boundary = bo.boundary; A.field_aligned = bo.field_aligned; A.max_key = &d->red[1];            // slot 1 of the module reductions: a maximum
    int rc = reset_reductions(d);
    if (rc) return rc;
    dim3 g1((d->P.ny + 127) / 128, d->P.nx), g2((d->P.ny + 255) / 256, d->P.nx);
    k_bo_mean<<<g1, 128, 0, d->stream>>>(d->P, A);
    unsigned long long h[4];
    if ((rc = read_reductions(d, h))) return rc;                                                     // all-gathered over the slabs
    const double init = -1.0 * bo.target;                                                            // :227
    bo.mean = init;
    if (h[1] != 0ULL) { const double best = bo_unkey(h[1]); if (init < best) bo.mean = best; }
    double accel = bo.max_accel;
    if (bo.dynamic) {                                                                                // :42-46
        accel = (bo.target - bo.mean) / bo.dynamic_time;
        accel = std::min(accel, bo.max_accel);
        accel = std::max(accel, 0.0);
    }
    bo.accel = accel;
    A.dt = dt; A.accel = accel;
    k_bo_apply<<<g2, 256, 0, d->stream>>>(d->P, A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if ((rc = launch_propagate(d, 0))) return rc;
    return after_module_propagate(d);
}
int ah_post(spruce_domain *d)
{
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_ambient_heating<<<grid, 256, 0, d->stream>>>(d->P, d->Pset.p[E_E], d->heating, &d->ctl->step, &d->ctl->done);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    int rc = launch_propagate(d, 0);                                                                // ambientheating.cpp:43-44
    if (rc) return rc;
    if ((rc = after_module_propagate(d))) return rc;
    return ms_feed(d, MS_RATE, d->heating, nullptr, d->ms_frac_ah);                                 // :45-48
}

// ---- artificial viscosity (source/modules/viscosity.cpp)
int evolved_slot(int var);
const double *materialise_var(spruce_domain *d, const PlaneSet &S, int var, double *scratch, int *rc)
{
    *rc = SPRUCE_OK;
    const int ev = evolved_slot(var);
    if (ev > 0) return S.p[ev];
    if (var == V_grav_x) return d->stat[S_GX];
    if (var == V_grav_y) return d->stat[S_GY];
    *rc = derive_to(d, var, scratch, &S);
    return scratch;
}
int visc_needs_dt_plane(const spruce_domain *d) { for (auto &t : d->visc) if (t.opt == 0 || t.opt == 2) return 1; return 0; }
int visc_refresh_dt(spruce_domain *d) { return visc_needs_dt_plane(d) ? derive_to(d, V_dt, d->dt_plane) : SPRUCE_OK; }
// constructSingleViscosityGrid :185-267 for term i on grid set S -> out
int exchange_plane(spruce_domain *d, double *plane);
// keep_dq: this evaluation is one the reference stores in m_grids_dqdt[i] (the masked RHS form :116, the second evaluation of an rk2 sub-cycle :148)
int visc_term(spruce_domain *d, const PlaneSet &S, int i, double *out, int masked, int keep_dq = 0)
{
    auto &t = d->visc[i];
    int rc;
    if (t.strength_plane && !t.halo_done) {          // static boundary profile: its halo rows travel once (gradient correction differentiates it)
        if ((rc = exchange_plane(d, t.strength_plane))) return rc;
        t.halo_done = true;
    }
    const double *q = materialise_var(d, S, t.var_diff, d->vscratch[7], &rc);
    if (rc) return rc;
    ViscArgs A{};
    A.q = q; A.n = S.p[E_N];
    A.dt_plane = (t.opt == 0 || t.opt == 2) ? d->dt_plane : nullptr;
    A.dt_min_bits = &d->ctl->dtmin_bits;
    A.strength_plane = (t.opt == 2 || t.opt == 3) ? t.strength_plane : nullptr;
    A.strength = t.strength;
    const bool evol_mom = (t.var_evol == V_mom_x || t.var_evol == V_mom_y || t.var_evol == V_mom_z);
    const bool diff_vel = (t.var_diff == V_v_x || t.var_diff == V_v_y || t.var_diff == V_v_z);
    A.scale_mode = (evol_mom && diff_vel) ? 1 : (t.var_evol == V_thermal_energy && t.var_diff == V_temp) ? 2 : 0;
    A.gradient_correction = d->visc_gradient_correction; A.masked = masked; A.out = out;
    if (d->visc_output) { A.lap_out = t.o_lap; A.dtg_out = t.o_dt; A.dq_out = keep_dq ? t.o_dqdt : nullptr; d->visc_evaluated = true; }
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_visc_term<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
// Module::computeTimeDerivativesModule for every RHS-form term (strength <= 1), evaluated on grid set S (viscosity.cpp:112-123)
int prepare_rhs_modules(spruce_domain *d, const PlaneSet &S)
{
    d->cur_nx = 0;
    for (size_t i = 0; i < d->visc.size(); i++) {
        if (d->visc[i].strength > 1.0) {
            // constructViscosityGrids evaluates EVERY term on this grid set (:269-276); a hyper-viscous term's result is dropped, but its laplacian / timescale planes
            // are now those of this evaluation -- only the output planes can tell
            if (d->visc_output) { int rc = visc_term(d, S, (int)i, nullptr, 0); if (rc) return rc; }
            continue;
        }
        if (d->cur_nx >= 4) return fail(SPRUCE_ERR_UNSUPPORTED, "at most 4 right-hand-side viscosity terms");
        int rc = visc_term(d, S, (int)i, d->vscratch[d->cur_nx], 1, 1);
        if (rc) return rc;
        d->cur_xterm[d->cur_nx] = d->vscratch[d->cur_nx];
        d->cur_xtarget[d->cur_nx] = evolved_slot(d->visc[i].var_evol);
        d->cur_nx++;
    }
    return SPRUCE_OK;
}
int launch_propagate(spruce_domain *d, int from_state);
// Viscosity::iterateModule :125-180 : sub-cycled hyper-viscous terms on the primary state
int av_iterate(spruce_domain *d, double dt)
{
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    const size_t plane_bytes = (size_t)d->P.nx * d->P.pitch * sizeof(double);
    for (size_t i = 0; i < d->visc.size(); i++) {
        const auto &t = d->visc[i];
        if (t.strength <= 1.0) continue;
        const int ev = evolved_slot(t.var_evol);
        double *evol = d->Pset.p[ev];
        const int ns = (int)(std::ceil(t.strength / d->visc_hv_epsilon) + 0.1);              // :130
        const double dts = dt / (double)ns;
        double *T[4] = {d->vscratch[0], d->vscratch[1], d->vscratch[2], d->vscratch[3]}, *init = d->vscratch[4];
        int rc;
        auto apply = [&](const double *base, double cc, int n_terms) -> int {
            AxpyArgs A{};
            A.base = base; A.t1 = T[0]; A.t2 = n_terms == 4 ? T[1] : nullptr; A.t3 = n_terms == 4 ? T[2] : nullptr; A.t4 = n_terms == 4 ? T[3] : nullptr;
            A.c = cc; A.out = evol; A.base_is_n = (ev == E_N); A.comb_out = d->visc_output ? t.o_dqdt : nullptr;
            k_visc_apply<<<grid, 256, 0, d->stream>>>(d->P, A);
            d->launches++;
            CUDA_TRY(cudaGetLastError());
            if (ev == E_N) d->raw_rho = true;
            int rp = launch_propagate(d, 0);
            return rp ? rp : after_module_propagate(d);
        };
        auto term = [&](double *out, int keep_dq = 0) -> int { int r_ = visc_refresh_dt(d); return r_ ? r_ : visc_term(d, d->Pset, (int)i, out, 0, keep_dq); };
        for (int sc = 0; sc < ns; sc++) {
            if (d->visc_hv_integrator == SPRUCE_TI_EULER) {
                if ((rc = term(T[0])) || (rc = apply(evol, dts, 1))) return rc;
            } else {
                CUDA_TRY(cudaMemcpyAsync(init, evol, plane_bytes, cudaMemcpyDeviceToDevice, d->stream));
                if (d->visc_hv_integrator == SPRUCE_TI_RK2) {
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5 * dts, 1))) return rc;
                    if ((rc = term(T[0], 1)) || (rc = apply(init, dts, 1))) return rc;
                } else {
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5 * dts, 1))) return rc;
                    std::swap(T[0], T[1]);            // keep dqdt1 in T[1] while T[0] is reused as the working term
                    if ((rc = term(T[0])) || (rc = apply(init, 0.5 * dts, 1))) return rc;
                    std::swap(T[0], T[2]);
                    if ((rc = term(T[0])) || (rc = apply(init, dts, 1))) return rc;
                    std::swap(T[0], T[3]);
                    if ((rc = term(T[0]))) return rc;
                    // now: T[1] = dqdt1, T[2] = dqdt2, T[3] = dqdt3, T[0] = dqdt4  -> (d1 + 2 d2 + 2 d3 + d4)/6
                    double *d1 = T[1], *d2 = T[2], *d3 = T[3], *d4 = T[0];
                    T[0] = d1; T[1] = d2; T[2] = d3; T[3] = d4;
                    if ((rc = apply(init, dts, 4))) return rc;
                }
            }
        }
        if ((rc = launch_propagate(d, 0)) || (rc = after_module_propagate(d))) return rc;    // :178
    }
    return SPRUCE_OK;
}


// ---- peer-store transport -------------------------------------------------------------------------------------------
size_t seg_flags_bytes() { return (sizeof(PeerFlags) + 255) & ~(size_t)255; }
PeerFlags *seg_flags(void *seg) { return (PeerFlags *)seg; }
double *seg_buf(const spruce_domain *d, void *seg, int side, int parity)
{
    return (double *)((char *)seg + seg_flags_bytes()) + ((size_t)side * 2 + parity) * d->halo_doubles;
}
int ensure_segment(spruce_domain *d)
{
    if (d->seg) return SPRUCE_OK;
    d->halo_doubles = (size_t)NEV * HALO * d->P.pitch;
    d->seg_bytes = seg_flags_bytes() + 4 * d->halo_doubles * sizeof(double);
    CUDA_TRY(cudaMalloc(&d->seg, d->seg_bytes));
    CUDA_TRY(cudaMemset(d->seg, 0, d->seg_bytes));
    CUDA_TRY(cudaMalloc(&d->push_counter, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(d->push_counter, 0, sizeof(unsigned int)));
    int lo_pri = 0, hi_pri = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    CUDA_TRY(cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi_pri));
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_main, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_comm, cudaEventDisableTiming));
    if (const char *ov = getenv("SPRUCE_HALO_OVERLAP")) d->overlap = atoi(ov) != 0;
    return SPRUCE_OK;
}
void ring_neighbours(const spruce_domain *d, int *lo, int *hi)
{
    const int r = d->cfg.rank, w = d->cfg.n_ranks;
    *lo = r > 0 ? r - 1 : (d->P.xper ? w - 1 : -1);
    *hi = r < w - 1 ? r + 1 : (d->P.xper ? 0 : -1);
}
// push my edge rows of `U` into the neighbours' segments, then wait for theirs and copy them into my halo rows
int peer_exchange(spruce_domain *d, double *const *U, cudaStream_t st)
{
    if (!st) st = d->stream;
    if (!d->peers_connected) return fail(SPRUCE_ERR_STATE, "peer transport used before spruce_mgpu_ipc_connect");
    const unsigned long long q = ++d->halo_seq;
    const int par = (int)(q & 1);
    int lo, hi;
    ring_neighbours(d, &lo, &hi);
    PushArgs S{};
    for (int v = 0; v < NEV; v++) S.U[v] = U[v];
    if (lo >= 0) { S.peer_lo_buf = seg_buf(d, d->peer_seg[lo], 1, par); S.peer_lo_flag = &seg_flags(d->peer_seg[lo])->halo_seq[1][0]; }   // I am its UPPER neighbour
    if (hi >= 0) { S.peer_hi_buf = seg_buf(d, d->peer_seg[hi], 0, par); S.peer_hi_flag = &seg_flags(d->peer_seg[hi])->halo_seq[0][0]; }
    S.seq = q; S.counter = d->push_counter; S.done_ptr = &d->ctl->done;
    dim3 grid((d->P.pitch + 255) / 256, NEV * HALO);
    k_halo_push<<<grid, 256, 0, st>>>(d->P, S);
    PullArgs R{};
    for (int v = 0; v < NEV; v++) R.U[v] = U[v];
    if (lo >= 0) { R.lo_buf = seg_buf(d, d->seg, 0, par); R.lo_flag = &seg_flags(d->seg)->halo_seq[0][0]; }
    if (hi >= 0) { R.hi_buf = seg_buf(d, d->seg, 1, par); R.hi_flag = &seg_flags(d->seg)->halo_seq[1][0]; }
    R.seq = q; R.error = &seg_flags(d->seg)->error; R.done_ptr = &d->ctl->done;
    k_halo_pull<<<grid, 256, 0, st>>>(d->P, R);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int peer_red_allgather(spruce_domain *d)
{
    if (!d->peers_connected) return fail(SPRUCE_ERR_STATE, "module reductions on a slab need spruce_mgpu_ipc_connect");
    RedGatherArgs A{};
    A.red = d->red; A.mine = seg_flags(d->seg); A.rank = d->cfg.rank; A.world = d->cfg.n_ranks; A.seq = ++d->red_seq;
    for (int r = 0; r < d->cfg.n_ranks; r++) A.peer[r] = seg_flags(d->peer_seg[r]);
    k_red_publish<<<1, MAX_RANKS, 0, d->stream>>>(A);
    k_red_collect<<<1, MAX_RANKS, 0, d->stream>>>(A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int peer_dt_allgather(spruce_domain *d)
{
    DtGatherArgs A{};
    A.ctl = d->ctl; A.mine = seg_flags(d->seg); A.rank = d->cfg.rank; A.world = d->cfg.n_ranks; A.seq = ++d->dt_seq;
    for (int r = 0; r < d->cfg.n_ranks; r++) A.peer[r] = seg_flags(d->peer_seg[r]);
    k_dt_publish<<<1, MAX_RANKS, 0, d->stream>>>(A);
    k_dt_collect<<<1, MAX_RANKS, 0, d->stream>>>(A);
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

// ghost zones of the set a stage produced, then (slab decomposition) its halo rows from the ring neighbours
int finish_stage(spruce_domain *d, const PlaneSet &U, int primary)
{
    int rc = launch_ghosts(d, U, primary);
    if (!rc) rc = moc_limit(d, U, primary);
    if (rc || d->cfg.n_ranks == 1) return rc;
    return peer_exchange(d, U.p, nullptr);
}
// One RK stage of a slab.  When no ghost-zone pass follows the stage (periodic y and no physical x boundary that writes ghost cells),
// the first and last row chunk run on the high-priority communication stream, followed there by the push / pull of the halo rows,
// while the chunks in between run on the main stream: the exchange hides behind the interior compute.
int stage_and_exchange(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode)
{
    NvtxRange range_stage(primary ? "rk stage (primary) + exchange" : "rk stage + exchange");
    int rc;
    const bool split = can_split(d, primary);
    if (!split) {
        if ((rc = launch_stage(d, S, B, D, coef, primary, kmode))) return rc;
        return finish_stage(d, D, primary);
    }
    if (primary && !d->fused_ctl) { k_dtmin_reset<<<1, 1, 0, d->stream>>>(d->ctl, dt_prune_enabled(d)); d->launches++; }
    CUDA_TRY(cudaEventRecord(d->ev_main, d->stream));
    CUDA_TRY(cudaStreamWaitEvent(d->comm_stream, d->ev_main, 0));
    MARK(d, "  comm: start", d->comm_stream);
    if ((rc = launch_stage(d, S, B, D, coef, primary, kmode, 1))) return rc;
    MARK(d, "  comm: edge rows done", d->comm_stream);
    if ((rc = peer_exchange(d, D.p, d->comm_stream))) return rc;
    MARK(d, "  comm: push + pull done", d->comm_stream);
    CUDA_TRY(cudaEventRecord(d->ev_comm, d->comm_stream));
    if ((rc = launch_stage(d, S, B, D, coef, primary, kmode, 2))) return rc;
    MARK(d, "  main: interior rows done", d->stream);
    CUDA_TRY(cudaStreamWaitEvent(d->stream, d->ev_comm, 0));
    MARK(d, "  main: joined", d->stream);
    return SPRUCE_OK;
}


// ---- PhysicalViscosity::iterateModule (physicalviscosity.cpp:150-245)
int exchange_planes4(spruce_domain *d, double *a, double *b, double *c, double *e)
{
    if (d->cfg.n_ranks == 1) return SPRUCE_OK;
    double *v[NEV] = {a, b, c, e, a, b, c, e};
    return peer_exchange(d, v, nullptr);
}
PvArgs pv_args(spruce_domain *d)
{
    auto &pv = d->pv;
    PvArgs A{};
    for (int k = 0; k < 3; k++) { A.bh[k] = pv.bh[k]; A.mom[k] = d->Pset.p[E_MX + k]; A.v[k] = pv.v[0][k]; }
    A.T = pv.T[0];
    A.n = d->Pset.p[E_N]; A.cg = pv.cg; A.e = d->Pset.p[E_E];
    A.coeff = pv.coeff; A.heating_on = pv.heating_on; A.force_on = pv.force_on; A.gc = pv.gc; A.red = d->red; A.fast = d->fast_interior ? 1 : 0;
    return A;
}
// the sub-cycles of iterateModule (physicalviscosity.cpp:148-245) for pv.nsub sub-cycles, from the velocity / temperature / b_hat planes pv_iterate derived
int pv_substeps(spruce_domain *d, double dt)
{
    int rc;
    auto &pv = d->pv;
    PvArgs A = pv_args(d);
    const dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    const double dts = dt / pv.nsub;
    if (pv.output) {                                                    // avg_heating / avg_force start every iterate at zero (:151-152)
        for (int k = 0; k < 4; k++) { A.avg[k] = pv.avg[k]; CUDA_TRY(cudaMemsetAsync(pv.avg[k] - d->row_off, 0, d->plane_doubles * sizeof(double), d->stream)); }
        A.nsub = (double)pv.nsub;
    }
    if (d->ms_on) { A.ms_i = d->ms_cum[1]; A.ms_e = d->ms_cum[0]; A.ms_f = d->ms_frac_pv; }
    int cur = 0;
    auto stage = [&](int from, int to, double half, int final_stage) -> int {
        for (int k = 0; k < 3; k++) { A.v[k] = pv.v[from][k]; A.v_out[k] = pv.v[to][k]; }
        A.T = pv.T[from]; A.T_out = pv.T[to]; A.dt = dts; A.half = half; A.final_stage = final_stage;
        k_pv_stage<<<grid, 128, 0, d->stream>>>(d->P, A);
        d->launches++;
        CUDA_TRY(cudaGetLastError());
        return A.diag_only ? SPRUCE_OK : exchange_planes4(d, pv.v[to][0], pv.v[to][1], pv.v[to][2], pv.T[to]);
    };
    if (!pv.inactive && (pv.heating_on || pv.force_on)) {
        for (int s = 0; s < pv.nsub; s++) {
            if (pv.integrator == SPRUCE_TI_EULER) { if ((rc = stage(cur, cur ^ 1, 1.0, 1))) return rc; cur ^= 1; }
            else {                                                      // rk2: half step into the other set, full step back
                if ((rc = stage(cur, cur ^ 1, 0.5, 0))) return rc;
                if ((rc = stage(cur ^ 1, cur, 1.0, 1))) return rc;
            }
        }
    } else if (pv.inactive && pv.output && (pv.heating_on || pv.force_on)) {
        // inactive_mode: the reference still evaluates heating and force in every sub-cycle, for the output planes only (:163-171, :199-223); the state does not
        // change in between, so one evaluation is added nsub times
        A.diag_only = 1; A.repeat = pv.nsub;
        if ((rc = stage(0, 1, 1.0, 1))) return rc;
    }
    if ((rc = launch_propagate(d, 0))) return rc;                       // :244
    return after_module_propagate(d);
}
int pv_iterate(spruce_domain *d, double dt)
{
    int rc;
    auto &pv = d->pv;
    const int vars_v[3] = {V_v_x, V_v_y, V_v_z}, vars_b[3] = {V_b_hat_x, V_b_hat_y, V_b_hat_z};
    for (int k = 0; k < 3; k++) { if ((rc = derive_to(d, vars_v[k], pv.v[0][k]))) return rc; if ((rc = derive_to(d, vars_b[k], pv.bh[k]))) return rc; }
    if ((rc = derive_to(d, V_temp, pv.T[0]))) return rc;
    if (!pv.cg_halo_done) {                          // the coefficient plane is static: its halo rows travel once (gradient correction differentiates it)
        if ((rc = exchange_plane(d, pv.cg))) return rc;
        pv.cg_halo_done = true;
    }
    // computeViscousSubcycles :66-80 -- evaluated here, in iterate, on the state earlier modules have already changed
    const PvArgs A = pv_args(d);
    if ((rc = reset_reductions(d))) return rc;
    const dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_pv_count<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    unsigned long long h[4];
    if ((rc = read_reductions(d, h))) return rc;
    pv.nsub = (int)(dt / (pv.epsilon * bits_to_double(h[0]))) + 1;
    return pv_substeps(d, dt);
}

// after the step's last stage: all-gather of the slabs' minima, then the check that the skip test's window held (else: every cell)
int finish_dt(spruce_domain *d)
{
    int rc;
    const bool multi = d->cfg.n_ranks > 1;
    if (multi && (rc = peer_dt_allgather(d))) return rc;
    if (!dt_prune_enabled(d)) return SPRUCE_OK;
    k_dt_validate<<<1, 1, 0, d->stream>>>(d->ctl);
    DtFullArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.ctl = d->ctl;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx < 64 ? d->P.nx : 64);
    k_dt_full<<<grid, 256, 0, d->stream>>>(d->P, A);            // returns at once unless the window was missed
    d->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (multi && (rc = peer_dt_allgather(d))) return rc;
    return SPRUCE_OK;
}

// the fused step control of plain runs (mhd_kernels.cuh: k_step_open / k_step_mid / k_step_close)
void fill_gather(spruce_domain *d, DtGatherArgs &A, unsigned long long seq)
{
    A.ctl = d->ctl; A.rank = d->cfg.rank; A.world = d->cfg.n_ranks; A.seq = seq;
    A.mine = d->cfg.n_ranks > 1 ? seg_flags(d->seg) : nullptr;
    for (int r = 0; r < d->cfg.n_ranks && d->cfg.n_ranks > 1; r++) A.peer[r] = seg_flags(d->peer_seg[r]);
}
int finish_dt_fused(spruce_domain *d, int next_slot)
{
    DtGatherArgs G{};
    fill_gather(d, G, ++d->dt_seq);          // the same sequence as peer_dt_allgather (set-up, module runs): consecutive gathers alternate the buffer parity
    k_step_mid<<<1, MAX_RANKS, 0, d->stream>>>(G);
    DtFullArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.ctl = d->ctl;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx < 64 ? d->P.nx : 64);
    k_dt_full<<<grid, 256, 0, d->stream>>>(d->P, A);            // returns at once unless the window was missed
    k_step_close<<<1, MAX_RANKS, 0, d->stream>>>(G, d->dt_hist, next_slot, dt_prune_enabled(d));
    d->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

#include "anomres_host.cuh"

// one advanceTime (evolution.cpp:59-82) worth of launches
// plain run: no module hook, no viscosity term, no open_moc strip kernel, dt skip test in use -> the step control is fused (and the step size of
// the next step is fixed by this step's closing kernel)
bool plain_run(const spruce_domain *d)
{
    return d->fuse_ctl_enabled && d->module_order.empty() && d->visc.empty() && !d->moc_any && dt_prune_enabled(d) && d->cfg.time_integrator != SPRUCE_TI_RK4
           && !(d->moc_lim.b_on || d->moc_lim.mom_on);
}
int enqueue_step_plain(spruce_domain *d, int hist_slot, bool first, bool last)
{
    NvtxRange range_step("spruce step (plain)");
    int rc;
    d->timeline_on = hist_slot < d->timeline_steps;
    MARK(d, "step begin", d->stream);
    if (first) { k_step_open<<<1, 1, 0, d->stream>>>(d->ctl, d->dt_hist, hist_slot, dt_prune_enabled(d)); d->launches++; }
    d->fused_ctl = true;
    if (d->cfg.time_integrator == SPRUCE_TI_EULER) {
        rc = stage_and_exchange(d, d->Pset, d->Pset, d->Mset, 1.0, 1, KM_NONE);
        if (!rc) std::swap(d->Pset, d->Mset);
    } else {
        rc = stage_and_exchange(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_NONE);
        if (!rc) rc = stage_and_exchange(d, d->Mset, d->Pset, d->Pset, 1.0, 1, KM_NONE);
    }
    d->fused_ctl = false;
    if (rc) return rc;
    MARK(d, "stages done (main)", d->stream);
    rc = finish_dt_fused(d, last ? -1 : hist_slot + 1);
    MARK(d, "step control done", d->stream);
    d->timeline_on = false;
    return rc;
}

int enqueue_step(spruce_domain *d, int hist_slot)
{
    NvtxRange range_step("spruce step");
    int rc;
    k_step_begin<<<1, 1, 0, d->stream>>>(d->ctl, d->dt_hist, hist_slot);
    d->launches++;
    double step_time = 0.0, step_size = 0.0;                             // host copies for the post-iterate hooks
    if (dev_subcycles(d)) {
        // counts, sub-cycle step and step size stay on the device: nothing here waits for the host
        NvtxRange range_dev("modules: preIterate + iterate (device-resident plan)");
        if ((rc = plan_subcycles(d))) return rc;
        for (int m : d->module_order) {                                  // iterateModules, evolution.cpp:66
            if (m == spruce_domain::MOD_TC && (rc = tc_iterate(d, 0.0, true))) return rc;
            if (m == spruce_domain::MOD_RL && (rc = rl_iterate(d, 0.0, true))) return rc;
        }
    } else if (!d->module_order.empty()) {
        // the module hooks need the step size on the host (sub-cycle counts decide how many kernels are launched)
        StepCtl h;
        CUDA_TRY(cudaMemcpyAsync(&h, d->ctl, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
        CUDA_TRY(cudaStreamSynchronize(d->stream));
        if (h.done) return SPRUCE_OK;
        const double step = h.step;
        step_time = h.time; step_size = h.step;
        if (d->ar.on && !d->ar.ready && (rc = ar_setup_run(d))) return rc;   // setupModule works on the state before the first step
        {
            NvtxRange range_pre("modules: preIterate");
            for (int m : d->module_order) {                              // preIterateModules, evolution.cpp:65
                if (m == spruce_domain::MOD_TC && (rc = tc_count(d, step, &d->tc_nsub))) return rc;
                if (m == spruce_domain::MOD_RL && (rc = rl_count(d, step, &d->rl_nsub))) return rc;
                if (m == spruce_domain::MOD_FH && (rc = fh_pre(d))) return rc;
            }
        }
        NvtxRange range_it("modules: iterate");
        for (int m : d->module_order) {                                  // iterateModules, evolution.cpp:66
            if (m == spruce_domain::MOD_TC && (rc = tc_iterate(d, step))) return rc;
            if (m == spruce_domain::MOD_RL && (rc = rl_iterate(d, step))) return rc;
            if (m == spruce_domain::MOD_AV && (rc = av_iterate(d, step))) return rc;
            if (m == spruce_domain::MOD_PV && (rc = pv_iterate(d, step))) return rc;
            if (m == spruce_domain::MOD_FH && (rc = fh_iterate(d, step))) return rc;
            if (m == spruce_domain::MOD_AR && (rc = ar_iterate(d, step))) return rc;
        }
    }
    if (!d->visc.empty() && (rc = visc_refresh_dt(d))) return rc;     // Viscosity reads the PRIMARY state's dt plane
    if (!d->module_order.empty()) {
        // The modules use the planes of Mset as scratch (tc_iterate, dc_post, fh_pre).  The 2-D instance of the stage kernel neither reads nor
        // writes mom_z / bi_z -- it relies on those planes being zero in EVERY set: euler swaps Mset in as the primary state, and the open_moc
        // strip kernel reads all eight planes of the stage copy.  Restore the invariant before the stages run.
        const ActiveList L = active_quantities(d);
        if (L.n == 6 && L.q == XY_LIST_2D) {
            CUDA_TRY(cudaMemsetAsync(d->Mset.p[E_MZ] - d->row_off, 0, d->plane_doubles * sizeof(double), d->stream));
            CUDA_TRY(cudaMemsetAsync(d->Mset.p[E_BZ] - d->row_off, 0, d->plane_doubles * sizeof(double), d->stream));
        }
    }
    const int ti = d->cfg.time_integrator;
    if (ti == SPRUCE_TI_EULER) {                                        // evolution.cpp:84-88
        if ((rc = stage_and_exchange(d, d->Pset, d->Pset, d->Mset, 1.0, 1, KM_NONE))) return rc;   // ghost zones / halos of Mset ...
        std::swap(d->Pset, d->Mset);                                    // ... which now becomes the primary set (D never aliases S: ping-pong)
    } else if (ti == SPRUCE_TI_RK2) {                                   // evolution.cpp:90-101
        if ((rc = stage_and_exchange(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_NONE))) return rc;
        if ((rc = stage_and_exchange(d, d->Mset, d->Pset, d->Pset, 1.0, 1, KM_NONE))) return rc;
    } else {                                                            // evolution.cpp:103-124
        if ((rc = ensure_rk4(d))) return rc;
        if ((rc = stage_and_exchange(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_STORE_K1))) return rc;
        if ((rc = stage_and_exchange(d, d->Mset, d->Pset, d->M2set, 0.5, 0, KM_STORE_K2))) return rc;
        if ((rc = stage_and_exchange(d, d->M2set, d->Pset, d->Mset, 1.0, 0, KM_ADD_K2))) return rc;
        if ((rc = stage_and_exchange(d, d->Mset, d->Pset, d->Pset, 1.0, 1, KM_FINAL))) return rc;
    }
    if ((rc = finish_dt(d))) return rc;                                 // global min(dt) for the next step (evolution.cpp:62)
    NvtxRange range_post("modules: postIterate");
    for (int m : d->module_order) {                                      // postIterateModules, evolution.cpp:74
        if (m == spruce_domain::MOD_AH && (rc = ah_post(d))) return rc;
        if (m == spruce_domain::MOD_DC && (rc = dc_post(d, step_size))) return rc;
        if (m == spruce_domain::MOD_BO && (rc = bo_post(d, step_size))) return rc;
        if (m >= spruce_domain::MOD_SRC0 && (rc = src_post(d, d->sources[m - spruce_domain::MOD_SRC0], step_time, step_size))) return rc;
    }
    k_step_end<<<1, 1, 0, d->stream>>>(d->ctl);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

struct NameMap { const char *name; int var; };
const NameMap kVarNames[] = {
    {"rho", V_rho}, {"temp", V_temp}, {"mom_x", V_mom_x}, {"mom_y", V_mom_y}, {"mom_z", V_mom_z}, {"bi_x", V_bi_x}, {"bi_y", V_bi_y},
    {"bi_z", V_bi_z}, {"grav_x", V_grav_x}, {"grav_y", V_grav_y}, {"n", V_n}, {"press", V_press}, {"thermal_energy", V_thermal_energy},
    {"v_x", V_v_x}, {"v_y", V_v_y}, {"v_z", V_v_z}, {"kinetic_energy", V_kinetic_energy}, {"b_x", V_b_x}, {"b_y", V_b_y}, {"b_z", V_b_z},
    {"b_mag", V_b_mag}, {"b_hat_x", V_b_hat_x}, {"b_hat_y", V_b_hat_y}, {"b_hat_z", V_b_hat_z}, {"dt", V_dt}};

int var_index(const char *name)
{
    for (const auto &m : kVarNames) if (!strcmp(m.name, name)) return m.var;
    return -1;
}
// evolved plane index for a variable, or -1
int evolved_slot(int var)
{
    switch (var) {
    case V_rho: return E_N; case V_mom_x: return E_MX; case V_mom_y: return E_MY; case V_mom_z: return E_MZ;
    case V_thermal_energy: return E_E; case V_bi_x: return E_BX; case V_bi_y: return E_BY; case V_bi_z: return E_BZ;
    default: return -1;
    }
}
int static_slot(const char *name)
{
    if (!strcmp(name, "be_x")) return S_BEX; if (!strcmp(name, "be_y")) return S_BEY; if (!strcmp(name, "be_z")) return S_BEZ;
    if (!strcmp(name, "grav_x")) return S_GX; if (!strcmp(name, "grav_y")) return S_GY;
    return -1;
}

int h2d_plane_begin(spruce_domain *d, double *dev, const double *host)
{
    CUDA_TRY(cudaMemcpy2DAsync(dev, d->P.pitch * sizeof(double), host, d->P.ny * sizeof(double), d->P.ny * sizeof(double), d->P.nx,
                               cudaMemcpyHostToDevice, d->stream));
    return SPRUCE_OK;
}
int h2d_plane_end(spruce_domain *d)
{
    CUDA_TRY(cudaStreamSynchronize(d->stream));          // the caller's buffer may be reused once the call returns
    return SPRUCE_OK;
}
int h2d_plane(spruce_domain *d, double *dev, const double *host)
{
    const int rc = h2d_plane_begin(d, dev, host);
    return rc ? rc : h2d_plane_end(d);
}
int d2h_plane(spruce_domain *d, double *host, const double *dev)
{
    CUDA_TRY(cudaMemcpy2DAsync(host, d->P.ny * sizeof(double), dev, d->P.pitch * sizeof(double), d->P.ny * sizeof(double), d->P.nx,
                               cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return SPRUCE_OK;
}

#include "ideal2f_host.cuh"
#include "mhd2e_host.cuh"
static_assert((int)SPRUCE_TI_EULER == (int)e2::TI2_EULER && (int)SPRUCE_TI_RK2 == (int)e2::TI2_RK2 && (int)SPRUCE_TI_RK4 == (int)e2::TI2_RK4, "integrator codes");

}  // namespace

extern "C" {

const char *spruce_last_error(void) { return g_err; }
int spruce_abi_version(void) { return SPRUCE_ABI_VERSION; }

int spruce_domain_create(const spruce_config *cfg, spruce_domain **out)
{
    if (!cfg || !out) return fail(SPRUCE_ERR_ARG, "null argument");
    if (cfg->abi_version != SPRUCE_ABI_VERSION) return fail(SPRUCE_ERR_ARG, "ABI version mismatch: header %d, library %d", cfg->abi_version, SPRUCE_ABI_VERSION);
    const bool two_fluid = (cfg->equation_set == SPRUCE_EQS_IDEAL_2F), two_energy = (cfg->equation_set == SPRUCE_EQS_IDEAL_MHD_2E);
    if (cfg->equation_set != SPRUCE_EQS_IDEAL_MHD && !two_fluid && !two_energy)
        return fail(SPRUCE_ERR_UNSUPPORTED, "equation set %d is not built (ideal_mhd, ideal_mhd_2E and ideal_2F only)", cfg->equation_set);
    if (two_fluid) { int rc2 = tf_check_boundaries(*cfg); if (rc2) return rc2; }
    if (two_energy) { int rc2 = e2_check(*cfg); if (rc2) return rc2; }
    // "Grid too small for ghost zones", plasmadomain.cpp:140
    if (cfg->xdim <= 2 * HALO || cfg->ydim <= 2 * HALO) return fail(SPRUCE_ERR_ARG, "Grid too small for ghost zones");
    const int bcs[4] = {cfg->x_bound_1, cfg->x_bound_2, cfg->y_bound_1, cfg->y_bound_2};
    for (int b : bcs) {
        if (b < 0 || b > SPRUCE_BC_OPEN_UCNP) return fail(SPRUCE_ERR_ARG, "Boundary cond'n must be defined");
    }
    bool moc_any = false;
    for (int b : bcs) if (b == SPRUCE_BC_OPEN_MOC) moc_any = true;
    if (moc_any) {
        if (two_fluid) return fail(SPRUCE_ERR_UNSUPPORTED, "open_moc boundaries exist for ideal_mhd only (idealmhd.cpp:306)");
        for (int a = 0; a < 2; a++)       // a periodic side opposite an open_moc side is not a configuration the reference can run
            if ((bcs[2 * a] == SPRUCE_BC_PERIODIC) != (bcs[2 * a + 1] == SPRUCE_BC_PERIODIC)) return fail(SPRUCE_ERR_ARG, "periodic boundaries come in pairs");
    }
    if (cfg->time_integrator < 0 || cfg->time_integrator > SPRUCE_TI_RK4) return fail(SPRUCE_ERR_ARG, "invalid time integrator");
    if (cfg->n_ranks < 1 || cfg->nx_local < 1 || cfg->row0 < 0 || cfg->row0 + cfg->nx_local > cfg->xdim) return fail(SPRUCE_ERR_ARG, "bad slab [%d,%d) of %d", cfg->row0, cfg->row0 + cfg->nx_local, cfg->xdim);
    if (cfg->n_ranks > 1 && cfg->nx_local < 2 * HALO) return fail(SPRUCE_ERR_ARG, "slab thinner than the halo");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(SPRUCE_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback (%s)", cudaGetErrorString(e));
    if (cfg->device >= 0) CUDA_TRY(cudaSetDevice(cfg->device));

    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<0, 0ULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(NTR, 0)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<6, XY_LIST_2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(6, 0)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<6, XY_LIST_2D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(6, 1)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<6, XY_LIST_2D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(6, 2)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<6, XY_LIST_2D, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(6, 3)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<12, XY_LIST_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(NTR, 0)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<12, XY_LIST_FULL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(NTR, 1)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<12, XY_LIST_FULL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(NTR, 2)));
    CUDA_TRY(cudaFuncSetAttribute(k_mhd_stage_xy<12, XY_LIST_FULL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes(NTR, 3)));
    spruce_domain *d = new spruce_domain();
    d->cfg = *cfg;
    if (const char *fc = getenv("SPRUCE_FUSED_CTL")) d->fuse_ctl_enabled = atoi(fc) != 0;
    if (const char *fi = getenv("SPRUCE_FAST_INTERIOR")) d->fast_interior = atoi(fi) != 0;
    if (const char *ds = getenv("SPRUCE_DEVICE_SUBCYCLES")) d->dev_sub = atoi(ds) != 0;
    if (const char *tp = getenv("SPRUCE_TC_TWO_PASS")) d->tc_two_pass = atoi(tp) != 0;
    if (const char *tb = getenv("SPRUCE_TC_BUDGET")) { const int v = atoi(tb); if (v >= 1) d->tc_budget = v; }      // the first advance's budget
    if (const char *tl = getenv("SPRUCE_TIMELINE")) d->timeline_steps = atoi(tl);
    if (const char *sl = getenv("SPRUCE_STATIC_LISTS")) d->static_lists = atoi(sl) != 0;
    if (const char *sv = getenv("SPRUCE_STAGE_VARIANTS")) d->stage_variants = atoi(sv) != 0 ? 1 : 0;
    if (const char *sb = getenv("SPRUCE_VEC_ROWS")) d->vec_rows = atoi(sb) != 0;
    if (const char *cr = getenv("SPRUCE_CHUNK_ROWS")) { const int v = atoi(cr); if (v >= 8) d->chunk_rows_override = v; }
    if (const char *ar = getenv("SPRUCE_ARITH")) {
        if (!strcmp(ar, "relaxed")) d->relaxed = true;
        else if (strcmp(ar, "exact")) { delete d; return fail(SPRUCE_ERR_ARG, "SPRUCE_ARITH must be exact or relaxed"); }
    }
    d->moc_any = moc_any;
    DomainParams &P = d->P;
    P.nx = cfg->nx_local; P.ny = cfg->ydim; P.pitch = (cfg->ydim + 15) & ~15;
    P.gnx = cfg->xdim; P.row0 = cfg->row0;
    iteration_bounds(*cfg, P);
    P.xper = (cfg->x_bound_1 == SPRUCE_BC_PERIODIC && cfg->x_bound_2 == SPRUCE_BC_PERIODIC);
    P.yper = (cfg->y_bound_1 == SPRUCE_BC_PERIODIC && cfg->y_bound_2 == SPRUCE_BC_PERIODIC);
    P.xwrap = (P.xper && cfg->n_ranks == 1);
    P.bc_x1 = cfg->x_bound_1; P.bc_x2 = cfg->x_bound_2; P.bc_y1 = cfg->y_bound_1; P.bc_y2 = cfg->y_bound_2;
    P.m_i = cfg->ion_mass; P.rm_i = 1.0 / cfg->ion_mass;
    P.gamma = cfg->adiabatic_index; P.gm1 = cfg->adiabatic_index - 1.0;
    P.n_min = cfg->density_min; P.T_min = cfg->temp_min; P.e_min = cfg->thermal_energy_min;
    P.fourpi = 4.0 * kPI; P.rfourpi = 1.0 / (4.0 * kPI);
    P.epsilon = cfg->epsilon;

    d->plane_doubles = (size_t)(P.nx + 2 * HALO) * P.pitch;
    d->row_off = (size_t)HALO * P.pitch;
    int rc = SPRUCE_OK;
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(SPRUCE_ERR_CUDA, "cudaStreamCreate failed"); }
    if (!rc) {                                                        // base arena: evolved sets + static + scratch (+ two-fluid extras)
        d->pool_planes = two_fluid ? (2 * 14 + 2 + 4 + NSTATIC + 2) : two_energy ? (2 * 7 + 2 + NSTATIC + 2) : (2 * NEV + NSTATIC + 2);
        if (cudaMalloc(&d->pool, d->pool_planes * d->plane_doubles * sizeof(double)) != cudaSuccess) { cudaGetLastError(); d->pool = nullptr; d->pool_planes = 0; }
        else d->allocs.push_back(d->pool);
    }
    if (!rc && two_fluid) rc = tf_create(d);
    if (!rc && two_energy) rc = e2_create(d);
    if (!rc && !two_fluid && !two_energy) rc = alloc_set(d, d->Pset);
    if (!rc && !two_fluid && !two_energy) rc = alloc_set(d, d->Mset);
    for (int v = 0; v < NSTATIC && !rc; v++) rc = alloc_plane(d, &d->stat[v]);
    if (!rc) rc = alloc_plane(d, &d->scratch_temp);
    if (!rc) rc = alloc_plane(d, &d->scratch_out);
    d->strip_pitch = (P.ny > P.nx ? P.ny : P.nx) + 16;
    for (int s = 0; s < 4 && !rc; s++) {
        const size_t n = 4 * (size_t)d->strip_pitch;
        if (cudaMalloc(&d->strip[s], n * sizeof(double)) != cudaSuccess) rc = fail(SPRUCE_ERR_CUDA, "cudaMalloc failed");
        else { cudaMemsetAsync(d->strip[s], 0, n * sizeof(double), d->stream); d->allocs.push_back(d->strip[s]); }
    }
    if (!rc && cudaMalloc(&d->ctl, sizeof(StepCtl)) != cudaSuccess) rc = fail(SPRUCE_ERR_CUDA, "cudaMalloc failed");
    if (!rc && cudaMalloc(&d->red, 4 * sizeof(unsigned long long)) != cudaSuccess) rc = fail(SPRUCE_ERR_CUDA, "cudaMalloc failed");
    if (!rc && d->dev_sub) {
        SubPlan sp{}; sp.tc_last = -1; sp.rl_last = -1;
        if (cudaMalloc(&d->plan, sizeof(SubPlan)) != cudaSuccess || cudaMemcpy(d->plan, &sp, sizeof(sp), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(SPRUCE_ERR_CUDA, "cudaMalloc failed");
    }
    if (!rc) {
        StepCtl h{};
        h.step = 0.0; h.time = cfg->time; h.max_time = -1.0; h.epsilon = cfg->epsilon; h.iter = 0; h.done = 0;
        h.dtmin_bits = 0x7FEFFFFFFFFFFFFFULL;
        h.need_full = 0; h.inv_thr = 0.0; h.thr_bits = 0x7FF0000000000000ULL;
        h.prune_factor = 1.05;                     // window of the dt skip test: cells certainly above 1.05 x the last minimum are not evaluated
        if (const char *pf = getenv("SPRUCE_DT_PRUNE")) h.prune_factor = atof(pf);
        if (cudaMemcpy(d->ctl, &h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(SPRUCE_ERR_CUDA, "cudaMemcpy failed");
    }
    if (rc) { spruce_domain_destroy(d); return rc; }
    *out = d;
    return SPRUCE_OK;
}

void spruce_domain_destroy(spruce_domain *d)
{
    if (!d) return;
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (double *p : d->allocs) cudaFree(p);
    if (d->tab_dev) cudaFree(d->tab_dev);
    if (d->ctl) cudaFree(d->ctl);
    if (d->red) cudaFree(d->red);
    if (d->plan) cudaFree(d->plan);
    if (d->moc_base) cudaFree(d->moc_base);
    if (d->dt_hist) cudaFree(d->dt_hist);
    for (int r = 0; r < MAX_RANKS; r++) if (d->peer_seg[r] && d->peer_seg[r] != d->seg) cudaIpcCloseMemHandle(d->peer_seg[r]);
    if (d->seg) cudaFree(d->seg);
    if (d->push_counter) cudaFree(d->push_counter);
    if (d->comm_stream) { cudaStreamSynchronize(d->comm_stream); cudaStreamDestroy(d->comm_stream); }
    if (d->ev_main) cudaEventDestroy(d->ev_main);
    if (d->ev_comm) cudaEventDestroy(d->ev_comm);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d->tf;
    delete d->e2;
    delete d;
}

int spruce_set_cell_sizes(spruce_domain *d, const double *d_x, size_t n_x, const double *d_y, size_t n_y)
{
    CHECK_DOM(d);
    if (!d_x || !d_y || (int)n_x != d->cfg.xdim || (int)n_y != d->cfg.ydim) return fail(SPRUCE_ERR_ARG, "d_x needs xdim=%d and d_y ydim=%d entries", d->cfg.xdim, d->cfg.ydim);
    for (size_t k = 0; k < n_x; k++) if (!(d_x[k] > 0.0)) return fail(SPRUCE_ERR_ARG, "d_x[%zu] is not positive", k);
    for (size_t k = 0; k < n_y; k++) if (!(d_y[k] > 0.0)) return fail(SPRUCE_ERR_ARG, "d_y[%zu] is not positive", k);
    d->dxg.assign(d_x, d_x + n_x);
    d->dyg.assign(d_y, d_y + n_y);
    int rc = upload_tables(d);
    if (rc) return rc;
    build_ghost_proto(d);
    d->have_geom = true;
    return SPRUCE_OK;
}

int spruce_grid_upload(spruce_domain *d, const char *name, const double *host, size_t count)
{
    CHECK_DOM(d);
    if (!name || !host) return fail(SPRUCE_ERR_ARG, "null argument");
    if (count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "plane <%s>: expected %zu values, got %zu", name, (size_t)d->P.nx * d->P.ny, count);
    if (!strcmp(name, "pos_x") || !strcmp(name, "pos_y") || !strcmp(name, "d_x") || !strcmp(name, "d_y")) return SPRUCE_OK; // host-only grids
    if (d->tf) return tf_upload(d, name, host);
    if (d->e2) return e2_upload(d, name, host);
    // destination first (every name check before any copy is started) ...
    double *dst = nullptr;
    const int s = static_slot(name);
    if (s >= 0) dst = d->stat[s];
    else {
        const int var = var_index(name);
        if (var < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);   // equationset.cpp:181
        if (var == V_temp) dst = d->scratch_temp;
        else {
            const int ev = evolved_slot(var);
            if (ev < 0) return fail(SPRUCE_ERR_ARG, "<%s> is a derived variable and cannot be uploaded", name);
            if (ev == E_N) d->raw_rho = true;
            dst = d->Pset.p[ev];
        }
    }
    // ... then the copy is started, and while it is in flight the zero-plane bookkeeping scans the host plane (host_scan.hpp): planes whose transport can be
    // skipped exactly
    int rc = h2d_plane_begin(d, dst, host);
    if (rc) return rc;
    {
        const char *tracked[7] = {"mom_z", "bi_z", "be_x", "be_y", "be_z", "grav_x", "grav_y"};
        for (int b = 0; b < 7; b++) if (!strcmp(name, tracked[b])) {
            if (host_plane_nonzero(host, count)) d->nonzero_mask |= (1u << b); else d->nonzero_mask &= ~(1u << b);
        }
    }
    return h2d_plane_end(d);
}

int spruce_grid_download(spruce_domain *d, const char *name, double *host, size_t count)
{
    CHECK_DOM(d);
    if (!name || !host) return fail(SPRUCE_ERR_ARG, "null argument");
    if (count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "plane <%s>: expected %zu values, got %zu", name, (size_t)d->P.nx * d->P.ny, count);
    if (d->tf) return tf_download(d, name, host);
    if (d->e2) return e2_download(d, name, host);
    const int s = static_slot(name);
    if (s >= 0) return d2h_plane(d, host, d->stat[s]);
    const int var = var_index(name);
    if (var < 0) return fail(SPRUCE_ERR_ARG, "Variable name <%s> not recognized", name);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "download of <%s> before spruce_eqs_setup", name);
    const int ev = evolved_slot(var);
    if (ev > 0) return d2h_plane(d, host, d->Pset.p[ev]);
    DeriveArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = d->Pset.p[v];
    for (int v = 0; v < NSTATIC; v++) A.st[v] = d->stat[v];
    A.out = d->scratch_out; A.which = var;
    dim3 grid((d->P.ny + 255) / 256, d->P.nx);
    k_mhd_derive<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return d2h_plane(d, host, d->scratch_out);
}

int spruce_eqs_setup(spruce_domain *d)
{
    CHECK_DOM(d);
    if (!d->have_geom) return fail(SPRUCE_ERR_STATE, "spruce_set_cell_sizes must precede spruce_eqs_setup");
    d->raw_rho = true;
    // Ideal2F defaults to use_sub_cycling = true, whose time derivatives leave six 1x1 grids that abort at the ghost-zone mask
    // multiply (ideal2F.cpp:73-94, grid.cpp:84): only the non-sub-cycled Maxwell update can run at all
    if (d->tf && d->tf->use_sub_cycling) return fail(SPRUCE_ERR_UNSUPPORTED, "ideal_2F: use_sub_cycling = true aborts in the reference (size-mismatched grids); set use_sub_cycling = false");
    int rc = d->tf ? tf_launch_propagate(d, 1) : d->e2 ? e2_launch_propagate(d, 1) : launch_propagate(d, 1);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    d->is_setup = true;
    return SPRUCE_OK;
}

int spruce_eqs_propagate_changes(spruce_domain *d)
{
    CHECK_DOM(d);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "propagate before setup");
    int rc = d->tf ? tf_launch_propagate(d, 0) : d->e2 ? e2_launch_propagate(d, 0) : launch_propagate(d, 0);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return SPRUCE_OK;
}

int spruce_next_step_size(spruce_domain *d, double *step)
{
    CHECK_DOM(d);
    if (!d->is_setup || !step) return fail(SPRUCE_ERR_STATE, "step size before setup");
    StepCtl h;
    CUDA_TRY(cudaMemcpyAsync(&h, d->ctl, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    double m; memcpy(&m, &h.dtmin_bits, sizeof(m));
    *step = d->cfg.epsilon * m;
    return SPRUCE_OK;
}

// which storage currently plays the primary state (an euler step exchanges the roles of two sets on the host)
static const void *primary_id(const spruce_domain *d)
{
    if (d->tf) return d->tf->P.p[0];
    if (d->e2) return &d->e2->order[0] + d->e2->order[0];
    return d->Pset.p[0];
}
static void undo_euler_swap(spruce_domain *d)
{
    if (d->tf) std::swap(d->tf->P, d->tf->M);
    else if (d->e2) std::swap(d->e2->order[0], d->e2->order[1]);
    else std::swap(d->Pset, d->Mset);
}

int spruce_advance(spruce_domain *d, int n_steps, double max_time, double *dt_used, int *steps_done)
{
    CHECK_DOM(d);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "advance before setup");
    if (n_steps < 0) return fail(SPRUCE_ERR_ARG, "negative step count");
    if (d->cfg.n_ranks > 1 && !d->peers_connected) return fail(SPRUCE_ERR_STATE, "a slab of a decomposed domain advances either through spruce_mgpu_stage (caller-owned exchange) or, after spruce_mgpu_ipc_connect, through spruce_advance");
    if ((size_t)n_steps > d->dt_hist_cap) {
        if (d->dt_hist) cudaFree(d->dt_hist);
        d->dt_hist_cap = (size_t)n_steps + 64;
        CUDA_TRY(cudaMalloc(&d->dt_hist, d->dt_hist_cap * sizeof(double)));
    }
    StepCtl h0;
    CUDA_TRY(cudaMemcpyAsync(&h0, d->ctl, sizeof(h0), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    CUDA_TRY(cudaMemcpyAsync(&d->ctl->max_time, &max_time, sizeof(double), cudaMemcpyHostToDevice, d->stream));
    NvtxRange range_adv("spruce_advance");
    const bool plain = !d->tf && !d->e2 && plain_run(d);
    StepCtl h1;
    for (int first = 0;;) {
        int swaps = 0;
        for (int s = first; s < n_steps; s++) {
            const void *before = primary_id(d);
            int rc = d->tf ? tf_enqueue_step(d, s) : d->e2 ? e2_enqueue_step(d, s) : plain ? enqueue_step_plain(d, s, s == 0, s == n_steps - 1) : enqueue_step(d, s);
            if (rc) return rc;
            if (primary_id(d) != before) swaps++;
        }
        CUDA_TRY(cudaMemcpyAsync(&h1, d->ctl, sizeof(h1), cudaMemcpyDeviceToHost, d->stream));
        CUDA_TRY(cudaStreamSynchronize(d->stream));
        // steps enqueued after the run had stopped (max_time, sub-cycle budget) were no-ops on the device, but an euler step also exchanges the roles of the
        // two sets on the host: an odd number of such exchanges would leave the host naming the stale set as the primary state
        if (d->cfg.time_integrator == SPRUCE_TI_EULER && ((swaps - ((int)(h1.iter - h0.iter) - first)) & 1)) undo_euler_swap(d);
        if (!dev_subcycles(d)) break;
        // device-resident sub-cycle plan: the counts of the last planned step, the budget of the next advance; a step that needed more conduction sub-cycles than
        // were enqueued stopped the run before it changed anything (done = 3): raise the budget and enqueue the remaining steps again
        SubPlan sp;
        CUDA_TRY(cudaMemcpy(&sp, d->plan, sizeof(sp), cudaMemcpyDeviceToHost));
        if (h1.done != DONE_SUBCYCLE_BUDGET) {
            if (sp.tc_last >= 0) { d->tc_nsub = sp.tc_last; d->rl_nsub = sp.rl_last; }
            d->tc_budget = std::max(4, sp.tc_max + std::max(2, sp.tc_max / 4));
            sp.tc_max = 0;
            CUDA_TRY(cudaMemcpy(d->plan, &sp, sizeof(sp), cudaMemcpyHostToDevice));
            break;
        }
        if (sp.tc_need > (1 << 24)) return fail(SPRUCE_ERR_STATE, "thermal_conduction asks for %d sub-cycles in one step", sp.tc_need);
        d->tc_budget = sp.tc_need + std::max(2, sp.tc_need / 4);
        d->replans++;
        first = (int)(h1.iter - h0.iter);
        const int zero = 0;
        CUDA_TRY(cudaMemcpy(&d->ctl->done, &zero, sizeof(int), cudaMemcpyHostToDevice));
    }
    dump_marks(d);
    const int done = (int)(h1.iter - h0.iter);
    if (steps_done) *steps_done = done;
    if (dt_used && done > 0) CUDA_TRY(cudaMemcpy(dt_used, d->dt_hist, (size_t)done * sizeof(double), cudaMemcpyDeviceToHost));
    if (h1.done) { int zero = 0; CUDA_TRY(cudaMemcpy(&d->ctl->done, &zero, sizeof(int), cudaMemcpyHostToDevice)); }
    return SPRUCE_OK;
}

int spruce_get_time(spruce_domain *d, double *time, int64_t *iter)
{
    CHECK_DOM(d);
    StepCtl h;
    CUDA_TRY(cudaMemcpyAsync(&h, d->ctl, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    if (time) *time = h.time;
    if (iter) *iter = h.iter;
    return SPRUCE_OK;
}

int spruce_eqs_time_derivatives(spruce_domain *d, double *k_out, size_t count)
{
    CHECK_DOM(d);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "time derivatives before setup");
    if (d->tf) return tf_time_derivatives(d, k_out, count);
    if (d->e2) return e2_time_derivatives(d, k_out, count);
    const size_t np = (size_t)d->P.nx * d->P.ny;
    if (!k_out || count != NEV * np) return fail(SPRUCE_ERR_ARG, "k_out needs %zu values", NEV * np);
    int rc = ensure_rk4(d);
    if (rc) return rc;
    if ((rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 0.0, 0, KM_EXPORT))) return rc;
    for (int v = 0; v < NEV; v++) if ((rc = d2h_plane(d, k_out + v * np, d->K1set.p[v]))) return rc;
    return SPRUCE_OK;
}

int spruce_operator(spruce_domain *d, const char *op, int index, const double *q, const double *vel, double *out, size_t count)
{
    CHECK_DOM(d);
    if (!op || !q || !out) return fail(SPRUCE_ERR_ARG, "null argument");
    if (!d->have_geom) return fail(SPRUCE_ERR_STATE, "operators need the cell sizes");
    if (d->cfg.n_ranks > 1) return fail(SPRUCE_ERR_UNSUPPORTED, "stand-alone operators are single-rank");
    if (count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "plane size mismatch");
    const int code = !strcmp(op, "derivative1D") ? 0 : !strcmp(op, "secondDerivative1D") ? 1 : !strcmp(op, "laplacian") ? 2 : !strcmp(op, "transportDerivative1D") ? 3 : -1;
    if (code < 0) return fail(SPRUCE_ERR_ARG, "unknown operator <%s>", op);
    if (code == 3 && !vel) return fail(SPRUCE_ERR_ARG, "transportDerivative1D needs a velocity plane");
    if (index != 0 && index != 1) return fail(SPRUCE_ERR_ARG, "This function assumes two dimensions");
    int rc = ensure_rk4(d);                     // K1 planes double as operand scratch
    if (rc) return rc;
    if ((rc = h2d_plane(d, d->K1set.p[0], q))) return rc;
    if (vel && (rc = h2d_plane(d, d->K1set.p[1], vel))) return rc;
    OpArgs A{};
    A.q = d->K1set.p[0]; A.vel = d->K1set.p[1]; A.out = d->K1set.p[2]; A.op = code; A.index = index;
    dim3 grid((d->P.ny + 127) / 128, d->P.nx);
    k_operator<<<grid, 128, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return d2h_plane(d, out, d->K1set.p[2]);
}

// The two-operand operators of PlasmaDomain (derivs.cpp:216-220, 407-414, 471-474): each is the sum / difference of two single-direction passes of
// the operators above, formed on the host with the same IEEE addition the reference's Grid::operator+ / operator- performs per cell.
int spruce_operator2(spruce_domain *d, const char *op, const double *a, const double *b, const double *c, double *out, size_t count)
{
    CHECK_DOM(d);
    if (!op || !a || !b || !out) return fail(SPRUCE_ERR_ARG, "null argument");
    const int code = !strcmp(op, "divergence2D") ? 0 : !strcmp(op, "curl2D") ? 1 : !strcmp(op, "transportDivergence2D") ? 2 : -1;
    if (code < 0) return fail(SPRUCE_ERR_ARG, "unknown operator <%s>", op);
    if (code == 2 && !c) return fail(SPRUCE_ERR_ARG, "transportDivergence2D needs both velocity planes");
    if (count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "plane size mismatch");
    std::vector<double> first(count);
    int rc;
    if (code == 0) {          // derivative1D(a_x, 0) + derivative1D(a_y, 1)
        if ((rc = spruce_operator(d, "derivative1D", 0, a, nullptr, first.data(), count))) return rc;
        if ((rc = spruce_operator(d, "derivative1D", 1, b, nullptr, out, count))) return rc;
        for (size_t k = 0; k < count; k++) out[k] = first[k] + out[k];
    } else if (code == 1) {   // derivative1D(y, 0) - derivative1D(x, 1)
        if ((rc = spruce_operator(d, "derivative1D", 0, b, nullptr, first.data(), count))) return rc;
        if ((rc = spruce_operator(d, "derivative1D", 1, a, nullptr, out, count))) return rc;
        for (size_t k = 0; k < count; k++) out[k] = first[k] - out[k];
    } else {                  // transportDerivative1D(q, vel[0], 0) + transportDerivative1D(q, vel[1], 1)
        if ((rc = spruce_operator(d, "transportDerivative1D", 0, a, b, first.data(), count))) return rc;
        if ((rc = spruce_operator(d, "transportDerivative1D", 1, a, c, out, count))) return rc;
        for (size_t k = 0; k < count; k++) out[k] = first[k] + out[k];
    }
    return SPRUCE_OK;
}

int spruce_module_thermal_conduction(spruce_domain *d, int flux_saturation, int time_integrator, double epsilon, double dt_subcycle_min, double weakening_factor)
{
    CHECK_DOM(d);
    NOT_2F(d, "thermal_conduction");
    if (time_integrator < 0 || time_integrator > SPRUCE_TI_RK4) return fail(SPRUCE_ERR_ARG, "Invalid time integrator given for Thermal Conduction module");
    int rc = ensure_rk4(d);   // not needed for memory, keeps scratch planes uniform
    (void)rc;
    d->tc.flux_saturation = flux_saturation ? 1 : 0;
    d->tc.kappa = weakening_factor * kKappa0;
    d->tc.dt_subcycle_min = dt_subcycle_min;
    d->tc_integrator = time_integrator; d->tc_epsilon = epsilon;
    d->module_order.push_back(spruce_domain::MOD_TC);
    return SPRUCE_OK;
}
int spruce_module_radiative_losses(spruce_domain *d, int time_integrator, double cutoff_ramp, double cutoff_temp, double epsilon, int prevent_subcycling)
{
    CHECK_DOM(d);
    NOT_2F(d, "radiative_losses");
    if (time_integrator < 0 || time_integrator > SPRUCE_TI_RK4) return fail(SPRUCE_ERR_ARG, "Invalid time integrator given for Radiative Losses module");
    d->rl.integrator = time_integrator; d->rl.cutoff_ramp = cutoff_ramp; d->rl.cutoff_temp = cutoff_temp; d->rl.epsilon = epsilon;
    d->rl.prevent_subcycling = prevent_subcycling ? 1 : 0;
    d->module_order.push_back(spruce_domain::MOD_RL);
    return SPRUCE_OK;
}
int spruce_module_ambient_heating(spruce_domain *d, const double *heating, size_t count)
{
    CHECK_DOM(d);
    NOT_2F(d, "ambient_heating");
    if (!heating || count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "heating plane needs %zu values", (size_t)d->P.nx * d->P.ny);
    if (!d->heating) { int rc = alloc_plane(d, &d->heating); if (rc) return rc; }
    int rc = h2d_plane(d, d->heating, heating);
    if (rc) return rc;
    d->module_order.push_back(spruce_domain::MOD_AH);
    return SPRUCE_OK;
}
// ---- pointwise solar source terms -------------------------------------------------------------------------------------
namespace {
solar::Geom template_geom(const spruce_domain *d)
{
    solar::Geom g;
    g.nx_local = d->P.nx; g.ny = d->P.ny; g.row0 = d->P.row0; g.xdim = d->cfg.xdim; g.ydim = d->cfg.ydim;
    g.x_periodic = d->cfg.x_bound_1 == SPRUCE_BC_PERIODIC; g.y_periodic = d->cfg.y_bound_1 == SPRUCE_BC_PERIODIC;
    return g;
}
int add_source(spruce_domain *d, spruce_domain::SourceTerm &m, const std::vector<double> *p0, const std::vector<double> *p1)
{
    const std::vector<double> *pl[2] = {p0, p1};
    for (int k = 0; k < 2; k++) {
        if (!pl[k]) continue;
        int rc = alloc_plane(d, &m.plane[k]);
        if (rc) return rc;
        if ((rc = h2d_plane(d, m.plane[k], pl[k]->data()))) return rc;
    }
    d->sources.push_back(m);
    d->module_order.push_back(spruce_domain::MOD_SRC0 + (int)d->sources.size() - 1);
    return SPRUCE_OK;
}
}  // namespace

int spruce_module_ambient_heating_sink(spruce_domain *d, const double *reduction, size_t count)
{
    CHECK_DOM(d);
    NOT_2F(d, "ambient_heating_sink");
    const size_t np = (size_t)d->P.nx * d->P.ny;
    if (!reduction || count != np) return fail(SPRUCE_ERR_ARG, "reduction plane needs %zu values", np);
    spruce_domain::SourceTerm m; m.kind = SRC_SINK; m.ms_fraction = 0.5;
    const std::vector<double> p(reduction, reduction + np);
    return add_source(d, m, &p, nullptr);
}
int spruce_module_localized_heating(spruce_domain *d, double start_time, double duration, double max_heating_rate, double stddev_x, double stddev_y,
                                    double center_x, double center_y, double ramp_time)
{
    CHECK_DOM(d);
    NOT_2F(d, "localized_heating");
    spruce_domain::SourceTerm m; m.kind = SRC_HEATING; m.start = start_time; m.duration = duration; m.ramp_time = ramp_time;
    std::vector<double> p;
    solar::positive_template(template_geom(d), max_heating_rate, stddev_x, stddev_y, center_x, center_y, p);
    return add_source(d, m, &p, nullptr);
}
int spruce_module_mass_injection(spruce_domain *d, double start_time, double duration, double max_injection_rate, double stddev_x, double stddev_y,
                                 double center_x, double center_y)
{
    CHECK_DOM(d);
    NOT_2F(d, "mass_injection");
    spruce_domain::SourceTerm m; m.kind = SRC_MASS; m.start = start_time; m.duration = duration;
    std::vector<double> p;
    solar::positive_template(template_geom(d), max_injection_rate, stddev_x, stddev_y, center_x, center_y, p);
    return add_source(d, m, &p, nullptr);
}
int spruce_module_momentum_injection(spruce_domain *d, double start_time, double duration, double max_accel, double stddev_x, double stddev_y, double center_x,
                                     double center_y, double dir_x, double dir_y, double template_angle, int oscillatory, double oscillation_period)
{
    CHECK_DOM(d);
    NOT_2F(d, "momentum_injection");
    if (dir_x == 0.0 && dir_y == 0.0) return fail(SPRUCE_ERR_ARG, "Momentum Injection module must be given a nonzero acceleration direction");
    if (!(template_angle > -90.0 && template_angle < 90.0)) return fail(SPRUCE_ERR_ARG, "GaussianGridRotated requires angles between -90 deg and +90 deg");
    spruce_domain::SourceTerm m; m.kind = SRC_MOMENTUM; m.start = start_time; m.duration = duration; m.max_accel = max_accel;
    m.oscillatory = oscillatory ? 1 : 0; m.period = oscillation_period;
    std::vector<double> p[2];
    solar::momentum_templates(template_geom(d), stddev_x, stddev_y, center_x, center_y, dir_x, dir_y, template_angle, p[0], p[1]);
    return add_source(d, m, &p[0], &p[1]);
}
int spruce_module_div_cleaning(spruce_domain *d, double epsilon, double time_scale)
{
    CHECK_DOM(d);
    NOT_2F(d, "div_cleaning");
    if (!(epsilon > 0.0) || !(time_scale > 0.0)) return fail(SPRUCE_ERR_ARG, "div_cleaning needs epsilon > 0 and time_scale > 0");
    d->dc.epsilon = epsilon; d->dc.time_scale = time_scale;
    d->module_order.push_back(spruce_domain::MOD_DC);
    return SPRUCE_OK;
}
int spruce_module_field_heating(spruce_domain *d, double coeff, double current_pow, double b_pow, double n_pow, double roc_pow, int inactive_mode)
{
    CHECK_DOM(d);
    NOT_2F(d, "field_heating");
    d->fh.coeff = coeff; d->fh.current_pow = current_pow; d->fh.b_pow = b_pow; d->fh.n_pow = n_pow; d->fh.roc_pow = roc_pow; d->fh.inactive = inactive_mode ? 1 : 0;
    if (!d->fh.H) { int rc = alloc_plane(d, &d->fh.H); if (rc) return rc; }
    d->module_order.push_back(spruce_domain::MOD_FH);
    return SPRUCE_OK;
}
int spruce_module_anomalous_resistivity(spruce_domain *d, const double *pos_x, const double *pos_y, size_t count, const double *p, int n_params)
{
    CHECK_DOM(d);
    NOT_2F(d, "anomalous_resistivity");
    if (d->cfg.n_ranks > 1) return fail(SPRUCE_ERR_UNSUPPORTED, "anomalous_resistivity runs on a whole domain only (its flood fill and null-point search are global)");
    const size_t np = (size_t)d->cfg.xdim * d->cfg.ydim;
    if (!pos_x || !pos_y || count != np) return fail(SPRUCE_ERR_ARG, "pos_x / pos_y need the %zu values of the whole domain", np);
    if (!p || n_params != SPRUCE_AR_N_PARAMS) return fail(SPRUCE_ERR_ARG, "anomalous_resistivity takes %d parameters", SPRUCE_AR_N_PARAMS);
    if (d->ar.on) return fail(SPRUCE_ERR_ARG, "anomalous_resistivity is already configured");
    auto &A = d->ar;
    ar::Params &q = A.s.p;
    q.time_scale = p[0]; q.frob_coeff = p[1]; q.sigma = p[2]; q.safety = p[3]; q.smoothing = p[4] != 0.0; q.integrator = (int)p[5]; q.flood_fill = p[6] != 0.0;
    q.max_radius = p[7]; q.argmin_radius = p[8]; q.min_current = p[9]; q.ramp_length = p[10]; q.threshold = p[11]; q.model = (int)p[12]; q.gradient_correction = p[13] != 0.0;
    q.model_params[0] = p[14]; q.model_params[1] = p[15]; q.model_params[2] = p[16];
    if (q.integrator < 0 || q.integrator > 2) return fail(SPRUCE_ERR_ARG, "Invalid time integrator for anomalous resistivity module");
    if (q.model < 0 || q.model > 2) return fail(SPRUCE_ERR_ARG, "anomalous_resistivity: resistivity_model must be time_scale, syntelis_19 or ys_94");
    if (q.smoothing && !(q.sigma > 0.0)) return fail(SPRUCE_ERR_ARG, "anomalous_resistivity: smoothing_sigma must be positive");
    A.s.time_scale = q.time_scale;
    int rc;
    for (int k = 0; k < ar::P_COUNT; k++) if ((rc = alloc_plane(d, &A.planes[k]))) return rc;
    if ((rc = alloc_plane(d, &A.px)) || (rc = alloc_plane(d, &A.py)) || (rc = alloc_plane(d, &A.dtp))) return rc;
    if ((rc = h2d_plane(d, A.px, pos_x)) || (rc = h2d_plane(d, A.py, pos_y))) return rc;
    A.hpx.assign(pos_x, pos_x + np); A.hpy.assign(pos_y, pos_y + np);
    if (q.smoothing) {
        A.s.kr = ar::smoothing_radius(q.sigma);
        std::vector<double> w((size_t)(2 * A.s.kr + 1) * (2 * A.s.kr + 1));
        ar::smoothing_kernel(q.sigma, A.s.kr, w.data());
        CUDA_TRY(cudaMalloc(&A.kernel, w.size() * sizeof(double)));
        d->allocs.push_back(A.kernel);
        CUDA_TRY(cudaMemcpy(A.kernel, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
        A.s.kernel = A.kernel;
    }
    A.on = true; A.ready = false;
    d->module_order.push_back(spruce_domain::MOD_AR);
    // setupModule (:18-44) works on the set-up state: the template of iteration 0's output frame is this one.  A caller that configures modules
    // before spruce_eqs_setup gets it at the start of the first step instead.
    if (d->is_setup && (rc = ar_setup_run(d))) return rc;
    return SPRUCE_OK;
}
int spruce_module_anomalous_resistivity_state(spruce_domain *d, int *null_i, int *null_j, int *num_subcycles)
{
    CHECK_DOM(d);
    if (!d->ar.on) return fail(SPRUCE_ERR_ARG, "anomalous_resistivity is not configured");
    if (null_i) *null_i = d->ar.s.null_i;
    if (null_j) *null_j = d->ar.s.null_j;
    if (num_subcycles) *num_subcycles = d->ar.s.nsub;
    return SPRUCE_OK;
}
int spruce_module_boundary_outflow(spruce_domain *d, const double *pos_x, const double *pos_y, size_t count, double max_accel, double falloff_length, int boundary,
                                   int falloff_shape, double feather_length, int field_aligned_mode, int dynamic_mode, double dynamic_time, double dynamic_target_speed)
{
    CHECK_DOM(d);
    NOT_2F(d, "boundary_outflow");
    const size_t np = (size_t)d->cfg.xdim * d->cfg.ydim;                    // the GLOBAL position planes, on every slab (extrema and windows are global)
    if (!pos_x || !pos_y || count != np) return fail(SPRUCE_ERR_ARG, "pos_x / pos_y need the %zu values of the whole domain", np);
    if (boundary < 0 || boundary > 3) return fail(SPRUCE_ERR_ARG, "BoundaryOutflow boundary config must be {x,y}_bound_{1,2}");
    if (falloff_shape < 0 || falloff_shape > 2) return fail(SPRUCE_ERR_ARG, "BoundaryOutflow shape must be exp or gaussian or flat");
    auto &bo = d->bo;
    bo.max_accel = max_accel; bo.boundary = boundary; bo.field_aligned = field_aligned_mode ? 1 : 0; bo.dynamic = dynamic_mode ? 1 : 0;
    bo.dynamic_time = dynamic_time; bo.target = dynamic_target_speed;
    const int bcs[4] = {d->cfg.x_bound_1, d->cfg.x_bound_2, d->cfg.y_bound_1, d->cfg.y_bound_2};
    const solar::Window w = solar::outflow_bounds(d->cfg.xdim, d->cfg.ydim, bcs, boundary, SPRUCE_BC_PERIODIC, SPRUCE_BC_OPEN_MOC, HALO);
    std::vector<double> t;
    solar::outflow_template(d->cfg.xdim, d->cfg.ydim, w, pos_x, pos_y, falloff_length, feather_length, boundary, falloff_shape, t);
    const solar::Window m = solar::outflow_mean_window(d->cfg.ydim, w, pos_x, pos_y, falloff_length, feather_length, boundary);
    bo.win[0] = m.xl; bo.win[1] = m.xu; bo.win[2] = m.yl; bo.win[3] = m.yu;
    int rc;
    if (!bo.tmpl && (rc = alloc_plane(d, &bo.tmpl))) return rc;
    if ((rc = h2d_plane(d, bo.tmpl, t.data() + (size_t)d->P.row0 * d->cfg.ydim))) return rc;        // this slab's rows
    d->module_order.push_back(spruce_domain::MOD_BO);
    return SPRUCE_OK;
}
int spruce_module_boundary_outflow_state(spruce_domain *d, double *mean_outflow, double *curr_accel)
{
    CHECK_DOM(d);
    if (mean_outflow) *mean_outflow = d->bo.mean;
    if (curr_accel) *curr_accel = d->bo.accel;
    return SPRUCE_OK;
}
int spruce_module_physical_viscosity(spruce_domain *d, double coeff, const double *coeff_plane, size_t count, double epsilon, int heating_on, int force_on,
                                     int gradient_correction, int time_integrator, int inactive_mode)
{
    CHECK_DOM(d);
    NOT_2F(d, "physical_viscosity");
    if (!coeff_plane || count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "coefficient plane needs %zu values", (size_t)d->P.nx * d->P.ny);
    if (time_integrator != SPRUCE_TI_EULER && time_integrator != SPRUCE_TI_RK2) return fail(SPRUCE_ERR_ARG, "Invalid time integrator given for Physical Viscosity module");
    auto &pv = d->pv;
    int rc;
    if (!pv.cg) {
        if ((rc = alloc_plane(d, &pv.cg))) return rc;
        for (int s = 0; s < 2; s++) { for (int k = 0; k < 3; k++) if ((rc = alloc_plane(d, &pv.v[s][k]))) return rc; if ((rc = alloc_plane(d, &pv.T[s]))) return rc; }
        for (int k = 0; k < 3; k++) if ((rc = alloc_plane(d, &pv.bh[k]))) return rc;
    }
    if ((rc = h2d_plane(d, pv.cg, coeff_plane))) return rc;
    pv.coeff = coeff; pv.epsilon = epsilon; pv.heating_on = heating_on ? 1 : 0; pv.force_on = force_on ? 1 : 0;
    pv.gc = gradient_correction ? 1 : 0; pv.integrator = time_integrator; pv.inactive = inactive_mode ? 1 : 0;
    d->module_order.push_back(spruce_domain::MOD_PV);
    return SPRUCE_OK;
}
int spruce_module_viscosity(spruce_domain *d, int hv_time_integrator, double hv_epsilon, int gradient_correction)
{
    CHECK_DOM(d);
    if (d->tf) return fail(SPRUCE_ERR_UNSUPPORTED, "artificial_viscosity is not available with the ideal_2F equation set");
    if (hv_time_integrator < 0 || hv_time_integrator > SPRUCE_TI_RK4) return fail(SPRUCE_ERR_ARG, "Invalid hyperviscous time integrator given for Viscosity module");
    if (d->e2) return e2_visc_config(d, hv_time_integrator, hv_epsilon, gradient_correction);       // ideal_mhd_2E: mhd2e_host.cuh
    d->visc_hv_integrator = hv_time_integrator; d->visc_hv_epsilon = hv_epsilon; d->visc_gradient_correction = gradient_correction ? 1 : 0;
    for (int k = 0; k < 8; k++) if (!d->vscratch[k]) { int rc = alloc_plane(d, &d->vscratch[k]); if (rc) return rc; }
    if (!d->dt_plane) { int rc = alloc_plane(d, &d->dt_plane); if (rc) return rc; }
    d->module_order.push_back(spruce_domain::MOD_AV);
    return SPRUCE_OK;
}
int spruce_module_viscosity_term(spruce_domain *d, const char *visc_opt, double strength, const char *var_to_diff, const char *var_to_evol,
                                 const char *species, const double *strength_plane, size_t count)
{
    CHECK_DOM(d);
    if (d->tf) return fail(SPRUCE_ERR_UNSUPPORTED, "artificial_viscosity is not available with the ideal_2F equation set");
    if (!visc_opt || !var_to_diff || !var_to_evol) return fail(SPRUCE_ERR_ARG, "null argument");
    if (d->e2) return e2_visc_add_term(d, visc_opt, strength, var_to_diff, var_to_evol, species, strength_plane, count);
    if (!d->dt_plane) return fail(SPRUCE_ERR_STATE, "spruce_module_viscosity must precede its terms");
    spruce_domain::ViscTerm t{};
    t.opt = !strcmp(visc_opt, "local") ? 0 : !strcmp(visc_opt, "global") ? 1 : !strcmp(visc_opt, "boundary") ? 2 : !strcmp(visc_opt, "boundary_global") ? 3 : -1;
    if (t.opt < 0) return fail(SPRUCE_ERR_ARG, "Viscosity option must be global, local, or boundary.");
    t.strength = strength;
    t.var_diff = var_index(var_to_diff);
    t.var_evol = var_index(var_to_evol);
    if (t.var_diff < 0) return fail(SPRUCE_ERR_ARG, "Each variable to differentiate must be a valid variable within the chosen equation set.");
    if (t.var_evol < 0 || evolved_slot(t.var_evol) < 0) return fail(SPRUCE_ERR_ARG, "Each variable to evolve must be a valid evolved variable within the chosen equation set.");
    t.species = (species && species[0]) ? species[0] : 'i';
    if (t.opt >= 2) {
        if (!strength_plane || count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "boundary viscosity needs its strength profile (%zu values)", (size_t)d->P.nx * d->P.ny);
        int rc = alloc_plane(d, &t.strength_plane);
        if (rc) return rc;
        if ((rc = h2d_plane(d, t.strength_plane, strength_plane))) return rc;
    }
    // a viscosity term may feed the z system from any variable: mom_z / bi_z can then leave zero even if they were uploaded as zero planes
    if (evolved_slot(t.var_evol) == E_MZ) d->nonzero_mask |= 0x01u;
    if (evolved_slot(t.var_evol) == E_BZ) d->nonzero_mask |= 0x02u;
    d->visc.push_back(t);
    return SPRUCE_OK;
}
int spruce_eqs_ideal_mhd_options(spruce_domain *d, double global_viscosity)
{
    CHECK_DOM(d);
    if (d->tf || d->e2) return fail(SPRUCE_ERR_STATE, "ideal_mhd options on a domain with another equation set");
    d->global_viscosity = global_viscosity;          // only the open_moc boundary reads it (idealmhd.cpp:90)
    return SPRUCE_OK;
}
int spruce_eqs_ideal_mhd_moc_limiting(spruce_domain *d, int b_limiting, double b_lower_lim, double b_upper_lim, int mom_limiting, double mom_lower_lim, double mom_upper_lim)
{
    CHECK_DOM(d);
    if (d->tf || d->e2) return fail(SPRUCE_ERR_STATE, "ideal_mhd options on a domain with another equation set");
    if (b_limiting && !(b_lower_lim <= 1.0 && b_upper_lim >= 1.0)) return fail(SPRUCE_ERR_ARG, "MoC B field thresholds: lower <= 1.0 <= upper");      // idealmhd.cpp:21,25
    if (mom_limiting && !(mom_lower_lim <= 1.0 && mom_upper_lim >= 1.0)) return fail(SPRUCE_ERR_ARG, "MoC momentum thresholds: lower <= 1.0 <= upper");   // :29,33
    if (b_limiting && d->cfg.y_bound_2 == SPRUCE_BC_OPEN_MOC && d->cfg.xdim - 2 - HALO >= d->cfg.ydim)
        return fail(SPRUCE_ERR_ARG, "moc_b_limiting on y_bound_2 reads column xdim-2-N_GHOST (idealmhd.cpp:154): outside this grid, the reference aborts");
    d->moc_lim.b_on = b_limiting ? 1 : 0; d->moc_lim.b_lo = b_lower_lim; d->moc_lim.b_hi = b_upper_lim;
    d->moc_lim.mom_on = mom_limiting ? 1 : 0; d->moc_lim.mom_lo = mom_lower_lim; d->moc_lim.mom_hi = mom_upper_lim;
    return SPRUCE_OK;
}
int spruce_eqs_ideal2f_options(spruce_domain *d, int use_sub_cycling, int remove_curl_terms)
{
    CHECK_DOM(d);
    if (!d->tf) return fail(SPRUCE_ERR_STATE, "ideal_2F options on a domain with another equation set");
    d->tf->use_sub_cycling = use_sub_cycling ? 1 : 0;
    d->tf->remove_curl_terms = remove_curl_terms ? 1 : 0;
    return SPRUCE_OK;
}
int spruce_module_eic_thermalization(spruce_domain *d)
{
    CHECK_DOM(d);
    // setupModule looks its grids up by name in the equation set and aborts when one is missing (eic_thermalization.cpp:13-24)
    // -- ideal_2F and ideal_mhd_2E have all four (n, e_temp, e_thermal_energy, i_thermal_energy), ideal_mhd has no e_temp
    if (d->e2) { d->e2->g.eic = 1; return SPRUCE_OK; }
    if (!d->tf) return fail(SPRUCE_ERR_ARG, "Grid <e_temp> was not found within the EquationSet.");
    d->tf->eic = 1;
    return SPRUCE_OK;
}
// inactive_mode = true of thermal_conduction / radiative_losses: the module is evaluated (sub-cycle count, output_to_file planes, cumulative planes of multispecies_mode)
// and nothing is applied to the state (thermalconduction.cpp:109, radiativelosses.cpp:98)
int spruce_module_inactive_mode(spruce_domain *d, const char *module, int on)
{
    CHECK_DOM(d);
    if (!module) return fail(SPRUCE_ERR_ARG, "null argument");
    NOT_2F(d, "inactive_mode");
    int rc;
    if (on && !d->old_e && (rc = alloc_plane(d, &d->old_e))) return rc;
    if (!strcmp(module, "thermal_conduction")) d->tc_inactive = on != 0;
    else if (!strcmp(module, "radiative_losses")) d->rl_inactive = on != 0;
    else return fail(SPRUCE_ERR_ARG, "module <%s> has no inactive_mode here", module);
    return SPRUCE_OK;
}
// multispecies_mode = true (fileio.cpp:322): the three cumulative planes exist from now on, zero (plasmadomain.cpp:55-59)
int spruce_multispecies_mode(spruce_domain *d, int on)
{
    CHECK_DOM(d);
    NOT_2F(d, "multispecies_mode");
    int rc;
    if (on) {
        for (int k = 0; k < 3; k++) if (!d->ms_cum[k] && (rc = alloc_plane(d, &d->ms_cum[k]))) return rc;
        if (!d->old_e && (rc = alloc_plane(d, &d->old_e))) return rc;
    }
    d->ms_on = on != 0;
    return SPRUCE_OK;
}
// the reset after every stored frame (evolution.cpp:36-41)
int spruce_multispecies_reset(spruce_domain *d)
{
    CHECK_DOM(d);
    if (!d->ms_on) return SPRUCE_OK;
    for (int k = 0; k < 3; k++) CUDA_TRY(cudaMemsetAsync(d->ms_cum[k] - d->row_off, 0, d->plane_doubles * sizeof(double), d->stream));
    return SPRUCE_OK;
}
// ms_electron_heating_fraction of a module (thermalconduction.cpp:27,40 and the like): 0 <= f <= 1; call after the module is configured
int spruce_module_ms_fraction(spruce_domain *d, const char *module, double f)
{
    CHECK_DOM(d);
    if (!module) return fail(SPRUCE_ERR_ARG, "null argument");
    if (!(f >= 0.0 && f <= 1.0)) return fail(SPRUCE_ERR_ARG, "%s MS electron heating fraction must be between 0 and 1", module);
    if (!strcmp(module, "thermal_conduction")) d->ms_frac_tc = f;
    else if (!strcmp(module, "radiative_losses")) d->ms_frac_rl = f;
    else if (!strcmp(module, "ambient_heating")) d->ms_frac_ah = f;
    else if (!strcmp(module, "physical_viscosity")) d->ms_frac_pv = f;
    else if (!strcmp(module, "ambient_heating_sink") || !strcmp(module, "localized_heating")) {
        const int kind = !strcmp(module, "localized_heating") ? SRC_HEATING : SRC_SINK;
        bool any = false;
        for (auto &m : d->sources) if (m.kind == kind) { m.ms_fraction = f; any = true; }
        if (!any) return fail(SPRUCE_ERR_STATE, "ms_electron_heating_fraction of <%s> before the module is configured", module);
    } else return fail(SPRUCE_ERR_ARG, "module <%s> has no ms_electron_heating_fraction", module);
    return SPRUCE_OK;
}
int spruce_module_output_to_file(spruce_domain *d, const char *module, int on)
{
    CHECK_DOM(d);
    if (!module) return fail(SPRUCE_ERR_ARG, "null argument");
    NOT_2F(d, "module diagnostic planes");
    int rc;
    if (on && !d->old_e && (rc = alloc_plane(d, &d->old_e))) return rc;
    if (!strcmp(module, "thermal_conduction")) {
        d->tc_output = on != 0;
        if (on && !d->tc_avg && ((rc = alloc_plane(d, &d->tc_avg)) || (rc = alloc_plane(d, &d->tc_sat)))) return rc;
    } else if (!strcmp(module, "radiative_losses")) {
        d->rl_output = on != 0;
        if (on && !d->rl_avg && (rc = alloc_plane(d, &d->rl_avg))) return rc;
    } else if (!strcmp(module, "anomalous_resistivity")) {                                          // anomalousresistivity.cpp:320-329
        d->ar.output = on != 0;
        if (on && !d->ar.avg && ((rc = alloc_plane(d, &d->ar.avg)) || (rc = alloc_plane(d, &d->ar.prod)))) return rc;
    } else if (!strcmp(module, "artificial_viscosity")) {                                           // viscosity.cpp:351-376; call after the terms are configured
        d->visc_output = on != 0;
        if (on) for (auto &t : d->visc) if (!t.o_lap && ((rc = alloc_plane(d, &t.o_dqdt)) || (rc = alloc_plane(d, &t.o_lap)) || (rc = alloc_plane(d, &t.o_dt)))) return rc;
    } else if (!strcmp(module, "physical_viscosity")) {                                             // physicalviscosity.cpp:292-308
        d->pv.output = on != 0;
        if (on && !d->pv.avg[0]) for (int k = 0; k < 4; k++) if ((rc = alloc_plane(d, &d->pv.avg[k]))) return rc;
    } else return fail(SPRUCE_ERR_ARG, "no diagnostic planes for module <%s>", module);
    return SPRUCE_OK;
}
int spruce_module_output(spruce_domain *d, const char *name, double *host, size_t count)
{
    CHECK_DOM(d);
    if (!name || !host) return fail(SPRUCE_ERR_ARG, "null argument");
    if (count != (size_t)d->P.nx * d->P.ny) return fail(SPRUCE_ERR_ARG, "plane size mismatch");
    const double *src = !strcmp(name, "thermal_conduction") ? d->tc_avg : !strcmp(name, "flux_saturation") ? d->tc_sat : !strcmp(name, "rad") ? d->rl_avg : nullptr;
    if (d->ms_on) {                                                                                  // multispecies_mode: what has accumulated since the last reset
        const char *msn[3] = {"cumulative_electron_heating", "cumulative_ion_heating", "cumulative_joule_heating"};
        for (int k = 0; k < 3; k++) if (!strcmp(name, msn[k])) src = d->ms_cum[k];
    }
    if (d->pv.output && d->pv.avg[0]) {                                                             // zero planes before the first step, like the reference's (:38-39, :295)
        const char *pvn[4] = {"viscous_heating", "viscous_force_x", "viscous_force_y", "viscous_force_z"};
        for (int k = 0; k < 4; k++) if (!strcmp(name, pvn[k])) src = d->pv.avg[k];
    }
    if (d->visc_output) {                                                                            // "visc_dqdt:<i>", "visc_lap:<i>", "visc_dt:<i>" of term i (config order)
        const char *pre[3] = {"visc_dqdt:", "visc_lap:", "visc_dt:"};
        for (int w = 0; w < 3; w++) if (!strncmp(name, pre[w], strlen(pre[w]))) {
            const int i = atoi(name + strlen(pre[w]));
            if (i < 0 || i >= (int)d->visc.size() || !d->visc[i].o_lap) return fail(SPRUCE_ERR_ARG, "no viscosity term <%s>", name);
            src = w == 0 ? d->visc[i].o_dqdt : w == 1 ? d->visc[i].o_lap : d->visc[i].o_dt;
        }
        if (!strncmp(name, "visc_str:", 9)) {                                                         // m_grids_strength[i] (:195-196): zero until the first evaluation of any term
            const int i = atoi(name + 9);
            if (i < 0 || i >= (int)d->visc.size()) return fail(SPRUCE_ERR_ARG, "no viscosity term <%s>", name);
            const auto &t = d->visc[i];
            if (d->visc_evaluated && (t.opt == 2 || t.opt == 3)) return d2h_plane(d, host, t.strength_plane);
            const double v = d->visc_evaluated ? t.strength : 0.0;
            for (size_t k = 0; k < count; k++) host[k] = v;
            return SPRUCE_OK;
        }
    }
    if (!strcmp(name, "field_heating") && d->fh.H) src = d->fh.H;                                  // fieldheating.cpp:73-80: mask*(dt*heating) of the last step, zero before the first
    if (d->ar.on && d->ar.output) {
        if (!strcmp(name, "anomalous_template")) src = d->ar.planes[ar::P_TMPL];
        else if (!strcmp(name, "joule_heating")) src = d->ar.avg;
        else if (!strcmp(name, "anomalous_diffusivity")) {                                            // anomalous_template*diffusivity as they stand after the last evaluation
            const dim3 grid((d->P.ny + 255) / 256, d->P.nx);
            k_plane_product<<<grid, 256, 0, d->stream>>>(d->P, d->ar.prod, d->ar.planes[ar::P_TMPL], d->ar.planes[ar::P_DIFF]);
            d->launches++;
            CUDA_TRY(cudaGetLastError());
            src = d->ar.prod;
        }
    }
    if (!src) return fail(SPRUCE_ERR_STATE, "diagnostic plane <%s> is not enabled (spruce_module_output_to_file)", name);
    return d2h_plane(d, host, src);
}
int spruce_module_subcycles(spruce_domain *d, const char *which, int *count)
{
    CHECK_DOM(d);
    if (!which || !count) return fail(SPRUCE_ERR_ARG, "null argument");
    if (!strcmp(which, "thermal_conduction")) *count = d->tc_nsub;
    else if (!strcmp(which, "radiative_losses")) *count = d->rl_nsub;
    else if (!strcmp(which, "physical_viscosity")) *count = d->pv.nsub;
    else if (!strcmp(which, "div_cleaning")) *count = d->dc.nsub;
    else if (!strcmp(which, "anomalous_resistivity")) *count = d->ar.s.nsub;
    // the device-resident sub-cycle plan (SPRUCE_DEVICE_SUBCYCLES=1): is it in use for the configured module set, the conduction sub-cycles the next advance enqueues per
    // step, and how often an advance had to raise that number and enqueue again
    else if (!strcmp(which, "device_plan")) *count = dev_subcycles(d) ? 1 : 0;
    else if (!strcmp(which, "device_plan_budget")) *count = d->tc_budget;
    else if (!strcmp(which, "device_plan_replans")) *count = d->replans;
    else return fail(SPRUCE_ERR_ARG, "no sub-cycling module named <%s>", which);
    return SPRUCE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU: the caller owns the communicator (torch.distributed / NCCL) and moves the staging buffers
// ---------------------------------------------------------------------------------------------------------------
static PlaneSet *set_by_id(spruce_domain *d, int which) { return which == 0 ? &d->Pset : which == 1 ? &d->Mset : which == 2 ? &d->M2set : nullptr; }

int spruce_halo_buffers(spruce_domain *d, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes)
{
    CHECK_DOM(d);
    NOT_2F(d, "slab decomposition");
    if (!d->halo[0]) {
        d->halo_doubles = (size_t)NEV * HALO * d->P.pitch;
        for (int k = 0; k < 4; k++) {
            CUDA_TRY(cudaMalloc(&d->halo[k], d->halo_doubles * sizeof(double)));
            CUDA_TRY(cudaMemsetAsync(d->halo[k], 0, d->halo_doubles * sizeof(double), d->stream));
            d->allocs.push_back(d->halo[k]);
        }
    }
    if (send_lo) *send_lo = d->halo[0];
    if (send_hi) *send_hi = d->halo[1];
    if (recv_lo) *recv_lo = d->halo[2];
    if (recv_hi) *recv_hi = d->halo[3];
    if (bytes) *bytes = d->halo_doubles * sizeof(double);
    return SPRUCE_OK;
}
static int halo_copy(spruce_domain *d, int which, int unpack)
{
    PlaneSet stat_view;                    // which == 3: the static planes (be_x, be_y, be_z are transported: they need halo rows too)
    for (int v = 0; v < NEV; v++) stat_view.p[v] = d->stat[v < NSTATIC ? v : 0];
    PlaneSet *s = which == 3 ? &stat_view : set_by_id(d, which);
    if (!s || !s->p[0]) return fail(SPRUCE_ERR_ARG, "state set %d does not exist", which);
    int rc = spruce_halo_buffers(d, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    HaloArgs A{};
    for (int v = 0; v < NEV; v++) A.U[v] = s->p[v];
    A.lo = unpack ? d->halo[2] : d->halo[0];
    A.hi = unpack ? d->halo[3] : d->halo[1];
    A.unpack = unpack;
    dim3 grid((d->P.pitch + 255) / 256, NEV * HALO);
    k_halo_copy<<<grid, 256, 0, d->stream>>>(d->P, A);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int spruce_mgpu_pack(spruce_domain *d, int which) { CHECK_DOM(d); NOT_2F(d, "slab decomposition"); return halo_copy(d, which, 0); }
int spruce_mgpu_unpack(spruce_domain *d, int which) { CHECK_DOM(d); NOT_2F(d, "slab decomposition"); return halo_copy(d, which, 1); }

int spruce_mgpu_n_stages(spruce_domain *d, int *n)
{
    CHECK_DOM(d);
    if (!n) return fail(SPRUCE_ERR_ARG, "null argument");
    *n = d->cfg.time_integrator == SPRUCE_TI_EULER ? 1 : d->cfg.time_integrator == SPRUCE_TI_RK2 ? 2 : 4;
    return SPRUCE_OK;
}
// Runs RK stage `stage` and packs the halo rows of the set it produced; *out_set (returned through the status-free
// convention: the id is also the return value of spruce_mgpu_stage_output) tells the caller which set to unpack into.
static int stage_output_set(const spruce_domain *d, int stage)
{
    const int ti = d->cfg.time_integrator;
    if (ti == SPRUCE_TI_EULER) return 0;
    if (ti == SPRUCE_TI_RK2) return stage == 0 ? 1 : 0;
    return stage == 0 ? 1 : stage == 1 ? 2 : stage == 2 ? 1 : 0;
}
int spruce_mgpu_stage(spruce_domain *d, int stage)
{
    if (d && (d->moc_lim.b_on || d->moc_lim.mom_on)) return fail(SPRUCE_ERR_UNSUPPORTED, "the open_moc limiters are not available through spruce_mgpu_stage (use the peer transport)");
    CHECK_DOM(d);
    NOT_2F(d, "slab decomposition");
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "stage before setup");
    d->in_mgpu_stage_api = true;            // the caller reduces dt itself: every cell's dt is evaluated (no skip test)
    int rc;
    const int ti = d->cfg.time_integrator;
    if (ti == SPRUCE_TI_EULER) {
        if (stage != 0) return fail(SPRUCE_ERR_ARG, "euler has one stage");
        if ((rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 1.0, 1, KM_NONE))) return rc;
        std::swap(d->Pset, d->Mset);
        if ((rc = launch_ghosts(d, d->Pset, 1))) return rc;
    } else if (ti == SPRUCE_TI_RK2) {
        if (stage == 0) { if ((rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_NONE)) || (rc = launch_ghosts(d, d->Mset, 0))) return rc; }
        else if (stage == 1) { if ((rc = launch_stage(d, d->Mset, d->Pset, d->Pset, 1.0, 1, KM_NONE)) || (rc = launch_ghosts(d, d->Pset, 1))) return rc; }
        else return fail(SPRUCE_ERR_ARG, "rk2 has two stages");
    } else {
        if ((rc = ensure_rk4(d))) return rc;
        if (stage == 0) { if ((rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_STORE_K1)) || (rc = launch_ghosts(d, d->Mset, 0))) return rc; }
        else if (stage == 1) { if ((rc = launch_stage(d, d->Mset, d->Pset, d->M2set, 0.5, 0, KM_STORE_K2)) || (rc = launch_ghosts(d, d->M2set, 0))) return rc; }
        else if (stage == 2) { if ((rc = launch_stage(d, d->M2set, d->Pset, d->Mset, 1.0, 0, KM_ADD_K2)) || (rc = launch_ghosts(d, d->Mset, 0))) return rc; }
        else if (stage == 3) { if ((rc = launch_stage(d, d->Mset, d->Pset, d->Pset, 1.0, 1, KM_FINAL)) || (rc = launch_ghosts(d, d->Pset, 1))) return rc; }
        else return fail(SPRUCE_ERR_ARG, "rk4 has four stages");
    }
    return halo_copy(d, stage_output_set(d, stage), 0);
}
int spruce_mgpu_stage_output(spruce_domain *d, int stage, int *which)
{
    CHECK_DOM(d);
    if (!which) return fail(SPRUCE_ERR_ARG, "null argument");
    *which = stage_output_set(d, stage);
    return SPRUCE_OK;
}
int spruce_mgpu_dt_min_ptr(spruce_domain *d, void **device_double)
{
    CHECK_DOM(d);
    if (!device_double) return fail(SPRUCE_ERR_ARG, "null argument");
    *device_double = (void *)&d->ctl->dtmin_bits;     // bits of a positive double: min over doubles == min over bit patterns
    return SPRUCE_OK;
}
int spruce_mgpu_begin_step(spruce_domain *d)
{
    CHECK_DOM(d);
    NOT_2F(d, "slab decomposition");
    if ((size_t)1 > d->dt_hist_cap) { d->dt_hist_cap = 64; CUDA_TRY(cudaMalloc(&d->dt_hist, d->dt_hist_cap * sizeof(double))); }
    k_step_begin<<<1, 1, 0, d->stream>>>(d->ctl, d->dt_hist, 0);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}
int spruce_mgpu_end_step(spruce_domain *d)
{
    CHECK_DOM(d);
    NOT_2F(d, "slab decomposition");
    k_step_end<<<1, 1, 0, d->stream>>>(d->ctl);
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPRUCE_OK;
}

// ---- peer-store transport: export my segment, map the neighbours', exchange the initial halos
int spruce_mgpu_ipc_export(spruce_domain *d, void *handle64)
{
    CHECK_DOM(d);
    if (!handle64) return fail(SPRUCE_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int rc = ensure_segment(d);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, d->seg));
    memcpy(handle64, &h, sizeof(h));
    return SPRUCE_OK;
}
int spruce_mgpu_ipc_connect(spruce_domain *d, const void *handles, int n_handles)
{
    CHECK_DOM(d);
    if (!handles || n_handles != d->cfg.n_ranks || n_handles > MAX_RANKS) return fail(SPRUCE_ERR_ARG, "need one 64-byte handle per rank (%d ranks, at most %d)", d->cfg.n_ranks, MAX_RANKS);
    int rc = ensure_segment(d);
    if (rc) return rc;
    for (int r = 0; r < n_handles; r++) {
        if (r == d->cfg.rank) { d->peer_seg[r] = d->seg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof(h));
        CUDA_TRY(cudaIpcOpenMemHandle(&d->peer_seg[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    d->peers_connected = true;
    return SPRUCE_OK;
}
// after spruce_eqs_setup on every rank: halo rows of the static planes and of the primary state, global dt minimum
int spruce_mgpu_initial_exchange(spruce_domain *d)
{
    CHECK_DOM(d);
    if (!d->is_setup) return fail(SPRUCE_ERR_STATE, "initial exchange before setup");
    if (d->tf) { int rc2 = tf_initial_exchange(d); if (rc2) return rc2; CUDA_TRY(cudaStreamSynchronize(d->stream)); return SPRUCE_OK; }
    if (d->e2) { int rc2 = e2_initial_exchange(d); if (rc2) return rc2; CUDA_TRY(cudaStreamSynchronize(d->stream)); return SPRUCE_OK; }
    double *stat_view[NEV];
    for (int v = 0; v < NEV; v++) stat_view[v] = d->stat[v < NSTATIC ? v : 0];
    int rc;
    if ((rc = peer_exchange(d, stat_view, nullptr))) return rc;
    if ((rc = peer_exchange(d, d->Pset.p, nullptr))) return rc;
    if ((rc = peer_dt_allgather(d))) return rc;
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return SPRUCE_OK;
}

// Zero-plane knowledge must be GLOBAL with a slab decomposition (a neighbour's halo rows may be non-zero): every rank
// reports its local mask and sets the OR over all ranks.  bit 0 mom_z, 1 bi_z, 2 be_x, 3 be_y, 4 be_z.
int spruce_plane_activity(spruce_domain *d, int *local_mask, int set_global_mask)
{
    CHECK_DOM(d);
    if (local_mask) *local_mask = (int)(d->nonzero_mask & 0x1Fu);
    if (set_global_mask >= 0) d->nonzero_mask = ((unsigned)set_global_mask & 0x1Fu) | (d->nonzero_mask & 0x60u);
    return SPRUCE_OK;
}

int spruce_stream(spruce_domain *d, void **stream) { CHECK_DOM(d); if (!stream) return fail(SPRUCE_ERR_ARG, "null"); *stream = (void *)d->stream; return SPRUCE_OK; }
int spruce_synchronize(spruce_domain *d) { CHECK_DOM(d); CUDA_TRY(cudaStreamSynchronize(d->stream)); return SPRUCE_OK; }
int spruce_launch_count(spruce_domain *d, int64_t *count) { CHECK_DOM(d); if (count) *count = d->launches; return SPRUCE_OK; }

int spruce_time_stage_kernel(spruce_domain *d, int reps, float *ms_mean)
{
    CHECK_DOM(d);
    NOT_2F(d, "the stage-kernel timer");
    if (!d->is_setup || reps < 1 || !ms_mean) return fail(SPRUCE_ERR_STATE, "stage timing needs a set-up domain");
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
    int rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_NONE);   // warm-up
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(a, d->stream));
    for (int k = 0; k < reps; k++) if ((rc = launch_stage(d, d->Pset, d->Pset, d->Mset, 0.5, 0, KM_NONE))) return rc;
    CUDA_TRY(cudaEventRecord(b, d->stream));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_mean = ms / reps;
    return SPRUCE_OK;
}

}  // extern "C"
