/* spruce_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the
 * product path (spruce_b200/).  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg.
 *
 * A plain-C, op-by-op CPU restatement of the reference's (gszypko/spruce) per-timestep advance for
 * 2-D ideal MHD on a non-uniform rectilinear grid, written from the reference's behaviour:
 *   every full-plane operator materialises its result (zeros outside its index range) exactly like the
 *   reference's Grid temporaries, every arithmetic operation is a separately rounded IEEE-754 double
 *   operation in the reference's order (compile with -ffp-contract=off; the reference binary has no FMA).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file bit-for-bit against outputs of the
 * unmodified reference binary (oracle/_ref/run) committed under tests/golden/ (generator:
 * tests/golden/make_golden.py).  The reference itself ships no tests or golden vectors.
 *
 * The two-fluid equation set (Ideal2F) and EIC thermalization are restated in ideal2f_oracle.inc (included at the end of
 * this file) and pinned the same way against the tf_*.npz fixtures.
 *
 * Reference citations are given per function as file:line relative to the reference tree.
 * Layout: plane[i*ny + j], i = x index, j = y index (source/mhd/grid.cpp:516-526).
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define N_GHOST 2                 /* source/constants.hpp:4 */
#define K_B 1.3807e-16            /* source/constants.hpp:8 */
#define M_ELECTRON 9.1094e-28     /* source/constants.hpp:9 */
#define KAPPA_0 1.0e-6            /* source/constants.hpp:15 */
#define PI 3.14159265358979323846 /* source/constants.hpp:16 */
#define FOURPI (4.0 * PI)

/* plasmadomain.hpp:22 enum order */
enum { BC_PERIODIC = 0, BC_OPEN = 1, BC_FIXED = 2, BC_REFLECT = 3, BC_OPEN_MOC = 4, BC_OPEN_UCNP = 5 };
/* plasmadomain.hpp:29 */
enum { TI_EULER = 0, TI_RK2 = 1, TI_RK4 = 2 };
/* idealmhd.hpp:24-26 */
enum { V_rho, V_temp, V_mom_x, V_mom_y, V_mom_z, V_bi_x, V_bi_y, V_bi_z, V_grav_x, V_grav_y,
       V_n, V_press, V_thermal_energy, V_v_x, V_v_y, V_v_z, V_kinetic_energy,
       V_b_x, V_b_y, V_b_z, V_b_mag, V_b_hat_x, V_b_hat_y, V_b_hat_z, V_dt, NV };
/* idealmhd.hpp:32-34 */
static const int EVOLVED[8] = { V_rho, V_mom_x, V_mom_y, V_mom_z, V_thermal_energy, V_bi_x, V_bi_y, V_bi_z };
#define NEV 8

typedef struct {
    /* thermal_conduction (thermalconduction.hpp) */
    int tc_on, tc_flux_saturation, tc_integrator; double tc_epsilon, tc_dt_subcycle_min, tc_weakening; int tc_nsub;
    double *tc_avg, *tc_sat, *rl_avg;   /* output_to_file planes (thermalconduction.cpp:101-104, radiativelosses.cpp:93), kept for the tests */
    /* radiative_losses (radiativelosses.hpp) */
    int rl_on, rl_integrator, rl_prevent_subcycling; double rl_cutoff_ramp, rl_cutoff_temp, rl_epsilon; int rl_nsub;
    /* ambient_heating (ambientheating.hpp) */
    int ah_on; double *ah_heating;
    /* artificial_viscosity (source/modules/viscosity.hpp) : up to 8 terms */
    int av_on, av_nterms, av_hv_integrator, av_gradient_correction; double av_hv_epsilon;
    int av_opt[8];            /* 0 local, 1 global, 2 boundary, 3 boundary_global */
    double av_strength[8];
    int av_diff[8], av_evol[8];   /* variable indices (equation-set numbering) */
    int av_species[8];        /* 'i' or 'e' */
    double *av_strength_grid[8];  /* boundary options: profile built by the caller (host libm exp) */
    double *av_out[4][8];         /* the planes fileOutput appends per term (viscosity.cpp:351-376): [0] m_grids_dqdt, [1] m_grids_lap, [2] m_grids_strength, [3] m_grids_dt --
                                   * whatever the LAST evaluation of the term left there (zero planes before the first one, :99-102) */
    /* physical_viscosity (source/modules/solar/physicalviscosity.hpp) */
    int pv_on, pv_heating_on, pv_force_on, pv_gc, pv_integrator, pv_inactive, pv_nsub; double pv_coeff, pv_epsilon; double *pv_cg;
    /* multispecies_mode (plasmadomain.hpp:134-135): cumulative electron / ion / joule heating between outputs, fed by the modules with their
     * ms_electron_heating_fraction; ms_frac is indexed by module: 1 tc, 2 rl, 3 ah, 5 pv, 6 ambient_heating_sink, 7 localized_heating */
    int ms_on; double ms_frac[8]; double *ms_cum[3];
    int tc_inactive, rl_inactive;       /* inactive_mode of thermal_conduction / radiative_losses: everything is evaluated (output planes, cumulative planes), nothing applied */
    double *pv_avg[4];                  /* output_to_file planes: viscous_heating, viscous_force_x/y/z (physicalviscosity.cpp:151-152,166,170,218,222,292-308) */
    void *anom;                   /* anomalous_resistivity (anomalous_resistivity_oracle.inc); order id 13 */
    void *small[8]; int n_small;  /* small solar modules (solar_small_modules_oracle.inc); order id = 100 + index */
    int order[16]; int n_modules; /* module ids in config order: 1=tc 2=rl 3=ah 4=av 5=pv, 100+k = small module k */
} modules_t;

typedef struct oracle {
    int nx, ny, n;
    int xb1, xb2, yb1, yb2, integrator;
    int xl, xu, yl, yu;
    double m_i, gamma, epsilon, n_min, T_min, e_min, open_strength, open_decay;
    double global_viscosity;    /* IdealMHD m_global_viscosity (idealmhd.hpp:48): only the open_moc boundary uses it */
    int moc_b_limiting, moc_mom_limiting;                      /* idealmhd.hpp:59-64 */
    double moc_b_lim[2], moc_mom_lim[2];                       /* lower, upper */
    double *dx, *dy, *bex, *bey, *bez, *posx, *posy, *mask;
    double *g[NV];
    double t; int iter;
    modules_t mod;
} oracle;

/* std::min / std::max semantics (NaN behaviour matters: SURVEY Q22) */
static inline double smin(double a, double b) { return (b < a) ? b : a; }
static inline double smax(double a, double b) { return (a < b) ? b : a; }

static double *pl_new(const oracle *o) { return (double *)calloc((size_t)o->n, sizeof(double)); }
static double *pl_dup(const oracle *o, const double *a) { double *r = (double *)malloc(sizeof(double) * o->n); memcpy(r, a, sizeof(double) * o->n); return r; }
#define IDX(i, j) ((size_t)(i) * ny + (j))

/* ------------------------------------------------------------------ operators (source/mhd/derivs.cpp) */
static inline int xper(const oracle *o) { return o->xb1 == BC_PERIODIC && o->xb2 == BC_PERIODIC; }
static inline int yper(const oracle *o) { return o->yb1 == BC_PERIODIC && o->yb2 == BC_PERIODIC; }

/* derivs.cpp:477-487 */
static inline double b_interp(const oracle *o, const double *q, int i1, int j1, int i2, int j2)
{
    int ny = o->ny;
    double a = q[IDX(i1, j1)], b = q[IDX(i2, j2)], da, db;
    if (i1 == i2) { da = 0.5 * o->dy[IDX(i1, j1)]; db = 0.5 * o->dy[IDX(i2, j2)]; }
    else          { da = 0.5 * o->dx[IDX(i1, j1)]; db = 0.5 * o->dx[IDX(i2, j2)]; }
    return (a * db + b * da) / (db + da);
}
/* derivs.cpp:490-499 */
static inline double b_extrap(const oracle *o, const double *q, int i1, int j1, int i2, int j2)
{
    int ny = o->ny;
    double a = q[IDX(i1, j1)], b = q[IDX(i2, j2)], da, db;
    if (i1 == i2) { da = 0.5 * o->dy[IDX(i1, j1)]; db = 0.5 * o->dy[IDX(i2, j2)]; }
    else          { da = 0.5 * o->dx[IDX(i1, j1)]; db = 0.5 * o->dx[IDX(i2, j2)]; }
    return a + (b - a) * (da + 2.0 * db) / (da + db);
}

/* derivs.cpp:10-73 : Barton upwind face values; out has (nx+1-index) x (ny+index) entries */
static void upwind_surface(const oracle *o, const double *q, const double *vel, int index, double *surf)
{
    int nx = o->nx, ny = o->ny;
    int sx = nx + 1 - index, sy = ny + index;
    memset(surf, 0, sizeof(double) * (size_t)sx * sy);
#pragma omp parallel for
    for (int i = o->xl; i <= o->xu + 1 - index; i++) {
        for (int j = o->yl; j <= o->yu + index; j++) {
            int i2 = i, j2 = j, i0, i1, i3, j0, j1, j3;
            if (index == 0) {
                j0 = j1 = j3 = j2; i0 = i2 - 2; i1 = i2 - 1; i3 = i2 + 1;
                if (xper(o)) { if (i2 == nx) continue; i0 = (i0 + nx) % nx; i1 = (i1 + nx) % nx; i3 = (i3 + nx) % nx; }
            } else {
                i0 = i1 = i3 = i2; j0 = j2 - 2; j1 = j2 - 1; j3 = j2 + 1;
                if (yper(o)) { if (j2 == ny) continue; j0 = (j0 + ny) % ny; j1 = (j1 + ny) % ny; j3 = (j3 + ny) % ny; }
            }
            double d1, d2, d3, r;
            d2 = b_interp(o, q, i1, j1, i2, j2);
            double vf = b_interp(o, vel, i1, j1, i2, j2);
            double qc = q[IDX(i2, j2)], qm = q[IDX(i1, j1)];
            if (vf > 0.0) {
                d3 = qm; d1 = b_extrap(o, q, i0, j0, i1, j1);
                r = (qc <= qm) ? smin(d3, smax(d1, d2)) : smax(d3, smin(d1, d2));
            } else if (vf < 0.0) {
                d3 = qc; d1 = b_extrap(o, q, i3, j3, i2, j2);
                r = (qc <= qm) ? smax(d3, smin(d1, d2)) : smin(d3, smax(d1, d2));
            } else r = d2;
            surf[(size_t)i2 * sy + j2] = r;
        }
    }
}

/* derivs.cpp:122-162 (multiply_vel = true) */
static void transport_derivative1D(const oracle *o, const double *q, const double *vel, int index, double *out)
{
    int nx = o->nx, ny = o->ny;
    int sy = ny + index;
    double *surf = (double *)malloc(sizeof(double) * (size_t)(nx + 1) * (ny + 1));
    upwind_surface(o, q, vel, index, surf);
    const double *den = index == 0 ? o->dx : o->dy; /* denom = d_x*(1-index) + d_y*index : exact (x*1 + y*0) */
    memset(out, 0, sizeof(double) * o->n);
#pragma omp parallel for
    for (int i = o->xl; i <= o->xu; i++)
        for (int j = o->yl; j <= o->yu; j++) {
            int i0, i2, j0, j2;
            if (index == 0) { j0 = j2 = j; i0 = i - 1; i2 = i + 1; if (xper(o)) { i0 = (i0 + nx) % nx; i2 = (i2 + nx) % nx; } }
            else            { i0 = i2 = i; j0 = j - 1; j2 = j + 1; if (yper(o)) { j0 = (j0 + ny) % ny; j2 = (j2 + ny) % ny; } }
            out[IDX(i, j)] = (surf[(size_t)i2 * sy + j2] * b_interp(o, vel, i, j, i2, j2)
                              - surf[(size_t)i * sy + j] * b_interp(o, vel, i0, j0, i, j)) / den[IDX(i, j)];
        }
    free(surf);
}

/* derivs.cpp:216-220 */
static void transport_divergence2D(const oracle *o, const double *q, const double *vx, const double *vy, double *out)
{
    double *a = pl_new(o), *b = pl_new(o);
    transport_derivative1D(o, q, vx, 0, a);
    transport_derivative1D(o, q, vy, 1, b);
    for (int k = 0; k < o->n; k++) out[k] = a[k] + b[k];
    free(a); free(b);
}

/* derivs.cpp:223-264 */
static void derivative1D(const oracle *o, const double *q, int index, double *out)
{
    int nx = o->nx, ny = o->ny;
    memset(out, 0, sizeof(double) * o->n);
#pragma omp parallel for
    for (int i = o->xl; i <= o->xu; i++)
        for (int j = o->yl; j <= o->yu; j++) {
            int i0, i2, j0, j2; double den;
            if (index == 0) { j0 = j2 = j; i0 = i - 1; i2 = i + 1; if (xper(o)) { i0 = (i0 + nx) % nx; i2 = (i2 + nx) % nx; } den = o->dx[IDX(i, j)]; }
            else            { i0 = i2 = i; j0 = j - 1; j2 = j + 1; if (yper(o)) { j0 = (j0 + ny) % ny; j2 = (j2 + ny) % ny; } den = o->dy[IDX(i, j)]; }
            out[IDX(i, j)] = (b_interp(o, q, i, j, i2, j2) - b_interp(o, q, i0, j0, i, j)) / den;
        }
}

/* derivs.cpp:417-455 */
static void second_derivative1D(const oracle *o, const double *q, int index, double *out)
{
    int nx = o->nx, ny = o->ny;
    memset(out, 0, sizeof(double) * o->n);
    const double *d = index == 0 ? o->dx : o->dy;
#pragma omp parallel for
    for (int i = o->xl; i <= o->xu; i++)
        for (int j = o->yl; j <= o->yu; j++) {
            int i0, i2, j0, j2;
            if (index == 0) { j0 = j2 = j; i0 = i - 1; i2 = i + 1; if (xper(o)) { i0 = (i0 + nx) % nx; i2 = (i2 + nx) % nx; } }
            else            { i0 = i2 = i; j0 = j - 1; j2 = j + 1; if (yper(o)) { j0 = (j0 + ny) % ny; j2 = (j2 + ny) % ny; } }
            double h = 0.5 * d[IDX(i, j)]; /* (0.5*d_x*(1-index) + 0.5*d_y*index).square() */
            double den = h * h;
            out[IDX(i, j)] = (b_interp(o, q, i, j, i2, j2) - 2.0 * q[IDX(i, j)] + b_interp(o, q, i0, j0, i, j)) / den;
        }
}

/* derivs.cpp:458-462 */
static void laplacian(const oracle *o, const double *q, double *out)
{
    double *a = pl_new(o), *b = pl_new(o);
    second_derivative1D(o, q, 0, a); second_derivative1D(o, q, 1, b);
    for (int k = 0; k < o->n; k++) out[k] = a[k] + b[k];
    free(a); free(b);
}

/* ------------------------------------------------------------------ IdealMHD (source/equationsets/idealmhd.cpp) */

static int moc_any(const oracle *o);
static void moc_add_characteristic_evolution(const oracle *o, double *const *G, double **k);

/* idealmhd.cpp:42-105 (without an open_moc side the characteristic terms are +0) */
static void ideal_mhd_rhs(const oracle *o, double *const *G, double **k /* NEV planes, allocated */)
{
    const int n = o->n;
    const double *vx = G[V_v_x], *vy = G[V_v_y];
    double *t = pl_new(o), *u = pl_new(o);
    double *dbyx = pl_new(o), *dbxy = pl_new(o), *czx = pl_new(o), *czy = pl_new(o), *cdb = pl_new(o);
    /* continuity :52 */
    transport_divergence2D(o, G[V_rho], vx, vy, t);
    for (int c = 0; c < n; c++) k[0][c] = t[c] * -1.0;
    /* magnetic forces :54-60 */
    derivative1D(o, G[V_bi_y], 0, dbyx); derivative1D(o, G[V_bi_x], 1, dbxy);
    for (int c = 0; c < n; c++) cdb[c] = (dbyx[c] - dbxy[c]) / FOURPI;
    derivative1D(o, G[V_bi_z], 1, czx);                       /* curlZ: {d/dy z, -d/dx z}  derivs.cpp:465-469 */
    derivative1D(o, G[V_bi_z], 0, t);
    for (int c = 0; c < n; c++) czy[c] = t[c] * -1.0;
    double *pgx = pl_new(o), *pgy = pl_new(o);
    derivative1D(o, G[V_press], 0, pgx); derivative1D(o, G[V_press], 1, pgy);
    /* momentum :62-73 */
    double *tm = pl_new(o);
    transport_divergence2D(o, G[V_mom_x], vx, vy, tm);
    for (int c = 0; c < n; c++) {
        double ext_x = (cdb[c] * -1.0) * o->bey[c];
        double int_x = (cdb[c] * -1.0) * G[V_bi_y][c];
        double bzi = G[V_bi_z][c] / FOURPI, bze = o->bez[c] / FOURPI;
        /* CrossProduct2DZ(a, bz) = CrossProductZ2D(-1.0*bz, a) = { -(-bz)*a_y, (-bz)*a_x }  grid.cpp:455-468 */
        double iz_x = ((bzi * -1.0) * -1.0) * czy[c];
        double ez_x = ((bze * -1.0) * -1.0) * czy[c];
        k[1][c] = (((((tm[c] * -1.0) - pgx[c]) + G[V_rho][c] * G[V_grav_x][c]) + ext_x) + int_x) + iz_x + ez_x;
    }
    transport_divergence2D(o, G[V_mom_y], vx, vy, tm);
    for (int c = 0; c < n; c++) {
        double ext_y = cdb[c] * o->bex[c];
        double int_y = cdb[c] * G[V_bi_x][c];
        double bzi = G[V_bi_z][c] / FOURPI, bze = o->bez[c] / FOURPI;
        double iz_y = (bzi * -1.0) * czx[c];
        double ez_y = (bze * -1.0) * czx[c];
        k[2][c] = (((((tm[c] * -1.0) - pgy[c]) + G[V_rho][c] * G[V_grav_y][c]) + ext_y) + int_y) + iz_y + ez_y;
    }
    transport_divergence2D(o, G[V_mom_z], vx, vy, tm);
    for (int c = 0; c < n; c++) {
        double fze = (czx[c] * o->bey[c] - czy[c] * o->bex[c]) / FOURPI;            /* CrossProduct2D grid.cpp:448-452 */
        double fzi = (czx[c] * G[V_bi_y][c] - czy[c] * G[V_bi_x][c]) / FOURPI;
        k[3][c] = ((tm[c] * -1.0) + fze) + fzi;
    }
    /* energy :75-76 */
    double *dvxx = pl_new(o), *dvyy = pl_new(o);
    derivative1D(o, vx, 0, dvxx); derivative1D(o, vy, 1, dvyy);
    transport_divergence2D(o, G[V_thermal_energy], vx, vy, tm);
    for (int c = 0; c < n; c++) k[4][c] = (tm[c] * -1.0) - G[V_press][c] * (dvxx[c] + dvyy[c]);
    /* induction :78-86 */
    const double *bi[3] = { G[V_bi_x], G[V_bi_y], G[V_bi_z] };
    const double *be[3] = { o->bex, o->bey, o->bez };
    const double *vv[3] = { G[V_v_x], G[V_v_y], G[V_v_z] };
    double *te = pl_new(o);
    for (int a = 0; a < 3; a++) {
        transport_divergence2D(o, bi[a], vx, vy, tm);
        transport_divergence2D(o, be[a], vx, vy, te);
        derivative1D(o, vv[a], 0, t); derivative1D(o, vv[a], 1, u);
        for (int c = 0; c < n; c++)
            k[5 + a][c] = (((tm[c] * -1.0) - te[c]) + (bi[0][c] + be[0][c]) * t[c]) + (bi[1][c] + be[1][c]) * u[c];
    }
    /* mask multiply and (zero) characteristic add :99-103 */
    if (moc_any(o)) {
        for (int v = 0; v < NEV; v++) for (int c = 0; c < n; c++) k[v][c] *= o->mask[c];
        moc_add_characteristic_evolution(o, G, k);
    } else {
        for (int v = 0; v < NEV; v++) for (int c = 0; c < n; c++) { k[v][c] *= o->mask[c]; k[v][c] += 0.0; }
    }
    free(t); free(u); free(dbyx); free(dbxy); free(czx); free(czy); free(cdb); free(pgx); free(pgy); free(tm); free(te); free(dvxx); free(dvyy);
}

/* idealmhd.cpp:234-239 */
static void enforce_minimums(const oracle *o, double **G)
{
    for (int c = 0; c < o->n; c++) {
        G[V_rho][c] = smax(G[V_rho][c] / o->m_i, o->n_min) * o->m_i;
        G[V_thermal_energy][c] = smax(G[V_thermal_energy][c], o->e_min);
    }
}

/* evolution.cpp:158-224.  NOTE: reads and writes the PRIMARY state P, whatever set is being propagated. */
static void open_bc(const oracle *o, double **P, int i1, int i2, int i3, int j1, int j2, int j3)
{
    int ny = o->ny;
    int xbnd = (j1 == j2);
    const double *d = xbnd ? o->dx : o->dy;
    double delta_last = d[IDX(i3, j3)];
    double dist23 = 0.5 * (d[IDX(i2, j2)] + d[IDX(i3, j3)]);
    double dist12 = 0.5 * (d[IDX(i1, j1)] + d[IDX(i2, j2)]);
    double scale_2 = pow(o->open_decay, dist23 / delta_last);
    double scale_1 = pow(o->open_decay, dist12 / delta_last);
    double *rho = P[V_rho], *e = P[V_thermal_energy], *mx = P[V_mom_x], *my = P[V_mom_y];
    rho[IDX(i1, j1)] = scale_1 * rho[IDX(i3, j3)]; rho[IDX(i2, j2)] = scale_2 * rho[IDX(i3, j3)];
    e[IDX(i1, j1)] = scale_1 * e[IDX(i3, j3)];     e[IDX(i2, j2)] = scale_2 * e[IDX(i3, j3)];
    double press = e[IDX(i3, j3)] * (o->gamma - 1.0);
    double c_s = 0.0, c_new = sqrt(o->gamma * press / rho[IDX(i3, j3)]);
    if (c_new > c_s) c_s = c_new;
    double vel_x = mx[IDX(i3, j3)] / rho[IDX(i3, j3)], vel_y = my[IDX(i3, j3)] / rho[IDX(i3, j3)];
    double boost = o->open_strength * c_s;
    if (i2 > i1 || j2 > j1) boost *= -1.0;
    double dist = dist23;
    if (xbnd) {
        double bv = (i1 > i2) ? smax(0.0, vel_x + boost) : smin(0.0, vel_x + boost);
        double gv = (dist * bv - 0.5 * d[IDX(i2, j2)] * vel_x) / (0.5 * d[IDX(i3, j3)]);
        mx[IDX(i1, j1)] = rho[IDX(i1, j1)] * gv;    mx[IDX(i2, j2)] = rho[IDX(i2, j2)] * gv;
        my[IDX(i1, j1)] = rho[IDX(i1, j1)] * vel_y; my[IDX(i2, j2)] = rho[IDX(i2, j2)] * vel_y;
    } else {
        double bv = (j1 > j2) ? smax(0.0, vel_y + boost) : smin(0.0, vel_y + boost);
        double gv = (dist * bv - 0.5 * d[IDX(i2, j2)] * vel_y) / (0.5 * d[IDX(i3, j3)]);
        mx[IDX(i1, j1)] = rho[IDX(i1, j1)] * vel_x; mx[IDX(i2, j2)] = rho[IDX(i2, j2)] * vel_x;
        my[IDX(i1, j1)] = rho[IDX(i1, j1)] * gv;    my[IDX(i2, j2)] = rho[IDX(i2, j2)] * gv;
    }
}
/* evolution.cpp:231-266 (writes PRIMARY, Q2) */
static void reflect_bc(const oracle *o, double **P, int i1, int i2, int i3, int j1, int j2, int j3)
{
    int ny = o->ny;
    P[V_thermal_energy][IDX(i1, j1)] = P[V_thermal_energy][IDX(i3, j3)]; P[V_thermal_energy][IDX(i2, j2)] = P[V_thermal_energy][IDX(i3, j3)];
    P[V_rho][IDX(i1, j1)] = P[V_rho][IDX(i3, j3)]; P[V_rho][IDX(i2, j2)] = P[V_rho][IDX(i3, j3)];
    int m[3] = { V_mom_x, V_mom_y, V_mom_z };
    for (int a = 0; a < 3; a++) { P[m[a]][IDX(i1, j1)] = 0.0; P[m[a]][IDX(i2, j2)] = 0.0; P[m[a]][IDX(i3, j3)] = 0.0; }
}
/* evolution.cpp:272-282 (writes PRIMARY, Q2) */
static void fixed_bc(const oracle *o, double **P, int i1, int i2, int i3, int j1, int j2, int j3)
{
    int ny = o->ny;
    int m[3] = { V_mom_x, V_mom_y, V_mom_z };
    for (int a = 0; a < 3; a++) { P[m[a]][IDX(i1, j1)] = 0.0; P[m[a]][IDX(i2, j2)] = 0.0; P[m[a]][IDX(i3, j3)] = 0.0; }
}
/* evolution.cpp:290-333 (writes the set being propagated) ; side: 0 xl, 1 xu, 2 yl, 3 yu */
static void ucnp_bc(const oracle *o, double **G, int side, int index)
{
    int ny = o->ny;
    static const int vars[] = { V_rho, V_thermal_energy, V_bi_x, V_bi_y, V_bi_z, V_mom_x, V_mom_y, V_mom_z };
    for (int g = 0; g < N_GHOST; g++) {
        int ix, iy, bx, by;
        switch (side) {
        case 0: ix = o->xl; iy = index; bx = g; by = index; break;
        case 1: ix = o->xu; iy = index; bx = o->xu + 1 + g; by = index; break;
        case 2: ix = index; iy = o->yl; bx = index; by = g; break;
        default: ix = index; iy = o->yu; bx = index; by = o->yu + 1 + g; break;
        }
        for (size_t v = 0; v < sizeof(vars) / sizeof(int); v++) G[vars[v]][IDX(bx, by)] = G[vars[v]][IDX(ix, iy)];
    }
}

/* evolution.cpp:126-152 */
static void update_ghost_zones(const oracle *o, double **G, double **P)
{
    int nx = o->nx, ny = o->ny;
    if (o->xb1 == BC_OPEN)      for (int j = o->yl; j <= o->yu; j++) open_bc(o, P, 0, 1, 2, j, j, j);
    else if (o->xb1 == BC_REFLECT) for (int j = o->yl; j <= o->yu; j++) reflect_bc(o, P, 0, 1, 2, j, j, j);
    else if (o->xb1 == BC_FIXED)   for (int j = 0; j < ny; j++) fixed_bc(o, P, 0, 1, 2, j, j, j);
    else if (o->xb1 == BC_OPEN_UCNP) for (int j = o->yl; j <= o->yu; j++) ucnp_bc(o, G, 0, j);
    if (o->xb2 == BC_OPEN)      for (int j = o->yl; j <= o->yu; j++) open_bc(o, P, nx - 1, nx - 2, nx - 3, j, j, j);
    else if (o->xb2 == BC_REFLECT) for (int j = o->yl; j <= o->yu; j++) reflect_bc(o, P, nx - 1, nx - 2, nx - 3, j, j, j);
    else if (o->xb2 == BC_FIXED)   for (int j = 0; j < ny; j++) fixed_bc(o, P, nx - 1, nx - 2, nx - 3, j, j, j);
    else if (o->xb2 == BC_OPEN_UCNP) for (int j = o->yl; j <= o->yu; j++) ucnp_bc(o, G, 1, j);
    if (o->yb1 == BC_OPEN)      for (int i = o->xl; i <= o->xu; i++) open_bc(o, P, i, i, i, 0, 1, 2);
    else if (o->yb1 == BC_REFLECT) for (int i = o->xl; i <= o->xu; i++) reflect_bc(o, P, i, i, i, 0, 1, 2);
    else if (o->yb1 == BC_FIXED)   for (int i = 0; i < nx; i++) fixed_bc(o, P, i, i, i, 0, 1, 2);
    else if (o->yb1 == BC_OPEN_UCNP) for (int i = o->xl; i <= o->xu; i++) ucnp_bc(o, G, 2, i);
    if (o->yb2 == BC_OPEN)      for (int i = o->xl; i <= o->xu; i++) open_bc(o, P, i, i, i, ny - 1, ny - 2, ny - 3);
    else if (o->yb2 == BC_REFLECT) for (int i = o->xl; i <= o->xu; i++) reflect_bc(o, P, i, i, i, ny - 1, ny - 2, ny - 3);
    else if (o->yb2 == BC_FIXED)   for (int i = 0; i < nx; i++) fixed_bc(o, P, i, i, i, ny - 1, ny - 2, ny - 3);
    else if (o->yb2 == BC_OPEN_UCNP) for (int i = o->xl; i <= o->xu; i++) ucnp_bc(o, G, 3, i);
}

/* applyMomThresholdingMoC / applyBThresholdingMoC (idealmhd.cpp:107-223): on every open_moc side, the two ghost layers and the first interior layer
 * are clamped between lower*ref and upper*ref, ref = the value in the second interior layer of the same line; sides in the order x1, x2, y1, y2 */
static double moc_clamp(double v, double ref, double lo, double hi)
{
    return (ref >= 0.0) ? smin(smax(v, lo), hi) : smax(smin(v, lo), hi);
}
static void moc_thresholding(const oracle *o, double **G)
{
    const int nx = o->nx, ny = o->ny;
    const int bcs[4] = { o->xb1, o->xb2, o->yb1, o->yb2 };
    if (o->moc_mom_limiting) {
        const int moms[3] = { V_mom_x, V_mom_y, V_mom_z };
        for (int q = 0; q < 3; q++) { double *g = G[moms[q]];
            for (int s = 0; s < 4; s++) { if (bcs[s] != BC_OPEN_MOC) continue;
                const int xside = s < 2, lower = (s % 2) == 0, nal = xside ? ny : nx, ncr = xside ? nx : ny;
                for (int a = 0; a < nal; a++) for (int k = 0; k < N_GHOST + 1; k++) {
                    const int e = lower ? k : ncr - 1 - k, r = lower ? N_GHOST + 1 : ncr - 2 - N_GHOST;
                    const size_t c = xside ? IDX(e, a) : IDX(a, e), cr = xside ? IDX(r, a) : IDX(a, r);
                    const double ref = g[cr];
                    g[c] = moc_clamp(g[c], ref, o->moc_mom_lim[0] * ref, o->moc_mom_lim[1] * ref);
                } } }
    }
    if (o->moc_b_limiting) {
        const int bis[3] = { V_bi_x, V_bi_y, V_bi_z };
        const double *bes[3] = { o->bex, o->bey, o->bez };
        for (int q = 0; q < 3; q++) { double *g = G[bis[q]]; const double *be = bes[q];
            for (int s = 0; s < 4; s++) { if (bcs[s] != BC_OPEN_MOC) continue;
                const int xside = s < 2, lower = (s % 2) == 0, nal = xside ? ny : nx, ncr = xside ? nx : ny;
                for (int a = 0; a < nal; a++) for (int k = 0; k < N_GHOST + 1; k++) {
                    /* reference typo (idealmhd.cpp:154): the y_bound_2 reference column of the B limiter is m_xdim-2-N_GHOST, not m_ydim-2-N_GHOST
                     * (it must lie inside the grid: the reference asserts otherwise) */
                    const int e = lower ? k : ncr - 1 - k, r = lower ? N_GHOST + 1 : (s == 3 ? nx : ncr) - 2 - N_GHOST;
                    if (s == 3 && (r < 0 || r >= ny)) continue;
                    const size_t c = xside ? IDX(e, a) : IDX(a, e), cr = xside ? IDX(r, a) : IDX(a, r);
                    const double ref = be[cr] + g[cr];
                    g[c] = moc_clamp(g[c], ref, o->moc_b_lim[0] * ref - be[c], o->moc_b_lim[1] * ref - be[c]);
                } } }
    }
}

/* idealmhd.cpp:241-277 */
static void recompute_derived(const oracle *o, double **G)
{
    if (o->moc_mom_limiting || o->moc_b_limiting) moc_thresholding(o, G);
    for (int c = 0; c < o->n; c++) {
        double n = smax(G[V_rho][c] / o->m_i, o->n_min);
        G[V_n][c] = n;
        double rho = n * o->m_i;
        G[V_rho][c] = rho;
        double vx = G[V_mom_x][c] / rho, vy = G[V_mom_y][c] / rho, vz = G[V_mom_z][c] / rho;
        G[V_v_x][c] = vx; G[V_v_y][c] = vy; G[V_v_z][c] = vz;
        G[V_kinetic_energy][c] = (rho * 0.5) * (vx * vx + vy * vy);
        double e = smax(G[V_thermal_energy][c], o->e_min);
        G[V_thermal_energy][c] = e;
        double p = e * (o->gamma - 1.0);
        G[V_press][c] = p;
        G[V_temp][c] = smax(p / (n * (2 * K_B)), o->T_min);
        double bx = o->bex[c] + G[V_bi_x][c], by = o->bey[c] + G[V_bi_y][c], bz = o->bez[c] + G[V_bi_z][c];
        G[V_b_x][c] = bx; G[V_b_y][c] = by; G[V_b_z][c] = bz;
        double bm = sqrt((bx * bx + by * by) + bz * bz);
        G[V_b_mag][c] = bm;
        if (bm == 0.0) { G[V_b_hat_x][c] = 0.0; G[V_b_hat_y][c] = 0.0; G[V_b_hat_z][c] = 0.0; }
        else { G[V_b_hat_x][c] = bx / bm; G[V_b_hat_y][c] = by / bm; G[V_b_hat_z][c] = bz / bm; }
    }
}

/* idealmhd.cpp:279-304 */
static void recompute_dt(const oracle *o, double **G)
{
    for (int c = 0; c < o->n; c++) {
        double rho = G[V_rho][c];
        double cs = sqrt(o->gamma * G[V_press][c] / rho);
        double cs2 = cs * cs;
        double va = G[V_b_mag][c] / sqrt(rho * FOURPI);
        double va2 = va * va;
        double s = cs2 + va2;
        double delta = sqrt(1.0 - ((cs2 * 4.0) * va2) / (s * s));
        double vfast = sqrt((s * 0.5) * (1.0 + delta));
        double vslow = sqrt((s * 0.5) * (1.0 - delta));
        double vmx = sqrt(G[V_v_x][c] * G[V_v_x][c]), vmy = sqrt(G[V_v_y][c] * G[V_v_y][c]);
        double M = smax(smax(smax(cs, va), vfast), vslow);
        G[V_dt][c] = 1.0 / ((vmx + M) / o->dx[c] + (vmy + M) / o->dy[c]);
    }
}

/* equationset.cpp:212-220 */
static void propagate_changes(const oracle *o, double **G, double **P)
{
    enforce_minimums(o, G);
    update_ghost_zones(o, G, P);
    recompute_derived(o, G);
    recompute_dt(o, G);
}

/* ------------------------------------------------------------------ modules */
static void av_rhs(const oracle *o, double *const *G, double **k);
static void module_rhs_hooks(const oracle *o, double *const *G, double **k) { if (o->mod.av_on) av_rhs(o, G, k); }

/* equationset.cpp:204-210 */
static void compute_time_derivatives(const oracle *o, double *const *G, double **k)
{
    ideal_mhd_rhs(o, G, k);
    module_rhs_hooks(o, G, k);
}
/* equationset.cpp:222-230 */
static void apply_time_derivatives(const oracle *o, double **G, double **P, double *const *k, double step)
{
    for (int v = 0; v < NEV; v++) { double *g = G[EVOLVED[v]]; for (int c = 0; c < o->n; c++) g[c] += k[v][c] * step; }
    propagate_changes(o, G, P);
}

static double **kalloc(const oracle *o) { double **k = (double **)malloc(sizeof(double *) * NEV); for (int v = 0; v < NEV; v++) k[v] = pl_new(o); return k; }
static void kfree(double **k) { for (int v = 0; v < NEV; v++) free(k[v]); free(k); }
static double **gcopy(const oracle *o, double *const *G) { double **r = (double **)malloc(sizeof(double *) * NV); for (int v = 0; v < NV; v++) r[v] = pl_dup(o, G[v]); return r; }
static void gfree(double **g) { for (int v = 0; v < NV; v++) free(g[v]); free(g); }

/* grid.cpp:71-82 (NaN-ignoring std::min) */
static double min_range(const oracle *o, const double *a, int il, int jl, int iu, int ju)
{
    int ny = o->ny; double m = 1.7976931348623157e308;
    for (int i = il; i <= iu; i++) for (int j = jl; j <= ju; j++) m = smin(m, a[IDX(i, j)]);
    return m;
}

static void propagate_changes(const oracle *o, double **G, double **P);
/* multispecies_mode: a module's energy input w goes to the cumulative planes as  ion += (1 - f) * w  (if f < 1),  electron += f * w  (if f > 0), f = the module's
 * ms_electron_heating_fraction (e.g. thermalconduction.cpp:105-108).  W(fr) is the per-cell expression with the fraction factor `fr` placed where the reference places it. */
#define MS_FEED(o, module, OP, W) do { if ((o)->mod.ms_on) { const double f_ = (o)->mod.ms_frac[module]; \
        if (f_ < 1.0) { const double fr = 1.0 - f_; for (int c = 0; c < (o)->n; c++) (o)->mod.ms_cum[1][c] OP (W); } \
        if (f_ > 0.0) { const double fr = f_; for (int c = 0; c < (o)->n; c++) (o)->mod.ms_cum[0][c] OP (W); } } } while (0)
#include "physical_viscosity_oracle.inc"
#include "moc_oracle.inc"
#include "ucnp_modules_oracle.inc"
#include "solar_small_modules_oracle.inc"
#include "anomalous_resistivity_oracle.inc"

/* ---- artificial viscosity (source/modules/viscosity.cpp) */
static int ev_index(int var) { for (int v = 0; v < NEV; v++) if (EVOLVED[v] == var) return v; return -1; }
static int is_mom(int v) { return v == V_mom_x || v == V_mom_y || v == V_mom_z; }
static int is_vel(int v) { return v == V_v_x || v == V_v_y || v == V_v_z; }
/* constructSingleViscosityGrid :185-267 ; G = the grid set the RHS is evaluated on; dt / dt_min always come from the PRIMARY state (Q13) */
static void av_single(const oracle *o, double *const *G, int i, double *out)
{
    int n = o->n;
    const modules_t *m = &o->mod;
    int xl = o->xl, yl = o->yl, xu = o->xu, yu = o->yu;
    double dt_min = min_range(o, o->g[V_dt], xl, yl, xu, yu);
    double *coef = pl_new(o), *lap = pl_new(o), *scale = pl_new(o);
    for (int c = 0; c < n; c++) {
        double str = (m->av_opt[i] == 0 || m->av_opt[i] == 1) ? m->av_strength[i] : m->av_strength_grid[i][c];
        double dtg = (m->av_opt[i] == 0 || m->av_opt[i] == 2) ? o->g[V_dt][c] : dt_min;
        coef[c] = (((str * 1.0) / (1.0 / (o->dx[c] * o->dx[c]) + 1.0 / (o->dy[c] * o->dy[c]))) / 2.) / dtg;       /* :213 */
        scale[c] = 1.0;
    }
    laplacian(o, G[m->av_diff[i]], lap);                                                                          /* :225 */
    if (m->av_out[1][i]) {                                                                                        /* m_grids_lap :225, m_grids_strength :195-196, m_grids_dt :209-210 */
        memcpy(m->av_out[1][i], lap, sizeof(double) * n);
        for (int c = 0; c < n; c++) {
            m->av_out[2][i][c] = (m->av_opt[i] == 0 || m->av_opt[i] == 1) ? m->av_strength[i] : m->av_strength_grid[i][c];
            m->av_out[3][i][c] = (m->av_opt[i] == 0 || m->av_opt[i] == 2) ? o->g[V_dt][c] : dt_min;
        }
    }
    if (is_mom(m->av_evol[i]) && is_vel(m->av_diff[i])) for (int c = 0; c < n; c++) scale[c] = G[V_n][c] * o->m_i;              /* :229-239 */
    if (m->av_evol[i] == V_thermal_energy && m->av_diff[i] == V_temp) for (int c = 0; c < n; c++) scale[c] = G[V_n][c] * (K_B / (o->gamma - 1));   /* :244-254 */
    if (m->av_gradient_correction) {                                                                               /* :261-265 */
        double *cs = pl_new(o), *a = pl_new(o), *b = pl_new(o), *cc = pl_new(o), *d = pl_new(o);
        for (int c = 0; c < n; c++) cs[c] = coef[c] * scale[c];
        derivative1D(o, cs, 0, a); derivative1D(o, G[m->av_diff[i]], 0, b); derivative1D(o, cs, 1, cc); derivative1D(o, G[m->av_diff[i]], 1, d);
        for (int c = 0; c < n; c++) out[c] = ((coef[c] * lap[c]) * scale[c] + a[c] * b[c]) + cc[c] * d[c];
        free(cs); free(a); free(b); free(cc); free(d);
    } else for (int c = 0; c < n; c++) out[c] = (coef[c] * lap[c]) * scale[c];
    free(coef); free(lap); free(scale);
}
/* computeTimeDerivativesModule :112-123 */
static void av_rhs(const oracle *o, double *const *G, double **k)
{
    double *dq = pl_new(o);
    for (int i = 0; i < o->mod.av_nterms; i++) {
        av_single(o, G, i, dq);                      /* constructViscosityGrids evaluates every term */
        if (o->mod.av_strength[i] <= 1.0) {
            double *kk = k[ev_index(o->mod.av_evol[i])], *keep = o->mod.av_out[0][i];
            for (int c = 0; c < o->n; c++) { const double t = dq[c] * o->mask[c]; if (keep) keep[c] = t; kk[c] += t; }       /* m_grids_dqdt[i] = dqdt * mask, :116-118 */
        }
    }
    free(dq);
}
/* iterateModule :125-180 : hyper-viscous terms (strength > 1) are sub-cycled on the primary state */
static void av_iterate(oracle *o, double dt)
{
    int n = o->n;
    double *dq = pl_new(o), *d2 = pl_new(o), *d3 = pl_new(o), *d4 = pl_new(o), *init = pl_new(o);
    for (int i = 0; i < o->mod.av_nterms; i++) {
        if (o->mod.av_strength[i] <= 1.0) continue;
        double *ge = o->g[o->mod.av_evol[i]];
        int ns = (int)(ceil(o->mod.av_strength[i] / o->mod.av_hv_epsilon) + 0.1);
        double dts = dt / (double)ns;
        for (int sc = 0; sc < ns; sc++) {
            if (o->mod.av_hv_integrator == TI_EULER) {
                av_single(o, o->g, i, dq);
                for (int c = 0; c < n; c++) ge[c] = ge[c] + (o->mask[c] * dts) * dq[c];
                propagate_changes(o, o->g, o->g);
            } else if (o->mod.av_hv_integrator == TI_RK2) {
                memcpy(init, ge, sizeof(double) * n);
                av_single(o, o->g, i, dq);
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * (0.5 * dts)) * dq[c];
                propagate_changes(o, o->g, o->g);
                av_single(o, o->g, i, dq);
                if (o->mod.av_out[0][i]) memcpy(o->mod.av_out[0][i], dq, sizeof(double) * n);                     /* m_grids_dqdt[i] = dqdt (unmasked), :148 */
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * dts) * dq[c];
                propagate_changes(o, o->g, o->g);
            } else {
                memcpy(init, ge, sizeof(double) * n);
                av_single(o, o->g, i, dq);
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * (0.5 * dts)) * dq[c];
                propagate_changes(o, o->g, o->g); av_single(o, o->g, i, d2);
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * (0.5 * dts)) * d2[c];
                propagate_changes(o, o->g, o->g); av_single(o, o->g, i, d3);
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * dts) * d3[c];
                propagate_changes(o, o->g, o->g); av_single(o, o->g, i, d4);
                if (o->mod.av_out[0][i]) for (int c = 0; c < n; c++) o->mod.av_out[0][i][c] = (((dq[c] + d2[c] * 2.0) + d3[c] * 2.0) + d4[c]) / 6.0;   /* :172 */
                for (int c = 0; c < n; c++) ge[c] = init[c] + (o->mask[c] * dts) * ((((dq[c] + d2[c] * 2.0) + d3[c] * 2.0) + d4[c]) / 6.0);
                propagate_changes(o, o->g, o->g);
            }
        }
        propagate_changes(o, o->g, o->g);
    }
    free(dq); free(d2); free(d3); free(d4); free(init);
}

/* ---- thermal conduction (source/modules/solar/thermalconduction.cpp) */
static void tc_field_aligned_flux(const oracle *o, double *fx, double *fy, const double *temp, double k0)
{   /* :154-178 */
    int ny = o->ny, n = o->n;
    const double *rho = o->g[V_rho], *bhx = o->g[V_b_hat_x], *bhy = o->g[V_b_hat_y];
    double *dTx = pl_new(o), *dTy = pl_new(o), *cx = pl_new(o), *cy = pl_new(o);
    derivative1D(o, temp, 0, dTx); derivative1D(o, temp, 1, dTy);
    for (int c = 0; c < n; c++) {
        double kmax = (((o->dx[c] * o->dy[c]) * K_B) * (rho[c] / o->m_i)) / o->mod.tc_dt_subcycle_min;
        double kap = smin(pow(temp[c], 5.0 / 2.0) * k0, kmax);
        cx[c] = (kap * -1.0) * dTx[c]; cy[c] = (kap * -1.0) * dTy[c];
    }
    for (int i = o->xl; i <= o->xu; i++) for (int j = o->yl; j <= o->yu; j++) {
        size_t c = IDX(i, j);
        double fm = cx[c] * bhx[c] + cy[c] * bhy[c];
        fx[c] = fm * bhx[c]; fy[c] = fm * bhy[c];
    }
    free(dTx); free(dTy); free(cx); free(cy);
}
static void tc_saturate_flux(const oracle *o, double *fx, double *fy, const double *temp)
{   /* :182-188 */
    const double *rho = o->g[V_rho];
    const double c1 = (1.0 / 6.0) * (3.0 / 2.0);
    const double sme = sqrt(M_ELECTRON);
    for (int c = 0; c < o->n; c++) {
        double sat = (((rho[c] / o->m_i) * c1) * pow(temp[c] * K_B, 1.5)) / sme;
        double fm = sqrt(fx[c] * fx[c] + fy[c] * fy[c]);
        double sc = sat / sqrt(sat * sat + fm * fm);
        fx[c] *= sc; fy[c] *= sc;
    }
}
static void tc_saturation_terms(const oracle *o, const double *temp, double *coef, double *add)
{   /* :211-227 */
    int n = o->n; double k0 = o->mod.tc_weakening * KAPPA_0;
    double *fx = pl_new(o), *fy = pl_new(o), *rx, *ry, *fm = pl_new(o), *dcx = pl_new(o), *dcy = pl_new(o);
    tc_field_aligned_flux(o, fx, fy, temp, k0);
    for (int c = 0; c < n; c++) fm[c] = sqrt(fx[c] * fx[c] + fy[c] * fy[c]);
    rx = pl_dup(o, fx); ry = pl_dup(o, fy);
    tc_saturate_flux(o, fx, fy, temp);
    for (int c = 0; c < n; c++) { double sfm = sqrt(fx[c] * fx[c] + fy[c] * fy[c]); coef[c] = (fm[c] != 0.0) ? sfm / fm[c] : 1.0; }
    derivative1D(o, coef, 0, dcx); derivative1D(o, coef, 1, dcy);
    for (int c = 0; c < n; c++) add[c] = (o->mask[c] * -1.0) * (dcx[c] * rx[c] + dcy[c] * ry[c]);
    free(fx); free(fy); free(rx); free(ry); free(fm); free(dcx); free(dcy);
}
static void tc_energy_derivative(const oracle *o, const double *T, const double *bhx, const double *bhy, double *out)
{   /* :114-132 */
    int n = o->n; double kap = o->mod.tc_weakening * KAPPA_0;
    double *Tx = pl_new(o), *Ty = pl_new(o), *Txx = pl_new(o), *Tyy = pl_new(o), *Txy = pl_new(o), *Tyx = pl_new(o);
    double *bxx = pl_new(o), *bxy = pl_new(o), *byx = pl_new(o), *byy = pl_new(o);
    derivative1D(o, T, 0, Tx); derivative1D(o, T, 1, Ty);
    second_derivative1D(o, T, 0, Txx); second_derivative1D(o, T, 1, Tyy);
    derivative1D(o, Tx, 1, Txy); derivative1D(o, Ty, 0, Tyx);
    derivative1D(o, bhx, 0, bxx); derivative1D(o, bhx, 1, bxy); derivative1D(o, bhy, 0, byx); derivative1D(o, bhy, 1, byy);
    for (int c = 0; c < n; c++) {
        double bg = bhx[c] * Tx[c] + bhy[c] * Ty[c];
        double t1x = bhx[c] * Txx[c] + bhy[c] * Txy[c], t1y = bhx[c] * Tyx[c] + bhy[c] * Tyy[c];
        double t2x = Tx[c] * bxx[c] + Ty[c] * bxy[c],   t2y = Tx[c] * byx[c] + Ty[c] * byy[c];
        double cu = byx[c] - bxy[c];
        double t3x = ((cu * -1.0) * -1.0) * Ty[c], t3y = (cu * -1.0) * Tx[c];
        double p15 = pow(T[c], 3.0 / 2.0), p25 = pow(T[c], 5.0 / 2.0);
        double ttx = ((p15 * (5.0 / 2.0)) * Tx[c]) * bg + p25 * ((t1x + t2x) + t3x);
        double tty = ((p15 * (5.0 / 2.0)) * Ty[c]) * bg + p25 * ((t1y + t2y) + t3y);
        out[c] = (((p25 * bg) * (bxx[c] + byy[c])) + (bhx[c] * ttx + bhy[c] * tty)) * kap;
    }
    if (o->mod.tc_flux_saturation) {
        double *coef = pl_new(o), *add = pl_new(o);
        tc_saturation_terms(o, T, coef, add);
        for (int c = 0; c < n; c++) out[c] = coef[c] * out[c] + add[c];
        free(coef); free(add);
    }
    free(Tx); free(Ty); free(Txx); free(Tyy); free(Txy); free(Tyx); free(bxx); free(bxy); free(byx); free(byy);
}
static int tc_number_subcycles(oracle *o, double dt)
{   /* :135-149 */
    int n = o->n; double kap = o->mod.tc_weakening * KAPPA_0;
    const double *rho = o->g[V_rho], *T = o->g[V_temp];
    double *dts = pl_new(o);
    if (!o->mod.tc_flux_saturation) {
        for (int c = 0; c < n; c++) dts[c] = ((((rho[c] / o->m_i) * (K_B / kap)) * o->dx[c]) * o->dy[c]) / pow(T[c], 2.5);
    } else {
        double *Tx = pl_new(o), *Ty = pl_new(o), *fx = pl_new(o), *fy = pl_new(o);
        derivative1D(o, T, 0, Tx); derivative1D(o, T, 1, Ty);
        double mx = -1.7976931348623157e308;
        for (int c = 0; c < n; c++) { Tx[c] = Tx[c] * o->g[V_b_hat_x][c] + Ty[c] * o->g[V_b_hat_y][c]; mx = smax(mx, fabs(Tx[c])); }
        if (mx == 0.0) { free(Tx); free(Ty); free(fx); free(fy); free(dts); return 0; }
        tc_field_aligned_flux(o, fx, fy, T, kap);            /* saturatedKappa :192-204 */
        tc_saturate_flux(o, fx, fy, T);
        for (int c = 0; c < n; c++) {
            double km = fabs(sqrt(fx[c] * fx[c] + fy[c] * fy[c]) / Tx[c]);
            dts[c] = (((K_B / km) * (rho[c] / o->m_i)) * o->dx[c]) * o->dy[c];
        }
        free(Tx); free(Ty); free(fx); free(fy);
    }
    double md = smax(o->mod.tc_epsilon * min_range(o, dts, o->xl, o->yl, o->xu, o->yu), o->mod.tc_dt_subcycle_min);
    free(dts);
    return (int)(dt / md) + 1;
}
static void temp_from_energy(const oracle *o, const double *e, const double *nn, double *T)
{
    for (int c = 0; c < o->n; c++) T[c] = smax((e[c] * (o->gamma - 1.0)) / (nn[c] * (2.0 * K_B)), o->T_min);
}
static void tc_iterate(oracle *o, double dt)
{   /* :47-112 */
    int n = o->n, ns = o->mod.tc_nsub;
    double *e = pl_dup(o, o->g[V_thermal_energy]), *T = pl_dup(o, o->g[V_temp]), *nn = pl_dup(o, o->g[V_n]);
    double *bhx = pl_dup(o, o->g[V_b_hat_x]), *bhy = pl_dup(o, o->g[V_b_hat_y]);
    double *k1 = pl_new(o), *k2 = pl_new(o), *k3 = pl_new(o), *k4 = pl_new(o), *im = pl_new(o), *imT = pl_new(o);
    double dts = dt / (double)ns;
    if (!o->mod.tc_avg) { o->mod.tc_avg = pl_new(o); o->mod.tc_sat = pl_new(o); }
    if (o->mod.tc_flux_saturation) { double *add = pl_new(o); tc_saturation_terms(o, T, o->mod.tc_sat, add); free(add); }      /* :53-59: sat_terms of the entry temperature */
    for (int s = 0; s < ns; s++) {
        if (o->mod.tc_integrator == TI_EULER) {
            tc_energy_derivative(o, T, bhx, bhy, k1);
            for (int c = 0; c < n; c++) e[c] = smax(e[c] + (o->mask[c] * dts) * k1[c], o->e_min);
            temp_from_energy(o, e, nn, T);
        } else if (o->mod.tc_integrator == TI_RK2) {
            tc_energy_derivative(o, T, bhx, bhy, k1);
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * (0.5 * dts)) * k1[c], o->e_min);
            temp_from_energy(o, im, nn, imT);
            tc_energy_derivative(o, imT, bhx, bhy, k2);
            for (int c = 0; c < n; c++) e[c] = smax(e[c] + (o->mask[c] * dts) * k2[c], o->e_min);
            temp_from_energy(o, e, nn, T);
        } else {
            tc_energy_derivative(o, T, bhx, bhy, k1);
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * (0.5 * dts)) * k1[c], o->e_min);
            temp_from_energy(o, im, nn, imT); tc_energy_derivative(o, imT, bhx, bhy, k2);
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * (0.5 * dts)) * k2[c], o->e_min);
            temp_from_energy(o, im, nn, imT); tc_energy_derivative(o, imT, bhx, bhy, k3);
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * dts) * k3[c], o->e_min);
            temp_from_energy(o, im, nn, imT); tc_energy_derivative(o, imT, bhx, bhy, k4);
            for (int c = 0; c < n; c++) e[c] = smax(e[c] + ((o->mask[c] * dts) * (((k1[c] + k2[c] * 2.0) + k3[c] * 2.0) + k4[c])) / 6.0, o->e_min);
            temp_from_energy(o, e, nn, T);
        }
    }
    for (int c = 0; c < n; c++) o->mod.tc_avg[c] = (e[c] - o->g[V_thermal_energy][c]) / dt;                                  /* :102 */
    MS_FEED(o, 1, +=, (e[c] - o->g[V_thermal_energy][c]) * fr);                                                              /* :105-108 */
    if (!o->mod.tc_inactive) {                                                                                               /* :109 */
    memcpy(o->g[V_thermal_energy], e, sizeof(double) * n);
    propagate_changes(o, o->g, o->g);
    }
    free(e); free(T); free(nn); free(bhx); free(bhy); free(k1); free(k2); free(k3); free(k4); free(im); free(imT);
}

/* ---- radiative losses (source/modules/solar/radiativelosses.cpp) */
static void rl_losses(const oracle *o, const double *T, const double *nn, double *out)
{   /* :110-158 */
    int ny = o->ny;
    memset(out, 0, sizeof(double) * o->n);
    for (int i = o->xl; i <= o->xu; i++) for (int j = o->yl; j <= o->yu; j++) {
        size_t c = IDX(i, j);
        if (T[c] < o->mod.rl_cutoff_temp) { out[c] = 0.0; continue; }
        double lt = log10(T[c]), chi, alpha;
        if (lt <= 4.97) { chi = 1.09e-31; alpha = 2.0; }
        else if (lt <= 5.67) { chi = 8.87e-17; alpha = -1.0; }
        else if (lt <= 6.18) { chi = 1.90e-22; alpha = 0.0; }
        else if (lt <= 6.55) { chi = 3.53e-13; alpha = -1.5; }
        else if (lt <= 6.90) { chi = 3.46e-25; alpha = 1.0 / 3.0; }
        else if (lt <= 7.63) { chi = 5.49e-16; alpha = -1.0; }
        else { chi = 1.96e-27; alpha = 0.5; }
        double r = pow(nn[c], 2.0) * chi * pow(T[c], alpha);
        if (T[c] < o->mod.rl_cutoff_temp + o->mod.rl_cutoff_ramp) { double ramp = (T[c] - o->mod.rl_cutoff_temp) / o->mod.rl_cutoff_ramp; r *= ramp; }
        if (o->mod.rl_prevent_subcycling) {
            double e = o->g[V_thermal_energy][c], dtc = o->g[V_dt][c];
            if (0.1 * o->mod.rl_epsilon * (e / r) < o->epsilon * dtc) r = 0.1 * o->mod.rl_epsilon * e / (o->epsilon * dtc);
        }
        out[c] = r;
    }
}
static int rl_number_subcycles(oracle *o, double dt)
{   /* :161-166 */
    double *L = pl_new(o);
    rl_losses(o, o->g[V_temp], o->g[V_n], L);
    double mx = -1.7976931348623157e308, mn = 1.7976931348623157e308;
    for (int c = 0; c < o->n; c++) mx = smax(mx, L[c]);
    if (mx == 0.0) { free(L); return 0; }
    for (int c = 0; c < o->n; c++) mn = smin(mn, fabs(o->g[V_thermal_energy][c] / L[c]));
    free(L);
    double sdt = o->mod.rl_epsilon * mn;
    return (int)(dt / sdt) + 1;
}
static void rl_iterate(oracle *o, double dt)
{   /* :45-101 */
    int n = o->n, ns = o->mod.rl_nsub;
    double *e = pl_dup(o, o->g[V_thermal_energy]), *T = pl_dup(o, o->g[V_temp]), *nn = pl_dup(o, o->g[V_n]);
    double *k1 = pl_new(o), *k2 = pl_new(o), *k3 = pl_new(o), *k4 = pl_new(o), *im = pl_new(o), *imT = pl_new(o);
    double dts = dt / (double)ns;
    for (int s = 0; s < ns; s++) {
        if (o->mod.rl_integrator == TI_EULER) {
            rl_losses(o, T, nn, k1);
            for (int c = 0; c < n; c++) e[c] = smax(e[c] - (o->mask[c] * dts) * k1[c], o->e_min);
            temp_from_energy(o, e, nn, T);
        } else if (o->mod.rl_integrator == TI_RK2) {
            rl_losses(o, T, nn, k1);
            for (int c = 0; c < n; c++) im[c] = smax(e[c] - (o->mask[c] * (0.5 * dts)) * k1[c], o->e_min);
            temp_from_energy(o, im, nn, imT);
            rl_losses(o, imT, nn, k2);
            for (int c = 0; c < n; c++) e[c] = smax(e[c] - (o->mask[c] * dts) * k2[c], o->e_min);
            temp_from_energy(o, e, nn, T);
        } else {
            rl_losses(o, T, nn, k1); for (int c = 0; c < n; c++) k1[c] = k1[c] * -1.0;
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * (0.5 * dts)) * k1[c], o->e_min);
            temp_from_energy(o, im, nn, imT); rl_losses(o, imT, nn, k2); for (int c = 0; c < n; c++) k2[c] = k2[c] * -1.0;
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * (0.5 * dts)) * k2[c], o->e_min);
            temp_from_energy(o, im, nn, imT); rl_losses(o, imT, nn, k3); for (int c = 0; c < n; c++) k3[c] = k3[c] * -1.0;
            for (int c = 0; c < n; c++) im[c] = smax(e[c] + (o->mask[c] * dts) * k3[c], o->e_min);
            temp_from_energy(o, im, nn, imT); rl_losses(o, imT, nn, k4); for (int c = 0; c < n; c++) k4[c] = k4[c] * -1.0;
            for (int c = 0; c < n; c++) e[c] = smax(e[c] + ((o->mask[c] * dts) * (((k1[c] + k2[c] * 2.0) + k3[c] * 2.0) + k4[c])) / 6.0, o->e_min);
            temp_from_energy(o, e, nn, T);
        }
    }
    if (!o->mod.rl_avg) o->mod.rl_avg = pl_new(o);
    for (int c = 0; c < n; c++) o->mod.rl_avg[c] = (e[c] - o->g[V_thermal_energy][c]) / dt;                                  /* radiativelosses.cpp:93 */
    MS_FEED(o, 2, +=, (e[c] - o->g[V_thermal_energy][c]) * fr);                                                              /* :94-97 */
    if (!o->mod.rl_inactive) {                                                                                               /* :98 */
    memcpy(o->g[V_thermal_energy], e, sizeof(double) * n);
    propagate_changes(o, o->g, o->g);
    }
    free(e); free(T); free(nn); free(k1); free(k2); free(k3); free(k4); free(im); free(imT);
}

/* ---- ambient heating (source/modules/solar/ambientheating.cpp:42-49) */
static void ah_post_iterate(oracle *o, double dt)
{
    for (int c = 0; c < o->n; c++) o->g[V_thermal_energy][c] += o->mod.ah_heating[c] * dt;
    propagate_changes(o, o->g, o->g);
    MS_FEED(o, 3, +=, ((o->mask[c] * fr) * o->mod.ah_heating[c]) * dt);                                                      /* :45-48 */
}

/* ------------------------------------------------------------------ time loop (source/mhd/evolution.cpp) */
/* evolution.cpp:59-82 ; returns the step size used */
static double advance_time(oracle *o)
{
    /* :62; dt bounds = interior, widened by the ghost zone on open_moc sides (plasmadomain.cpp:155-159) */
    double step = o->epsilon * min_range(o, o->g[V_dt], o->xl - (o->xb1 == BC_OPEN_MOC ? N_GHOST : 0), o->yl - (o->yb1 == BC_OPEN_MOC ? N_GHOST : 0),
                                         o->xu + (o->xb2 == BC_OPEN_MOC ? N_GHOST : 0), o->yu + (o->yb2 == BC_OPEN_MOC ? N_GHOST : 0));
    for (int m = 0; m < o->mod.n_modules; m++) {                                         /* preIterate :65 */
        if (o->mod.order[m] == 1) o->mod.tc_nsub = tc_number_subcycles(o, step);
        if (o->mod.order[m] == 2) o->mod.rl_nsub = rl_number_subcycles(o, step);
        if (o->mod.order[m] >= 100) small_module_pre(o, (small_module *)o->mod.small[o->mod.order[m] - 100]);
    }
    for (int m = 0; m < o->mod.n_modules; m++) {                                         /* iterate :66 */
        if (o->mod.order[m] == 1) tc_iterate(o, step);
        if (o->mod.order[m] == 2) rl_iterate(o, step);
        if (o->mod.order[m] == 4) av_iterate(o, step);
        if (o->mod.order[m] == 5) pv_iterate(o, step);
        if (o->mod.order[m] >= 100) small_module_iterate(o, (small_module *)o->mod.small[o->mod.order[m] - 100], step);
        if (o->mod.order[m] == 13) ar_iterate(o, (anom_res *)o->mod.anom, step);
    }
    double **k1 = kalloc(o);
    if (o->integrator == TI_EULER) {                                                     /* :84-88 */
        compute_time_derivatives(o, o->g, k1);
        apply_time_derivatives(o, o->g, o->g, k1, step);
    } else if (o->integrator == TI_RK2) {                                                /* :90-101 */
        compute_time_derivatives(o, o->g, k1);
        double **mid = gcopy(o, o->g);
        apply_time_derivatives(o, mid, o->g, k1, 0.5 * step);
        compute_time_derivatives(o, mid, k1);
        apply_time_derivatives(o, o->g, o->g, k1, step);
        gfree(mid);
    } else {                                                                             /* :103-124 */
        double **k2 = kalloc(o), **k3 = kalloc(o), **k4 = kalloc(o);
        compute_time_derivatives(o, o->g, k1);
        double **rk = gcopy(o, o->g); apply_time_derivatives(o, rk, o->g, k1, 0.5 * step); compute_time_derivatives(o, rk, k2); gfree(rk);
        rk = gcopy(o, o->g); apply_time_derivatives(o, rk, o->g, k2, 0.5 * step); compute_time_derivatives(o, rk, k3); gfree(rk);
        rk = gcopy(o, o->g); apply_time_derivatives(o, rk, o->g, k3, step); compute_time_derivatives(o, rk, k4); gfree(rk);
        for (int v = 0; v < NEV; v++) for (int c = 0; c < o->n; c++) k1[v][c] = (k1[v][c] + k4[v][c]) / 6.0 + (k2[v][c] + k3[v][c]) / 3.0;
        apply_time_derivatives(o, o->g, o->g, k1, step);
        kfree(k2); kfree(k3); kfree(k4);
    }
    kfree(k1);
    for (int m = 0; m < o->mod.n_modules; m++) {                                         /* postIterate :74 */
        if (o->mod.order[m] == 3) ah_post_iterate(o, step);
        if (o->mod.order[m] >= 100) small_module_post(o, (small_module *)o->mod.small[o->mod.order[m] - 100], step);
    }
    o->t += step; o->iter++;
    return step;
}

/* ------------------------------------------------------------------ C entry points (ctypes) */
oracle *oracle_create(int nx, int ny, int xb1, int xb2, int yb1, int yb2, int integrator,
                      double m_i, double gamma, double epsilon, double n_min, double T_min, double e_min,
                      double open_strength, double open_decay)
{
    oracle *o = (oracle *)calloc(1, sizeof(oracle));
    /* ms_electron_heating_fraction defaults: thermalconduction.hpp:34, radiativelosses.hpp:31 (1.0); ambientheating.hpp:28, ambientheatingsink.hpp:28 (0.5);
     * physicalviscosity.hpp:33, localizedheating.hpp:28 (0.0) */
    o->mod.ms_frac[1] = 1.0; o->mod.ms_frac[2] = 1.0; o->mod.ms_frac[3] = 0.5; o->mod.ms_frac[6] = 0.5;
    o->nx = nx; o->ny = ny; o->n = nx * ny;
    o->xb1 = xb1; o->xb2 = xb2; o->yb1 = yb1; o->yb2 = yb2; o->integrator = integrator;
    o->m_i = m_i; o->gamma = gamma; o->epsilon = epsilon; o->n_min = n_min; o->T_min = T_min; o->e_min = e_min;
    o->open_strength = open_strength; o->open_decay = open_decay;
    /* plasmadomain.cpp:138-161 */
    o->xl = (xb1 == BC_PERIODIC) ? 0 : N_GHOST; o->xu = (xb2 == BC_PERIODIC) ? nx - 1 : nx - N_GHOST - 1;
    o->yl = (yb1 == BC_PERIODIC) ? 0 : N_GHOST; o->yu = (yb2 == BC_PERIODIC) ? ny - 1 : ny - N_GHOST - 1;
    double **dom[] = { &o->dx, &o->dy, &o->bex, &o->bey, &o->bez, &o->posx, &o->posy, &o->mask };
    for (size_t k = 0; k < sizeof(dom) / sizeof(dom[0]); k++) *dom[k] = pl_new(o);
    for (int v = 0; v < NV; v++) o->g[v] = pl_new(o);
    for (int i = o->xl; i <= o->xu; i++) for (int j = o->yl; j <= o->yu; j++) o->mask[(size_t)i * ny + j] = 1.0;
    return o;
}
void oracle_destroy(oracle *o)
{
    double *dom[] = { o->dx, o->dy, o->bex, o->bey, o->bez, o->posx, o->posy, o->mask };
    for (size_t k = 0; k < sizeof(dom) / sizeof(dom[0]); k++) free(dom[k]);
    for (int v = 0; v < NV; v++) free(o->g[v]);
    free(o->mod.ah_heating);
    free(o->mod.pv_cg);
    for (int k = 0; k < 4; k++) free(o->mod.pv_avg[k]);
    for (int w = 0; w < 4; w++) for (int i = 0; i < 8; i++) free(o->mod.av_out[w][i]);
    for (int k = 0; k < 3; k++) free(o->mod.ms_cum[k]);
    free(o);
}
/* which: 0 d_x 1 d_y 2 be_x 3 be_y 4 be_z 5 pos_x 6 pos_y 7 mask(read only) ; 100+v equation-set variable v */
double *oracle_plane(oracle *o, int which)
{
    double *dom[] = { o->dx, o->dy, o->bex, o->bey, o->bez, o->posx, o->posy, o->mask };
    if (which >= 100 && which < 100 + NV) return o->g[which - 100];
    if (which >= 0 && which < 8) return dom[which];
    return NULL;
}
/* equationset.cpp:96-104 + idealmhd.cpp:226-232 */
void oracle_setup(oracle *o)
{
    double **G = o->g;
    for (int c = 0; c < o->n; c++) {
        G[V_n][c] = smax(G[V_rho][c] / o->m_i, o->n_min);
        G[V_press][c] = (G[V_n][c] * 2.0) * K_B * smax(G[V_temp][c], o->T_min);
        G[V_thermal_energy][c] = G[V_press][c] / (o->gamma - 1.0);
    }
    propagate_changes(o, G, G);
}
void oracle_propagate(oracle *o) { propagate_changes(o, o->g, o->g); }
void oracle_set_thermal_conduction(oracle *o, int flux_saturation, int integrator, double epsilon, double dt_subcycle_min, double weakening)
{
    o->mod.tc_on = 1; o->mod.tc_flux_saturation = flux_saturation; o->mod.tc_integrator = integrator;
    o->mod.tc_epsilon = epsilon; o->mod.tc_dt_subcycle_min = dt_subcycle_min; o->mod.tc_weakening = weakening;
    o->mod.order[o->mod.n_modules++] = 1;
}
void oracle_set_radiative_losses(oracle *o, int integrator, double cutoff_ramp, double cutoff_temp, double epsilon, int prevent_subcycling)
{
    o->mod.rl_on = 1; o->mod.rl_integrator = integrator; o->mod.rl_cutoff_ramp = cutoff_ramp; o->mod.rl_cutoff_temp = cutoff_temp;
    o->mod.rl_epsilon = epsilon; o->mod.rl_prevent_subcycling = prevent_subcycling;
    o->mod.order[o->mod.n_modules++] = 2;
}
/* ambientheating.cpp:28-40 ; exp_mode: heating = mask*base*exp(-pos_y/H) (optionally max with the split profile) */
void oracle_set_ambient_heating(oracle *o, double heating_rate, int exp_mode, double exp_base, double exp_scale_height,
                                int split_exp_mode, double split_scale_height, double split_start_height)
{
    o->mod.ah_on = 1; o->mod.ah_heating = pl_new(o);
    for (int c = 0; c < o->n; c++) {
        if (!exp_mode) o->mod.ah_heating[c] = o->mask[c] * heating_rate;
        else {
            double h = (o->mask[c] * exp_base) * exp((o->posy[c] * -1.0) / exp_scale_height);
            if (split_exp_mode) {
                double sb = exp_base * exp((exp_scale_height - split_scale_height) * split_start_height / (exp_scale_height * split_scale_height));
                double h2 = (o->mask[c] * sb) * exp((o->posy[c] * -1.0) / split_scale_height);
                h = smax(h, h2);
            }
            o->mod.ah_heating[c] = h;
        }
    }
    o->mod.order[o->mod.n_modules++] = 3;
}
void oracle_set_viscosity(oracle *o, int hv_integrator, double hv_epsilon, int gradient_correction)
{
    o->mod.av_on = 1; o->mod.av_hv_integrator = hv_integrator; o->mod.av_hv_epsilon = hv_epsilon; o->mod.av_gradient_correction = gradient_correction;
    o->mod.order[o->mod.n_modules++] = 4;
}
/* strength_grid: nx*ny profile for the boundary options (copied), NULL otherwise */
void oracle_add_viscosity_term(oracle *o, int opt, double strength, int var_diff, int var_evol, int species, const double *strength_grid)
{
    int i = o->mod.av_nterms++;
    o->mod.av_opt[i] = opt; o->mod.av_strength[i] = strength; o->mod.av_diff[i] = var_diff; o->mod.av_evol[i] = var_evol; o->mod.av_species[i] = species;
    o->mod.av_strength_grid[i] = NULL;
    if (strength_grid) { o->mod.av_strength_grid[i] = pl_dup(o, strength_grid); }
    for (int w = 0; w < 4; w++) o->mod.av_out[w][i] = pl_new(o);
}
/* test accessor: plane `which` (0 dqdt, 1 lap, 2 str, 3 dt) of viscosity term i as fileOutput would write it now */
int oracle_viscosity_output(const oracle *o, int which, int i, double *out)
{
    if (which < 0 || which > 3 || i < 0 || i >= o->mod.av_nterms) return 0;
    memcpy(out, o->mod.av_out[which][i], sizeof(double) * o->n);
    return 1;
}
/* kind: 6 ambient_heating_sink, 7 localized_heating, 8 mass_injection, 9 momentum_injection, 10 div_cleaning, 11 field_heating; p: see small_module_setup */
void oracle_add_small_module(oracle *o, int kind, const double *p, int np)
{
    small_module *m = (small_module *)calloc(1, sizeof(small_module));
    m->kind = kind;
    for (int k = 0; k < np && k < 12; k++) m->p[k] = p[k];
    small_module_setup(o, m);
    o->mod.small[o->mod.n_small] = m;
    o->mod.order[o->mod.n_modules++] = 100 + o->mod.n_small++;
}
/* test accessor: template / coefficient plane k of small module idx; returns 0 when that plane does not exist (yet: the Gaussian templates
 * of kinds 7 and 8 are built by the first preIterateModule) */
int oracle_small_module_plane(oracle *o, int idx, int k, double *out)
{
    if (idx < 0 || idx >= o->mod.n_small || k < 0 || k > 1) return 0;
    const small_module *m = (const small_module *)o->mod.small[idx];
    if (!m->plane[k]) return 0;
    memcpy(out, m->plane[k], sizeof(double) * o->n);
    return 1;
}
/* test accessor: output_to_file plane k (F_x, F_y, dP_x, dP_y) of the coulomb_explosion module idx; returns -1 for another kind, else 1 + (the reference would have aborted) */
int oracle_coulomb_plane(oracle *o, int idx, int k, double *out)
{
    if (idx < 0 || idx >= o->mod.n_small || k < 0 || k > 3) return -1;
    const small_module *m = (const small_module *)o->mod.small[idx];
    if (m->kind != 14) return -1;
    memcpy(out, ((const coulomb_explosion_t *)m->ext)->var[k], sizeof(double) * o->n);
    return 1 + m->flag[0];
}
/* test accessor: BoundaryOutflow::computeMeanOutflow of small module idx on the current state */
double oracle_outflow_mean(oracle *o, int idx)
{
    if (idx < 0 || idx >= o->mod.n_small) return 0.0;
    const small_module *m = (const small_module *)o->mod.small[idx];
    return m->kind == 12 ? outflow_mean(o, m) : 0.0;
}
/* p: time_scale, frobenius_metric_coeff, smoothing_sigma, safety_factor, metric_smoothing, time_integrator, flood_fill (1) / frobenius (0), flood_fill_max_radius,
 *    flood_fill_argmin_radius, flood_fill_min_current, flood_fill_current_ramp_length, flood_fill_threshold, resistivity_model (0 time_scale, 1 syntelis_19, 2 ys_94),
 *    gradient_correction, model parameter 0..2.  The module's setupModule needs the populated state: call after oracle_setup. */
void oracle_set_anomalous_resistivity(oracle *o, const double *p)
{
    anom_res *A = (anom_res *)calloc(1, sizeof(anom_res));
    A->time_scale = p[0]; A->frob_coeff = p[1]; A->sigma = p[2]; A->safety = p[3]; A->smoothing = (int)p[4]; A->integrator = (int)p[5]; A->flood_fill = (int)p[6];
    A->max_radius = p[7]; A->argmin_radius = p[8]; A->min_current = p[9]; A->ramp_length = p[10]; A->threshold = p[11]; A->model = (int)p[12]; A->gc = (int)p[13];
    A->params[0] = p[14]; A->params[1] = p[15]; A->params[2] = p[16];
    o->mod.anom = A;
    ar_setup(o, A);
    o->mod.order[o->mod.n_modules++] = 13;
}
/* test accessors: one iterateModule(dt) of anomalous_resistivity on the current planes WITHOUT the write-back + propagateChanges; out = bi_x, bi_y, bi_z,
 * thermal_energy (4 planes).  raw_commit: store the result into the planes as it is (no floors, no boundary pass), so that a second call continues from it. */
void oracle_anomalous_core(oracle *o, double dt, double *out, int raw_commit)
{
    anom_res *A = (anom_res *)o->mod.anom;
    ar_capture = out;
    ar_iterate(o, A, dt);
    ar_capture = NULL;
    if (raw_commit) {
        const size_t n = o->n;
        memcpy(o->g[V_bi_x], out, sizeof(double) * n); memcpy(o->g[V_bi_y], out + n, sizeof(double) * n); memcpy(o->g[V_bi_z], out + 2 * n, sizeof(double) * n);
        memcpy(o->g[V_thermal_energy], out + 3 * n, sizeof(double) * n);
    }
}
/* multispecies_mode = true (fileio.cpp:322; planes zeroed at construction, plasmadomain.cpp:55-59) and the modules' ms_electron_heating_fraction */
void oracle_set_multispecies(oracle *o, int on)
{
    o->mod.ms_on = on;
    for (int k = 0; k < 3; k++) { if (!o->mod.ms_cum[k]) o->mod.ms_cum[k] = pl_new(o); memset(o->mod.ms_cum[k], 0, sizeof(double) * o->n); }
}
/* inactive_mode of a module: 1 thermal_conduction, 2 radiative_losses */
void oracle_set_module_inactive(oracle *o, int module, int on) { if (module == 1) o->mod.tc_inactive = on; else if (module == 2) o->mod.rl_inactive = on; }
void oracle_set_ms_fraction(oracle *o, int module, double f) { if (module >= 0 && module < 8) o->mod.ms_frac[module] = f; }
/* which: 0 cumulative_electron_heating, 1 cumulative_ion_heating, 2 cumulative_joule_heating */
int oracle_ms_plane(const oracle *o, int which, double *out)
{
    if (!o->mod.ms_on || which < 0 || which > 2) return 0;
    memcpy(out, o->mod.ms_cum[which], sizeof(double) * o->n);
    return 1;
}
/* the reset after every stored frame (evolution.cpp:36-41) */
void oracle_ms_reset(oracle *o) { if (o->mod.ms_on) for (int k = 0; k < 3; k++) memset(o->mod.ms_cum[k], 0, sizeof(double) * o->n); }
/* test accessor: PhysicalViscosity::iterateModule alone (sub-cycle count, sub-cycles, closing propagateChanges) */
void oracle_physical_viscosity_iterate(oracle *o, double dt) { pv_iterate(o, dt); }
void oracle_anomalous_iterate(oracle *o, double dt) { ar_iterate(o, (anom_res *)o->mod.anom, dt); }
void oracle_anomalous_state(const oracle *o, int *null_ij, double *tmpl)
{
    const anom_res *A = (const anom_res *)o->mod.anom;
    null_ij[0] = A->null_i; null_ij[1] = A->null_j;
    if (tmpl) memcpy(tmpl, A->tmpl, sizeof(double) * o->n);
}
void oracle_anomalous_diffusivity(const oracle *o, double *out) { memcpy(out, ((const anom_res *)o->mod.anom)->diffusivity, sizeof(double) * o->n); }
/* test accessor: output_to_file plane of thermal_conduction / radiative_losses / physical_viscosity after the last step: 0 thermal_conduction, 1 flux_saturation, 2 rad,
 * 3 viscous_heating, 4-6 viscous_force_x/y/z */
int oracle_module_output(const oracle *o, int which, double *out)
{
    const double *p = which == 0 ? o->mod.tc_avg : which == 1 ? o->mod.tc_sat : which == 2 ? o->mod.rl_avg : (which >= 3 && which <= 6) ? o->mod.pv_avg[which - 3] : NULL;
    if (!p) return 0;
    memcpy(out, p, sizeof(double) * o->n);
    return 1;
}
int oracle_anomalous_subcycles(const oracle *o) { return o->mod.anom ? ((anom_res *)o->mod.anom)->nsub : 0; }
void oracle_set_global_viscosity(oracle *o, double v) { o->global_viscosity = v; }
/* test accessor: applyMomThresholdingMoC + applyBThresholdingMoC on the current planes, nothing else */
void oracle_apply_moc_thresholding(oracle *o) { moc_thresholding(o, o->g); }
/* test accessor: SGFilter::singleVarSavitzkyGolay on an arbitrary plane (for checking the product's host-resident filter) */
void oracle_sg_filter(oracle *o, double *plane) { sg_filter_plane(o, plane); }
/* moc_b_limiting / moc_mom_limiting with their bounds (idealmhd.cpp:17-36; defaults 0.1 and 10.0, idealmhd.hpp:59-64) */
void oracle_set_moc_limiting(oracle *o, int b_on, double b_lo, double b_hi, int mom_on, double mom_lo, double mom_hi)
{
    o->moc_b_limiting = b_on; o->moc_b_lim[0] = b_lo; o->moc_b_lim[1] = b_hi;
    o->moc_mom_limiting = mom_on; o->moc_mom_lim[0] = mom_lo; o->moc_mom_lim[1] = mom_hi;
}
/* test accessor: only the small solar modules' hooks of one phase on the current planes (0 preIterate, 1 iterate, 2 postIterate); time is not advanced */
void oracle_small_module_hooks(oracle *o, int phase, double step)
{
    for (int m = 0; m < o->mod.n_modules; m++) {
        if (o->mod.order[m] < 100) continue;
        small_module *sm = (small_module *)o->mod.small[o->mod.order[m] - 100];
        if (phase == 0) small_module_pre(o, sm);
        else if (phase == 1) small_module_iterate(o, sm, step);
        else small_module_post(o, sm, step);
    }
}
void oracle_set_time(oracle *o, double t) { o->t = t; }
double oracle_step(oracle *o) { return advance_time(o); }
void oracle_run(oracle *o, int nsteps, double *dt_out) { for (int s = 0; s < nsteps; s++) { double d = advance_time(o); if (dt_out) dt_out[s] = d; } }
double oracle_time(const oracle *o) { return o->t; }
int oracle_subcycles(const oracle *o, int which) { return which == 1 ? o->mod.tc_nsub : which == 5 ? o->mod.pv_nsub : o->mod.rl_nsub; }
void oracle_set_physical_viscosity(oracle *o, double coeff, const double *coeff_plane, double epsilon, int heating_on, int force_on, int gradient_correction,
                                   int integrator, int inactive_mode)
{
    o->mod.pv_on = 1; o->mod.pv_coeff = coeff; o->mod.pv_cg = pl_dup(o, coeff_plane); o->mod.pv_epsilon = epsilon; o->mod.pv_heating_on = heating_on;
    o->mod.pv_force_on = force_on; o->mod.pv_gc = gradient_correction; o->mod.pv_integrator = integrator; o->mod.pv_inactive = inactive_mode; o->mod.pv_nsub = 1;
    o->mod.order[o->mod.n_modules++] = 5;
}
/* one evaluation of the ideal-MHD right-hand side on the primary state; k: 8 planes [NEV][n] contiguous */
void oracle_rhs(oracle *o, double *k)
{
    double *kk[NEV]; for (int v = 0; v < NEV; v++) kk[v] = k + (size_t)v * o->n;
    compute_time_derivatives(o, o->g, kk);
}
/* operator known-answer entry points: op 0 derivative1D, 1 secondDerivative1D, 2 laplacian, 3 transportDerivative1D(q, vel) */
void oracle_operator(oracle *o, int op, int index, const double *q, const double *vel, double *out)
{
    if (op == 0) derivative1D(o, q, index, out);
    else if (op == 1) second_derivative1D(o, q, index, out);
    else if (op == 2) laplacian(o, q, out);
    else if (op == 3) transport_derivative1D(o, q, vel, index, out);
}

/* ---------------------------------------------------------------------------------------------------------
 * Two-fluid equation set (Ideal2F) + EIC thermalization: separate restatement on top of the operators above.
 * --------------------------------------------------------------------------------------------------------- */
#include "ideal2f_oracle.inc"
#include "ideal_mhd2e_oracle.inc"
