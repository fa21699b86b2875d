// anomres_host_check.cpp -- TEST INFRASTRUCTURE.  Runs the product's anomalous_resistivity module (spruce_b200/csrc/anomres_cells.hpp: the functors AND the
// sequence of passes the device executor instantiates) on the HOST: "launch a kernel over the cells" becomes "loop over the cells".
// tests/test_anomres_host_check.py compares the result bit for bit with the CPU restatement that is pinned to the reference.  Nothing in the product links this.
#include "../../spruce_b200/csrc/anomres_cells.hpp"
#include <vector>

using namespace spruce::ar;

struct HostExec {
    Geom g;
    std::vector<std::vector<double>> planes;
    int reverse = 0;                                       // visit the cells backwards: the passes must not depend on the order
    double *plane(int slot) { return planes[slot].data(); }
    double host_px(int i, int j) const { return g.px[at(g, i, j)]; }
    double host_py(int i, int j) const { return g.py[at(g, i, j)]; }
    template <class F> void cells(const F &f)
    {
        if (!reverse) { for (int i = 0; i < g.nx; i++) for (int j = 0; j < g.ny; j++) f(i, j); }
        else { for (int i = g.nx - 1; i >= 0; i--) for (int j = g.ny - 1; j >= 0; j--) f(i, j); }
    }
    template <class F> int reduce_min(const F &f, double *m)
    {
        double v = kHuge;
        if (!reverse) { for (int i = 0; i < g.nx; i++) for (int j = 0; j < g.ny; j++) v = smin2(v, f(i, j)); }
        else { for (int i = g.nx - 1; i >= 0; i--) for (int j = g.ny - 1; j >= 0; j--) v = smin2(v, f(i, j)); }
        *m = v;
        return 0;
    }
};

// in: bi_x, bi_y, bi_z, thermal_energy, be_x, be_y, be_z, n, dt, pos_x, pos_y (11 planes); p[17] laid out like oracle_set_anomalous_resistivity
// out: 4 planes after `iters` iterateModule(dt) calls (each continuing from the raw result of the one before), template and diffusivity planes, null point, nsub
extern "C" int anomres_host_run(const double *const *in, const double *dx, const double *dy, int nx, int ny, const int *bounds, const int *per, const int *moc_ext,
                                const double *p, double epsilon, double dt, int iters, int reverse, double *out, double *tmpl, double *diff, int *null_ij, int *nsub)
{
    HostExec x;
    x.reverse = reverse;
    x.g = Geom{nx, ny, ny, bounds[0], bounds[1], bounds[2], bounds[3], per[0], per[1], dx, dy, in[9], in[10]};
    x.planes.assign(P_COUNT, std::vector<double>((size_t)nx * ny, 0.0));
    State s{};
    s.p.time_scale = p[0]; s.p.frob_coeff = p[1]; s.p.sigma = p[2]; s.p.safety = p[3]; s.p.smoothing = (int)p[4]; s.p.integrator = (int)p[5]; s.p.flood_fill = (int)p[6];
    s.p.max_radius = p[7]; s.p.argmin_radius = p[8]; s.p.min_current = p[9]; s.p.ramp_length = p[10]; s.p.threshold = p[11]; s.p.model = (int)p[12];
    s.p.gradient_correction = (int)p[13]; s.p.model_params[0] = p[14]; s.p.model_params[1] = p[15]; s.p.model_params[2] = p[16];
    s.time_scale = s.p.time_scale;
    std::vector<double> w;
    if (s.p.smoothing) { s.kr = smoothing_radius(s.p.sigma); w.resize((size_t)(2 * s.kr + 1) * (2 * s.kr + 1)); smoothing_kernel(s.p.sigma, s.kr, w.data()); s.kernel = w.data(); }
    const size_t n = (size_t)nx * ny;
    int rc;
    if ((rc = setup(x, x.g, s, in[4], in[5], in[0], in[1], in[2]))) return rc;
    for (int q = 0; q < 4; q++) std::copy(in[q], in[q] + n, x.plane(P_BIX + q));
    for (int it = 0; it < iters; it++)
        if ((rc = iterate(x, x.g, s, in[4], in[5], in[6], in[7], in[8], moc_ext, epsilon, dt))) return rc;
    for (int q = 0; q < 4; q++) std::copy(x.plane(P_BIX + q), x.plane(P_BIX + q) + n, out + q * n);
    std::copy(x.plane(P_TMPL), x.plane(P_TMPL) + n, tmpl);
    std::copy(x.plane(P_DIFF), x.plane(P_DIFF) + n, diff);
    null_ij[0] = s.null_i; null_ij[1] = s.null_j; *nsub = s.nsub;
    return 0;
}
