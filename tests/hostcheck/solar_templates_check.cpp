// solar_templates_check.cpp -- TEST INFRASTRUCTURE.  Exposes the product's host-side template builders (spruce_b200/csrc/solar_templates.hpp,
// the functions the spruce_module_localized_heating / mass_injection / momentum_injection entry points call) so that
// tests/test_solar_templates_host_check.py can compare them with the pinned CPU restatement without a GPU.
#include "../../spruce_b200/csrc/solar_templates.hpp"
#include <cstring>

static spruce::solar::Geom geom(int xdim, int ydim, int row0, int nx_local, int xper, int yper)
{
    spruce::solar::Geom g;
    g.nx_local = nx_local; g.ny = ydim; g.row0 = row0; g.xdim = xdim; g.ydim = ydim; g.x_periodic = xper != 0; g.y_periodic = yper != 0;
    return g;
}
extern "C" void tmpl_positive(int xdim, int ydim, int row0, int nx_local, int xper, int yper, double peak, double sx, double sy, double cx, double cy, double *out)
{
    std::vector<double> p;
    spruce::solar::positive_template(geom(xdim, ydim, row0, nx_local, xper, yper), peak, sx, sy, cx, cy, p);
    std::memcpy(out, p.data(), p.size() * sizeof(double));
}
extern "C" void tmpl_momentum(int xdim, int ydim, int row0, int nx_local, int xper, int yper, double sx, double sy, double cx, double cy, double dir_x, double dir_y,
                              double angle, double *out_x, double *out_y)
{
    std::vector<double> px, py;
    spruce::solar::momentum_templates(geom(xdim, ydim, row0, nx_local, xper, yper), sx, sy, cx, cy, dir_x, dir_y, angle, px, py);
    std::memcpy(out_x, px.data(), px.size() * sizeof(double));
    std::memcpy(out_y, py.data(), py.size() * sizeof(double));
}
// BoundaryOutflow: accel_template and the window of computeMeanOutflow; bc codes as in include/spruce_b200.h (periodic 0, open_moc 4)
extern "C" void tmpl_outflow(int xdim, int ydim, const int *bc, const double *x, const double *y, double length, double feather, int boundary, int shape,
                             double *out, int *win)
{
    const spruce::solar::Window w = spruce::solar::outflow_bounds(xdim, ydim, bc, boundary, 0, 4, 2);
    std::vector<double> t;
    spruce::solar::outflow_template(xdim, ydim, w, x, y, length, feather, boundary, shape, t);
    std::memcpy(out, t.data(), t.size() * sizeof(double));
    const spruce::solar::Window m = spruce::solar::outflow_mean_window(ydim, w, x, y, length, feather, boundary);
    win[0] = m.xl; win[1] = m.xu; win[2] = m.yl; win[3] = m.yu;
}
