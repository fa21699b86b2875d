// ideal2f_sides_check.cpp -- TEST INFRASTRUCTURE.  The product's ordered two-fluid boundary passes (spruce_b200/csrc/ideal2f_sides.cuh) compiled for the
// host: applied to arbitrary planes, whole domain or slab by slab, for tests/test_ideal2f_sides_host_check.py to compare with the CPU restatement.
#include "../../spruce_b200/csrc/ideal2f_sides.cuh"
#include <cmath>
#include <cstring>
#include <vector>
using namespace spruce::tf2;

// planes[14][nx*ny] in place.  n_ranks slabs, each with two halo rows per side (NaN: the passes must never touch them), addressed by global row.
extern "C" int tf2_host_sides(double *const *planes, int nx, int ny, const int *bc, int n_ranks)
{
    const int H = 2;
    std::vector<Geo> g(n_ranks);
    std::vector<std::vector<double>> loc[NV];
    for (int v = 0; v < NV; v++) loc[v].resize(n_ranks);
    for (int r = 0; r < n_ranks; r++) {
        Geo &q = g[r];
        q.nx = nx; q.ny = ny; q.pitch = ny;
        q.row0 = (int)((long long)nx * r / n_ranks); q.nxl = (int)((long long)nx * (r + 1) / n_ranks) - q.row0;
        for (int s = 0; s < 4; s++) q.bc[s] = bc[s];
        q.xl = bc[0] == BCT_PERIODIC ? 0 : 2; q.xu = bc[1] == BCT_PERIODIC ? nx - 1 : nx - 3;
        q.yl = bc[2] == BCT_PERIODIC ? 0 : 2; q.yu = bc[3] == BCT_PERIODIC ? ny - 1 : ny - 3;
        for (int v = 0; v < NV; v++) {
            loc[v][r].assign((size_t)(q.nxl + 2 * H) * ny, std::nan(""));
            std::memcpy(&loc[v][r][(size_t)H * ny], planes[v] + (size_t)q.row0 * ny, (size_t)q.nxl * ny * sizeof(double));
        }
    }
    for (int side = 0; side < 4; side++) for (int r = 0; r < n_ranks; r++) {           // sides in order; within a side every slab does its part
        Planes P;
        for (int v = 0; v < NV; v++) P.u[v] = loc[v][r].data() + (size_t)H * ny - (long long)g[r].row0 * ny;
        for (int t = 0; t < side_length(g[r], side); t++) side_line(g[r], P, P, side, t);
    }
    for (int r = 0; r < n_ranks; r++) {
        for (int v = 0; v < NV; v++) {
            for (int k = 0; k < H * ny; k++) if (!std::isnan(loc[v][r][k]) || !std::isnan(loc[v][r][(size_t)(H + g[r].nxl) * ny + k])) return 1;   // a halo row was written
            std::memcpy(planes[v] + (size_t)g[r].row0 * ny, &loc[v][r][(size_t)H * ny], (size_t)g[r].nxl * ny * sizeof(double));
        }
    }
    return 0;
}
