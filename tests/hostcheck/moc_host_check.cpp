// moc_host_check.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-cell open_moc arithmetic (spruce_b200/csrc/moc_kernels.cuh,
// namespace spruce::moc -- the very functions the CUDA kernel k_moc_stage calls) with the host compiler, so that
// tests/test_moc_host_check.py can compare it bit for bit against the CPU oracle without a GPU.  Nothing in the product links this.
#include "../../spruce_b200/csrc/moc_kernels.cuh"

extern "C" int moc_host_terms(const double *const *planes /* n mx my mz e bix biy biz bex bey bez gx gy */, const double *dx, const double *dy,
                              int nx, int ny, const int *bc, double m_i, double gamma, double visc, double *k_out /* [8][nx*ny] */, unsigned char *owned)
{
    spruce::moc::Field F;
    F.n = planes[0]; F.mx = planes[1]; F.my = planes[2]; F.mz = planes[3]; F.e = planes[4]; F.bix = planes[5]; F.biy = planes[6]; F.biz = planes[7];
    F.bex = planes[8]; F.bey = planes[9]; F.bez = planes[10]; F.gx = planes[11]; F.gy = planes[12];
    F.dx = dx; F.dy = dy; F.nx = nx; F.ny = ny; F.pitch = ny;
    for (int s = 0; s < 4; s++) F.bc[s] = bc[s];
    F.m_i = m_i; F.gamma = gamma; F.gm1 = gamma - 1.0; F.visc = visc; F.x_halo = 0;
    const size_t n = (size_t)nx * ny;
    int count = 0;
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
        double k[8];
        const bool own = spruce::moc::moc_cell_terms(F, i, j, k);
        owned[(size_t)i * ny + j] = own ? 1 : 0;
        if (!own) continue;
        count++;
        for (int v = 0; v < 8; v++) k_out[v * n + (size_t)i * ny + j] = k[v];
    }
    return count;
}

// one euler step of the evolved ghost cells (D = P + step*k(P), floors, pointwise zeroing) and their new dt
extern "C" int moc_host_euler(const double *const *planes, const double *dx, const double *dy, int nx, int ny, const int *bc, double m_i, double gamma, double visc,
                              double n_min, double e_min, double step, double *out /* [8][nx*ny] */, double *dt_out, unsigned char *owned)
{
    spruce::moc::Field F;
    F.n = planes[0]; F.mx = planes[1]; F.my = planes[2]; F.mz = planes[3]; F.e = planes[4]; F.bix = planes[5]; F.biy = planes[6]; F.biz = planes[7];
    F.bex = planes[8]; F.bey = planes[9]; F.bez = planes[10]; F.gx = planes[11]; F.gy = planes[12];
    F.dx = dx; F.dy = dy; F.nx = nx; F.ny = ny; F.pitch = ny;
    for (int s = 0; s < 4; s++) F.bc[s] = bc[s];
    F.m_i = m_i; F.gamma = gamma; F.gm1 = gamma - 1.0; F.visc = visc; F.x_halo = 0;
    const spruce::moc::Floors fl{n_min, e_min};
    const size_t n = (size_t)nx * ny;
    int count = 0;
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
        const size_t q = (size_t)i * ny + j;
        double k[8];
        const bool own = spruce::moc::moc_cell_terms(F, i, j, k);
        owned[q] = own ? 1 : 0;
        if (!own) continue;
        count++;
        double base[8];
        for (int v = 0; v < 8; v++) base[v] = planes[v][q];
        const spruce::moc::Updated u = spruce::moc::advance_cell(F, fl, base, k, 1.0 * step, true, i, j);
        const double vals[8] = {u.n, u.mx, u.my, u.mz, u.e, u.bx, u.by, u.bz};
        for (int v = 0; v < 8; v++) out[v * n + q] = vals[v];
        dt_out[q] = spruce::moc::in_dt_bounds(F, i, j) && !spruce::moc::in_interior(F, i, j)
                        ? spruce::moc::cell_dt_plain(F, u.n, u.mx, u.my, u.e, F.bex[q] + u.bx, F.bey[q] + u.by, F.bez[q] + u.bz, dx[i], dy[j]) : -1.0;
    }
    return count;
}

// the strip kernels' thread -> cell mapping: visits[i*ny + j] = number of threads that own cell (i, j); returns the thread count
extern "C" int moc_host_thread_visits(int nx, int ny, const int *bc, int *visits)
{
    spruce::moc::Field F{};
    F.nx = nx; F.ny = ny; F.pitch = ny;
    for (int s = 0; s < 4; s++) F.bc[s] = bc[s];
    const int T = spruce::moc::n_threads(nx, ny);
    for (int t = 0; t < T + 64; t++) {                  // past-the-end threads of the last block must map to nothing
        int side, i, j;
        if (!spruce::moc::thread_cell(nx, ny, 0, nx, t, &side, &i, &j)) continue;
        if (t >= T) return -1;
        if (i < 0 || i >= nx || j < 0 || j >= ny) return -2;
        if (spruce::moc::thread_owns(F, side, i, j)) visits[(size_t)i * ny + j]++;
    }
    return T;
}


// The same right-hand side evaluated the way the slabs of a decomposed run evaluate it: every rank sees only its rows plus two halo rows on each
// side (filled from the ring neighbour when x is periodic, NaN beyond a physical x boundary), addresses them by global row through shifted
// pointers exactly as moc_field() in moc_stage.cuh does, and runs the strip kernels' thread loop.  visits counts owners per cell over all ranks.
#include <cmath>
#include <vector>
extern "C" int moc_host_terms_slabs(const double *const *planes, const double *dx, const double *dy, int nx, int ny, const int *bc, double m_i, double gamma, double visc,
                                    int n_ranks, double *k_out, int *visits)
{
    const int H = spruce::moc::NG, APR = 3;
    const bool xper = bc[0] == 0;
    const size_t n = (size_t)nx * ny;
    for (int rank = 0; rank < n_ranks; rank++) {
        const int row0 = (int)((long long)nx * rank / n_ranks), nxl = (int)((long long)nx * (rank + 1) / n_ranks) - row0;
        std::vector<std::vector<double>> loc(13, std::vector<double>((size_t)(nxl + 2 * H) * ny, std::nan("")));
        for (int v = 0; v < 13; v++) for (int r = -H; r < nxl + H; r++) {
            int g = row0 + r;
            if (g < 0 || g >= nx) { if (!xper) continue; g = (g + nx) % nx; }
            for (int j = 0; j < ny; j++) loc[v][(size_t)(r + H) * ny + j] = planes[v][(size_t)g * ny + j];
        }
        std::vector<double> dxl(nxl + 2 * APR, 1.0);
        for (int k = 0; k < nxl + 2 * APR; k++) { int g = row0 + k - APR; if (g < 0 || g >= nx) { if (!xper) continue; g = (g + nx) % nx; } dxl[k] = dx[g]; }
        spruce::moc::Field F;
        const double *base[13];
        for (int v = 0; v < 13; v++) base[v] = loc[v].data() + (size_t)H * ny - (long long)row0 * ny;      // global row indexing
        F.n = base[0]; F.mx = base[1]; F.my = base[2]; F.mz = base[3]; F.e = base[4]; F.bix = base[5]; F.biy = base[6]; F.biz = base[7];
        F.bex = base[8]; F.bey = base[9]; F.bez = base[10]; F.gx = base[11]; F.gy = base[12];
        F.dx = dxl.data() + APR - row0; F.dy = dy; F.nx = nx; F.ny = ny; F.pitch = ny;
        for (int s = 0; s < 4; s++) F.bc[s] = bc[s];
        F.m_i = m_i; F.gamma = gamma; F.gm1 = gamma - 1.0; F.visc = visc;
        F.x_halo = (xper && n_ranks > 1) ? 1 : 0;
        const int T = spruce::moc::n_threads(nxl, ny);
        for (int t = 0; t < T + 64; t++) {
            int side, i, j;
            if (!spruce::moc::thread_cell(nxl, ny, row0, nx, t, &side, &i, &j)) continue;
            if (t >= T) return -1;
            if (i < row0 || i >= row0 + nxl || j < 0 || j >= ny) return -2;                                // a rank only ever writes its own rows
            if (!spruce::moc::thread_owns(F, side, i, j)) continue;
            visits[(size_t)i * ny + j]++;
            double k[8];
            spruce::moc::moc_cell_terms(F, i, j, k);
            for (int v = 0; v < 8; v++) k_out[v * n + (size_t)i * ny + j] = k[v];
        }
    }
    return 0;
}

// the limiter passes on arbitrary planes: mutable[6] = mom_x, mom_y, mom_z, bi_x, bi_y, bi_z (in place); be[3]
extern "C" void moc_host_limit(double *const *mut, const double *const *be, int nx, int ny, const int *bc, int b_on, double b_lo, double b_hi, int mom_on, double mom_lo, double mom_hi)
{
    spruce::moc::Field F{};
    F.nx = nx; F.ny = ny; F.pitch = ny; F.x_halo = 0;
    for (int s = 0; s < 4; s++) F.bc[s] = bc[s];
    F.bex = be[0]; F.bey = be[1]; F.bez = be[2];
    const spruce::moc::Mutable U{mut[0], mut[1], mut[2], mut[3], mut[4], mut[5]};
    const spruce::moc::Limits L{b_on, mom_on, b_lo, b_hi, mom_lo, mom_hi};
    for (int s = 0; s < 4; s++) for (int a = 0; a < (s < 2 ? ny : nx); a++) spruce::moc::limit_line(F, U, L, s, a);    // sides in order, lines in any order
}
