// anomres_host.cuh -- device executor of the anomalous_resistivity module: the passes of anomres_cells.hpp (the functors and their sequence, proven
// on the host by tests/test_anomres_host_check.py) as kernel launches, one thread per cell; minima come back through the domain's reduction slot.
// Included by capi.cu inside its anonymous namespace.  A whole domain on one rank only.
// STATUS: written after the round-1 GPU budget was spent; the launch side has not run on a GPU yet (tests/test_zz_gpu_unvalidated.py).
#pragma once

template <class F>
__global__ void __launch_bounds__(128) k_ar_cells(int ny, const F f)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ny) f((int)blockIdx.y, j);
}
template <class F>
__global__ void __launch_bounds__(128) k_ar_min(int ny, const F f, unsigned long long *out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double v = ar::kHuge;
    if (j < ny) v = f((int)blockIdx.y, j);
    block_min_to_global(v, out);
}

struct ArExec {
    spruce_domain *d;
    int err = SPRUCE_OK;
    double *plane(int slot) { return d->ar.planes[slot]; }
    double host_px(int i, int j) const { return d->ar.hpx[(size_t)i * d->P.ny + j]; }
    double host_py(int i, int j) const { return d->ar.hpy[(size_t)i * d->P.ny + j]; }
    dim3 grid() const { return dim3((d->P.ny + 127) / 128, d->P.nx); }
    template <class F> void cells(const F &f)
    {
        if (err) return;
        k_ar_cells<F><<<grid(), 128, 0, d->stream>>>(d->P.ny, f);
        d->launches++;
        if (cudaGetLastError() != cudaSuccess) err = fail(SPRUCE_ERR_CUDA, "anomalous_resistivity: kernel launch failed");
    }
    template <class F> int reduce_min(const F &f, double *m)
    {
        if (err) return err;
        const unsigned long long init = 0x7FEFFFFFFFFFFFFFULL;
        unsigned long long h = 0;
        CUDA_TRY(cudaMemcpyAsync(d->red, &init, sizeof(init), cudaMemcpyHostToDevice, d->stream));
        k_ar_min<F><<<grid(), 128, 0, d->stream>>>(d->P.ny, f, d->red);
        d->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&h, d->red, sizeof(h), cudaMemcpyDeviceToHost, d->stream));
        CUDA_TRY(cudaStreamSynchronize(d->stream));
        *m = bits_to_double(h);
        return SPRUCE_OK;
    }
};

inline ar::Geom ar_geom(const spruce_domain *d)
{
    return ar::Geom{d->P.nx, d->P.ny, d->P.pitch, d->P.xl, d->P.xu, d->P.yl, d->P.yu, d->P.xper, d->P.yper, d->P.tx.d, d->P.ty.d, d->ar.px, d->ar.py};
}

// AnomalousResistivity::setupModule (anomalousresistivity.cpp:18-44) on the current primary state
int ar_setup_run(spruce_domain *d)
{
    ArExec x{d};
    const ar::Geom g = ar_geom(d);
    int rc = ar::setup(x, g, d->ar.s, d->stat[S_BEX], d->stat[S_BEY], d->Pset.p[E_BX], d->Pset.p[E_BY], d->Pset.p[E_BZ]);
    if (rc) return rc;
    if (x.err) return x.err;
    d->ar.ready = true;
    return SPRUCE_OK;
}
// AnomalousResistivity::iterateModule (:107-178)
int ar_iterate(spruce_domain *d, double dt)
{
    int rc;
    ArExec x{d};
    const ar::Geom g = ar_geom(d);
    const int src[4] = {E_BX, E_BY, E_BZ, E_E};
    for (int q = 0; q < 4; q++) x.cells(ar::Copy{g, x.plane(ar::P_BIX + q), d->Pset.p[src[q]]});
    if ((rc = derive_to(d, V_dt, d->ar.dtp))) return rc;
    const int moc_ext[4] = {d->P.bc_x1 == SPRUCE_BC_OPEN_MOC, d->P.bc_x2 == SPRUCE_BC_OPEN_MOC, d->P.bc_y1 == SPRUCE_BC_OPEN_MOC, d->P.bc_y2 == SPRUCE_BC_OPEN_MOC};
    if ((rc = ar::iterate(x, g, d->ar.s, d->stat[S_BEX], d->stat[S_BEY], d->stat[S_BEZ], d->Pset.p[E_N], d->ar.dtp, moc_ext, d->P.epsilon, dt))) return rc;
    if (d->ar.output) {                                                                                                  // :167, against the still untouched primary plane
        const dim3 grid256((d->P.ny + 255) / 256, d->P.nx);
        k_avg_change<<<grid256, 256, 0, d->stream>>>(d->P, d->ar.avg, x.plane(ar::P_E), d->Pset.p[E_E], dt);
        d->launches++;
    }
    if ((rc = ms_feed(d, MS_JOULE, x.plane(ar::P_E), d->Pset.p[E_E], 0.0))) return rc;                                  // m_cumulative_joule_heating, :168-170
    for (int q = 0; q < 4; q++) x.cells(ar::Copy{g, d->Pset.p[src[q]], x.plane(ar::P_BIX + q)});                       // :175-176
    if (x.err) return x.err;
    if ((rc = launch_propagate(d, 0))) return rc;                                                                        // :177
    return after_module_propagate(d);
}
