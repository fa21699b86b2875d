// stage_relaxed.cu -- the fused Runge-Kutta stage kernel in RELAXED arithmetic: the same source as the exact build (mhd_stage_xy.cuh), compiled
// as its own translation unit WITH fused multiply-add contraction (-fmad=true) and with a/b for table divisors as one multiplication by the
// correctly rounded reciprocal (SPRUCE_RELAXED in exact_math.cuh: <= 1.5 ulp instead of 5 operations for the correctly rounded quotient).
//
// Why it exists: the north star states its parity bar as "all state fields within <= 1e-9 relative L-infinity after 100 steps, since FMA
// contraction and reduction order differ"; the default build goes further (every operation individually rounded, bit-identical fields and
// step-size history) and pays for it in FP64 issue slots.  This unit is the opt-in other end of that trade:
// SPRUCE_ARITH=relaxed at domain creation.  Results then agree with the reference to rounding-error growth, not bit for bit, and the step-size
// history agrees to ~1e-15 relative, not bit for bit.  Everything else (propagate, ghost passes, modules, dt minimum) stays the exact code.
// Every symbol of the included headers lands in namespace spruce_relaxed, so nothing here can be confused with the exact instances at link time.
// Validated on a B200 within the north star 1e-9 bar (tests/test_gpu_extended.py).
#define SPRUCE_RELAXED 1
#define spruce spruce_relaxed
#include "mhd_stage_xy.cuh"
#undef spruce

namespace R = spruce_relaxed;

namespace {
template <int LN, unsigned long long LQ, int VAR>
cudaError_t launch_one(dim3 grid, cudaStream_t st, const R::DomainParams &P, const R::StageArgs &A, const R::ActiveList &L)
{
    const size_t smem = R::xy_smem_bytes(R::xy_rows(LN), VAR);
    static bool configured = false;               // one attribute call per instance and process
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(R::k_mhd_stage_xy<LN, LQ, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    R::k_mhd_stage_xy<LN, LQ, VAR><<<grid, R::XY_NT, smem, st>>>(P, A, L);
    return cudaGetLastError();
}
template <int LN, unsigned long long LQ>
cudaError_t launch_var(int var, dim3 grid, cudaStream_t st, const R::DomainParams &P, const R::StageArgs &A, const R::ActiveList &L)
{
    if (var == 1) return launch_one<LN, LQ, 1>(grid, st, P, A, L);
    if (var == 2) return launch_one<LN, LQ, 2>(grid, st, P, A, L);
    if (var == 3) return launch_one<LN, LQ, 3>(grid, st, P, A, L);
    return launch_one<LN, LQ, 0>(grid, st, P, A, L);
}
}  // namespace

// P, A, L: the exact build's DomainParams / StageArgs / ActiveList (same definitions, hence the same layout).  list: 6 = the 2-D list, 12 = the full
// list (capi.cu: active_quantities); var: the compile-time integrator stage (0 = run-time).  Returns a cudaError_t.
extern "C" __attribute__((visibility("hidden"))) int spruce_relaxed_launch_stage(unsigned gx, unsigned gy, void *stream, const void *P, const void *A, const void *L, int list, int var)
{
    const R::DomainParams &p = *static_cast<const R::DomainParams *>(P);
    const R::StageArgs &a = *static_cast<const R::StageArgs *>(A);
    const R::ActiveList &l = *static_cast<const R::ActiveList *>(L);
    const dim3 grid(gx, gy);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (list == 6) return (int)launch_var<6, R::XY_LIST_2D>(var, grid, st, p, a, l);
    if (list == 12) return (int)launch_var<12, R::XY_LIST_FULL>(var, grid, st, p, a, l);
    return (int)cudaErrorInvalidValue;
}
