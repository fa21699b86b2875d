// fast_instances.cu -- one kernel per stencil body and instance (general / FAST), so that `cuobjdump -sass` can count each on its own
// (scripts/fast_instances_sass.sh -> profiles/r2_fast_instances_static_sass.txt).  Not part of the library.
#include "../../spruce_b200/csrc/mhd_kernels.cuh"
#include "../../spruce_b200/csrc/module_kernels.cuh"
#include "../../spruce_b200/csrc/ideal2f_kernels.cuh"
using namespace spruce;
#define CELL const int j = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y; const size_t off = (size_t)r * P.pitch + j
template <bool FAST> __device__ __forceinline__ void tc_body(const DomainParams &P, const TcStageArgs &A, double *o) { CELL; o[off] = tc_energy_derivative<FAST>(P, A.C, A.F, r, j); }
__global__ void __launch_bounds__(128) tc_general(const __grid_constant__ DomainParams P, const __grid_constant__ TcStageArgs A, double *o) { tc_body<false>(P, A, o); }
__global__ void __launch_bounds__(128) tc_fast(const __grid_constant__ DomainParams P, const __grid_constant__ TcStageArgs A, double *o) { tc_body<true>(P, A, o); }
template <bool FAST> __device__ __forceinline__ void tf_body(const DomainParams &P, const TfArgs &A, double *o)
{ CELL; double k[NEV2]; for (int v = 0; v < NEV2; v++) k[v] = 0; tf_rhs<FAST>(P, A, r, j, off, k); for (int v = 0; v < NEV2; v++) o[(size_t)v * 16777216 + off] = k[v]; }
__global__ void __launch_bounds__(128) tf_general(const __grid_constant__ DomainParams P, const __grid_constant__ TfArgs A, double *o) { tf_body<false>(P, A, o); }
__global__ void __launch_bounds__(128) tf_fast(const __grid_constant__ DomainParams P, const __grid_constant__ TfArgs A, double *o) { tf_body<true>(P, A, o); }
template <bool FAST> __device__ __forceinline__ void pv_body(const DomainParams &P, const PvArgs &A, double *o)
{ CELL; double h = 0, f[3] = {0, 0, 0}; pv_cell<FAST>(P, A, r, j, true, 1.0, h, f); o[off] = h; o[16777216 + off] = f[0]; o[2 * 16777216 + off] = f[1]; o[3 * 16777216 + off] = f[2]; }
__global__ void __launch_bounds__(128) pv_general(const __grid_constant__ DomainParams P, const __grid_constant__ PvArgs A, double *o) { pv_body<false>(P, A, o); }
__global__ void __launch_bounds__(128) pv_fast(const __grid_constant__ DomainParams P, const __grid_constant__ PvArgs A, double *o) { pv_body<true>(P, A, o); }
