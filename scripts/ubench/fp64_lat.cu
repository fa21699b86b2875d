// FP64 pipe micro-benchmark for B200: dependent-chain latency and throughput vs warps/SMSP and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = a + k + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fma(x[k], b, a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / (iters * 8.0);
}
template <int ILP> void run(int warps_per_sm, double *d)
{
    // one CTA per SM, warps_per_sm warps
    int iters = 2000;
    chain<ILP><<<148, 32 * warps_per_sm>>>(d, 1.0, 0.999999, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chain<ILP><<<148, 32 * warps_per_sm>>>(d, 1.0, 0.999999, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double inst = 148.0 * warps_per_sm * iters * 8.0 * ILP;           // warp-instructions
    printf("warps/SM %2d ILP %d : %.2f cycles per dependent step (ILP instr); warp-instr/clk/SM %.3f  (%.2f ms)\n", warps_per_sm, ILP, cyc,
           ILP * warps_per_sm / cyc, ms);
}
int main()
{
    double *d; cudaMalloc(&d, 148 * 1024 * 8);
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); }
    return 0;
}
