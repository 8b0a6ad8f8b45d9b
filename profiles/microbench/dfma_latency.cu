// DFMA dependent-issue latency and throughput vs. independent chains / resident warps on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(int iters, double *sink, long long *cyc)
{
    double a[CH];
    const double m = 1.0000001, c = 1e-9;
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = (threadIdx.x + i) * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
void run(int warps_per_sm)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<CH><<<148, 32 * warps_per_sm>>>(iters, sink, cyc);
    k<CH><<<148, 32 * warps_per_sm>>>(iters, sink, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / iters / CH;      // cycles per DFMA per warp
    double per_smsp = per / ((warps_per_sm + 3) / 4);  // cycles per warp-DFMA per SMSP
    printf("chains=%d warps/SM=%2d  cycles per DFMA (one warp) = %6.2f   per SMSP issue interval = %5.2f\n", CH,
           warps_per_sm, per, per_smsp);
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
