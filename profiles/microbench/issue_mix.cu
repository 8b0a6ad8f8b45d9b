// Does a DFMA (half-rate on sm_100a: 16 FP64 lanes per SM sub-partition) hold the warp scheduler's issue port for one
// cycle or for two?  Per loop iteration every thread issues ND independent DFMAs and NX independent FFMA / IMAD / LDS;
// if the port is free in a DFMA's second cycle, ND DFMAs + ND other instructions still take 2*ND cycles per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NF, int NI, int NL>
__global__ void k(int iters, double *sink, long long *cyc)
{
    __shared__ double sm[8 * 256];
    double a[8];
    float f[8];
    int n[8];
    double l[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-2f + i; n[i] = threadIdx.x + i; l[i] = 0; sm[i * 256 + threadIdx.x] = i; }
    const double m = 1.0000001, c = 1e-9;
    const float mf = 1.0001f, cf = 1e-5f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < ND) a[i] = fma(a[i], m, c);
            if (i < NF) f[i] = fmaf(f[i], mf, cf);
            if (i < NI) n[i] = n[i] * 3 + it;
            if (i < NL) l[i] += sm[i * 256 + ((threadIdx.x + it) & 255)];
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + f[i] + n[i] + l[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ND, int NF, int NI, int NL>
void run(int warps_per_sm)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<ND, NF, NI, NL><<<148, 32 * warps_per_sm>>>(iters, sink, cyc);
    k<ND, NF, NI, NL><<<148, 32 * warps_per_sm>>>(iters, sink, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter_smsp = (double)h / iters;      // cycles per loop iteration as seen by one warp
    const int wps = (warps_per_sm + 3) / 4;
    printf("warps/SMSP %d  DFMA %d FFMA %d IMAD %d LDS %d : %7.2f cycles per iteration per warp = %6.2f per SMSP-iteration  (DFMA pipe floor %d, one-port floor %d, two-cycle-port floor %d)\n",
           wps, ND, NF, NI, NL, per_iter_smsp, per_iter_smsp / wps, 2 * ND, ND + NF + NI + NL, 2 * ND + NF + NI + NL);
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    for (int w : {12, 16}) {
        run<8, 0, 0, 0>(w);
        run<8, 8, 0, 0>(w);
        run<8, 0, 8, 0>(w);
        run<8, 4, 4, 0>(w);
        run<8, 8, 8, 0>(w);
        run<4, 8, 8, 0>(w);
        run<8, 0, 0, 4>(w);
        run<8, 4, 0, 4>(w);
        run<0, 8, 8, 0>(w);
    }
    return 0;
}
