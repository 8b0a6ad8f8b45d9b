// DFMA issue rate vs. where its operands come from: (A) one fresh register operand + two loop-invariant ones that the
// operand-reuse cache can hold, (B) three distinct register operands per instruction, (C) two registers + a
// constant-bank operand, (D) DMUL/DADD with two distinct registers.  12 and 16 warps per SM, 8 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.0000001, 1.0000002, 1.0000003, 1.0000004, 1.0000005, 1.0000006, 1.0000007, 1.0000008};

template <int MODE>
__global__ void k(int iters, double *sink, long long *cyc, const double *init)
{
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = init[i] + threadIdx.x * 1e-3; b[i] = init[8 + i]; c[i] = init[16 + i]; }
    const double m = init[24], d = init[25];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = fma(a[i], m, d);
            if (MODE == 1) a[i] = fma(b[i], c[(i + 3) & 7], a[i]);
            if (MODE == 2) a[i] = fma(b[i], kc[i], a[i]);
            if (MODE == 3) a[i] = a[i] * b[i];
            if (MODE == 4) a[i] = fma(a[i], b[i], c[(i + 3) & 7]);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + b[i] + c[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(int warps_per_sm, const char *what, const double *init)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<MODE><<<148, 32 * warps_per_sm>>>(iters, sink, cyc, init);
    k<MODE><<<148, 32 * warps_per_sm>>>(iters, sink, cyc, init);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SMSP %d  %-52s : %5.2f cycles per instruction per SMSP\n", (warps_per_sm + 3) / 4, what,
           (double)h / iters / 8 / ((warps_per_sm + 3) / 4));
    cudaFree(sink); cudaFree(cyc);
}
int main()
{
    double hinit[32], *init;
    for (int i = 0; i < 32; i++) hinit[i] = 1.0 + 1e-9 * i;
    cudaMalloc(&init, sizeof hinit); cudaMemcpy(init, hinit, sizeof hinit, cudaMemcpyHostToDevice);
    for (int w : {12, 16}) {
        run<0>(w, "DFMA a = a*m + d (two loop-invariant operands)", init);
        run<1>(w, "DFMA a = b[i]*c[j] + a (three distinct registers)", init);
        run<4>(w, "DFMA a = a*b[i] + c[j] (three distinct registers)", init);
        run<2>(w, "DFMA a = b[i]*const[i] + a (constant-bank operand)", init);
        run<3>(w, "DMUL a = a*b[i]", init);
    }
    return 0;
}
