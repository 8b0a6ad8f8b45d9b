// Issue rate of the XU-pipe instructions the env step uses (per SM sub-partition, cycles per warp instruction):
// MUFU.RCP64H / MUFU.RSQ64H (the FP64 reciprocal / rsqrt seeds), MUFU.RCP / RSQ / EX2 / LG2 (float), F2F.F64.F32,
// F2F.F32.F64, I2F.F64, and — for comparison — DFMA.   8 independent chains, 12 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rates xu_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int iters, double *sink, long long *cyc, const double *init)
{
    double a[8]; float f[8];
    for (int i = 0; i < 8; i++) { a[i] = init[i] + threadIdx.x * 1e-3; f[i] = (float)a[i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r; }
            if (MODE == 1) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r; }
            if (MODE == 2) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[i])); f[i] = r; }
            if (MODE == 3) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[i])); f[i] = r; }
            if (MODE == 4) { double r; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(r) : "f"(f[i])); a[i] = r; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(r)); }
            if (MODE == 5) a[i] = fma(a[i], 1.0000001, 1e-9);
            if (MODE == 6) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[i])); f[i] = r; }
            if (MODE == 7) { int v = __float2int_rn(f[i]); f[i] = __int2float_rn(v + it); }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + f[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(int warps_per_sm, const char *what, int per_iter, const double *init)
{
    double *sink; long long *cyc, h;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    k<MODE><<<148, 32 * warps_per_sm>>>(iters, sink, cyc, init);
    k<MODE><<<148, 32 * warps_per_sm>>>(iters, sink, cyc, init);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SMSP %d  %-44s : %6.2f cycles per warp instruction per SMSP\n", (warps_per_sm + 3) / 4, what,
           (double)h / iters / per_iter / ((warps_per_sm + 3) / 4));
    cudaFree(sink); cudaFree(cyc);
}
int main()
{
    double hinit[32], *init;
    for (int i = 0; i < 32; i++) hinit[i] = 1.0 + 0.01 * i;
    cudaMalloc(&init, sizeof hinit); cudaMemcpy(init, hinit, sizeof hinit, cudaMemcpyHostToDevice);
    for (int w : {4, 12}) {
        run<0>(w, "MUFU.RCP64H (rcp.approx.ftz.f64)", 8, init);
        run<1>(w, "MUFU.RSQ64H (rsqrt.approx.ftz.f64)", 8, init);
        run<2>(w, "MUFU.RCP (rcp.approx.ftz.f32)", 8, init);
        run<3>(w, "MUFU.EX2", 8, init);
        run<6>(w, "MUFU.SQRT (sqrt.approx.ftz.f32)", 8, init);
        run<4>(w, "F2F.F64.F32 + F2F.F32.F64 (pair)", 16, init);
        run<7>(w, "F2I + I2F (pair, 32-bit)", 16, init);
        run<5>(w, "DFMA", 8, init);
    }
    return 0;
}
