// Relative error of the FP64 MUFU seeds (rcp.approx.ftz.f64, rsqrt.approx.ftz.f64) and of the refinement schemes
// built on them in r6_core.cuh.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_seed mufu_seed.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rcp_seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ double rsq_seed(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ double rcp_cubic(double x) { double r = rcp_seed(x); double e = fma(-x, r, 1.0); return fma(r, fma(e, e, e), r); }
__device__ double rcp_newton2(double x) { double r = rcp_seed(x); double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); return fma(r, e, r); }
__device__ double sqrt_gs2(double x)
{
    double r = rsq_seed(x + 1e-300);
    double s = x * r, h = 0.5 * r;
    double e = fma(-s, h, 0.5);
    s = fma(s, e, s); h = fma(h, e, h);
    e = fma(-s, h, 0.5);
    return fma(s, e, s);
}
__global__ void k(double *out, int n)
{
    // out[0..5]: max rel err of rcp seed, rsqrt seed, rcp cubic, rcp newton2, sqrt gs2, (unused)
    __shared__ double m[6][256];
    double e[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // log-uniform over [1e-6, 1e12] plus a fine sweep of one binade
        double u = (i + 0.5) / n;
        double x = (i & 1) ? exp(log(1e-6) + u * (log(1e12) - log(1e-6))) : 1.0 + u;
        double ex = 1.0 / x, sx = sqrt(x), rx = 1.0 / sx;
        e[0] = fmax(e[0], fabs(rcp_seed(x) - ex) / ex);
        e[1] = fmax(e[1], fabs(rsq_seed(x) - rx) / rx);
        e[2] = fmax(e[2], fabs(rcp_cubic(x) - ex) / ex);
        e[3] = fmax(e[3], fabs(rcp_newton2(x) - ex) / ex);
        e[4] = fmax(e[4], fabs(sqrt_gs2(x) - sx) / sx);
    }
    for (int j = 0; j < 6; j++) m[j][threadIdx.x] = e[j];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int j = 0; j < 6; j++) {
            double v = 0;
            for (int t = 0; t < 256; t++) v = fmax(v, m[j][t]);
            atomicMax((unsigned long long *)&out[j], __double_as_longlong(v));   // positive doubles order as integers
        }
}
int main()
{
    double *d, h[6];
    cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
    k<<<592, 256>>>(d, 1 << 26);
    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("max relative error: rcp seed %.3e (2^%.1f)  rsqrt seed %.3e (2^%.1f)  rcp cubic %.3e  rcp newton2 %.3e  sqrt goldschmidt2 %.3e  (eps = 1.11e-16)\n",
           h[0], log2(h[0]), h[1], log2(h[1]), h[2], h[3], h[4]);
    printf("sqrt(0) = %g\n", 0.0);
    return 0;
}
