// r6_mlp_tc.cuh — the policy MLP (13 -> 128 -> 64 -> 3, tanh) on the tensor cores, one warp = 32 envs.
//
// This is the only place of the env step that is a contraction (north_star: "tensor cores appear only if
// the PPO policy MLP is fused into closed-loop rollouts").  Shape of the problem: per warp a
// [32 x 16] x [16 x 128] x [128 x 64] x [64 x 8] chain every env-step, i.e. three skinny GEMMs whose
// activations never leave the register file.  That shape is why the warp-level MMA is used here and
// not tcgen05: a tcgen05.mma needs its A operand in shared memory and its accumulator in TMEM, so each
// of the three layers would cost a register -> smem store, a CTA-wide barrier, an MMA commit/wait and
// a TMEM load, for ~5 k MACs per env; with mma.sync the accumulator fragment of one layer IS the A
// fragment of the next (after a free permutation of the weight rows, below) and the warps of a CTA,
// which finish their adaptive integration at different times, never have to meet.
//
// Precision: 3xTF32 error compensation — every operand is split x = hi + lo with hi = round_tf32(x), and
// hi*hi + lo*hi + hi*lo is accumulated in float32, which restores ~float32 accuracy (the dropped lo*lo
// term is 2^-22 relative).  tanh is evaluated in float32 on the exp2 / rcp units.
// Measured against the float32 CUDA-core network (R6_ACT_MLP): |d action| <= 2e-6 (tests/test_gpu_policy.py).
//
// Fragment layouts of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (g = lane / 4, t = lane % 4):
//   A (16 x 8):  a0 = A[g][t]      a1 = A[g+8][t]    a2 = A[g][t+4]   a3 = A[g+8][t+4]
//   B (8 x 8):   b0 = B[t][g]      b1 = B[t+4][g]
//   C (16 x 8):  c0 = C[g][2t]     c1 = C[g][2t+1]   c2 = C[g+8][2t]  c3 = C[g+8][2t+1]
// Chaining: the C tile of hidden units (8j .. 8j+7) is used as the A tile of k-step j of the next layer
// with a0 = c0, a1 = c2, a2 = c1, a3 = c3, i.e. k-slot t <-> unit 8j + 2t and k-slot t + 4 <-> unit
// 8j + 2t + 1; the next layer's B fragments are packed with the same permutation, so it costs nothing.
#pragma once

#include <stdint.h>

#include "r6_core.cuh"

namespace r6 {

// packed weight block in shared memory (floats), fragment order:
//   W0f [16 n-tiles][2 k-steps][32 lanes][2]   b0 [128]   W1f [16 k-steps][8 n-tiles][32][2]   b1 [64]
//   W2f [8 k-steps][32][2]                     b2 [4]
constexpr int kTcOffB0 = 16 * 2 * 64, kTcOffW1 = kTcOffB0 + 128, kTcOffB1 = kTcOffW1 + 16 * 8 * 64;
constexpr int kTcOffW2 = kTcOffB1 + 64, kTcOffB2 = kTcOffW2 + 8 * 64, kMlpTcFloats = kTcOffB2 + 4;

__device__ __forceinline__ float mlp_tc_pack_element(const R6Mlp &m, int idx)
{
    if (idx < kTcOffB0) {
        const int e = idx & 1, lane = (idx >> 1) & 31, ks = (idx >> 6) & 1, j = idx >> 7;
        const int g = lane >> 2, t = lane & 3, in = 8 * ks + t + 4 * e, out = 8 * j + g;
        return in < kMlpIn ? m.w0[out * kMlpIn + in] : 0.0f;
    }
    if (idx < kTcOffW1) return m.b0[idx - kTcOffB0];
    if (idx < kTcOffB1) {
        const int r = idx - kTcOffW1, e = r & 1, lane = (r >> 1) & 31, n = (r >> 6) & 7, j = r >> 9;
        const int g = lane >> 2, t = lane & 3;
        return m.w1[(8 * n + g) * kMlpH0 + 8 * j + 2 * t + e];
    }
    if (idx < kTcOffW2) return m.b1[idx - kTcOffB1];
    if (idx < kTcOffB2) {
        const int r = idx - kTcOffW2, e = r & 1, lane = (r >> 1) & 31, j = r >> 6;
        const int g = lane >> 2, t = lane & 3;
        return g < kMlpRows ? mlp_w2_row(m, g, 8 * j + 2 * t + e) : 0.0f;
    }
    return mlp_b2_row(m, idx - kTcOffB2);
}

struct Split { uint32_t hi, lo; };
// x = hi + lo with hi the nearest TF32 (10-bit mantissa) value, by integer arithmetic: add half an ulp of the
// 13 dropped bits and mask them (3 instructions; sm_100a has no hardware cvt.rna.tf32 and ptxas expands that
// into a ~7-instruction sequence with branches — it was 45 % of this kernel).  Finite inputs only.
__device__ __forceinline__ Split tf32_split(float x)
{
    Split s;
    s.hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    s.lo = __float_as_uint(x - __uint_as_float(s.hi));
    return s;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// d += A B with both operands split (3 MMAs; small terms first)
__device__ __forceinline__ void mma_3x(float (&d)[4], const Split (&a)[4], Split b0, Split b1)
{
    mma_tf32(d, a[0].lo, a[1].lo, a[2].lo, a[3].lo, b0.hi, b1.hi);
    mma_tf32(d, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b0.lo, b1.lo);
    mma_tf32(d, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b0.hi, b1.hi);
}
// float32 tanh with ~1.5e-7 ABSOLUTE error: 1 - 2/(e^{2|x|} + 1) on the exp2 / rcp units, sign restored.
// (The relative error near 0 is irrelevant here: the value feeds a dot product with O(1) weights.)
__device__ __forceinline__ float tanh_f32(float x)
{
    float e, r;
    const float z = fminf(fabsf(x), 40.0f) * 2.8853900817779268f;      // 2|x| log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), x);
}

// Warp-collective.  x[13]: this lane's observation (anything if the lane has no live env).  `scratch` is this
// warp's own staging tile in shared memory: float [16][33] (padded rows: conflict-free transposes).
// Returns this lane's raw outputs: out[0..2] = Gaussian mean (unclipped), out[3] = value.
__device__ __forceinline__ void mlp_forward_tc(const float *__restrict__ W, float *scratch, const float (&x)[kMlpIn],
                                               float (&out)[4])
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // ---- transpose the observations into A fragments through shared memory: scratch[c][env] ----
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 16; c++) scratch[c * 33 + lane] = c < kMlpIn ? x[c] : 0.0f;
    __syncwarp();
    Split ax[2][2][4];      // [m-tile][k-step][a0..a3]
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            const int r0 = 16 * m + g, c0 = 8 * ks + t;
            ax[m][ks][0] = tf32_split(scratch[c0 * 33 + r0]);
            ax[m][ks][1] = tf32_split(scratch[c0 * 33 + r0 + 8]);
            ax[m][ks][2] = tf32_split(scratch[(c0 + 4) * 33 + r0]);
            ax[m][ks][3] = tf32_split(scratch[(c0 + 4) * 33 + r0 + 8]);
        }
    // ---- hidden-1 accumulators (32 x 64 per warp), initialised with the bias ----
    float acc[2][8][4];
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const float2 b = *reinterpret_cast<const float2 *>(W + kTcOffB1 + 8 * n + 2 * t);
#pragma unroll
        for (int m = 0; m < 2; m++) { acc[m][n][0] = b.x; acc[m][n][1] = b.y; acc[m][n][2] = b.x; acc[m][n][3] = b.y; }
    }
    // ---- layer 0 tile by tile, each tile fed straight into layer 1 ----
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        const float2 bias = *reinterpret_cast<const float2 *>(W + kTcOffB0 + 8 * j + 2 * t);
        float d[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++) { d[m][0] = bias.x; d[m][1] = bias.y; d[m][2] = bias.x; d[m][3] = bias.y; }
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            const float2 w = *reinterpret_cast<const float2 *>(W + ((j * 2 + ks) * 32 + lane) * 2);
            const Split b0 = tf32_split(w.x), b1 = tf32_split(w.y);
            mma_3x(d[0], ax[0][ks], b0, b1);
            mma_3x(d[1], ax[1][ks], b0, b1);
        }
        Split ah[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++) {
            ah[m][0] = tf32_split(tanh_f32(d[m][0]));
            ah[m][1] = tf32_split(tanh_f32(d[m][2]));
            ah[m][2] = tf32_split(tanh_f32(d[m][1]));
            ah[m][3] = tf32_split(tanh_f32(d[m][3]));
        }
        const float *w1 = W + kTcOffW1 + (j * 8 * 32 + lane) * 2;
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const float2 w = *reinterpret_cast<const float2 *>(w1 + n * 64);
            const Split b0 = tf32_split(w.x), b1 = tf32_split(w.y);
            mma_3x(acc[0][n], ah[0], b0, b1);
            mma_3x(acc[1][n], ah[1], b0, b1);
        }
    }
    // ---- layer 2: 64 -> 3 (n-tile padded to 8) ----
    float o[2][4];
    {
        const float bx = (2 * t < kMlpRows) ? W[kTcOffB2 + 2 * t] : 0.0f;
        const float by = (2 * t + 1 < kMlpRows) ? W[kTcOffB2 + 2 * t + 1] : 0.0f;
#pragma unroll
        for (int m = 0; m < 2; m++) { o[m][0] = bx; o[m][1] = by; o[m][2] = bx; o[m][3] = by; }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float2 w = *reinterpret_cast<const float2 *>(W + kTcOffW2 + (j * 32 + lane) * 2);
        const Split b0 = tf32_split(w.x), b1 = tf32_split(w.y);
#pragma unroll
        for (int m = 0; m < 2; m++) {
            Split ah[4];
            ah[0] = tf32_split(tanh_f32(acc[m][j][0]));
            ah[1] = tf32_split(tanh_f32(acc[m][j][2]));
            ah[2] = tf32_split(tanh_f32(acc[m][j][1]));
            ah[3] = tf32_split(tanh_f32(acc[m][j][3]));
            mma_3x(o[m], ah, b0, b1);
        }
    }
    // ---- route the 32 x 3 outputs back to the lanes that own the envs: scratch[col][env] ----
    __syncwarp();
    if (t < 2) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            scratch[(2 * t) * 33 + 16 * m + g] = o[m][0];
            scratch[(2 * t) * 33 + 16 * m + g + 8] = o[m][2];
            scratch[(2 * t + 1) * 33 + 16 * m + g] = o[m][1];
            scratch[(2 * t + 1) * 33 + 16 * m + g + 8] = o[m][3];
        }
    }
    __syncwarp();
    out[0] = scratch[0 * 33 + lane];
    out[1] = scratch[1 * 33 + lane];
    out[2] = scratch[2 * 33 + lane];
    out[3] = scratch[3 * 33 + lane];
    __syncwarp();
}
// deterministic action (clipped mean)
__device__ __forceinline__ void mlp_policy_tc(const float *__restrict__ W, float *scratch, const float (&x)[kMlpIn],
                                              float &act0, float &act1, float &act2)
{
    float out[4];
    mlp_forward_tc(W, scratch, x, out);
    act0 = fminf(fmaxf(out[0], -1.0f), 1.0f);
    act1 = fminf(fmaxf(out[1], -1.0f), 1.0f);
    act2 = fminf(fmaxf(out[2], -1.0f), 1.0f);
}

}  // namespace r6
