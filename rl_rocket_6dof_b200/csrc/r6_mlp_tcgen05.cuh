// r6_mlp_tcgen05.cuh — the policy MLP (13 -> 128 -> 64 -> 3, tanh) on the 5th-generation tensor cores:
// tcgen05.mma (kind::tf32, M = 128 envs per CTA, operands in shared memory, accumulators in TMEM), tcgen05.ld for
// the epilogues, one elected thread issuing the MMAs, mbarrier completion.  This is the FAST policy mode of
// r6_policy (tensor_cores = 2): single-pass TF32, |d action| ~1e-3 against the float32 network, for rollouts where
// the policy's own exploration noise dwarfs that; the faithful modes are the FMA network and the 3xTF32 MMA tiles
// (r6_mlp_tc.cuh).  It exists as its own uniform kernel because a tcgen05 MMA is a CTA-wide operation: all 128
// rows of the A tile must be in shared memory and every thread meets at a barrier per layer, which the divergent
// integrator cannot afford but a dedicated policy kernel can.
//
// Per tile of 128 envs (thread t <-> env row t <-> TMEM lane t):
//   X [128x16]  -> smem  | MMA  D0[128x128] = X  W0^T           (2 x K8)
//   D0 cols  0..63  -> +b0, tanh -> Hs [128x64] | MMA D1[128x64]  = Hs W1[:,  0: 64]^T   (8 x K8)
//   D0 cols 64..127 -> +b0, tanh -> Hs          | MMA D1        += Hs W1[:, 64:128]^T   (8 x K8)
//   D1 -> +b1, tanh -> Hs [128x64]              | MMA D2[128x16] = Hs W2^T              (8 x K8)
//   D2 cols 0..2 -> +b2, clip -> actions
// The 128-wide hidden layer is fed to layer 1 in two K-halves so that the A staging tile is 32 KB instead of 64:
// 84 KB of shared memory and 256 TMEM columns per CTA => two CTAs per SM, one running MMAs while the other is in
// an epilogue.
//
// Shared-memory operand layout (both operands K-major, SWIZZLE_NONE "interleaved" canonical layout): element
// (row r, k) of a tile with K_tot columns lives at byte  (r % 8) * 16 + (r / 8) * SBO + (k / 4) * 128 + (k % 4) * 4,
// SBO = (K_tot / 4) * 128: 8-row x 16-byte core matrices, consecutive K chunks 128 B apart (LBO), 8-row groups SBO
// apart.  One MMA consumes K = 8 (two chunks), so k-block kb starts 256 B further.
#pragma once

#include <stdint.h>

#include "r6_core.cuh"

namespace r6 {
namespace tc5 {

constexpr int kTile = 128;                       // envs per CTA tile = MMA M
constexpr uint32_t kTmemCols = 256;              // D0: [0,128)  D1: [128,192)  D2: [192,208)
constexpr uint32_t kColD0 = 0, kColD1 = 128, kColD2 = 192;
// byte offsets inside the dynamic shared memory block
constexpr int kOffW0 = 0;                        // [128][16]   8 KB
constexpr int kOffW1 = kOffW0 + 128 * 16 * 4;    // [64][128]  32 KB
constexpr int kOffW2 = kOffW1 + 64 * 128 * 4;    // [16][64]    4 KB
constexpr int kOffX = kOffW2 + 16 * 64 * 4;      // [128][16]   8 KB
constexpr int kOffH = kOffX + 128 * 16 * 4;      // [128][64]  32 KB
constexpr int kOffBias = kOffH + 128 * 64 * 4;   // b0[128] b1[64] b2[4]
constexpr int kOffBar = kOffBias + (128 + 64 + 4) * 4;
constexpr int kOffTmemPtr = kOffBar + 8;
constexpr int kSmemBytes = kOffTmemPtr + 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int tile_off(int r, int k, int ktot) { return (r & 7) * 16 + (r >> 3) * (ktot * 32) + (k >> 2) * 128 + (k & 3) * 4; }
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units,
// version 1 (Blackwell), SWIZZLE_NONE
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t instr_desc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// tanh on the MUFU unit (tanh.approx.f32, relative error 2^-11): the same precision class as the TF32 operands it
// is rounded to right afterwards, one XU instruction instead of two plus five FP32 ones (the epilogue of this
// kernel is bound by the XU pipe).
__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tanh(D[:, col0 .. col0+63] + bias) of this thread's row -> staging tile Hs [128][64] (TF32-rounded)
__device__ __forceinline__ void epilogue_to_h(uint32_t tmem_row, uint32_t col0, const float *bias, char *Hs, int row)
{
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tmem_ld32(tmem_row + col0 + cc, v);
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
            float4 h;
            h.x = round_tf32(tanh_fast(v[q] + bias[cc + q]));
            h.y = round_tf32(tanh_fast(v[q + 1] + bias[cc + q + 1]));
            h.z = round_tf32(tanh_fast(v[q + 2] + bias[cc + q + 2]));
            h.w = round_tf32(tanh_fast(v[q + 3] + bias[cc + q + 3]));
            *reinterpret_cast<float4 *>(Hs + tile_off(row, cc + q, 64)) = h;
        }
    }
}

// issue nk K-blocks of D[tmem] (+)= A[a_base ..] B[b_base ..]^T ; both tiles advance 256 B per K-block
__device__ __forceinline__ void issue_mmas(uint32_t d_tmem, uint32_t a_base, uint32_t a_sbo, uint32_t b_base, uint32_t b_sbo,
                                           int nk, int n, bool accumulate_first)
{
    const uint32_t idesc = instr_desc(n);
    for (int kb = 0; kb < nk; kb++)
        mma_tf32(d_tmem, smem_desc(a_base + kb * 256, 128, a_sbo), smem_desc(b_base + kb * 256, 128, b_sbo), idesc,
                 accumulate_first || kb > 0);
}

}  // namespace tc5

// ------------------------------------------------------------------------------------------------------------------
// The FAITHFUL policy mode on the 5th-generation tensor cores (r6_policy tensor_cores = 3): 3xTF32 error compensation
// — every operand split x = hi + lo, hi = round_tf32(x), lo = round_tf32(x - hi), and hi*hi + lo*hi + hi*lo chained
// into ONE TMEM accumulator (three tcgen05.mma per K-block, the dropped lo*lo term is 2^-22 relative) — with an
// float32-accurate tanh, so the actions agree with the float32 FMA network to ~1e-6 (tests/test_gpu_policy.py) instead
// of the 1e-3 of the single-pass mode above.
//
// The split operands need twice the shared memory (weights hi + lo: 88 KB), which rules out two CTAs per SM — and one
// CTA of four epilogue warps cannot keep the XU / FP32 pipes busy while its MMAs are in flight.  So ONE CTA per SM
// carries TWO independent 128-env tile groups (warps 0-7 and 8-15; two threads per env row, each taking half of the
// columns of an epilogue) over one resident copy of the weights: each group
// has its own activation tiles, TMEM columns, mbarrier and named barrier, its own elected MMA-issuing thread, and walks
// its own tiles; while one group waits for its MMAs the other is in an epilogue.  The 16-wide input tile aliases the
// first 8 KB of the group's hidden-activation tile (it is dead once layer 0 has been committed and waited for).
//   shared memory: W0h W0l (8+8) | W1h W1l (32+32) | W2h W2l (4+4) | group 0: Hh Hl (32+32) | group 1: Hh Hl | bias
//   = 88 + 128 KB + 1 KB;  TMEM: 512 columns, group g at column 256 g: D0 [0,128) D1 [128,192) D2 [192,208).
namespace tc5x3 {
using namespace tc5;
constexpr int kThreads = 512;           // two tile groups x (128 rows x 2 column halves)
constexpr int kGroupThreads = 256;
constexpr int kOffW0h = 0, kOffW0l = 8 << 10, kOffW1h = 16 << 10, kOffW1l = 48 << 10, kOffW2h = 80 << 10, kOffW2l = 84 << 10;
constexpr int kOffGroup = 88 << 10, kGroupBytes = 64 << 10, kOffHl = 32 << 10;       // per group: Hh at +0, Hl at +32 KB
constexpr int kOffBias3 = kOffGroup + 2 * kGroupBytes;                               // b0[128] b1[64] b2[4]
constexpr int kOffBar3 = kOffBias3 + (128 + 64 + 4) * 4;                             // two mbarriers
constexpr int kOffTmemPtr3 = kOffBar3 + 16;
constexpr int kSmemBytes3 = kOffTmemPtr3 + 16;
constexpr uint32_t kTmemCols3 = 512, kGroupCols = 256;

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = round_tf32(x);
    lo = round_tf32(x - hi);
}
// tanh(x) = 1 - 2 / (exp(2x) + 1) on the exp2 / rcp units: absolute error ~1.5e-7 (the float32 ulp at 1 is 6e-8),
// saturates correctly (exp -> inf => 1, exp -> 0 => -1)
__device__ __forceinline__ float tanh_f32(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));      // exp(2x) = 2^(2x log2 e)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory"); }

// tanh(D[:, col0 + 32 half .. + 31] + bias) of this thread's row, split -> the group's Hh / Hl tiles [128][64]; the two
// threads of a row (warps w and w + 4 reach the same 32 TMEM lanes) take 32 of the 64 columns each
__device__ __forceinline__ void epilogue_to_h3(uint32_t tmem_row, uint32_t col0, const float *bias, char *Hh, char *Hl, int row, int half)
{
    {
        const int cc = 32 * half;
        float v[32];
        tmem_ld32(tmem_row + col0 + cc, v);
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
            float4 h, l;
            split_tf32(tanh_f32(v[q] + bias[cc + q]), h.x, l.x);
            split_tf32(tanh_f32(v[q + 1] + bias[cc + q + 1]), h.y, l.y);
            split_tf32(tanh_f32(v[q + 2] + bias[cc + q + 2]), h.z, l.z);
            split_tf32(tanh_f32(v[q + 3] + bias[cc + q + 3]), h.w, l.w);
            const int off = tile_off(row, cc + q, 64);
            *reinterpret_cast<float4 *>(Hh + off) = h;
            *reinterpret_cast<float4 *>(Hl + off) = l;
        }
    }
}
// nk K-blocks of D (+)= (Ah + Al)(Bh + Bl)^T without the lo*lo term; hi*hi first, then the two corrections
__device__ __forceinline__ void issue_mmas3(uint32_t d_tmem, uint32_t ah, uint32_t al, uint32_t a_sbo, uint32_t bh, uint32_t bl,
                                            uint32_t b_sbo, int nk, int n, bool accumulate_first)
{
    const uint32_t idesc = instr_desc(n);
    for (int kb = 0; kb < nk; kb++) {
        const uint64_t dah = smem_desc(ah + kb * 256, 128, a_sbo), dal = smem_desc(al + kb * 256, 128, a_sbo);
        const uint64_t dbh = smem_desc(bh + kb * 256, 128, b_sbo), dbl = smem_desc(bl + kb * 256, 128, b_sbo);
        mma_tf32(d_tmem, dah, dbh, idesc, accumulate_first || kb > 0);
        mma_tf32(d_tmem, dal, dbh, idesc, true);
        mma_tf32(d_tmem, dah, dbl, idesc, true);
    }
}
}  // namespace tc5x3

// ------------------------------------------------------------------------------------------------------------------
// The same faithful 3xTF32 policy with the ACTIVATIONS IN TENSOR MEMORY (tcgen05.mma with the A operand read from TMEM,
// only the weights come from shared memory).  Why: with both operands in shared memory an M = 128, N = 64, K = 8 TF32
// MMA reads 4 KB of A + 2 KB of B for a 32-cycle tensor-pipe slot — 48 cycles of shared-memory bandwidth, and the
// 3xTF32 chain reads each activation tile twice — so the MMAs of namespace tc5x3 ran at the shared-memory rate, not the
// tensor-pipe floor, and competed with the epilogue's own stores.  Here an epilogue thread owns TMEM lane = env row:
// it reads its 32 accumulator columns (tcgen05.ld), applies bias + tanh, splits, and stores hi IN PLACE over the
// accumulator columns it just read and lo into a 64-column side tile (tcgen05.st) — which is exactly the K-major A
// layout the MMA wants (row = lane, k = column).  No activation ever touches shared memory, no generic->async proxy
// fence is needed, and the per-MMA shared-memory read is B only (2 KB at N = 64: under the floor).
//   TMEM, per group (256 columns):  D0 [0,128)   D1 [128,192)   L [192,256)
//     layer 0 : A = X   hi D1[0:16)  lo L[0:16)            -> D0[0:128)
//     layer 1a: A = H0a hi D0[0:64)  (in place) lo L[0:64) -> D1
//     layer 1b: A = H0b hi D0[64:128)           lo L[0:64) -> D1 +=
//     layer 2 : A = H1  hi D1[0:64)  (in place) lo L[0:64) -> D0[0:16)   (D0 is dead by then)
//   shared memory: the split weights (88 KB) + biases + barriers only.
namespace tc5ts {
using namespace tc5x3;
constexpr uint32_t kColL = 192;
constexpr int kOffBiasT = 88 << 10;
constexpr int kOffBarT = kOffBiasT + (128 + 64 + 4) * 4;
constexpr int kOffTmemPtrT = kOffBarT + 16;
constexpr int kSmemBytesT = kOffTmemPtrT + 16;

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tanh(D[:, src + 32 half .. + 31] + bias) of this thread's row: hi over the columns just read, lo -> L[32 half ..].
// `enter` / `leave` bracket the arithmetic only (the XU / FP32 phase two tile groups take turns in); the TMEM load before
// it and the TMEM stores after it overlap with the other group's turn.
template <class Enter, class Leave>
__device__ __forceinline__ void epilogue_in_tmem(uint32_t tmem_row, uint32_t src, uint32_t lo_dst, const float *bias, int half,
                                                 Enter enter, Leave leave)
{
    const int cc = 32 * half;
    float v[32], l[32];
    tmem_ld32(tmem_row + src + cc, v);
    enter();
#pragma unroll
    for (int q = 0; q < 32; q++) tc5x3::split_tf32(tc5x3::tanh_f32(v[q] + bias[cc + q]), v[q], l[q]);
    leave();
    tmem_st32(tmem_row + src + cc, v);
    tmem_st32(tmem_row + lo_dst + cc, l);
    tmem_st_wait();
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}
// nk K-blocks of D (+)= (Ah + Al)(Bh + Bl)^T without lo*lo; A tiles are TMEM columns (8 per K-block), B in smem.
// Called by a whole CONVERGED warp with warp-uniform arguments: the descriptors are then computed on the uniform
// datapath and each tcgen05.mma is issued by one elected lane straight from uniform registers.  (Issued from a single
// thread under a divergent branch, every MMA was wrapped in an ELECT / 4 x R2UR / branch "waterfall" of ~75 cycles —
// three times the 32-cycle tensor-pipe slot of an N = 64 MMA, so the issue rate, not the tensor pipe, set the pace.)
__device__ __forceinline__ void issue_mmas3_ts(uint32_t d_tmem, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, uint32_t b_sbo,
                                               int nk, int n, bool accumulate_first)
{
    const uint32_t idesc = instr_desc(n);
#pragma unroll
    for (int kb = 0; kb < nk; kb++) {
        const uint64_t dbh = smem_desc(bh + kb * 256, 128, b_sbo), dbl = smem_desc(bl + kb * 256, 128, b_sbo);
        if (elect_one()) {
            mma_tf32_ts(d_tmem, ah + 8 * kb, dbh, idesc, accumulate_first || kb > 0);
            mma_tf32_ts(d_tmem, al + 8 * kb, dbh, idesc, true);
            mma_tf32_ts(d_tmem, ah + 8 * kb, dbl, idesc, true);
        }
    }
}
}  // namespace tc5ts
}  // namespace r6
