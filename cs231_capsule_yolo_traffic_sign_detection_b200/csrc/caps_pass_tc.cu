// caps_pass_tc.cu -- tcgen05 / TMEM version of the pass kernel (sm_100a only).
//
// Same contract as k_pass in caps_kernels.cuh (modes A-uniform / A / L), but the K=8 prediction
// contraction u_hat[b,i,j,:] = u[b,i,:] . W[i,j,:,:] runs on the 5th-gen tensor cores:
//
//   per input capsule i:   D[128 samples x (JW capsules x DD dims)] = A_i[128 x 8] * B_i[JW*DD x 8]^T
//   (DD, JW) = (16, 8), (24, 4), (32, 4), (48, 2): N = 128, 96, 128, 96; other D >= 9 run zero-padded to the next DD
//
// as three kind::tf32 MMAs (3xTF32: lo*hi + hi*lo + hi*hi; measured 1e-7 relative error, i.e.
// fp32-grade, tools/probe_tc.cu), accumulated in TMEM.  Operands are pre-split into tf32 hi/lo
// halves and pre-arranged in the canonical no-swizzle K-major core-matrix layout by the prep
// kernels below, so that one pipeline stage is two cp.async.bulk copies (8 KB of A, 8 KB of B).
//
// Warp roles (384 threads):  warps 0-7 epilogue, warps 8-9 bulk-copy producers, warps 10-11 MMA issuers.
//   producer : waits smem_empty[s], arms smem_full[s] with expect_tx, issues the two bulk copies
//   MMA      : waits smem_full[s] and tmem_empty[t], issues 3 tcgen05.mma, commits to
//              smem_empty[s] (operands consumed) and tmem_full[t] (accumulator ready)
//   epilogue : warp w reads TMEM lanes 32*(w%4).. (= its 32 samples) and the 64 (48) columns of its JW/2
//              capsules with one tcgen05.ld.32x32b.x64 (x32 + x16), releases tmem_empty[t], and does the
//              per-sample part on the FMA pipe: acc += coef * u_hat (A modes) or
//              out = u_hat . X (L mode).  lane <-> sample, exactly like the FFMA kernel.
// Ring depths: `ns` smem stages (default 10 = 160 KB: the bulk copies have ~2000 cycles of latency to
// cover at ~250 cycles per stage), 4 TMEM accumulators (4 x 128 = all 512 columns).
#include "caps_internal.h"
#include "caps_tc_common.cuh"

namespace caps {
namespace {
using namespace tc;

constexpr int kTcMaxStages = 12;      // smem ring depth is a launch parameter (<= 12 x 16 KB)
#ifndef CAPS_TC_ACCUM_LOG2
#define CAPS_TC_ACCUM_LOG2 2
#endif
constexpr int kTcAccumLog2 = CAPS_TC_ACCUM_LOG2;
constexpr int kTcAccum = 1 << kTcAccumLog2;           // TMEM ring (4 x 128 columns); power of two (index = n & 3, phase = (n >> kTcAccumLog2) & 1)
constexpr int kTcN = 128;              // TMEM columns per accumulator (the MMA's N is 128 or 96, see k_pass_tc)
constexpr int kTcABytes = 2 * 2 * 128 * 16;      // [hi/lo][kq][128 rows][16 B] = 8 KB
constexpr int kTcBBytes = 2 * 2 * kTcN * 16;     // 8 KB
constexpr int kTcOperandBytes = kTcABytes + kTcBBytes;       // 16 KB of MMA operands per stage
constexpr int kTcCoefBytes = 4 * 8 * 32 * 4;                  // kModeA: [4 lane tiles][up to 8 capsules][32 lanes] coefficients
__host__ __device__ constexpr int tc_stage_bytes(int mode) { return kTcOperandBytes + (mode == kModeA ? kTcCoefBytes : 0); }
constexpr int kTcEpiWarps = 8;        // epilogue warps: 2 per TMEM lane quarter, half of the CTA's capsules each
constexpr int kTcIssuers = 2;         // producer warps and MMA-issuer warps: each takes every kTcIssuers-th stage, because one
                                      // thread's wait -> issue chain (~250 cycles: mbarrier.try_wait alone is ~100) is longer
                                      // than the 235 cycles of tensor work it feeds
constexpr int kTcThreads = 32 * (kTcEpiWarps + 2 * kTcIssuers);

// kModeA epilogue body as ONE asm block: the probe of the NEXT stage's tmem_full barrier is issued first, the 32 packed
// FMAs (acc += coef * u_hat on register pairs) run under its ~75-cycle round trip, and only then is the predicate read.
// A `selp` right behind `try_wait` stalls the in-order warp (tools/probe_overlap.cu: look-ahead probing written as two
// statements costs +100 cycles per stage; this form gains 7 %).
__device__ __forceinline__ uint32_t fused_fma_probe(uint64_t (&A)[32], const uint64_t (&U)[32], const uint64_t (&Cf)[32],
                                                    uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%97], %98;\n\t"
        "fma.rn.f32x2 %0, %65, %33, %0;\n\t"
        "fma.rn.f32x2 %1, %66, %34, %1;\n\t"
        "fma.rn.f32x2 %2, %67, %35, %2;\n\t"
        "fma.rn.f32x2 %3, %68, %36, %3;\n\t"
        "fma.rn.f32x2 %4, %69, %37, %4;\n\t"
        "fma.rn.f32x2 %5, %70, %38, %5;\n\t"
        "fma.rn.f32x2 %6, %71, %39, %6;\n\t"
        "fma.rn.f32x2 %7, %72, %40, %7;\n\t"
        "fma.rn.f32x2 %8, %73, %41, %8;\n\t"
        "fma.rn.f32x2 %9, %74, %42, %9;\n\t"
        "fma.rn.f32x2 %10, %75, %43, %10;\n\t"
        "fma.rn.f32x2 %11, %76, %44, %11;\n\t"
        "fma.rn.f32x2 %12, %77, %45, %12;\n\t"
        "fma.rn.f32x2 %13, %78, %46, %13;\n\t"
        "fma.rn.f32x2 %14, %79, %47, %14;\n\t"
        "fma.rn.f32x2 %15, %80, %48, %15;\n\t"
        "fma.rn.f32x2 %16, %81, %49, %16;\n\t"
        "fma.rn.f32x2 %17, %82, %50, %17;\n\t"
        "fma.rn.f32x2 %18, %83, %51, %18;\n\t"
        "fma.rn.f32x2 %19, %84, %52, %19;\n\t"
        "fma.rn.f32x2 %20, %85, %53, %20;\n\t"
        "fma.rn.f32x2 %21, %86, %54, %21;\n\t"
        "fma.rn.f32x2 %22, %87, %55, %22;\n\t"
        "fma.rn.f32x2 %23, %88, %56, %23;\n\t"
        "fma.rn.f32x2 %24, %89, %57, %24;\n\t"
        "fma.rn.f32x2 %25, %90, %58, %25;\n\t"
        "fma.rn.f32x2 %26, %91, %59, %26;\n\t"
        "fma.rn.f32x2 %27, %92, %60, %27;\n\t"
        "fma.rn.f32x2 %28, %93, %61, %28;\n\t"
        "fma.rn.f32x2 %29, %94, %62, %29;\n\t"
        "fma.rn.f32x2 %30, %95, %63, %30;\n\t"
        "fma.rn.f32x2 %31, %96, %64, %31;\n\t"
        "selp.u32 %32, 1, 0, p;\n\t}"
        : "+l"(A[0]), "+l"(A[1]), "+l"(A[2]), "+l"(A[3]), "+l"(A[4]), "+l"(A[5]), "+l"(A[6]), "+l"(A[7]), "+l"(A[8]), "+l"(A[9]), "+l"(A[10]), "+l"(A[11]), "+l"(A[12]), "+l"(A[13]), "+l"(A[14]), "+l"(A[15]), "+l"(A[16]), "+l"(A[17]), "+l"(A[18]), "+l"(A[19]), "+l"(A[20]), "+l"(A[21]), "+l"(A[22]), "+l"(A[23]), "+l"(A[24]), "+l"(A[25]), "+l"(A[26]), "+l"(A[27]), "+l"(A[28]), "+l"(A[29]), "+l"(A[30]), "+l"(A[31]), "=r"(ok)
        : "l"(U[0]), "l"(U[1]), "l"(U[2]), "l"(U[3]), "l"(U[4]), "l"(U[5]), "l"(U[6]), "l"(U[7]), "l"(U[8]), "l"(U[9]), "l"(U[10]), "l"(U[11]), "l"(U[12]), "l"(U[13]), "l"(U[14]), "l"(U[15]), "l"(U[16]), "l"(U[17]), "l"(U[18]), "l"(U[19]), "l"(U[20]), "l"(U[21]), "l"(U[22]), "l"(U[23]), "l"(U[24]), "l"(U[25]), "l"(U[26]), "l"(U[27]), "l"(U[28]), "l"(U[29]), "l"(U[30]), "l"(U[31]),
          "l"(Cf[0]), "l"(Cf[1]), "l"(Cf[2]), "l"(Cf[3]), "l"(Cf[4]), "l"(Cf[5]), "l"(Cf[6]), "l"(Cf[7]), "l"(Cf[8]), "l"(Cf[9]), "l"(Cf[10]), "l"(Cf[11]), "l"(Cf[12]), "l"(Cf[13]), "l"(Cf[14]), "l"(Cf[15]), "l"(Cf[16]), "l"(Cf[17]), "l"(Cf[18]), "l"(Cf[19]), "l"(Cf[20]), "l"(Cf[21]), "l"(Cf[22]), "l"(Cf[23]), "l"(Cf[24]), "l"(Cf[25]), "l"(Cf[26]), "l"(Cf[27]), "l"(Cf[28]), "l"(Cf[29]), "l"(Cf[30]), "l"(Cf[31]), "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

// kModeL counterpart: 8 packed partial sums, pair e of the 64 columns goes to partial e / 4 (8 consecutive dims each)
__device__ __forceinline__ uint32_t fused_dot_probe(uint64_t (&Dp)[8], const uint64_t (&U)[32], const uint64_t (&X)[32],
                                                    uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%73], %74;\n\t"
        "fma.rn.f32x2 %0, %9, %41, %0;\n\t"
        "fma.rn.f32x2 %0, %10, %42, %0;\n\t"
        "fma.rn.f32x2 %0, %11, %43, %0;\n\t"
        "fma.rn.f32x2 %0, %12, %44, %0;\n\t"
        "fma.rn.f32x2 %1, %13, %45, %1;\n\t"
        "fma.rn.f32x2 %1, %14, %46, %1;\n\t"
        "fma.rn.f32x2 %1, %15, %47, %1;\n\t"
        "fma.rn.f32x2 %1, %16, %48, %1;\n\t"
        "fma.rn.f32x2 %2, %17, %49, %2;\n\t"
        "fma.rn.f32x2 %2, %18, %50, %2;\n\t"
        "fma.rn.f32x2 %2, %19, %51, %2;\n\t"
        "fma.rn.f32x2 %2, %20, %52, %2;\n\t"
        "fma.rn.f32x2 %3, %21, %53, %3;\n\t"
        "fma.rn.f32x2 %3, %22, %54, %3;\n\t"
        "fma.rn.f32x2 %3, %23, %55, %3;\n\t"
        "fma.rn.f32x2 %3, %24, %56, %3;\n\t"
        "fma.rn.f32x2 %4, %25, %57, %4;\n\t"
        "fma.rn.f32x2 %4, %26, %58, %4;\n\t"
        "fma.rn.f32x2 %4, %27, %59, %4;\n\t"
        "fma.rn.f32x2 %4, %28, %60, %4;\n\t"
        "fma.rn.f32x2 %5, %29, %61, %5;\n\t"
        "fma.rn.f32x2 %5, %30, %62, %5;\n\t"
        "fma.rn.f32x2 %5, %31, %63, %5;\n\t"
        "fma.rn.f32x2 %5, %32, %64, %5;\n\t"
        "fma.rn.f32x2 %6, %33, %65, %6;\n\t"
        "fma.rn.f32x2 %6, %34, %66, %6;\n\t"
        "fma.rn.f32x2 %6, %35, %67, %6;\n\t"
        "fma.rn.f32x2 %6, %36, %68, %6;\n\t"
        "fma.rn.f32x2 %7, %37, %69, %7;\n\t"
        "fma.rn.f32x2 %7, %38, %70, %7;\n\t"
        "fma.rn.f32x2 %7, %39, %71, %7;\n\t"
        "fma.rn.f32x2 %7, %40, %72, %7;\n\t"
        "selp.u32 %8, 1, 0, p;\n\t}"
        : "+l"(Dp[0]), "+l"(Dp[1]), "+l"(Dp[2]), "+l"(Dp[3]), "+l"(Dp[4]), "+l"(Dp[5]), "+l"(Dp[6]), "+l"(Dp[7]), "=r"(ok)
        : "l"(U[0]), "l"(U[1]), "l"(U[2]), "l"(U[3]), "l"(U[4]), "l"(U[5]), "l"(U[6]), "l"(U[7]), "l"(U[8]), "l"(U[9]), "l"(U[10]), "l"(U[11]), "l"(U[12]), "l"(U[13]), "l"(U[14]), "l"(U[15]), "l"(U[16]), "l"(U[17]), "l"(U[18]), "l"(U[19]), "l"(U[20]), "l"(U[21]), "l"(U[22]), "l"(U[23]), "l"(U[24]), "l"(U[25]), "l"(U[26]), "l"(U[27]), "l"(U[28]), "l"(U[29]), "l"(U[30]), "l"(U[31]),
          "l"(X[0]), "l"(X[1]), "l"(X[2]), "l"(X[3]), "l"(X[4]), "l"(X[5]), "l"(X[6]), "l"(X[7]), "l"(X[8]), "l"(X[9]), "l"(X[10]), "l"(X[11]), "l"(X[12]), "l"(X[13]), "l"(X[14]), "l"(X[15]), "l"(X[16]), "l"(X[17]), "l"(X[18]), "l"(X[19]), "l"(X[20]), "l"(X[21]), "l"(X[22]), "l"(X[23]), "l"(X[24]), "l"(X[25]), "l"(X[26]), "l"(X[27]), "l"(X[28]), "l"(X[29]), "l"(X[30]), "l"(X[31]), "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// u [B][N][8] -> ua [ntq][N][hi/lo][kq][128 rows][4]  (row = sample within the 128-sample quad tile)
// and, when ut != nullptr, the plain lane-tile copy ut [nbt][N][kq][32][4] the gradient kernels read (same pass over u)
__global__ void k_prep_u_tc(const float* __restrict__ u, float* __restrict__ ua, float* __restrict__ ut, int B, int N, int ntq, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)ntq * N * 128) return;
    const int r = (int)(idx & 127);
    const long ti = idx >> 7;
    const int i = (int)(ti % N);
    const long tq = ti / N;
    const long b = tq * 128 + r;
    const long bt = tq * 4 + (r >> 5);
    float* dst = ua + (size_t)ti * 2048 + r * 4;
#pragma unroll
    for (int kq = 0; kq < 2; ++kq) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
        if (b < B) x = ldg4(u + ((size_t)b * N + i) * 8 + kq * 4);
        split4(x, hi, lo);
        st4(dst + (0 * 2 + kq) * 512, hi);
        st4(dst + (1 * 2 + kq) * 512, lo);
        if (ut != nullptr && bt < nbt) st4(ut + ((((size_t)bt * N + i) * 2 + kq) * kLanes + (r & 31)) * 4, x);
    }
}

// W [N][C][8][D] -> wb [N][JG][hi/lo][kq][128 rows][4]  (D = 16, 24, 32 or 48 -- the PADDED dimension, W being the padded
// copy when the public D is smaller; tc_jw(D) = 8, 4, 4, 2 capsules per group;
// row n = D*(j - tc_jw*jg) + d, rows >= tc_jw*D are zero; 4 = k % 4)
__global__ void k_prep_w_tc(const float* __restrict__ W, float* __restrict__ wb, int N, int C, int JG, int D) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * JG * 128) return;
    const int n = (int)(idx & 127);
    const long ig = idx >> 7;
    const int jg = (int)(ig % JG);
    const long i = ig / JG;
    const int JW = D == 16 ? 8 : D == 48 ? 2 : 4;
    const int j = n < JW * D ? jg * JW + n / D : C, d = n % D;
    float* dst = wb + (size_t)ig * 2048 + n * 4;
#pragma unroll
    for (int kq = 0; kq < 2; ++kq) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
        if (j < C) {
            const float* src = W + (((size_t)i * C + j) * 8 + kq * 4) * D + d;
            x = make_float4(__ldg(src), __ldg(src + D), __ldg(src + 2 * D), __ldg(src + 3 * D));
        }
        split4(x, hi, lo);
        st4(dst + (0 * 2 + kq) * 512, hi);
        st4(dst + (1 * 2 + kq) * 512, lo);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct PassTcParams {
    const float* ua;     // [ntq][N][2][2][128][4]
    const float* wb;     // [N][JG][2][2][128][4]
    const float* coef;   // kModeA: [nbt][N][C][32]
    const float* X;      // kModeL: [nbt][C][4][32][4]
    float* out;          // kModeL: [nbt][N][C][32];  kModeA*: part [IS][nbt][C][4][32][4]
    int N, C, JG, nbt, i_per_split, ns;
    int dbg;             // TIMING EXPERIMENTS ONLY (tuning knob "tcdbg"): 1 = skip the L-mode stores, 2 = skip the coefficient copies
};

// DD = (padded) class-capsule dimension: 16 (8 capsules per CTA, 4 per epilogue warp), 24 or 32 (4 per CTA, 2 per warp),
// 48 (2 per CTA, 1 per warp).  The MMA is M = 128, N = JW * DD (128, 96, 128, 96); accumulators stay 128 TMEM columns apart.
template <int MODE, int DD>
__global__ void __launch_bounds__(kTcThreads, 1) k_pass_tc(PassTcParams p) {
    constexpr int JW = DD == 16 ? 8 : DD == 48 ? 2 : 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int ns = p.ns;
    constexpr int kTcStageBytes = tc_stage_bytes(MODE);
    // shared-window addresses: stages [ns x 16 (20) KB], then the barriers (8 bytes each)
    const uint32_t stages = smem_u32(smem_raw);
    const uint32_t bars = stages + (uint32_t)ns * kTcStageBytes;
    const uint32_t smem_full = bars;                                   // [kTcMaxStages]
    const uint32_t smem_empty = bars + 8 * kTcMaxStages;               // [kTcMaxStages]
    const uint32_t tmem_full = bars + 16 * kTcMaxStages;               // [kTcAccum]
    const uint32_t tmem_empty = tmem_full + 8 * kTcAccum;              // [kTcAccum]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ns * kTcStageBytes + 16 * kTcMaxStages + 16 * kTcAccum);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jg = blockIdx.y, tq = blockIdx.z;
    const int i_begin = blockIdx.x * p.i_per_split;
    const int i_end = min(p.N, i_begin + p.i_per_split);
    const int n_i = max(i_end - i_begin, 0);

    if (threadIdx.x == 0) {
        // kModeA: the 8 epilogue warps read their coefficients out of the stage, so they release it too
        for (int s = 0; s < ns; ++s) { mbar_init(smem_full + 8 * s, 1); mbar_init(smem_empty + 8 * s, MODE == kModeA ? kTcEpiWarps + 1 : 1); }
        for (int t = 0; t < kTcAccum; ++t) { mbar_init(tmem_full + 8 * t, 1); mbar_init(tmem_empty + 8 * t, kTcEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcEpiWarps + kTcIssuers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp >= kTcEpiWarps && warp < kTcEpiWarps + kTcIssuers) {
        // ===== producer: converged warp, one elected lane arms the barrier and issues both copies =====
        // producer w takes stages n = w, w + NI, w + 2 NI, ...
        constexpr int NI = kTcIssuers;
        const int w = warp - kTcEpiWarps;
        if (w >= NI) goto role_done;
        const float* asrc = p.ua + ((size_t)tq * p.N + i_begin + w) * 2048;
        const float* bsrc = p.wb + ((size_t)(i_begin + w) * p.JG + jg) * 2048;
        const size_t bstep = (size_t)p.JG * 2048 * NI;
        int s = w;
        uint32_t ph = 1;                                 // parity to wait for on smem_empty (first lap passes)
        // kModeA: + one copy per valid lane tile of the quad: coef[tile][i][8 jg .. +nj][32] (nj * 128 bytes)
        const int nj = min(JW, p.C - jg * JW);
        const uint32_t cbytes = (uint32_t)nj * 128u;
        const size_t ctile = (size_t)p.N * p.C * kLanes;                 // coef elements per lane tile
        const float* csrc = MODE == kModeA ? p.coef + ((size_t)(tq * 4) * p.N + i_begin + w) * p.C * kLanes + (size_t)jg * JW * kLanes : nullptr;
        const size_t cstep = (size_t)p.C * kLanes * NI;
        const int nvt = (p.dbg & 2) ? 0 : min(4, p.nbt - tq * 4);        // valid lane tiles in this quad (>= 1)
        const uint32_t txbytes = (uint32_t)kTcOperandBytes + (MODE == kModeA ? (uint32_t)nvt * cbytes : 0u);
        for (int n = w; n < n_i; n += NI) {
            mbar_wait(smem_empty + 8 * s, ph);
            const uint32_t dst = stages + (uint32_t)s * kTcStageBytes, bar = smem_full + 8 * s;
            if (MODE == kModeA) {
                asm volatile(
                    "{\n\t"
                    ".reg .pred pe, p1, p2, p3;\n\t"
                    ".reg .b32 cd;\n\t"
                    "mov.u32 cd, %8;\n\t"
                    "elect.sync _|pe, 0xffffffff;\n\t"
                    "setp.gt.and.s32 p1, %14, 1, pe;\n\t"
                    "setp.gt.and.s32 p2, %14, 2, pe;\n\t"
                    "setp.gt.and.s32 p3, %14, 3, pe;\n\t"
                    "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
                    "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %4, [%0];\n\t"
                    "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%5], [%6], %7, [%0];\n\t"
                    "setp.gt.and.s32 p1, %14, 0, pe;\n\t"
                    "@p1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [cd], [%10], %9, [%0];\n\t"
                    "setp.gt.and.s32 p1, %14, 1, pe;\n\t"
                    "add.u32 cd, cd, 1024;\n\t"
                    "@p1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [cd], [%11], %9, [%0];\n\t"
                    "add.u32 cd, cd, 1024;\n\t"
                    "@p2 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [cd], [%12], %9, [%0];\n\t"
                    "add.u32 cd, cd, 1024;\n\t"
                    "@p3 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [cd], [%13], %9, [%0];\n\t"
                    "}"
                    ::"r"(bar), "r"(txbytes), "r"(dst), "l"(asrc), "r"((uint32_t)kTcABytes),
                      "r"(dst + kTcABytes), "l"(bsrc), "r"((uint32_t)kTcBBytes),
                      "r"(dst + kTcOperandBytes), "r"(cbytes), "l"(csrc), "l"(csrc + ctile), "l"(csrc + 2 * ctile), "l"(csrc + 3 * ctile),
                      "r"(nvt)
                    : "memory");
                csrc += cstep;
            } else {
                asm volatile(
                    "{\n\t"
                    ".reg .pred pe;\n\t"
                    "elect.sync _|pe, 0xffffffff;\n\t"
                    "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
                    "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %4, [%0];\n\t"
                    "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%5], [%6], %7, [%0];\n\t"
                    "}"
                    ::"r"(bar), "r"(txbytes), "r"(dst), "l"(asrc), "r"((uint32_t)kTcABytes),
                      "r"(dst + kTcABytes), "l"(bsrc), "r"((uint32_t)kTcBBytes)
                    : "memory");
            }
            asrc += 2048 * NI;
            bsrc += bstep;
            s += NI;
            if (s >= ns) { s -= ns; ph ^= 1; }
        }
    } else if (warp >= kTcEpiWarps + kTcIssuers) {
        // ===== MMA issuer: the whole warp runs the loop converged, one elected lane issues =====
        // kind::tf32, D = f32, A/B K-major, N = 128, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((JW * DD) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // descriptors differ between stages only in the 14-bit start-address field (bytes >> 4)
        const uint64_t desc0 = umma_desc(stages, 2048, 128);
        // issuer w takes stages n = w, w + NI, ...: stages are independent (own smem slot, own accumulator), and a
        // tcgen05.commit only tracks the MMAs of the thread that executes it.  Uniform mode: one dependent accumulation
        // chain per issuer, each in its own accumulator (a single issuer's wait -> issue chain took 312 cycles per stage).
        constexpr int NI = kTcIssuers;
        const int w = warp - kTcEpiWarps - kTcIssuers;
        if (w >= NI) goto role_done;
        int s = w;
        uint32_t sph = 0;
        for (int n = w; n < n_i; n += NI) {
            // uniform mode: issuer w accumulates ITS stages (n = w, w + 2, ...) in accumulator w; the epilogue adds the two
            const int t = (MODE == kModeAUniform) ? w : (n & (kTcAccum - 1));
            if (MODE != kModeAUniform) mbar_wait(tmem_empty + 8 * t, ((n >> kTcAccumLog2) & 1) ^ 1);
            mbar_wait(smem_full + 8 * s, sph);
            tc_fence_after();
            const uint64_t a_hi = desc0 + (uint64_t)((s * kTcStageBytes) >> 4);
            const uint64_t a_lo = a_hi + (4096 >> 4), b_hi = a_hi + (kTcABytes >> 4), b_lo = b_hi + (4096 >> 4);
            if (MODE == kModeAUniform)
                // uniform couplings: sum_i u_hat_i IS one long GEMM over (i,k): accumulate in TMEM across all the
                // stages, signal the epilogue once
                umma_stage(tmem_base + (uint32_t)(t * kTcN), a_hi, a_lo, b_hi, b_lo, idesc, smem_empty + 8 * s, tmem_full + 8 * t,
                           n >= NI, n + NI >= n_i);
            else
                umma_stage(tmem_base + (uint32_t)(t * kTcN), a_hi, a_lo, b_hi, b_lo, idesc, smem_empty + 8 * s, tmem_full + 8 * t);
            s += NI;
            if (s >= ns) { s -= ns; sph ^= 1; }
        }
    } else {
        // ===== epilogue: warp w -> samples of lane tile 4*tq + w%4, capsules j0 .. j0+JPW-1 =====
        constexpr int JPW = JW / (kTcEpiWarps / 4), NC = DD * JPW, D4 = DD / 4;
        static_assert(NC == 64 || NC == 48, "64 (48 for D = 24) accumulator columns per epilogue warp");
        const int q = warp & 3, jh = warp >> 2;
        const int tile = tq * 4 + q;
        const bool tvalid = tile < p.nbt;
        const int j0 = jg * JW + jh * JPW;
        float acc[JPW][DD];                 // A modes: running sums; L mode: the probe vectors X[b,j,:]
#pragma unroll
        for (int jj = 0; jj < JPW; ++jj)
#pragma unroll
            for (int d = 0; d < DD; ++d) acc[jj][d] = 0.f;
        if (MODE == kModeL && tvalid) {
#pragma unroll
            for (int jj = 0; jj < JPW; ++jj)
                if (j0 + jj < p.C) {
#pragma unroll
                    for (int dq = 0; dq < D4; ++dq) {
                        const float4 x = ldg4(p.X + ((((size_t)tile * p.C + j0 + jj) * D4 + dq) * kLanes + lane) * 4);
                        acc[jj][dq * 4 + 0] = x.x; acc[jj][dq * 4 + 1] = x.y; acc[jj][dq * 4 + 2] = x.z; acc[jj][dq * 4 + 3] = x.w;
                    }
                }
        }
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jh * NC);
        // kModeA: this warp's coefficients inside a stage: [q: 1 KB per lane tile][jh*JPW + jj][lane]
        const uint32_t coef_off = (uint32_t)kTcOperandBytes + (uint32_t)((q * 8 + jh * JPW) * kLanes + lane) * 4u;
        int s = 0;
        uint32_t sph = 0;
        if (MODE == kModeAUniform) {
            // one accumulator per issuer warp (each a dependent chain of its own stages): summed here, in fixed order
#pragma unroll
            for (int w = 0; w < kTcIssuers; ++w)
                if (n_i > w) {
                    mbar_wait(tmem_full + 8 * w, 0);
                    tc_fence_after();
                    float uh[NC];
                    tmem_ld<NC>(lane_base + (uint32_t)(w * kTcN), uh);
#pragma unroll
                    for (int jj = 0; jj < JPW; ++jj)
#pragma unroll
                        for (int d = 0; d < DD; ++d) acc[jj][d] += uh[jj * DD + d];
                }
        } else {
        bool ready = false;                  // kModeA: tmem_full of this stage was already seen complete by the previous one
        for (int n = 0; n < n_i; ++n) {
            // Every epilogue warp handles every stage, so the stage rate is bounded by this loop's serial chain of fixed
            // latencies (tools/probe_overlap.cu: a try_wait round trip is ~75 cycles even on a completed mbarrier, LDS
            // ~70, LDTM ~45; two warps per scheduler run their FMA bursts in phase).  Hence: ONE wait per stage --
            // tmem_full(n) is signalled by a commit behind MMAs whose issuer had already observed smem_full(n), so stage
            // n's bulk copies (the kModeA coefficients) have landed by then -- and the coefficient LDS rides under the LDTM.
            const int t = n & (kTcAccum - 1);
            const int i = i_begin + n;
            float cc[JPW];
#pragma unroll
            for (int jj = 0; jj < JPW; ++jj) cc[jj] = 0.f;
            if (!ready) mbar_wait(tmem_full + 8 * t, (n >> kTcAccumLog2) & 1);
            ready = false;
            if (MODE == kModeA && tvalid) {
                const uint32_t ca = stages + (uint32_t)s * kTcStageBytes + coef_off;
#pragma unroll
                for (int jj = 0; jj < JPW; ++jj)
                    if (j0 + jj < p.C) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cc[jj]) : "r"(ca + jj * 128) : "memory");
            }
            tc_fence_after();
            float uh[NC];
            tmem_ld<NC>(lane_base + (uint32_t)(t * kTcN), uh);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * t);  // accumulator t may be overwritten
            if (MODE == kModeL && NC == 64) {
                uint64_t U2[32], X2[32], D2[8];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int jj = (2 * e) / DD, d = (2 * e) % DD;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(U2[e]) : "f"(uh[2 * e]), "f"(uh[2 * e + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(X2[e]) : "f"(acc[jj][d]), "f"(acc[jj][d + 1]));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) D2[k] = 0ull;
                const int n1 = n + 1 < n_i ? n + 1 : n;
                ready = fused_dot_probe(D2, U2, X2, tmem_full + 8 * (n1 & (kTcAccum - 1)), (n1 >> kTcAccumLog2) & 1) != 0 && n + 1 < n_i;
                float part[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float lo, hi;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(D2[k]));
                    part[k] = lo + hi;
                }
                constexpr int GPC = 8 / JPW;                      // partial sums per capsule (2 or 4)
#pragma unroll
                for (int jj = 0; jj < JPW; ++jj) {
                    float dot = part[jj * GPC];
#pragma unroll
                    for (int k = 1; k < GPC; ++k) dot += part[jj * GPC + k];
                    if (tvalid && j0 + jj < p.C && !(p.dbg & 1)) p.out[(((size_t)tile * p.N + i) * p.C + j0 + jj) * kLanes + lane] = dot;
                }
            } else if (MODE == kModeL) {
#pragma unroll
                for (int jj = 0; jj < JPW; ++jj) {
                    // four partial sums in two FFMA2 chains (half the FMA instructions, chains of DD/4 instead of DD)
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int d = 0; d < DD; d += 4) {
                        ffma2(d0, d1, uh[jj * DD + d], uh[jj * DD + d + 1], acc[jj][d], acc[jj][d + 1]);
                        ffma2(d2, d3, uh[jj * DD + d + 2], uh[jj * DD + d + 3], acc[jj][d + 2], acc[jj][d + 3]);
                    }
                    const float dot = (d0 + d1) + (d2 + d3);
                    if (tvalid && j0 + jj < p.C && !(p.dbg & 1)) p.out[(((size_t)tile * p.N + i) * p.C + j0 + jj) * kLanes + lane] = dot;
                }
            } else if (NC == 64 && MODE == kModeA) {
                uint64_t A2[32], U2[32], C2[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int jj = (2 * e) / DD, d = (2 * e) % DD;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(A2[e]) : "f"(acc[jj][d]), "f"(acc[jj][d + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(U2[e]) : "f"(uh[2 * e]), "f"(uh[2 * e + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(C2[e]) : "f"(cc[jj]), "f"(cc[jj]));
                }
                const int n1 = n + 1 < n_i ? n + 1 : n;          // last stage: a harmless probe of the own (complete) barrier
                ready = fused_fma_probe(A2, U2, C2, tmem_full + 8 * (n1 & (kTcAccum - 1)), (n1 >> kTcAccumLog2) & 1) != 0 &&
                        n + 1 < n_i;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int jj = (2 * e) / DD, d = (2 * e) % DD;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[jj][d]), "=f"(acc[jj][d + 1]) : "l"(A2[e]));
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < JPW; ++jj) {
                    const float f = cc[jj];
#pragma unroll
                    for (int d = 0; d < DD; d += 2) ffma2(acc[jj][d], acc[jj][d + 1], f, f, uh[jj * DD + d], uh[jj * DD + d + 1]);
                }
            }
            if (MODE == kModeA) {
                // the FMAs above consumed cc, so the shared loads are complete: release our share of stage s
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_empty + 8 * s);
                if (++s == ns) { s = 0; sph ^= 1; }
            }
        }
        }
        if (MODE != kModeL && tvalid) {
#pragma unroll
            for (int jj = 0; jj < JPW; ++jj)
                if (j0 + jj < p.C) {
#pragma unroll
                    for (int dq = 0; dq < D4; ++dq)
                        st4(p.out + (((((size_t)blockIdx.x * p.nbt + tile) * p.C + j0 + jj) * D4 + dq) * kLanes + lane) * 4,
                            make_float4(acc[jj][dq * 4 + 0], acc[jj][dq * 4 + 1], acc[jj][dq * 4 + 2], acc[jj][dq * 4 + 3]));
                }
        }
    }

role_done:
    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + kTcIssuers) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace

int g_tc_stages = 10;
int g_tc_dbg = 0;

size_t tc_ua_floats(int B, int N) { return (size_t)cdiv(B > 0 ? B : 1, 128) * N * 2048; }
int tc_jw(int D) { return D == 16 ? 8 : D == 48 ? 2 : 4; }     // capsules per tcgen05 CTA (padded D = 16, 24, 32, 48)
size_t tc_wb_floats(int N, int C, int D) { return (size_t)N * cdiv(C, tc_jw(D)) * 2048; }

int launch_prep_u_tc(const Plan& pl, const float* u, float* ua, float* ut, cudaStream_t st) {
    const int ntq = cdiv(pl.B, 128);
    const long nu = (long)ntq * pl.N * 128;
    k_prep_u_tc<<<cdiv(nu, 256), 256, 0, st>>>(u, ua, ut, pl.B, pl.N, ntq, pl.nbt);
    LAUNCH_CHECK();
    return 0;
}

int launch_prep_w_tc(const Plan& pl, const float* W, float* wb, cudaStream_t st) {
    const int JG = cdiv(pl.C, tc_jw(pl.DP));              // W: the padded copy ([N][C][8][DP]) when D < DP
    const long nw = (long)pl.N * JG * 128;
    k_prep_w_tc<<<cdiv(nw, 256), 256, 0, st>>>(W, wb, pl.N, pl.C, JG, pl.DP);
    LAUNCH_CHECK();
    return 0;
}

int launch_pass_tc(const Plan& pl, int mode, const PassParams& pp, const float* ua, const float* wb, cudaStream_t st) {
    PassTcParams tp{};
    tp.ua = ua; tp.wb = wb; tp.coef = pp.coef; tp.X = pp.X; tp.out = pp.out;
    tp.N = pl.N; tp.C = pl.C; tp.JG = cdiv(pl.C, tc_jw(pl.DP)); tp.nbt = pl.nbt; tp.i_per_split = pl.i_per_split;
    tp.dbg = g_tc_dbg;
    tp.ns = g_tc_stages < 2 ? 2 : g_tc_stages > kTcMaxStages ? kTcMaxStages : g_tc_stages;
    while ((size_t)tp.ns * tc_stage_bytes(mode) + 512 > 227 * 1024) --tp.ns;      // 227 KB of dynamic smem per CTA
    const size_t smem = (size_t)tp.ns * tc_stage_bytes(mode) + 512;
    dim3 grid(pl.IS, tp.JG, cdiv(pl.nbt, 4)), block(kTcThreads);
#define CAPS_LAUNCH_TC_D(MODE, DD)                                                                       \
    {                                                                                                    \
        auto kern = k_pass_tc<MODE, DD>;                                                                 \
        CAPS_SET_SMEM(kern, smem);      /* per instantiation and per device */                           \
        kern<<<grid, block, smem, st>>>(tp);                                                             \
    }
#define CAPS_LAUNCH_TC(MODE)                                                                             \
    { if (pl.DP == 48) CAPS_LAUNCH_TC_D(MODE, 48) else if (pl.DP == 32) CAPS_LAUNCH_TC_D(MODE, 32)       \
      else if (pl.DP == 24) CAPS_LAUNCH_TC_D(MODE, 24) else CAPS_LAUNCH_TC_D(MODE, 16) }
    if (pl.DP != 16 && pl.DP != 24 && pl.DP != 32 && pl.DP != 48) return fail(CAPS_E_UNSUPPORTED, "tcgen05 pass kernel: padded D must be 16, 24, 32 or 48");
    if (mode == kModeAUniform) CAPS_LAUNCH_TC(kModeAUniform)
    else if (mode == kModeA) CAPS_LAUNCH_TC(kModeA)
    else CAPS_LAUNCH_TC(kModeL)
#undef CAPS_LAUNCH_TC
#undef CAPS_LAUNCH_TC_D
    LAUNCH_CHECK();
    return 0;
}

}  // namespace caps
