// caps_grad_mma.cu -- final backward pass with the two K=8 / D=16 contractions on tensor cores.
//
// Same contract as k_grad (caps_kernels.cuh):
//     G_bij = sum_m alpha^m_bij X^m_bj          (fp32 FMA pipe, lane <-> sample, X in registers)
//     dW_ij[k,d] = sum_b u_bi[k] G_bij[d]       (reduction over the batch)
//     du_bi[k]   = sum_j sum_d W_ij[k,d] G_bij[d]
// but the two contractions run as warp-level mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (lo*hi + hi*lo + hi*hi: fp32-grade accuracy), instead of 256 FFMA + a 124-shuffle transpose
// reduction per (i, j, 32 samples).  The batch reduction of dW happens INSIDE the MMA (the MMA's
// K dimension is the sample index), so there is no cross-lane reduction left at all.
//
// Why legacy mma.sync and not tcgen05 here: one side of both products is only 8 wide (k), and a
// tcgen05.mma costs ~62 cycles whatever its N (tools/probe_tc.cu), so 8/16-column UMMAs would
// waste the pipe; the G operand is produced per sample in registers, which is exactly the
// register-fragment model of mma.sync.  (The u_hat passes, whose N is 128, use tcgen05.)
//
// One warp <-> one output capsule j; the CTA owns dW[i-tile][8 capsules] completely (it loops over
// every 32-sample lane tile), so dW needs no atomics and is bit-reproducible.  Per (i, tile):
//   1. every lane builds G[16] for its sample and parks G and u in a per-warp shared tile
//   2. dW: 4 chunks of 8 samples:  C[d 16 x k 8] += A[d x b] * B[b x k]   (A = G^T, B = u)
//   3. du: 2 x 16 samples, 2 K-steps: C[b 16 x k 8] += A[b x d] * B[d x k] (A = G, B = W^T, whose
//      fragments are pre-split and pre-permuted in shared memory once per CTA)
//   4. du fragments are summed across the 8 warps (capsules) through shared memory; one partial
//      per j-group goes to global (k_reduce_du adds the j-groups in fixed order).
#include "caps_internal.h"

namespace caps {
namespace {

constexpr int kGmIT = 8;          // input capsules per CTA
constexpr int kGmDuBufs = 3;      // du round buffers: the consumers may run two rounds ahead of the service warp's reduction (2 buffers: 7.66 ms,
                                  // 3: 7.47, 4: 7.47 at the benchmark shape; the operand ring gets what is left: 7 stages, and 3 would do)
constexpr int kGmServiceWarps = 1;    // 1: one service warp fills the ring and sums the du fragments; 2: two warps -- measured SLOWER (9.3 vs 7.5 ms): ptxas budgets registers for 16 warps then (128 instead of 154)
constexpr bool g_use_ffma2 = true;   // G build as 8 packed FMAs per term (scalar-broadcast coefficient) instead of 16 FFMA
// Consumer warps (= output capsules) per CTA: 8 or 11.  11 + the producer warp = 384 threads is the most that still
// leaves 168 registers per thread (X^m alone takes 16 M); three warps per scheduler instead of two hide more of the
// LDS -> split -> HMMA chain, and C = 43 is 4 x 11 - 1.
__host__ __device__ constexpr int gm_dub(int) { return 2; }          // input capsules per du reduction round (kGmDuBufs round buffers)
__host__ __device__ constexpr int gm_stage_floats(int M, int JW) { return 272 + (M - 1) * JW * 32; }
__host__ __device__ constexpr int gm_fixed_floats(int JW) {
    return kGmIT * JW * 128 /* Wfrag */ + kGmIT * JW * 128 /* dWsm */ + JW * 32 * 16 /* Gs */ + kGmDuBufs * JW * gm_dub(JW) * 256 /* dusm */;
}
// operand ring depth (units in flight per CTA): what fits next to the fixed tiles, at most 8
__host__ __device__ constexpr int gm_stages(int M, int JW) {
    int ns = (227 * 1024 - 256 - 64 - gm_fixed_floats(JW) * 4) / (gm_stage_floats(M, JW) * 4 + 16);
    return ns > 8 ? 8 : ns;
}
// G tile of a warp: [32 samples][16 dims], no padding, with the 4-float column groups XOR-swizzled by sample bits 1
// and 2:  float (b, d) lives at b * 16 + (d ^ gm_sigma(b)).  That makes all three access patterns bank-conflict free:
// the 16-byte row stores (lane <-> b), the dW fragment reads (lane (g,t) <-> b = t, d = g) and the du fragment reads
// (ldmatrix rows b = g, 16-byte chunks of d).  No linear row stride can do both fragment patterns at once.
__device__ __forceinline__ int gm_sigma(int b) { return 8 * ((b >> 1) & 1) + 4 * ((b >> 2) & 1); }
constexpr int kGmUSkew = 144;     // floats between the two k-halves of the staged u tile (128 + 16: halves 16 banks apart)

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 3xTF32 split.  The tensor core reads only the top 19 bits of an operand (it truncates), so the
// "hi" half of x is x itself; lo = x - trunc(x) is exact in fp32 and costs one LOP3 + one FADD
// (cvt.rna.tf32 is a 5-instruction sequence in SASS).  Dropped terms (lo*lo and the truncation of
// lo) are ~2^-20 relative.
__device__ __forceinline__ uint32_t tf32_lo(float x) {
    return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}
// two fp32 FMAs in one instruction (Blackwell FFMA2): d = a * b + d on a register pair
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    uint64_t a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}
// the same split on a register pair: two LOP3 + one packed subtract (sub.f32x2) instead of two LOP3 + two FADD
__device__ __forceinline__ void tf32_lo2(uint32_t x0, uint32_t x1, uint32_t& l0, uint32_t& l1) {
    uint64_t x, t, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(x0), "r"(x1));
    asm("and.b64 %0, %1, 0xffffe000ffffe000;" : "=l"(t) : "l"(x));
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(t));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(l0), "=r"(l1) : "l"(d));
}
// c = a * b (no accumulator to clear first)
__device__ __forceinline__ void mma_tf32_m16n8k8_zero(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                      uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
// c (+)= a * b with a (4 regs) and b (2 regs) given in fp32; 3xTF32.  ZERO: c is written, not accumulated into.
template <bool ZERO>
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    uint32_t al[4], bl[2];
    tf32_lo2(a[0], a[1], al[0], al[1]);
    tf32_lo2(a[2], a[3], al[2], al[3]);
    tf32_lo2(b[0], b[1], bl[0], bl[1]);
    if (ZERO) mma_tf32_m16n8k8_zero(c, al[0], al[1], al[2], al[3], b[0], b[1]);
    else mma_tf32_m16n8k8(c, al[0], al[1], al[2], al[3], b[0], b[1]);
    mma_tf32_m16n8k8(c, a[0], a[1], a[2], a[3], bl[0], bl[1]);
    mma_tf32_m16n8k8(c, a[0], a[1], a[2], a[3], b[0], b[1]);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time: wrong for an event loop)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();      // ~4 s: a protocol bug, not a long wait
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}

// grid = (ceil(N/8), ceil(C/8)); block = 288 (8 consumer warps + 1 producer warp).  K = 8, D = 16 only.
//
// The per-(b,i) operands -- the u tile and the 2R-2 coefficient rows of the CTA's 8 capsules -- are
// streamed into a shared-memory ring by cp.async.bulk, gm_stages(M) units ahead.  With plain (even
// register-prefetched) loads the kernel ran at exactly one loaded-HBM latency (~1750 cycles) per unit
// whatever the math did (11 ms with ALL the math removed): 8 warps with one unit of loads in flight each
// is far too little memory-level parallelism.
// GEN = false: D == 16 exactly (the hot shapes): every dimension below is a compile-time constant.
template <int M, int JW, bool GEN>
__global__ void __launch_bounds__(32 * JW + 32 * kGmServiceWarps, 1) k_grad_mma(GradParams p) {
    const int pD = GEN ? p.D : 16, pDP = GEN ? p.DP : 16;
    constexpr int IT = kGmIT, DUB = gm_dub(JW), NT = 32 * JW;     // NT: consumer threads
    constexpr int NS = gm_stages(M, JW), SF = gm_stage_floats(M, JW);    // stage: u tile + up to M-1 coefficient rows
    static_assert(NS >= 3, "operand ring too shallow");
    extern __shared__ __align__(16) float smem[];
    float* Wfrag = smem;                                   // [IT][JW][ks 2][32 lanes][2]
    float* dWsm = Wfrag + IT * JW * 128;                   // [IT][JW][32 lanes][4]
    float* Gs = dWsm + IT * JW * 128;                      // [JW][32][16] swizzled
    float* dusm = Gs + JW * 32 * 16;                       // [buf 2][JW][DUB][half 2][32 lanes][4]
    float* ring = dusm + kGmDuBufs * JW * DUB * 256;                   // [NS][ u tile [kq 2][32][4], halves kGmUSkew apart | coef rows [M-1][JW][32] ]
    const uint32_t bars = smem_u32(ring + NS * SF);        // full[NS], empty[NS]
    const uint32_t bar_full = bars, bar_empty = bars + 8 * NS;
    const uint32_t bar_dufull = bars + 16 * NS, bar_dufree = bar_dufull + 8 * kGmDuBufs;      // [kGmDuBufs] each: du round buffers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;                 // mma fragment coordinates
    const int i0 = blockIdx.x * IT;
    const int ni = min(IT, p.N - i0);
    // D > 16: every capsule is two / three 16-dim "pseudo-capsules" (j, h): same coefficients, X / W / dW columns
    // 16 h .. 16 h + 15.  dW of the parts is independent and du is a sum over (j, d) anyway, so a warp simply owns one.
    // Columns at or beyond the padded dimension DP (D = 24, 21: the second part has 8 real columns) read as zero;
    // dW is only written for d < D.  W and the X arrays have rows of DP floats, dW (public) of D.
    const int DH = GEN ? (pDP + 15) >> 4 : 1;              // 1, 2, 3: pseudo-capsule index = DH j + h
    const int jp0 = blockIdx.y * JW;
    const int j = (jp0 + warp) / DH, h = (jp0 + warp) - j * DH;
    const bool jvalid = jp0 + warp < p.C * DH;
    const int njv = min(JW, p.C * DH - jp0);               // valid pseudo-capsules (warps) of this CTA
    const int jfirst = jp0 / DH;                           // the CTA's capsules jfirst .. jfirst + njr - 1: coefficient rows
    const int njr = min(p.C, (jp0 + JW - 1) / DH + 1) - jfirst;
    const int D4 = pDP >> 2;

    // W^T fragments for du:  B[d][k] = W[k][d];  b0 = (d = t + 8 ks, k = g), b1 = (d = t + 4 + 8 ks, k = g)
    for (int e = threadIdx.x; e < IT * JW * 64; e += blockDim.x) {
        const int l = e & 31, ks = (e >> 5) & 1, w = (e >> 6) % JW, il = e / (64 * JW);
        const int gg = l >> 2, tt = l & 3;
        float w0 = 0.f, w1 = 0.f;
        if (il < ni && w < njv) {
            const int jw = (jp0 + w) / DH, hw = (jp0 + w) - jw * DH;
            const float* row = p.W + (((size_t)(i0 + il) * p.C + jw) * 8 + gg) * pDP + hw * 16;   // W[i][j][k = gg][16 h ..]
            if (hw * 16 + tt + 8 * ks < pDP) w0 = __ldg(row + tt + 8 * ks);
            if (hw * 16 + tt + 4 + 8 * ks < pDP) w1 = __ldg(row + tt + 4 + 8 * ks);
        }
        // kept whole: the tensor core truncates its operands, so the "hi" half is w itself and lo = tf32_lo(w) at use
        *reinterpret_cast<float2*>(Wfrag + (size_t)((il * JW + w) * 2 + ks) * 64 + l * 2) = make_float2(w0, w1);
    }
    for (int e = threadIdx.x; e < IT * JW * 128; e += blockDim.x) dWsm[e] = 0.f;
    if (threadIdx.x == 0) {
        for (int q = 0; q < NS; ++q) { mbar_init(bar_full + 8 * q, 1); mbar_init(bar_empty + 8 * q, JW); }
        for (int q = 0; q < kGmDuBufs; ++q) { mbar_init(bar_dufull + 8 * q, JW); mbar_init(bar_dufree + 8 * q, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= JW) {
        // ===== service warps: (1) warp JW: one elected lane streams (tile, il) stages into the ring, NS ahead;
        // (2) warp JW + 1: sums the consumer warps' du fragments of each finished round and writes the CTA's partial --
        // so the consumers never meet at a CTA barrier.  With kGmServiceWarps == 1 one warp does both in an event
        // loop over the two non-blocking waits (the consumers then wait ~7 % of their time for a free du buffer).
        const bool do_fill = warp == JW, do_reduce = warp == JW + kGmServiceWarps - 1;
        const uint32_t ubytes = 2 * 32 * 4 * 4, cbytes = (uint32_t)njr * 128u;
        int ncoef = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) ncoef += p.coef[m] != nullptr;
        const uint32_t txbytes = ubytes + (uint32_t)ncoef * cbytes;
        int q = 0, ftile = do_fill ? 0 : p.nbt, fil = 0;    // next stage to fill
        uint32_t ph = 1;
        const int rounds = p.nbt * (IT / DUB);
        int rr = do_reduce ? 0 : rounds;                    // next round to reduce
        const int gg = lane >> 2, tt = lane & 3;
        const int kq = (2 * tt) >> 2, kk = (2 * tt) & 3;
        while (ftile < p.nbt || rr < rounds) {
            if (ftile < p.nbt && mbar_test(bar_empty + 8 * q, ph)) {
                if (lane == 0) {
                    const size_t ti = (size_t)ftile * p.N + i0 + fil;
                    const uint32_t dst = smem_u32(ring + q * SF), bar = bar_full + 8 * q;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(txbytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(p.ut + ti * 256), "r"(ubytes / 2), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst + kGmUSkew * 4), "l"(p.ut + ti * 256 + 128), "r"(ubytes / 2), "r"(bar) : "memory");
                    uint32_t cd = dst + 272 * 4;
#pragma unroll
                    for (int m = 0; m < M; ++m)
                        if (p.coef[m] != nullptr) {
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(cd), "l"(p.coef[m] + (ti * p.CS + jfirst) * kLanes), "r"(cbytes), "r"(bar) : "memory");
                            cd += JW * 128;
                        }
                }
                __syncwarp();
                if (++q == NS) { q = 0; ph ^= 1; }
                if (++fil == ni) { fil = 0; ++ftile; }
                continue;
            }
            if (rr < rounds && mbar_test(bar_dufull + 8 * (rr % kGmDuBufs), (rr / kGmDuBufs) & 1)) {
                const int tile = rr / (IT / DUB), ib = (rr % (IT / DUB)) * DUB;
                const float* dbuf = dusm + (size_t)(rr % kGmDuBufs) * JW * DUB * 256;
#pragma unroll
                for (int e = 0; e < DUB * 2; ++e) {         // lane <-> fragment lane; e = (ii, half)
                    const int half = e & 1, ii = e >> 1, il = ib + ii;
                    if (il < ni) {
                        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int w = 0; w < JW; ++w) {
                            const float4 x = *reinterpret_cast<const float4*>(dbuf + (size_t)(((w * DUB + ii) * 2 + half) * 32 + lane) * 4);
                            sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                        }
                        // fragment (mt = half): c0 (b = 16mt+gg, k = 2tt), c1 (.., k = 2tt+1), c2 (b + 8, k = 2tt), c3 (b + 8, 2tt+1)
                        const int b0 = 16 * half + gg;
                        float* dst = p.du_part + ((((size_t)blockIdx.y * p.nbt + tile) * p.N + i0 + il) * 2 + kq) * kLanes * 4;
                        *reinterpret_cast<float2*>(dst + b0 * 4 + kk) = make_float2(sum.x, sum.y);
                        *reinterpret_cast<float2*>(dst + (b0 + 8) * 4 + kk) = make_float2(sum.z, sum.w);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dufree + 8 * (rr % kGmDuBufs));
                ++rr;
                continue;
            }
            __nanosleep(32);
        }
        return;
    }

    float* Gw = Gs + warp * 32 * 16;
    // lane-invariant swizzled offsets into the G tile (floats)
    const int st_sw = gm_sigma(lane) >> 2;                                   // row store: group dq goes to dq ^ st_sw
    const int dwa0 = t * 16 + (g ^ gm_sigma(t));                             // dW A-fragment: (b = t, d = g); d + 8 is ^ 8
    const int dwa2 = (t + 4) * 16 + (g ^ gm_sigma(t + 4));                   //                (b = t + 4, d = g)
    // du A-fragment by ldmatrix.x4: this lane supplies the address of row (lane & 7) of matrix (lane >> 3):
    // matrices 0/1 = samples +0 / +8 of dims 8 ks + 0..3, matrices 2/3 = the same samples, dims 8 ks + 4..7
    const int lm_row = 8 * ((lane >> 3) & 1) + (lane & 7);
    const uint32_t lm_a0 = smem_u32(Gw + lm_row * 16 + (((lane >> 4) * 4) ^ gm_sigma(lm_row)));     // ks = 0
    const uint32_t lm_a1 = smem_u32(Gw + lm_row * 16 + (((lane >> 4) * 4 + 8) ^ gm_sigma(lm_row))); // ks = 1
    int sq = 0;                                             // ring position of the current unit
    int rnd = 0;                                            // du round counter (buffer = rnd & 1)
    uint32_t sph = 0;
    // X^m of (tile, j): 16 floats per term and lane; term 0 arrives pre-multiplied by its constant coupling 1/C
    // (caps_route_backward has k_dsquash do it), so nothing depends on the loads until the next unit.  The loads
    // for tile + 1 are issued in the middle of the tile's last unit (right after its G is built, when xr is dead)
    // so that their L2 latency hides under that unit's MMA phase instead of stalling all 8 warps once per tile.
    float xr[M][16];
    auto load_x = [&](int tile) {
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int dq = 0; dq < 4; ++dq) {
                if (GEN && h * 4 + dq >= D4) {              // beyond the padded dimension (warp-uniform)
                    xr[m][dq * 4 + 0] = 0.f; xr[m][dq * 4 + 1] = 0.f; xr[m][dq * 4 + 2] = 0.f; xr[m][dq * 4 + 3] = 0.f;
                    continue;
                }
                // volatile + "memory": must stay below the warp barrier it follows, or ptxas hoists it above the G
                // FMAs, needs a second register set for X and spills
                asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(xr[m][dq * 4 + 0]), "=f"(xr[m][dq * 4 + 1]), "=f"(xr[m][dq * 4 + 2]), "=f"(xr[m][dq * 4 + 3])
                             : "l"(p.X[m] + ((((size_t)tile * p.C + j) * D4 + h * 4 + dq) * kLanes + lane) * 4) : "memory");
            }
    };
    if (jvalid) load_x(0);
    for (int tile = 0; tile < p.nbt; ++tile) {
        for (int ib = 0; ib < IT; ib += DUB, ++rnd) {
            // this round's du buffer must have been drained by the service warp (kGmDuBufs rounds ago)
            mbar_wait(bar_dufree + 8 * (rnd % kGmDuBufs), ((rnd / kGmDuBufs) & 1) ^ 1);
            float* dbuf = dusm + (size_t)(rnd % kGmDuBufs) * JW * DUB * 256;
#pragma unroll 1
            for (int ii = 0; ii < DUB; ++ii) {
                const int il = ib + ii;
                float du0[4] = {0.f, 0.f, 0.f, 0.f}, du1[4] = {0.f, 0.f, 0.f, 0.f};       // du fragments of the two 16-sample halves
                const float* stg = ring + sq * SF;          // this unit's stage: u tile, then coefficient rows
                if (il < ni) mbar_wait(bar_full + 8 * sq, sph);
                if (jvalid && il < ni) {
                    // term 0 has the constant coupling 1/C (already folded into X^0 by its producer); terms 1..M-1 are
                    // the staged coefficient rows, in order (caps_route_backward builds the list that way)
                    float G[16];
#pragma unroll
                    for (int d = 0; d < 16; ++d) G[d] = xr[0][d];
                    const float* crow = stg + 272 + (j - jfirst) * 32 + lane;
#pragma unroll
                    for (int m = 1; m < M; ++m) {
                        const float al = crow[(m - 1) * JW * 32];
                        if (g_use_ffma2) {
#pragma unroll
                            for (int d = 0; d < 16; d += 2) ffma2(G[d], G[d + 1], al, al, xr[m][d], xr[m][d + 1]);
                        } else {
#pragma unroll
                            for (int d = 0; d < 16; ++d) G[d] = fmaf(al, xr[m][d], G[d]);
                        }
                    }
                    __syncwarp();                                   // previous unit's fragment reads are done
#pragma unroll
                    for (int dq = 0; dq < 4; ++dq)
                        st4(Gw + lane * 16 + ((dq ^ st_sw) * 4), make_float4(G[dq * 4], G[dq * 4 + 1], G[dq * 4 + 2], G[dq * 4 + 3]));
                    __syncwarp();
                    if (il == ni - 1 && tile + 1 < p.nbt) load_x(tile + 1);
                    // ---- dW[d][k] += sum_b G[b][d] u[b][k] : 4 chunks of 8 samples, two accumulators
                    float c0[4], c1[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t* gc = reinterpret_cast<const uint32_t*>(Gw + c * 128);
                        const uint32_t a[4] = {gc[dwa0], gc[dwa0 ^ 8], gc[dwa2], gc[dwa2 ^ 8]};
                        // u[b][k] straight out of the stage: (k >> 2) * kGmUSkew + b * 4 + (k & 3)
                        const uint32_t* ub = reinterpret_cast<const uint32_t*>(stg + (g >> 2) * kGmUSkew + (8 * c + t) * 4 + (g & 3));
                        const uint32_t b[2] = {ub[0], ub[16]};
                        if (c == 0) mma3<true>(c0, a, b);
                        else if (c == 1) mma3<true>(c1, a, b);
                        else if (c & 1) mma3<false>(c1, a, b);
                        else mma3<false>(c0, a, b);
                    }
                    float* dw = dWsm + (size_t)((il * JW + warp) * 32 + lane) * 4;     // lane-private: no hazard
                    float4 acc = *reinterpret_cast<float4*>(dw);
                    acc.x += c0[0] + c1[0]; acc.y += c0[1] + c1[1]; acc.z += c0[2] + c1[2]; acc.w += c0[3] + c1[3];
                    *reinterpret_cast<float4*>(dw) = acc;
                    // ---- du[b][k] = sum_d G[b][d] W[k][d] : 2 x 16 samples, 2 K-steps of 8 dims
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint2 bw = *reinterpret_cast<const uint2*>(Wfrag + (size_t)((il * JW + warp) * 2 + ks) * 64 + lane * 2);
                        const uint32_t b[2] = {bw.x, bw.y};
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            uint32_t a[4];
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                                         : "r"((ks ? lm_a1 : lm_a0) + mt * 1024) : "memory");
                            if (ks == 0) { if (mt) mma3<true>(du1, a, b); else mma3<true>(du0, a, b); }
                            else { if (mt) mma3<false>(du1, a, b); else mma3<false>(du0, a, b); }
                        }
                    }
                }
                if (il < ni) {                              // every consumer warp releases every stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8 * sq);
                    if (++sq == NS) { sq = 0; sph ^= 1; }
                }
                float* ds = dbuf + (size_t)((warp * DUB + ii) * 2) * 128 + lane * 4;   // [warp][ii][half][lane][4]
                st4(ds, make_float4(du0[0], du0[1], du0[2], du0[3]));
                st4(ds + 128, make_float4(du1[0], du1[1], du1[2], du1[3]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_dufull + 8 * (rnd % kGmDuBufs));        // release: this warp's fragments are visible
        }
    }

    // dW fragments -> public [N][C][8][16]:  c0 (d = g, k = 2t), c1 (d = g, k = 2t+1), c2 (d = g+8, k = 2t), c3 (d = g+8, k = 2t+1)
    // (each warp writes the fragments it accumulated itself: no barrier needed)
    if (jvalid)
        for (int il = 0; il < ni; ++il) {
            const float4 v = *reinterpret_cast<const float4*>(dWsm + (size_t)((il * JW + warp) * 32 + lane) * 4);
            float* dst = p.dW + ((size_t)(i0 + il) * p.C + j) * 8 * pD + h * 16;
            if (h * 16 + g < pD) {
                dst[(2 * t) * pD + g] = v.x;
                dst[(2 * t + 1) * pD + g] = v.y;
            }
            if (h * 16 + g + 8 < pD) {
                dst[(2 * t) * pD + g + 8] = v.z;
                dst[(2 * t + 1) * pD + g + 8] = v.w;
            }
        }
}

template <int M, int JW>
int launch_t(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    const size_t smem = ((size_t)gm_fixed_floats(JW) + (size_t)gm_stages(M, JW) * gm_stage_floats(M, JW)) * sizeof(float) +
                        16 * gm_stages(M, JW) + 16 * kGmDuBufs + 16;
    auto kern = pl.D == 16 ? k_grad_mma<M, JW, false> : k_grad_mma<M, JW, true>;
    if (pl.D == 16) CAPS_SET_SMEM(kern, smem); else CAPS_SET_SMEM(kern, smem);      // one cache per instantiation (and per device)
    dim3 grid(cdiv(pl.N, kGmIT), cdiv(pl.C * cdiv(pl.DP, 16), JW)), block(32 * JW + 32 * kGmServiceWarps);
    kern<<<grid, block, smem, st>>>(gp);
    LAUNCH_CHECK();
    return 0;
}

template <int JW>
int launch_m(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    switch (pl.M) {
        case 1: return launch_t<1, JW>(pl, gp, st);
        case 3: return launch_t<3, JW>(pl, gp, st);
        case 5: return launch_t<5, JW>(pl, gp, st);
        case 7: return launch_t<7, JW>(pl, gp, st);
        case 9: return launch_t<9, JW>(pl, gp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "R=%d unsupported", pl.R);
}

}  // namespace

int g_grad_jw = 0;   // tuning knob "gradjw": 0 = auto, 8 or 11

// output capsules per CTA of the mma gradient kernel.  11 never needs more j-groups than 8, and with 12 warps the
// service warp shares a scheduler with two consumers instead of three (JW = 8 measured 11.1 ms against 7.8 ms at
// C = 43); 8 stays selectable for experiments and as a second summation grouping in the tests.
int grad_mma_jw(const Plan&) { return g_grad_jw == 8 ? 8 : 11; }

// du partials the kernel writes (one per CTA row)
int grad_mma_parts(const Plan& pl) { return cdiv(pl.C * cdiv(pl.DP, 16), grad_mma_jw(pl)); }

// 9 <= D <= 48, K == 8, R <= 5.  Writes grad_mma_parts(pl) du partials.
int launch_grad_mma(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    return grad_mma_jw(pl) == 11 ? launch_m<11>(pl, gp, st) : launch_m<8>(pl, gp, st);
}

}  // namespace caps
