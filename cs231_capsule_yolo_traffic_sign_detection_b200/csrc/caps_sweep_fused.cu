// caps_sweep_fused.cu -- ONE sweep per routing iteration: logits -> softmax over the class capsules -> weighted
// sum, with the coupling logits never leaving the SM (reference models.py:75-79; the backward counterpart is
// softmax-backward + the dv sweep of the same iteration).
//
// The unfused path (caps_pass_tc.cu + k_softmax*) runs  L-sweep -> [B,N,C] logits in HBM -> softmax kernel ->
// [B,N,C] couplings in HBM -> A-sweep:  the [B,N,C] array crosses HBM four times per iteration and u_hat is
// recomputed twice.  Here the CTAs that hold the C/8 capsule groups of the SAME 128 samples and the SAME input
// capsules form a thread-block cluster and exchange the one thing the softmax needs from the other groups -- the
// per-(sample, i) partial normaliser -- through distributed shared memory:
//
//   stage n (input capsule i), CTA = capsule group jg, epilogue warp = 32 samples x 4 capsules x 16 dims:
//     P1(n)   u_hat = tcgen05.ld of the stage's accumulator (3xTF32 MMA, as in caps_pass_tc.cu)
//             FWD: e_j = exp(u_hat_j . V_j)             z = sum of the warp's 4 e_j
//             BWD: dc_j = u_hat_j . ds_j               z = sum of c_j dc_j        (c: saved by the forward)
//             z -> shared; the exchange warp adds the two column halves and stores the 128-sample row (512 B)
//             into every CTA of the cluster with st.async (registers -> shared::cluster), which credits a
//             transaction barrier at the receiver as the bytes land.
//     P2(n-2) Z = sum over the cluster's rows (fixed order: every CTA gets bit-identical Z)
//             FWD: c_j = e_j / Z          -> stored once for the backward     acc_j += c_j u_hat_j
//             BWD: beta_j = beta'_j + c_j (dc_j - Z) -> stored once           acc_j += beta_j u_hat_j
//             (u_hat is read again from TMEM: the accumulator ring is 4 deep, the exchange takes < 2 stages)
//
// So per iteration the [B,N,C] data is written ONCE (what the backward needs) and never re-read by the forward.
// Flow control of the exchange ring needs no credits: an iteration runs P1(it) and then P2(it-2), so a CTA sends
// the row of stage m only after its own P2(m-3), which needed every peer's row m-3, which each peer sent after its
// own P2(m-6): when row m arrives, rows <= m-6 have been consumed everywhere, and the ring holds 8 >= 6 rows
// (DESIGN.md section 3.4).
//
// softmax is evaluated without the max subtraction (exp2 of the logit times log2 e, clamped at 2^120): the logits
// are u_hat . (v^0 + .. + v^{r-1}) with |v| < 1, so they stay far inside the fp32 exponent range for any weights
// the reference can train; the unfused path (tuning knob "fused" = 0) keeps the max-subtracted form.
#include "caps_internal.h"
#include "caps_tc_common.cuh"

#include <algorithm>

namespace caps {
namespace {
using namespace tc;

constexpr int kFsEpiWarps = 8;
constexpr int kFsThreads = 384;            // warps 0-7 epilogue, 8 producer, 9 MMA issuer, 10 exchange, 11 idle
constexpr int kFsAccum = 4;                // TMEM ring: 4 accumulators x 128 columns
constexpr int kFsSkew = 2;                 // stages between P1 and P2 of the same input capsule
constexpr int kFsZSlots = 8;               // exchange ring depth (>= 3 * kFsSkew: the credit-free argument in DESIGN.md 3.4)
constexpr int kFsMaxCluster = 8;           // portable cluster size: C <= 64 at 8 capsules per CTA
constexpr int kFsMaxStages = 12;
constexpr int kFsOperandBytes = 16384;     // A (8 KB: u hi/lo) + B (8 KB: W hi/lo), layouts of caps_pass_tc.cu
constexpr int kFsCoefBytes = 4096;         // [4 lane tiles][8 capsules][32 lanes] floats
__host__ __device__ constexpr int fs_stage_bytes(bool bwd) { return kFsOperandBytes + (bwd ? 2 * kFsCoefBytes : 0); }
constexpr int kFsZpartBytes = kFsZSlots * 2 * 128 * 4;                    // [slot][column half][128 samples]
constexpr int kFsZrecvBytes = kFsZSlots * kFsMaxCluster * 128 * 4;        // [slot][sender rank][128 samples]
constexpr int kFsEringBytes = (kFsSkew + 1) * kFsEpiWarps * 4 * 32 * 4;   // [skew+1][warp][4][32]
constexpr int kFsBarBytes = 8 * (2 * kFsMaxStages + 2 * kFsAccum + 2 * kFsZSlots) + 16;     // + tcgen05.alloc slot + base slot
constexpr int kFsFixedBytes = kFsZpartBytes + kFsZrecvBytes + kFsEringBytes + kFsBarBytes;

struct FusedParams {
    const float* ua;        // [ntq][N][2][2][128][4]
    const float* wb;        // [N][JG][2][2][128][4]
    const float* X;         // FWD: sum of v so far; BWD: ds^r     [nbt][C][4][32][4]
    const float* coef_in;   // BWD: c^r                            [nbt][N][8 JG][32]  (rows padded to whole capsule groups;
                            //                                      the padding capsules hold exact zeros)
    const float* beta_in;   // BWD: beta^{r+1} or nullptr
    float* coef_out;        // FWD: c^r (nullptr: not wanted); BWD: beta^r
    float* part;            // [IS][nbt][C][4][32][4]
    int N, C, JG, nbt, i_per_split, ns;
    int dbg;                // TIMING EXPERIMENTS ONLY (tuning knob "fsdbg"): 1 = no exchange at all (Z = 1), 2 = no coefficient stores
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 16 bytes from this thread's registers into another CTA's shared memory; the receiver's transaction barrier is
// credited with the 16 bytes when they have landed (no shared-memory staging, no proxy fence on the sender's side)
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// Shared-memory map (byte offsets from the dynamic base; the fixed part first so that every address the epilogue
// touches is `one base register + compile-time constant + small ring offset`):
constexpr uint32_t kOffZpart = 0;                                   // [slot 8][column half 2][128 samples] floats
constexpr uint32_t kOffZrecv = kOffZpart + kFsZpartBytes;           // [slot 8][sender rank 8][128] floats; ranks >= JG stay zero
constexpr uint32_t kOffEring = kOffZrecv + kFsZrecvBytes;           // [eslot 3][column half 2][capsule 4][128] floats: P1 -> P2
constexpr uint32_t kOffBars = kOffEring + kFsEringBytes;
constexpr uint32_t kOffSmemFull = kOffBars;                         // [kFsMaxStages]
constexpr uint32_t kOffSmemEmpty = kOffSmemFull + 8 * kFsMaxStages; // [kFsMaxStages]
constexpr uint32_t kOffTmemFull = kOffSmemEmpty + 8 * kFsMaxStages; // [kFsAccum]
constexpr uint32_t kOffTmemEmpty = kOffTmemFull + 8 * kFsAccum;     // [kFsAccum]
constexpr uint32_t kOffZlocal = kOffTmemEmpty + 8 * kFsAccum;       // [kFsZSlots]: the 8 epilogue warps wrote their partials
constexpr uint32_t kOffZfull = kOffZlocal + 8 * kFsZSlots;          // [kFsZSlots]: every CTA's row has landed (transaction count)
constexpr uint32_t kOffTmemSlot = kOffZfull + 8 * kFsZSlots;        // tcgen05.alloc result
constexpr uint32_t kOffBaseSlot = kOffTmemSlot + 4;                 // the base address itself, read back (see below)
constexpr uint32_t kOffStages = (kFsFixedBytes + 1023) & ~1023u;    // [ns][A 8 KB | B 8 KB | (BWD) c 4 KB | beta' 4 KB]
static_assert(kOffBaseSlot + 4 <= kOffStages, "fixed region overflows");

template <bool BWD>
__global__ void __launch_bounds__(kFsThreads, 1) k_sweep_fused(FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t kStageBytes = fs_stage_bytes(BWD);
    const int ns = p.ns;
    const uint32_t base0 = smem_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jg = blockIdx.x, tq = blockIdx.z;                      // cluster = the JG CTAs along x: rank == jg
    const int i_begin = blockIdx.y * p.i_per_split;
    const int i_end = min(p.N, i_begin + p.i_per_split);
    const int n_i = max(i_end - i_begin, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(base0 + kOffSmemFull + 8 * s, 1); mbar_init(base0 + kOffSmemEmpty + 8 * s, BWD ? kFsEpiWarps + 1 : 1); }
        for (int t = 0; t < kFsAccum; ++t) { mbar_init(base0 + kOffTmemFull + 8 * t, 1); mbar_init(base0 + kOffTmemEmpty + 8 * t, kFsEpiWarps); }
        for (int z = 0; z < kFsZSlots; ++z) { mbar_init(base0 + kOffZlocal + 8 * z, kFsEpiWarps); mbar_init(base0 + kOffZfull + 8 * z, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base0 + kOffBaseSlot), "r"(base0) : "memory");
    }
    for (int e = threadIdx.x; e < kFsZrecvBytes / 4; e += kFsThreads) sts_f32(base0 + kOffZrecv + 4 * e, 0.f);   // absent ranks add 0
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base0 + kOffTmemSlot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // The shared window address of this CTA is (0x400 | cluster rank << 24): cheap to rebuild on the uniform datapath, so
    // ptxas rematerialises it (S2UR + ULEA, a dependent chain) at every use inside the register-starved epilogue loop.
    // Reading it back through shared memory makes it an ordinary value that lives in one register.
    uint32_t base, tmem_base;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(base) : "r"(base0 + kOffBaseSlot) : "memory");
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(base0 + kOffTmemSlot) : "memory");
    cluster_sync_all();                      // every CTA's barriers and zeroed rows are in place before any remote traffic

    if (warp >= kFsEpiWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const uint32_t stages = base + kOffStages;
        if (warp == 8) {
            // ===== producer: one elected lane arms smem_full[s] and issues the stage's bulk copies =====
            const float* asrc = p.ua + ((size_t)tq * p.N + i_begin) * 2048;
            const float* bsrc = p.wb + ((size_t)i_begin * p.JG + jg) * 2048;
            const size_t bstep = (size_t)p.JG * 2048;
            const size_t CS = (size_t)p.JG * 8;                               // capsules per coefficient row (padded)
            const uint32_t cbytes = 8u * 128u;                                // the CTA's 8 capsules x 32 lanes
            const size_t ctile = (size_t)p.N * CS * kLanes;                   // coefficient floats per lane tile
            const size_t coff = ((size_t)(tq * 4) * p.N + i_begin) * CS * kLanes + (size_t)jg * 8 * kLanes;
            const float* csrc = BWD ? p.coef_in + coff : nullptr;
            const float* esrc = (BWD && p.beta_in != nullptr) ? p.beta_in + coff : nullptr;
            const size_t cstep = CS * kLanes;
            const int nvt = min(4, p.nbt - tq * 4);                          // valid lane tiles of this quad (>= 1)
            const uint32_t txbytes = (uint32_t)kFsOperandBytes + (BWD ? (uint32_t)nvt * cbytes * (esrc ? 2u : 1u) : 0u);
            int s = 0;
            uint32_t ph = 1;
            for (int n = 0; n < n_i; ++n) {
                mbar_wait_i(base + kOffSmemEmpty + 8 * s, ph);
                if (elect_one()) {
                    const uint32_t dst = stages + (uint32_t)s * kStageBytes, bar = base + kOffSmemFull + 8 * s;
                    mbar_expect_tx(bar, txbytes);
                    bulk_g2s(dst, asrc, 8192, bar);
                    bulk_g2s(dst + 8192, bsrc, 8192, bar);
                    if (BWD) {
                        for (int tt = 0; tt < nvt; ++tt) {
                            bulk_g2s(dst + kFsOperandBytes + tt * 1024, csrc + tt * ctile, cbytes, bar);
                            if (esrc) bulk_g2s(dst + kFsOperandBytes + kFsCoefBytes + tt * 1024, esrc + tt * ctile, cbytes, bar);
                        }
                    }
                }
                __syncwarp();
                asrc += 2048;
                bsrc += bstep;
                if (BWD) { csrc += cstep; if (esrc) esrc += cstep; }
                if (++s == ns) { s = 0; ph ^= 1; }
            }
        } else if (warp == 9) {
            // ===== MMA issuer (converged warp, one elected lane inside umma_stage) =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc0 = umma_desc(stages, 2048, 128);
            int s = 0;
            uint32_t sph = 0;
            for (int n = 0; n < n_i; ++n) {
                const int t = n & (kFsAccum - 1);
                mbar_wait_i(base + kOffTmemEmpty + 8 * t, ((n >> 2) & 1) ^ 1);
                mbar_wait_i(base + kOffSmemFull + 8 * s, sph);
                tc_fence_after();
                const uint64_t a_hi = desc0 + (uint64_t)((s * kStageBytes) >> 4);
                const uint64_t a_lo = a_hi + (4096 >> 4), b_hi = a_hi + (8192 >> 4), b_lo = b_hi + (4096 >> 4);
                umma_stage(tmem_base + (uint32_t)(t * 128), a_hi, a_lo, b_hi, b_lo, idesc, base + kOffSmemEmpty + 8 * s, base + kOffTmemFull + 8 * t);
                if (++s == ns) { s = 0; sph ^= 1; }
            }
        } else if (warp == 10 && !(p.dbg & 1)) {
            // ===== exchange: lane l adds the two column halves of samples 4l..4l+3 and stores the 16 bytes straight into
            // every CTA's receive row (st.async: the data and the receiver's transaction count travel together) =====
            const uint32_t nrank = (uint32_t)p.JG;
            const uint32_t my_row = base + kOffZrecv + (uint32_t)jg * 512 + (uint32_t)lane * 16;       // + slot * 4096, in the receiver
            uint32_t raddr[kFsMaxCluster], rbar[kFsMaxCluster];
#pragma unroll
            for (int r = 0; r < kFsMaxCluster; ++r) {
                const uint32_t rr = (uint32_t)r < nrank ? (uint32_t)r : 0u;
                raddr[r] = mapa_u32(my_row, rr);
                rbar[r] = mapa_u32(base + kOffZfull, rr);
            }
            for (int n = 0; n < n_i; ++n) {
                const uint32_t slot = (uint32_t)n & (kFsZSlots - 1), par = ((uint32_t)n >> 3) & 1;
                mbar_wait_i(base + kOffZlocal + 8 * slot, par);
                const float4 a = lds_v4(base + kOffZpart + slot * 1024 + (uint32_t)lane * 16);
                const float4 b = lds_v4(base + kOffZpart + slot * 1024 + 512 + (uint32_t)lane * 16);
                const float4 z = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                if (lane == 0) mbar_expect_tx(base + kOffZfull + 8 * slot, nrank * 512u);
#pragma unroll
                for (int r = 0; r < kFsMaxCluster; ++r)
                    if ((uint32_t)r < nrank) st_async_v4(raddr[r] + slot * 4096, z, rbar[r] + 8 * slot);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== epilogue: warp w -> samples of lane tile 4 tq + (w & 3), capsules jg*8 + 4 (w >> 2) .. + 3 =====
        const int q = warp & 3, jh = warp >> 2;
        const int tile = tq * 4 + q;
        const bool tvalid = tile < p.nbt;
        const int j0 = jg * 8 + jh * 4;
        // No per-capsule predicates inside the loop: coefficient rows are padded to whole groups of 8 capsules, a capsule
        // beyond C has W = 0 (u_hat = 0), X = 0, and (FWD) its exponent is clamped to -200 instead of +120 so that its
        // e_j, its coupling and everything stored for it are exact zeros; (BWD) it then reads c = beta' = 0 back.
        float lim[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            lim[jj] = (j0 + jj < p.C) ? 120.f : -200.f;
            asm volatile("mov.b32 %0, %0;" : "+f"(lim[jj]));       // pin: ptxas would rebuild it from ctaid / tid every iteration
        }
        const bool has_beta = BWD && p.beta_in != nullptr;
        const bool no_xchg = (p.dbg & 1) != 0;
        const bool do_store = tvalid && p.coef_out != nullptr && !(p.dbg & 2);             // warp-uniform
        float X[4][16];                 // FWD: log2(e) * sum of v; BWD: ds
        float acc[4][16];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int d = 0; d < 16; ++d) { X[jj][d] = 0.f; acc[jj][d] = 0.f; }
        if (tvalid) {
            const float sc = BWD ? 1.f : 1.4426950408889634f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                if (j0 + jj < p.C) {
#pragma unroll
                    for (int dq = 0; dq < 4; ++dq) {
                        const float4 x = ldg4(p.X + ((((size_t)tile * p.C + j0 + jj) * 4 + dq) * kLanes + lane) * 4);
                        X[jj][dq * 4 + 0] = x.x * sc; X[jj][dq * 4 + 1] = x.y * sc; X[jj][dq * 4 + 2] = x.z * sc; X[jj][dq * 4 + 3] = x.w * sc;
                    }
                }
        }
        // one thread base (sample column q*32 + lane of every [..][128]-float row) + per-purpose constants
        const uint32_t tb = base + (uint32_t)(q * 32 + lane) * 4;
        const uint32_t zp_base = tb + kOffZpart + (uint32_t)jh * 512;                    // + zs * 1024
        const uint32_t zr_base = tb + kOffZrecv;                                         // + zs * 4096 + rank * 512
        const uint32_t er_base = tb + kOffEring + (uint32_t)jh * 2048;                   // + eslot * 4096 + jj * 512
        const uint32_t co_base = base + kOffStages + kFsOperandBytes + (uint32_t)((q * 8 + jh * 4) * kLanes + lane) * 4;   // + stage * kStageBytes + jj * 128
        uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jh * 64);
        asm volatile("mov.b32 %0, %0;" : "+r"(lane_base));          // pin (same reason)
        // the coefficient row this warp stores in P2 (c^r or beta^r of (tile, i, j0 + jj)): advances one input capsule per stage
        uintptr_t cptr = reinterpret_cast<uintptr_t>(p.coef_out) + ((((size_t)(tvalid ? tile : 0) * p.N + i_begin) * (p.JG * 8) + j0) * kLanes + lane) * 4;
        const uintptr_t cstep = (uintptr_t)p.JG * 8 * kLanes * 4;
        uint32_t st1 = 0, st2 = 0;                      // BWD: byte offset of the P1 / P2 stage in the ring
        uint32_t sb2 = 0;                               // BWD: smem_empty barrier offset of the P2 stage
        const uint32_t st_end = (uint32_t)ns * kStageBytes;
        uint32_t er1 = 0, er2 = 0;                      // ering slot offsets
        const int n_it = n_i + kFsSkew;
        for (int it = 0; it < n_it; ++it) {
            if (it < n_i) {
                // ---------------- P1(it): logits / dc, partial normaliser ----------------
                const uint32_t t = (uint32_t)it & (kFsAccum - 1), zs = (uint32_t)it & (kFsZSlots - 1);
                mbar_wait_i(base + kOffTmemFull + 8 * t, ((uint32_t)it >> 2) & 1);
                tc_fence_after();
                float z = 0.f;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float uh[32];
                    tmem_ld32(lane_base + t * 128 + hh * 32, uh);
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        const int jj = hh * 2 + j2;
                        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                        for (int d = 0; d < 16; d += 4) {
                            ffma2(d0, d1, uh[j2 * 16 + d], uh[j2 * 16 + d + 1], X[jj][d], X[jj][d + 1]);
                            ffma2(d2, d3, uh[j2 * 16 + d + 2], uh[j2 * 16 + d + 3], X[jj][d + 2], X[jj][d + 3]);
                        }
                        const float dot = (d0 + d1) + (d2 + d3);
                        float keep;
                        if (!BWD) {
                            keep = ex2_approx(fminf(dot, lim[jj]));
                            z += keep;
                        } else {
                            keep = dot;
                            z = fmaf(lds_f32(co_base + st1 + jj * 128), dot, z);
                        }
                        sts_f32(er_base + er1 + jj * 512, keep);
                    }
                }
                if (!no_xchg) {
                    sts_f32(zp_base + zs * 1024, z);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(base + kOffZlocal + 8 * zs);
                }
                if (BWD) { st1 += kStageBytes; if (st1 == st_end) st1 = 0; }
                er1 += 4096; if (er1 == (kFsSkew + 1) * 4096) er1 = 0;
            }
            if (it >= kFsSkew) {
                // ---------------- P2(it - skew): normalise, store, accumulate ----------------
                const uint32_t m = (uint32_t)(it - kFsSkew), t = m & (kFsAccum - 1), zs = m & (kFsZSlots - 1);
                float Z = 1.f;
                if (!no_xchg) {
                    mbar_wait_i(base + kOffZfull + 8 * zs, (m >> 3) & 1);
                    const uint32_t zr = zr_base + zs * 4096;
                    float zz[kFsMaxCluster];
#pragma unroll
                    for (int r = 0; r < kFsMaxCluster; ++r) zz[r] = lds_f32(zr + r * 512);
                    Z = ((zz[0] + zz[1]) + (zz[2] + zz[3])) + ((zz[4] + zz[5]) + (zz[6] + zz[7]));    // fixed order: same bits in every CTA
                }
                float f[4];
                const float rz = BWD ? 0.f : rcp_approx(Z);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float keep = lds_f32(er_base + er2 + jj * 512);
                    if (!BWD) {
                        f[jj] = keep * rz;
                    } else {
                        const float c = lds_f32(co_base + st2 + jj * 128);
                        const float bp = has_beta ? lds_f32(co_base + st2 + kFsCoefBytes + jj * 128) : 0.f;
                        f[jj] = fmaf(c, keep - Z, bp);
                    }
                }
                if (do_store) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) reinterpret_cast<float*>(cptr)[jj * kLanes] = f[jj];
                }
                tc_fence_after();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float uh[32];
                    tmem_ld32(lane_base + t * 128 + hh * 32, uh);
                    if (hh == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(base + kOffTmemEmpty + 8 * t);               // accumulator t may be overwritten
                            if (BWD) mbar_arrive(base + kOffSmemEmpty + sb2);        // and the coefficient rows of the P2 stage
                        }
                    }
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        const int jj = hh * 2 + j2;
#pragma unroll
                        for (int d = 0; d < 16; d += 2) ffma2(acc[jj][d], acc[jj][d + 1], f[jj], f[jj], uh[j2 * 16 + d], uh[j2 * 16 + d + 1]);
                    }
                }
                cptr += cstep;
                if (BWD) { st2 += kStageBytes; sb2 += 8; if (st2 == st_end) { st2 = 0; sb2 = 0; } }
                er2 += 4096; if (er2 == (kFsSkew + 1) * 4096) er2 = 0;
            }
        }
        if (tvalid) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                if (j0 + jj < p.C) {
#pragma unroll
                    for (int dq = 0; dq < 4; ++dq)
                        st4(p.part + (((((size_t)blockIdx.y * p.nbt + tile) * p.C + j0 + jj) * 4 + dq) * kLanes + lane) * 4,
                            make_float4(acc[jj][dq * 4 + 0], acc[jj][dq * 4 + 1], acc[jj][dq * 4 + 2], acc[jj][dq * 4 + 3]));
                }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    cluster_sync_all();                      // no CTA leaves while a peer's stores may still be on their way to it
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised variant (tuning knob "fsws", default on): the same stages, barriers and exchange, but P1 and P2
// are run by DIFFERENT warps -- 8 "logit" warps (P1) and 8 "accumulate" warps (P2) -- instead of one warp doing
// P1(it) and then P2(it-2).  The epilogue is a chain of fixed latencies (barrier probe, tcgen05.ld, FFMA2 chain,
// exchange, LDS) that two warps per scheduler cannot hide (ncu: 36 % issue utilisation); four warps per scheduler
// with half the instruction stream each can.  The register state splits exactly: the probe vectors X live only in
// the P1 warps, the accumulators only in the P2 warps (64 registers each), so 16 warps fit where 8 did.
//   P1 -> P2 hand-over of e_j / dc_j: through the `ering` rows in shared memory as before; the receiver's wait on
//   zfull[m] orders it (every P1 warp of every CTA arrived on its zlocal[m], a release, after storing its rows).
//   P1 runs at most kFsAccum stages ahead of P2 (it needs accumulator it, which the issuer refills only after P2 of
//   stage it-4 released it), so the ering holds kFsAccum slots and the credit-free argument for the 8-slot exchange
//   ring still holds: a CTA sends row m only after its own P2(m-4), which needed every peer's row m-4, which each peer
//   sent after its own P2(m-8).
constexpr int kWsThreads = 640;            // warps 0-7 P1, 8-15 P2, 16 producer, 17 MMA issuer, 18 exchange, 19 idle
constexpr int kWsEringSlots = kFsAccum;
constexpr int kWsEringBytes = kWsEringSlots * kFsEpiWarps * 4 * 32 * 4;
constexpr int kWsFixedBytes = kFsZpartBytes + kFsZrecvBytes + kWsEringBytes + kFsBarBytes;
constexpr uint32_t kWsOffZpart = 0;
constexpr uint32_t kWsOffZrecv = kWsOffZpart + kFsZpartBytes;
constexpr uint32_t kWsOffEring = kWsOffZrecv + kFsZrecvBytes;
constexpr uint32_t kWsOffBars = kWsOffEring + kWsEringBytes;
constexpr uint32_t kWsOffSmemFull = kWsOffBars;
constexpr uint32_t kWsOffSmemEmpty = kWsOffSmemFull + 8 * kFsMaxStages;
constexpr uint32_t kWsOffTmemFull = kWsOffSmemEmpty + 8 * kFsMaxStages;
constexpr uint32_t kWsOffTmemEmpty = kWsOffTmemFull + 8 * kFsAccum;
constexpr uint32_t kWsOffZlocal = kWsOffTmemEmpty + 8 * kFsAccum;
constexpr uint32_t kWsOffZfull = kWsOffZlocal + 8 * kFsZSlots;
constexpr uint32_t kWsOffTmemSlot = kWsOffZfull + 8 * kFsZSlots;
constexpr uint32_t kWsOffBaseSlot = kWsOffTmemSlot + 4;
constexpr uint32_t kWsOffStages = (kWsFixedBytes + 1023) & ~1023u;
static_assert(kWsOffBaseSlot + 4 <= kWsOffStages, "fixed region overflows");

template <bool BWD>
__global__ void __launch_bounds__(kWsThreads, 1) k_sweep_fused_ws(FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t kStageBytes = fs_stage_bytes(BWD);
    const int ns = p.ns;
    const uint32_t base0 = smem_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jg = blockIdx.x, tq = blockIdx.z;
    const int i_begin = blockIdx.y * p.i_per_split;
    const int i_end = min(p.N, i_begin + p.i_per_split);
    const int n_i = max(i_end - i_begin, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(base0 + kWsOffSmemFull + 8 * s, 1); mbar_init(base0 + kWsOffSmemEmpty + 8 * s, BWD ? kFsEpiWarps + 1 : 1); }
        for (int t = 0; t < kFsAccum; ++t) { mbar_init(base0 + kWsOffTmemFull + 8 * t, 1); mbar_init(base0 + kWsOffTmemEmpty + 8 * t, kFsEpiWarps); }
        for (int z = 0; z < kFsZSlots; ++z) { mbar_init(base0 + kWsOffZlocal + 8 * z, kFsEpiWarps); mbar_init(base0 + kWsOffZfull + 8 * z, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base0 + kWsOffBaseSlot), "r"(base0) : "memory");
    }
    for (int e = threadIdx.x; e < kFsZrecvBytes / 4; e += kWsThreads) sts_f32(base0 + kWsOffZrecv + 4 * e, 0.f);   // absent ranks add 0
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base0 + kWsOffTmemSlot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t base, tmem_base;            // read back through shared memory: see k_sweep_fused
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(base) : "r"(base0 + kWsOffBaseSlot) : "memory");
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(base0 + kWsOffTmemSlot) : "memory");
    cluster_sync_all();

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");      // 640 x 96 registers = 256 x 104 (P1) + 256 x 112 (P2) + 128 x 48
        const uint32_t stages = base + kWsOffStages;
        if (warp == 16) {
            // ===== producer =====
            const float* asrc = p.ua + ((size_t)tq * p.N + i_begin) * 2048;
            const float* bsrc = p.wb + ((size_t)i_begin * p.JG + jg) * 2048;
            const size_t bstep = (size_t)p.JG * 2048;
            const size_t CS = (size_t)p.JG * 8;
            const uint32_t cbytes = 8u * 128u;
            const size_t ctile = (size_t)p.N * CS * kLanes;
            const size_t coff = ((size_t)(tq * 4) * p.N + i_begin) * CS * kLanes + (size_t)jg * 8 * kLanes;
            const float* csrc = BWD ? p.coef_in + coff : nullptr;
            const float* esrc = (BWD && p.beta_in != nullptr) ? p.beta_in + coff : nullptr;
            const size_t cstep = CS * kLanes;
            const int nvt = min(4, p.nbt - tq * 4);
            const uint32_t txbytes = (uint32_t)kFsOperandBytes + (BWD ? (uint32_t)nvt * cbytes * (esrc ? 2u : 1u) : 0u);
            int s = 0;
            uint32_t ph = 1;
            for (int n = 0; n < n_i; ++n) {
                mbar_wait_i(base + kWsOffSmemEmpty + 8 * s, ph);
                if (elect_one()) {
                    const uint32_t dst = stages + (uint32_t)s * kStageBytes, bar = base + kWsOffSmemFull + 8 * s;
                    mbar_expect_tx(bar, txbytes);
                    bulk_g2s(dst, asrc, 8192, bar);
                    bulk_g2s(dst + 8192, bsrc, 8192, bar);
                    if (BWD) {
                        for (int tt = 0; tt < nvt; ++tt) {
                            bulk_g2s(dst + kFsOperandBytes + tt * 1024, csrc + tt * ctile, cbytes, bar);
                            if (esrc) bulk_g2s(dst + kFsOperandBytes + kFsCoefBytes + tt * 1024, esrc + tt * ctile, cbytes, bar);
                        }
                    }
                }
                __syncwarp();
                asrc += 2048;
                bsrc += bstep;
                if (BWD) { csrc += cstep; if (esrc) esrc += cstep; }
                if (++s == ns) { s = 0; ph ^= 1; }
            }
        } else if (warp == 17) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc0 = umma_desc(stages, 2048, 128);
            int s = 0;
            uint32_t sph = 0;
            for (int n = 0; n < n_i; ++n) {
                const int t = n & (kFsAccum - 1);
                mbar_wait_i(base + kWsOffTmemEmpty + 8 * t, ((n >> 2) & 1) ^ 1);
                mbar_wait_i(base + kWsOffSmemFull + 8 * s, sph);
                tc_fence_after();
                const uint64_t a_hi = desc0 + (uint64_t)((s * kStageBytes) >> 4);
                const uint64_t a_lo = a_hi + (4096 >> 4), b_hi = a_hi + (8192 >> 4), b_lo = b_hi + (4096 >> 4);
                umma_stage(tmem_base + (uint32_t)(t * 128), a_hi, a_lo, b_hi, b_lo, idesc, base + kWsOffSmemEmpty + 8 * s, base + kWsOffTmemFull + 8 * t);
                if (++s == ns) { s = 0; sph ^= 1; }
            }
        } else if (warp == 18) {
            // ===== exchange =====
            const uint32_t nrank = (uint32_t)p.JG;
            const uint32_t my_row = base + kWsOffZrecv + (uint32_t)jg * 512 + (uint32_t)lane * 16;
            uint32_t raddr[kFsMaxCluster], rbar[kFsMaxCluster];
#pragma unroll
            for (int r = 0; r < kFsMaxCluster; ++r) {
                const uint32_t rr = (uint32_t)r < nrank ? (uint32_t)r : 0u;
                raddr[r] = mapa_u32(my_row, rr);
                rbar[r] = mapa_u32(base + kWsOffZfull, rr);
            }
            for (int n = 0; n < n_i; ++n) {
                const uint32_t slot = (uint32_t)n & (kFsZSlots - 1), par = ((uint32_t)n >> 3) & 1;
                mbar_wait_i(base + kWsOffZlocal + 8 * slot, par);
                const float4 a = lds_v4(base + kWsOffZpart + slot * 1024 + (uint32_t)lane * 16);
                const float4 b = lds_v4(base + kWsOffZpart + slot * 1024 + 512 + (uint32_t)lane * 16);
                const float4 z = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                if (lane == 0) mbar_expect_tx(base + kWsOffZfull + 8 * slot, nrank * 512u);
#pragma unroll
                for (int r = 0; r < kFsMaxCluster; ++r)
                    if ((uint32_t)r < nrank) st_async_v4(raddr[r] + slot * 4096, z, rbar[r] + 8 * slot);
            }
        }
    } else {
        // warp w (mod 8) -> samples of lane tile 4 tq + (w & 3), capsules jg*8 + 4 ((w >> 2) & 1) .. + 3
        const int q = warp & 3, jh = (warp >> 2) & 1;
        const int tile = tq * 4 + q;
        const bool tvalid = tile < p.nbt;
        const int j0 = jg * 8 + jh * 4;
        const uint32_t tb = base + (uint32_t)(q * 32 + lane) * 4;
        const uint32_t er_base = tb + kWsOffEring + (uint32_t)jh * 2048;                 // + (stage & 3) * 4096 + jj * 512
        const uint32_t co_base = base + kWsOffStages + kFsOperandBytes + (uint32_t)((q * 8 + jh * 4) * kLanes + lane) * 4;   // + stage * kStageBytes + jj * 128
        uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jh * 64);
        asm volatile("mov.b32 %0, %0;" : "+r"(lane_base));
        const uint32_t st_end = (uint32_t)ns * kStageBytes;
        if (warp < kFsEpiWarps) {
            // ===================== P1 warps: logits / dc, partial normaliser =====================
            asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
            float lim[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                lim[jj] = (j0 + jj < p.C) ? 120.f : -200.f;
                asm volatile("mov.b32 %0, %0;" : "+f"(lim[jj]));
            }
            float X[4][16];                 // FWD: log2(e) * sum of v; BWD: ds
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int d = 0; d < 16; ++d) X[jj][d] = 0.f;
            if (tvalid) {
                const float sc = BWD ? 1.f : 1.4426950408889634f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (j0 + jj < p.C) {
#pragma unroll
                        for (int dq = 0; dq < 4; ++dq) {
                            const float4 x = ldg4(p.X + ((((size_t)tile * p.C + j0 + jj) * 4 + dq) * kLanes + lane) * 4);
                            X[jj][dq * 4 + 0] = x.x * sc; X[jj][dq * 4 + 1] = x.y * sc; X[jj][dq * 4 + 2] = x.z * sc; X[jj][dq * 4 + 3] = x.w * sc;
                        }
                    }
            }
            const uint32_t zp_base = tb + kWsOffZpart + (uint32_t)jh * 512;              // + zs * 1024
            uint32_t st1 = 0;
            for (int it = 0; it < n_i; ++it) {
                const uint32_t t = (uint32_t)it & (kFsAccum - 1), zs = (uint32_t)it & (kFsZSlots - 1);
                mbar_wait_i(base + kWsOffTmemFull + 8 * t, ((uint32_t)it >> 2) & 1);
                tc_fence_after();
                float z = 0.f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float uh[16];
                    tmem_ld16(lane_base + t * 128 + jj * 16, uh);
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int d = 0; d < 16; d += 4) {
                        ffma2(d0, d1, uh[d], uh[d + 1], X[jj][d], X[jj][d + 1]);
                        ffma2(d2, d3, uh[d + 2], uh[d + 3], X[jj][d + 2], X[jj][d + 3]);
                    }
                    const float dot = (d0 + d1) + (d2 + d3);
                    float keep;
                    if (!BWD) {
                        keep = ex2_approx(fminf(dot, lim[jj]));
                        z += keep;
                    } else {
                        keep = dot;
                        z = fmaf(lds_f32(co_base + st1 + jj * 128), dot, z);
                    }
                    sts_f32(er_base + t * 4096 + jj * 512, keep);
                }
                sts_f32(zp_base + zs * 1024, z);
                __syncwarp();
                if (lane == 0) mbar_arrive(base + kWsOffZlocal + 8 * zs);
                if (BWD) { st1 += kStageBytes; if (st1 == st_end) st1 = 0; }
            }
        } else {
            // ===================== P2 warps: normalise, store, accumulate =====================
            asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
            const bool has_beta = BWD && p.beta_in != nullptr;
            const bool do_store = tvalid && p.coef_out != nullptr && !(p.dbg & 2);             // warp-uniform
            const bool more_than_6 = p.JG > 6;
            float acc[4][16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int d = 0; d < 16; ++d) acc[jj][d] = 0.f;
            const uint32_t zr_base = tb + kWsOffZrecv;                                   // + zs * 4096 + rank * 512
            uintptr_t cptr = reinterpret_cast<uintptr_t>(p.coef_out) + ((((size_t)(tvalid ? tile : 0) * p.N + i_begin) * (p.JG * 8) + j0) * kLanes + lane) * 4;
            const uintptr_t cstep = (uintptr_t)p.JG * 8 * kLanes * 4;
            uint32_t st2 = 0, sb2 = 0;
            for (int m = 0; m < n_i; ++m) {
                const uint32_t t = (uint32_t)m & (kFsAccum - 1), zs = (uint32_t)m & (kFsZSlots - 1);
                // zfull[m] complete: every CTA's P1 warps are done with stage m (their rows are here), and each of them had
                // observed the accumulator's commit before it arrived -- the chain zlocal -> exchange -> zfull orders this
                // warp's tcgen05.ld behind that commit, so the accumulator barrier is not taken a second time
                mbar_wait_i(base + kWsOffZfull + 8 * zs, ((uint32_t)m >> 3) & 1);
                tc_fence_after();
                uint32_t ur[32];
                tmem_ld32_issue(lane_base + t * 128, ur);          // capsules 0 and 1 travel while Z is summed
                const uint32_t zr = zr_base + zs * 4096;
                float zz[kFsMaxCluster];
#pragma unroll
                for (int r = 0; r < 6; ++r) zz[r] = lds_f32(zr + r * 512);
                // rows of absent ranks are exact zeros: with <= 6 CTAs per cluster (C <= 48) their loads and adds are skipped,
                // (z4 + z5) + 0 == z4 + z5 bit for bit, so both branches give every CTA the same Z
                float z67 = 0.f;
                if (more_than_6) z67 = lds_f32(zr + 6 * 512) + lds_f32(zr + 7 * 512);
                const float Z = ((zz[0] + zz[1]) + (zz[2] + zz[3])) + ((zz[4] + zz[5]) + z67);    // fixed order: same bits in every CTA
                float f[4];
                const float rz = BWD ? 0.f : rcp_approx(Z);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float keep = lds_f32(er_base + t * 4096 + jj * 512);
                    if (!BWD) {
                        f[jj] = keep * rz;
                    } else {
                        const float c = lds_f32(co_base + st2 + jj * 128);
                        const float bp = has_beta ? lds_f32(co_base + st2 + kFsCoefBytes + jj * 128) : 0.f;
                        f[jj] = fmaf(c, keep - Z, bp);
                    }
                }
                if (do_store) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) reinterpret_cast<float*>(cptr)[jj * kLanes] = f[jj];
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (hh > 0) tmem_ld32_issue(lane_base + t * 128 + 32, ur);
                    tmem_ld32_wait(ur);
                    if (hh == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(base + kWsOffTmemEmpty + 8 * t);               // accumulator t may be overwritten
                            if (BWD) mbar_arrive(base + kWsOffSmemEmpty + sb2);        // and the coefficient rows of this stage
                        }
                    }
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        const int jj = hh * 2 + j2;
#pragma unroll
                        for (int d = 0; d < 16; d += 2)
                            ffma2(acc[jj][d], acc[jj][d + 1], f[jj], f[jj], __uint_as_float(ur[j2 * 16 + d]), __uint_as_float(ur[j2 * 16 + d + 1]));
                    }
                }
                cptr += cstep;
                if (BWD) { st2 += kStageBytes; sb2 += 8; if (st2 == st_end) { st2 = 0; sb2 = 0; } }
            }
            if (tvalid) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (j0 + jj < p.C) {
#pragma unroll
                        for (int dq = 0; dq < 4; ++dq)
                            st4(p.part + (((((size_t)blockIdx.y * p.nbt + tile) * p.C + j0 + jj) * 4 + dq) * kLanes + lane) * 4,
                                make_float4(acc[jj][dq * 4 + 0], acc[jj][dq * 4 + 1], acc[jj][dq * 4 + 2], acc[jj][dq * 4 + 3]));
                    }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    cluster_sync_all();
}

// couplings in the lane-tile layout [nbt][N][C][32] -> public [B][N][C] (tests and callers that ask for c_out)
__global__ void k_coef_public(const float* __restrict__ coef, float* __restrict__ c_pub, int B, int N, int C, int CS, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;
    const long b = (ti / N) * kLanes + lane;
    if (b >= B) return;
    const int i = (int)(ti % N);
    const float* src = coef + (size_t)ti * CS * kLanes + lane;
    float* dst = c_pub + ((size_t)b * N + i) * C;
    for (int j = 0; j < C; ++j) dst[j] = src[(size_t)j * kLanes];
}

struct ClusterCap { std::atomic<int> n[kMaxDevices][kFsMaxCluster + 1]; };      // 0 = not queried yet
ClusterCap g_cap[2][2];                                                          // [variant][bwd]

inline size_t fs_fixed_bytes(bool ws) { return ws ? (size_t)kWsOffStages : (size_t)kOffStages; }

int fs_stages(bool bwd, bool ws) {
    int ns = bwd ? 7 : 8;
    while (fs_fixed_bytes(ws) + (size_t)ns * fs_stage_bytes(bwd) > 227 * 1024) --ns;
    return ns;
}

const void* fs_kernel(bool bwd, bool ws) {
    if (ws) return bwd ? reinterpret_cast<const void*>(k_sweep_fused_ws<true>) : reinterpret_cast<const void*>(k_sweep_fused_ws<false>);
    return bwd ? reinterpret_cast<const void*>(k_sweep_fused<true>) : reinterpret_cast<const void*>(k_sweep_fused<false>);
}

int launch_k(const Plan& pl, const FusedParams& fp, int IS, bool bwd, bool ws, cudaStream_t st) {
    const void* kern = fs_kernel(bwd, ws);
    const size_t smem = fs_fixed_bytes(ws) + (size_t)fp.ns * fs_stage_bytes(bwd);
    {
        static SmemAttrCache cache[2][2];
        const int rc = ensure_dyn_smem(kern, smem, cache[ws ? 1 : 0][bwd ? 1 : 0]);
        if (rc) return rc;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fp.JG, IS, cdiv(pl.nbt, 4));
    cfg.blockDim = dim3(ws ? kWsThreads : kFsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = fp.JG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FusedParams arg = fp;
    void* args[1] = {&arg};
    CUDA_TRY(cudaLaunchKernelExC(&cfg, kern, args));
    return 0;
}

}  // namespace

int g_fs_dbg = 0;
int g_fs_ws = 1;         // tuning knob "fsws": 1 = warp-specialised epilogue (8 logit warps + 8 accumulate warps), 0 = 8 warps doing both

// The warp-specialised kernel re-partitions its register pool with setmaxnreg: 256 threads x 104 + 256 x 112 + 128 x 48 =
// 640 x 96.  That only adds up if ptxas gave the kernel exactly 96 registers per thread at launch; with fewer, an
// `inc` would wait for registers that never come.  Checked once per process; otherwise the 8-warp kernel runs.
bool fs_use_ws() {
    if (g_fs_ws == 0) return false;
    static const bool ok = [] {
        cudaFuncAttributes a0{}, a1{};
        if (cudaFuncGetAttributes(&a0, k_sweep_fused_ws<false>) != cudaSuccess || cudaFuncGetAttributes(&a1, k_sweep_fused_ws<true>) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a0.numRegs >= 96 && a1.numRegs >= 96;
    }();
    return ok;
}

// D == 16 (after padding), 8 capsules per CTA, one cluster of ceil(C/8) <= 8 CTAs per (128 samples, i range)
bool fused_shape_ok(int C, int DP, bool tc_ok) { return tc_ok && DP == 16 && cdiv(C, 8) <= kFsMaxCluster; }
bool fused_supported(const Plan& pl) { return pl.use_tc && pl.Reff > 1 && fused_shape_ok(pl.C, pl.DP, pl.tc_ok); }

// clusters of `jg` CTAs the device can run at once (GPC granularity: not simply SMs / jg)
int fused_cluster_capacity(int jg, bool bwd) {
    const bool ws = fs_use_ws();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || jg < 1 || jg > kFsMaxCluster) return 0;
    std::atomic<int>& slot = g_cap[ws ? 1 : 0][bwd ? 1 : 0].n[dev][jg];
    int n = slot.load(std::memory_order_relaxed);
    if (n > 0) return n;
    const size_t smem = fs_fixed_bytes(ws) + (size_t)fs_stages(bwd, ws) * fs_stage_bytes(bwd);
    const void* kern = fs_kernel(bwd, ws);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(jg, 64, 1);
    cfg.blockDim = dim3(ws ? kWsThreads : kFsThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = jg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); return 0; }
    slot.store(nc, std::memory_order_relaxed);
    return nc;
}

// number of splits of the i range: fill whole waves of the cluster capacity, keep >= 32 stages per CTA.
// Returns a count that tiles N exactly with splits of a multiple of 4 input capsules.
int fused_pick_splits(const Plan& pl, bool bwd, int forced) {
    auto effective = [&](int cand) { return cdiv(pl.N, cdiv(cdiv(pl.N, cand), 4) * 4); };
    const int max_is = std::max(1, std::min(kMaxSplits, pl.N / 32));
    if (forced > 0) return effective(std::min(forced, max_is));
    const int cap = fused_cluster_capacity(cdiv(pl.C, 8), bwd);
    const long per_split = cdiv(pl.nbt, 4);
    if (cap <= 0) return effective(std::min(max_is, std::max(1, cdiv(24, per_split))));
    int best_is = 1;
    double best = -1.0;
    for (int cand = 1; cand <= max_is; ++cand) {
        const int is = effective(cand);
        const long g = per_split * is;
        const double eff = (double)g / (double)(((g + cap - 1) / cap) * cap);
        if (eff > best + 0.04) { best = eff; best_is = is; }
    }
    return best_is;
}

int launch_sweep_fused(const Plan& pl, bool bwd, const float* ua, const float* wb, const float* X, const float* coef_in,
                       const float* beta_in, float* coef_out, float* part, int IS, cudaStream_t st) {
    FusedParams fp{};
    fp.ua = ua; fp.wb = wb; fp.X = X; fp.coef_in = coef_in; fp.beta_in = beta_in; fp.coef_out = coef_out; fp.part = part;
    fp.N = pl.N; fp.C = pl.C; fp.JG = cdiv(pl.C, 8); fp.nbt = pl.nbt;
    fp.i_per_split = cdiv(cdiv(pl.N, IS), 4) * 4;
    const bool ws = fs_use_ws();
    fp.ns = fs_stages(bwd, ws);
    fp.dbg = g_fs_dbg;
    if (cdiv(pl.N, fp.i_per_split) != IS) return fail(CAPS_E_BADARG, "fused sweep: %d splits do not tile N=%d", IS, pl.N);
    return launch_k(pl, fp, IS, bwd, ws, st);
}

int launch_coef_public(const Plan& pl, const float* coef, int CS, float* c_pub, cudaStream_t st) {
    const long n = (long)pl.nbt * pl.N * kLanes;
    k_coef_public<<<cdiv(n, 256), 256, 0, st>>>(coef, c_pub, pl.B, pl.N, pl.C, CS, pl.nbt);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace caps
