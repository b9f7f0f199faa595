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
//             z -> shared; the exchange warp adds the two column halves and sends the 128-sample row (512 B)
//             to every CTA of the cluster with cp.async.bulk (shared::cta -> shared::cluster), which completes
//             a transaction barrier at the receiver.
//     P2(n-2) Z = sum over the cluster's rows (fixed order: every CTA gets bit-identical Z)
//             FWD: c_j = e_j / Z          -> stored once for the backward     acc_j += c_j u_hat_j
//             BWD: beta_j = beta'_j + c_j (dc_j - Z) -> stored once           acc_j += beta_j u_hat_j
//             (u_hat is read again from TMEM: the accumulator ring is 4 deep, the exchange takes < 2 stages)
//
// So per iteration the [B,N,C] data is written ONCE (what the backward needs) and never re-read by the forward.
// Flow control of the exchange ring needs no credits: a CTA can only send the row of stage n+4 after every CTA
// of the cluster has consumed the row of stage n (its P1(n+4) follows its own P2(n+2), which needed every peer's
// row n+2, which each peer sent after its own P2(n)); see DESIGN.md section 3.4.
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
constexpr int kFsZSlots = 4;               // exchange ring depth (>= 2 * kFsSkew: the credit-free argument above)
constexpr int kFsMaxCluster = 8;           // portable cluster size: C <= 64 at 8 capsules per CTA
constexpr int kFsMaxStages = 12;
constexpr int kFsOperandBytes = 16384;     // A (8 KB: u hi/lo) + B (8 KB: W hi/lo), layouts of caps_pass_tc.cu
constexpr int kFsCoefBytes = 4096;         // [4 lane tiles][8 capsules][32 lanes] floats
__host__ __device__ constexpr int fs_stage_bytes(bool bwd) { return kFsOperandBytes + (bwd ? 2 * kFsCoefBytes : 0); }
constexpr int kFsZpartBytes = kFsZSlots * 2 * 128 * 4;
constexpr int kFsZcombBytes = kFsZSlots * 128 * 4;
constexpr int kFsZrecvBytes = kFsZSlots * kFsMaxCluster * 128 * 4;
constexpr int kFsEringBytes = (kFsSkew + 1) * kFsEpiWarps * 4 * 32 * 4;
constexpr int kFsBarBytes = 8 * (2 * kFsMaxStages + 2 * kFsAccum + 2 * kFsZSlots) + 16;
constexpr int kFsFixedBytes = kFsZpartBytes + kFsZcombBytes + kFsZrecvBytes + kFsEringBytes + kFsBarBytes;

struct FusedParams {
    const float* ua;        // [ntq][N][2][2][128][4]
    const float* wb;        // [N][JG][2][2][128][4]
    const float* X;         // FWD: sum of v so far; BWD: ds^r     [nbt][C][4][32][4]
    const float* coef_in;   // BWD: c^r                            [nbt][N][C][32]
    const float* beta_in;   // BWD: beta^{r+1} or nullptr
    float* coef_out;        // FWD: c^r (nullptr: not wanted); BWD: beta^r
    float* part;            // [IS][nbt][C][4][32][4]
    int N, C, JG, nbt, i_per_split, ns;
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

template <bool BWD>
__global__ void __launch_bounds__(kFsThreads, 1) k_sweep_fused(FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int kStageBytes = fs_stage_bytes(BWD);
    const int ns = p.ns;
    const uint32_t stages = smem_u32(smem_raw);
    const uint32_t zpart = stages + (uint32_t)ns * kStageBytes;      // [slot][column half][128 samples]
    const uint32_t zcomb = zpart + kFsZpartBytes;                    // [slot][128]: the row this CTA sends
    const uint32_t zrecv = zcomb + kFsZcombBytes;                    // [slot][sender rank][128]
    const uint32_t ering = zrecv + kFsZrecvBytes;                    // [skew+1][warp][4][32]: e_j (FWD) / dc_j (BWD) from P1 to P2
    const uint32_t bars = ering + kFsEringBytes;
    const uint32_t smem_full = bars;                                 // [kFsMaxStages]
    const uint32_t smem_empty = smem_full + 8 * kFsMaxStages;        // [kFsMaxStages]
    const uint32_t tmem_full = smem_empty + 8 * kFsMaxStages;        // [kFsAccum]
    const uint32_t tmem_empty = tmem_full + 8 * kFsAccum;            // [kFsAccum]
    const uint32_t zlocal = tmem_empty + 8 * kFsAccum;               // [kFsZSlots]: the 8 epilogue warps wrote their partials
    const uint32_t zfull = zlocal + 8 * kFsZSlots;                   // [kFsZSlots]: every CTA's row has landed (transaction count)
    const uint32_t tmem_slot = zfull + 8 * kFsZSlots;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - stages));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jg = blockIdx.x, tq = blockIdx.z;                      // cluster = the JG CTAs along x: rank == jg
    const int i_begin = blockIdx.y * p.i_per_split;
    const int i_end = min(p.N, i_begin + p.i_per_split);
    const int n_i = max(i_end - i_begin, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(smem_full + 8 * s, 1); mbar_init(smem_empty + 8 * s, BWD ? kFsEpiWarps + 1 : 1); }
        for (int t = 0; t < kFsAccum; ++t) { mbar_init(tmem_full + 8 * t, 1); mbar_init(tmem_empty + 8 * t, kFsEpiWarps); }
        for (int z = 0; z < kFsZSlots; ++z) { mbar_init(zlocal + 8 * z, kFsEpiWarps); mbar_init(zfull + 8 * z, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    cluster_sync_all();                      // every CTA's barriers are initialised before any remote traffic

    if (warp >= kFsEpiWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8) {
            // ===== producer: one elected lane arms smem_full[s] and issues the stage's bulk copies =====
            const float* asrc = p.ua + ((size_t)tq * p.N + i_begin) * 2048;
            const float* bsrc = p.wb + ((size_t)i_begin * p.JG + jg) * 2048;
            const size_t bstep = (size_t)p.JG * 2048;
            const int nj = min(8, p.C - jg * 8);
            const uint32_t cbytes = (uint32_t)nj * 128u;
            const size_t ctile = (size_t)p.N * p.C * kLanes;                  // coefficient floats per lane tile
            const size_t coff = ((size_t)(tq * 4) * p.N + i_begin) * p.C * kLanes + (size_t)jg * 8 * kLanes;
            const float* csrc = BWD ? p.coef_in + coff : nullptr;
            const float* esrc = (BWD && p.beta_in != nullptr) ? p.beta_in + coff : nullptr;
            const size_t cstep = (size_t)p.C * kLanes;
            const int nvt = min(4, p.nbt - tq * 4);                          // valid lane tiles of this quad (>= 1)
            const uint32_t txbytes = (uint32_t)kFsOperandBytes + (BWD ? (uint32_t)nvt * cbytes * (esrc ? 2u : 1u) : 0u);
            int s = 0;
            uint32_t ph = 1;
            for (int n = 0; n < n_i; ++n) {
                mbar_wait_i(smem_empty + 8 * s, ph);
                if (elect_one()) {
                    const uint32_t dst = stages + (uint32_t)s * kStageBytes, bar = smem_full + 8 * s;
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
                mbar_wait_i(tmem_empty + 8 * t, ((n >> 2) & 1) ^ 1);
                mbar_wait_i(smem_full + 8 * s, sph);
                tc_fence_after();
                const uint64_t a_hi = desc0 + (uint64_t)((s * kStageBytes) >> 4);
                const uint64_t a_lo = a_hi + (4096 >> 4), b_hi = a_hi + (8192 >> 4), b_lo = b_hi + (4096 >> 4);
                umma_stage(tmem_base + (uint32_t)(t * 128), a_hi, a_lo, b_hi, b_lo, idesc, smem_empty + 8 * s, tmem_full + 8 * t);
                if (++s == ns) { s = 0; sph ^= 1; }
            }
        } else if (warp == 10) {
            // ===== exchange: add the two column halves of a stage's partial normaliser, send the row to the cluster =====
            const uint32_t nrank = (uint32_t)p.JG;
            for (int n = 0; n < n_i; ++n) {
                const uint32_t slot = (uint32_t)n & (kFsZSlots - 1), par = ((uint32_t)n >> 2) & 1;
                mbar_wait_i(zlocal + 8 * slot, par);
                const uint32_t src = zpart + slot * 1024, dst = zcomb + slot * 512;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t o = (uint32_t)(lane + 32 * k) * 4;
                    sts_f32(dst + o, lds_f32(src + o) + lds_f32(src + 512 + o));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the bulk copy's reads
                __syncwarp();
                if (elect_one()) {
                    mbar_expect_tx(zfull + 8 * slot, nrank * 512u);
                    const uint32_t rdst = zrecv + (slot * kFsMaxCluster + (uint32_t)jg) * 512, rbar = zfull + 8 * slot;
                    for (uint32_t r = 0; r < nrank; ++r)
                        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(mapa_u32(rdst, r)), "r"(dst), "r"(512u), "r"(mapa_u32(rbar, r)) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== epilogue: warp w -> samples of lane tile 4 tq + (w & 3), capsules jg*8 + 4 (w >> 2) .. + 3 =====
        const int q = warp & 3, jh = warp >> 2;
        const int tile = tq * 4 + q;
        const bool tvalid = tile < p.nbt;
        const int j0 = jg * 8 + jh * 4;
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
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jh * 64);
        const uint32_t my_zpart = zpart + (uint32_t)(jh * 128 + q * 32 + lane) * 4;       // + slot * 1024
        const uint32_t my_zrecv = zrecv + (uint32_t)(q * 32 + lane) * 4;                   // + (slot * 8 + rank) * 512
        const uint32_t my_ering = ering + (uint32_t)((warp * 4) * 32 + lane) * 4;          // + eslot * 4096 + jj * 128
        const uint32_t my_coef = (uint32_t)kFsOperandBytes + (uint32_t)((q * 8 + jh * 4) * kLanes + lane) * 4;   // in a stage, + jj * 128
        int s1 = 0, s2 = 0;             // stage-ring positions of the P1 / P2 stage (BWD reads coefficients from the stage)
        int e1 = 0, e2 = 0;             // ering slots
        for (int it = 0; it < n_i + kFsSkew; ++it) {
            if (it < n_i) {
                // ---------------- P1(it): logits / dc, partial normaliser ----------------
                const int n = it, t = n & (kFsAccum - 1);
                mbar_wait_i(tmem_full + 8 * t, (n >> 2) & 1);
                tc_fence_after();
                float uh[64];
                tmem_ld64(lane_base + (uint32_t)(t * 128), uh);
                float z = 0.f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int d = 0; d < 16; d += 4) {
                        ffma2(d0, d1, uh[jj * 16 + d], uh[jj * 16 + d + 1], X[jj][d], X[jj][d + 1]);
                        ffma2(d2, d3, uh[jj * 16 + d + 2], uh[jj * 16 + d + 3], X[jj][d + 2], X[jj][d + 3]);
                    }
                    const float dot = (d0 + d1) + (d2 + d3);
                    float keep;
                    if (!BWD) {
                        keep = (j0 + jj < p.C) ? ex2_approx(fminf(dot, 120.f)) : 0.f;
                        z += keep;
                    } else {
                        keep = dot;
                        float c = 0.f;
                        if (tvalid && j0 + jj < p.C) c = lds_f32(stages + (uint32_t)s1 * kStageBytes + my_coef + jj * 128);
                        z = fmaf(c, dot, z);
                    }
                    sts_f32(my_ering + (uint32_t)e1 * 4096 + jj * 128, keep);
                }
                sts_f32(my_zpart + (uint32_t)(n & (kFsZSlots - 1)) * 1024, z);
                __syncwarp();
                if (lane == 0) mbar_arrive(zlocal + 8 * (n & (kFsZSlots - 1)));
                if (++s1 == ns) s1 = 0;
                if (++e1 == kFsSkew + 1) e1 = 0;
            }
            if (it >= kFsSkew) {
                // ---------------- P2(it - skew): normalise, store, accumulate ----------------
                const int m = it - kFsSkew, t = m & (kFsAccum - 1);
                const int i = i_begin + m;
                const uint32_t slot = (uint32_t)m & (kFsZSlots - 1);
                mbar_wait_i(zfull + 8 * slot, ((uint32_t)m >> 2) & 1);
                float Z = 0.f;
                for (int r = 0; r < p.JG; ++r) Z += lds_f32(my_zrecv + (slot * kFsMaxCluster + (uint32_t)r) * 512);
                float f[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float keep = lds_f32(my_ering + (uint32_t)e2 * 4096 + jj * 128);
                    if (!BWD) {
                        f[jj] = keep * rcp_approx(Z);
                    } else {
                        float c = 0.f, bp = 0.f;
                        if (tvalid && j0 + jj < p.C) {
                            c = lds_f32(stages + (uint32_t)s2 * kStageBytes + my_coef + jj * 128);
                            if (p.beta_in != nullptr) bp = lds_f32(stages + (uint32_t)s2 * kStageBytes + kFsCoefBytes + my_coef + jj * 128);
                        }
                        f[jj] = fmaf(c, keep - Z, bp);
                    }
                    if (tvalid && j0 + jj < p.C && p.coef_out != nullptr)
                        p.coef_out[(((size_t)tile * p.N + i) * p.C + j0 + jj) * kLanes + lane] = f[jj];
                }
                tc_fence_after();
                float uh[64];
                tmem_ld64(lane_base + (uint32_t)(t * 128), uh);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(tmem_empty + 8 * t);
                    if (BWD) mbar_arrive(smem_empty + 8 * s2);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int d = 0; d < 16; d += 2) ffma2(acc[jj][d], acc[jj][d + 1], f[jj], f[jj], uh[jj * 16 + d], uh[jj * 16 + d + 1]);
                if (++s2 == ns) s2 = 0;
                if (++e2 == kFsSkew + 1) e2 = 0;
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
    cluster_sync_all();                      // no CTA leaves while a peer's bulk copy may still read its shared memory
}

// couplings in the lane-tile layout [nbt][N][C][32] -> public [B][N][C] (tests and callers that ask for c_out)
__global__ void k_coef_public(const float* __restrict__ coef, float* __restrict__ c_pub, int B, int N, int C, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;
    const long b = (ti / N) * kLanes + lane;
    if (b >= B) return;
    const int i = (int)(ti % N);
    const float* src = coef + (size_t)ti * C * kLanes + lane;
    float* dst = c_pub + ((size_t)b * N + i) * C;
    for (int j = 0; j < C; ++j) dst[j] = src[(size_t)j * kLanes];
}

struct ClusterCap { std::atomic<int> n[kMaxDevices][kFsMaxCluster + 1]; };      // 0 = not queried yet
ClusterCap g_cap[2];

template <bool BWD>
int launch_t(const Plan& pl, const FusedParams& fp, int IS, cudaStream_t st) {
    auto kern = k_sweep_fused<BWD>;
    const size_t smem = (size_t)fp.ns * fs_stage_bytes(BWD) + kFsFixedBytes;
    CAPS_SET_SMEM(kern, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fp.JG, IS, cdiv(pl.nbt, 4));
    cfg.blockDim = dim3(kFsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = fp.JG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, fp));
    return 0;
}

int fs_stages(bool bwd) {
    int ns = bwd ? 7 : 8;
    while ((size_t)ns * fs_stage_bytes(bwd) + kFsFixedBytes > 227 * 1024) --ns;
    return ns;
}

}  // namespace

// D == 16 (after padding), 8 capsules per CTA, one cluster of ceil(C/8) <= 8 CTAs per (128 samples, i range)
bool fused_supported(const Plan& pl) {
    return pl.use_tc && pl.DP == 16 && pl.Reff > 1 && cdiv(pl.C, 8) <= kFsMaxCluster;
}

// clusters of `jg` CTAs the device can run at once (GPC granularity: not simply SMs / jg)
int fused_cluster_capacity(int jg, bool bwd) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || jg < 1 || jg > kFsMaxCluster) return 0;
    std::atomic<int>& slot = g_cap[bwd ? 1 : 0].n[dev][jg];
    int n = slot.load(std::memory_order_relaxed);
    if (n > 0) return n;
    const size_t smem = (size_t)fs_stages(bwd) * fs_stage_bytes(bwd) + kFsFixedBytes;
    const void* kern = bwd ? reinterpret_cast<const void*>(k_sweep_fused<true>) : reinterpret_cast<const void*>(k_sweep_fused<false>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(jg, 64, 1);
    cfg.blockDim = dim3(kFsThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = jg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = bwd ? cudaOccupancyMaxActiveClusters(&nc, k_sweep_fused<true>, &cfg)
                        : cudaOccupancyMaxActiveClusters(&nc, k_sweep_fused<false>, &cfg);
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
    fp.ns = fs_stages(bwd);
    if (cdiv(pl.N, fp.i_per_split) != IS) return fail(CAPS_E_BADARG, "fused sweep: %d splits do not tile N=%d", IS, pl.N);
    return bwd ? launch_t<true>(pl, fp, IS, st) : launch_t<false>(pl, fp, IS, st);
}

int launch_coef_public(const Plan& pl, const float* coef, float* c_pub, cudaStream_t st) {
    const long n = (long)pl.nbt * pl.N * kLanes;
    k_coef_public<<<cdiv(n, 256), 256, 0, st>>>(coef, c_pub, pl.B, pl.N, pl.C, pl.nbt);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace caps
