// caps_c1.cu -- the routing layer with ONE class capsule (the DarkCapsuleNet head: reference models.py:368-370, :398;
// routing batch = cells x images = 1568, N = 512 input capsules, K = 8, D = 5).
//
// With a single output capsule the softmax over j is identically 1, so every routing iteration returns
// squash(sum_i u_hat_i) whatever n_iter is (SURVEY 3.3 iv).  The whole layer is therefore one skinny GEMM per batch,
//     s[b, :] = u[b, (i,k)] . W[(i,k), :]          v = squash(s)                      (models.py:71, :76, :64-67)
// and its backward two more,
//     ds = squash'(s) dv        du[b, (i,k)] = ds[b, :] . W[(i,k), :]        dW[(i,k), :] = sum_b u[b, (i,k)] ds[b, :]
// all HBM-bound (u is read twice, du written once: 77 MB at the DarkCapsuleNet shape, ~12 us at the measured copy
// bandwidth).  The general kernels need 9 launches, 32-thread CTAs and the lane-tile re-layout of u for this shape
// (0.38 ms); the kernels here read u in place.  Two generations (tuning knob "c1v"): the first (below: k_c1_fwd / k_c1_bwd /
// k_c1_reduce, 0.080 ms per forward + loss + backward step at the DarkCapsuleNet shape) and the second (k_c1_fwd2 /
// k_c1_bwd2 further down, the default: 0.042 ms, 0.036 ms replayed from a CUDA graph), which the first cross-checks.
//   k_c1_fwd     one CTA per group of samples; W (d-major copy, <= 160 KB) staged in shared memory once per CTA;
//                thread <-> quads of (i,k); fixed-order block reduction; squash in the same kernel
//   k_c1_bwd     one CTA per 8 samples: ds (+ the margin-loss gradient) in the prologue; thread <-> quads of (i,k) with the
//                W rows and the dW accumulators in registers; du written, dW partial per CTA
//   k_c1_reduce  dW = sum of the per-CTA partials in fixed order (bit-reproducible)
#include "caps_internal.h"

#include <algorithm>

namespace caps {
namespace {

constexpr int kC1Threads = 256;
constexpr int kC1Tile = 8;            // samples per CTA of the backward kernel

// DP = 8 covers D <= 8.  W_t: [D][NK] (d-major copy, built in shared memory from the public [NK][D]).
template <int DMAX>
__global__ void __launch_bounds__(kC1Threads) k_c1_fwd(const float* __restrict__ u, const float* __restrict__ W,
                                                       float* __restrict__ v_pub, float* __restrict__ s_save,
                                                       int B, int NK, int D) {
    extern __shared__ __align__(16) float c1_smem[];
    float* Wt = c1_smem;                               // [D][NK]
    __shared__ float red[kC1Threads / 32][2][DMAX];
    for (int e = threadIdx.x; e < NK * D; e += kC1Threads) {
        const int ik = e / D, d = e - ik * D;
        Wt[d * NK + ik] = __ldg(W + e);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = NK >> 2;
    // two samples per pass: every shared-memory read of W serves both
    for (long b0 = (long)blockIdx.x * 2; b0 < B; b0 += (long)gridDim.x * 2) {
        const bool two = b0 + 1 < B;
        const float* ua = u + (size_t)b0 * NK;
        const float* ub = u + (size_t)(two ? b0 + 1 : b0) * NK;
        float acc[2][DMAX];
#pragma unroll
        for (int d = 0; d < DMAX; ++d) { acc[0][d] = 0.f; acc[1][d] = 0.f; }
        for (int q = threadIdx.x; q < nq; q += kC1Threads) {
            const float4 xa = ldg4(ua + 4 * q), xb = ldg4(ub + 4 * q);
#pragma unroll
            for (int d = 0; d < DMAX; ++d)
                if (d < D) {
                    const float4 w = *reinterpret_cast<const float4*>(Wt + d * NK + 4 * q);
                    acc[0][d] = fmaf(xa.x, w.x, acc[0][d]); acc[0][d] = fmaf(xa.y, w.y, acc[0][d]);
                    acc[0][d] = fmaf(xa.z, w.z, acc[0][d]); acc[0][d] = fmaf(xa.w, w.w, acc[0][d]);
                    acc[1][d] = fmaf(xb.x, w.x, acc[1][d]); acc[1][d] = fmaf(xb.y, w.y, acc[1][d]);
                    acc[1][d] = fmaf(xb.z, w.z, acc[1][d]); acc[1][d] = fmaf(xb.w, w.w, acc[1][d]);
                }
        }
        // fixed-order reduction: butterfly inside the warp, then the 8 warps in order
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int d = 0; d < DMAX; ++d) {
                float x = acc[h][d];
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (lane == 0) red[warp][h][d] = x;
            }
        __syncthreads();
        if (threadIdx.x < 2 && (threadIdx.x == 0 || two)) {
            const int h = threadIdx.x;
            float s[DMAX], vv[DMAX];
#pragma unroll
            for (int d = 0; d < DMAX; ++d) {
                float x = 0.f;
#pragma unroll
                for (int w = 0; w < kC1Threads / 32; ++w) x += red[w][h][d];
                s[d] = d < D ? x : 0.f;
            }
            squash_vec<DMAX>(s, vv);                   // reference models.py:64-67 (no epsilon: 0/0 stays NaN)
            const long b = b0 + h;
#pragma unroll
            for (int d = 0; d < DMAX; ++d) {
                s_save[b * DMAX + d] = s[d];
                if (d < D) v_pub[b * D + d] = vv[d];
            }
        }
        __syncthreads();
    }
}

// grid = ceil(B / kC1Tile).  dv = grad_v (+ margin gradient), ds = squash'(s) dv, then du and the CTA's dW partial.
template <int DMAX>
__global__ void __launch_bounds__(kC1Threads) k_c1_bwd(const float* __restrict__ u, const float* __restrict__ W,
                                                       const float* __restrict__ s_save, const float* __restrict__ grad_v,
                                                       const int64_t* __restrict__ y, float margin_scale,
                                                       const float* __restrict__ loss_grad, float* __restrict__ du,
                                                       float* __restrict__ dW_part, int B, int NK, int D) {
    __shared__ float ds_s[kC1Tile][DMAX];
    const long b_begin = (long)blockIdx.x * kC1Tile;
    const int nb = (int)min((long)kC1Tile, (long)B - b_begin);
    if (threadIdx.x < kC1Tile) {
        float s[DMAX], dv[DMAX], ds[DMAX];
#pragma unroll
        for (int d = 0; d < DMAX; ++d) { s[d] = 0.f; dv[d] = 0.f; ds[d] = 0.f; }
        if ((int)threadIdx.x < nb) {
            const long b = b_begin + threadIdx.x;
#pragma unroll
            for (int d = 0; d < DMAX; ++d) s[d] = s_save[b * DMAX + d];
            if (grad_v != nullptr) {
#pragma unroll
                for (int d = 0; d < DMAX; ++d)
                    if (d < D) dv[d] = __ldg(grad_v + b * D + d);
            }
            if (y != nullptr) {                         // margin-loss gradient on the single capsule (loss_fns.py:12-17)
                float vv[DMAX];
                squash_vec<DMAX>(s, vv);
                float m2 = 0.f;
#pragma unroll
                for (int d = 0; d < DMAX; ++d) m2 = fmaf(vv[d], vv[d], m2);
                const float m = sqrtf(m2);
                const bool hit = (y[b] == 0);
                const float lg = loss_grad != nullptr ? __ldg(loss_grad) : 1.f;
                const float dm = (hit ? -2.f * fmaxf(0.9f - m, 0.f) : fmaxf(m - 0.1f, 0.f)) * (margin_scale * lg);
                const float f = dm / m;
#pragma unroll
                for (int d = 0; d < DMAX; ++d) dv[d] = fmaf(f, vv[d], dv[d]);
            }
            squash_bwd_vec<DMAX>(s, dv, ds);
        }
#pragma unroll
        for (int d = 0; d < DMAX; ++d) ds_s[threadIdx.x][d] = (d < D) ? ds[d] : 0.f;
    }
    __syncthreads();
    const int nq = NK >> 2;
    for (int q = threadIdx.x; q < nq; q += kC1Threads) {
        float wr[4][DMAX], dw[4][DMAX];                 // W rows 4q..4q+3 and their gradient, in registers
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int d = 0; d < DMAX; ++d) {
                wr[e][d] = d < D ? __ldg(W + (size_t)(4 * q + e) * D + d) : 0.f;
                dw[e][d] = 0.f;
            }
        for (int t = 0; t < nb; ++t) {
            const size_t o = (size_t)(b_begin + t) * NK + 4 * q;
            const float4 x = ldg4(u + o);
            const float xe[4] = {x.x, x.y, x.z, x.w};
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < DMAX; ++d) {
                const float dsd = ds_s[t][d];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    g[e] = fmaf(dsd, wr[e][d], g[e]);
                    dw[e][d] = fmaf(xe[e], dsd, dw[e][d]);
                }
            }
            if (du != nullptr) st4(du + o, make_float4(g[0], g[1], g[2], g[3]));
        }
        float* dst = dW_part + ((size_t)blockIdx.x * NK + 4 * q) * D;
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int d = 0; d < DMAX; ++d)
                if (d < D) dst[e * D + d] = dw[e][d];
    }
}

// dW[e] = sum over the per-CTA partials, always in the same order: 8 slices of the partial range are summed
// concurrently (one per warp row of the block), then the 8 slice sums are added in slice order.
__global__ void __launch_bounds__(256) k_c1_reduce(const float* __restrict__ part, int nparts, float* __restrict__ dW, long n) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const long idx = (long)blockIdx.x * 32 + lane;
    const int per = (nparts + 7) / 8, p0 = slice * per, p1 = min(nparts, p0 + per);
    float x = 0.f;
    if (idx < n)
        for (int p = p0; p < p1; ++p) x += part[(size_t)p * n + idx];
    red[slice][lane] = x;
    __syncthreads();
    if (slice == 0 && idx < n) {
        float t = red[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += red[q][lane];
        dW[idx] = t;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Second generation (tuning knob "c1v" = 2, the default): the same three skinny GEMMs organised for memory-level
// parallelism, two launches instead of three.
//   k_c1_fwd2   CTA <-> a contiguous group of samples, the (i,k) range split over its 256 threads; a thread keeps the W
//               rows of one quad of (i,k) in registers and streams 8 samples' u against them (8 independent 16-byte
//               loads in flight per thread); one CTA-wide fixed-order reduction per 8 samples; squash in the same kernel.
//               No transposed W copy in shared memory (80 KB per CTA in k_c1_fwd).
//   k_c1_bwd2   CTA <-> 32 columns of (i,k) (8 quads, 128 contiguous bytes of every u row) for ALL samples: thread <->
//               (quad, sample slot), W rows and the dW accumulators of the quad in registers.  Every CTA recomputes ds
//               for the whole batch into shared memory (a few thousand squash' evaluations), so dW needs no per-CTA
//               partials and no reduction kernel: the 32 sample slots are summed in fixed order inside the CTA.
template <int D>
__device__ __forceinline__ void c1_load_w(const float* __restrict__ W, int q, float (&wr)[4][D]) {
    float flat[4 * D];                                     // rows 4q .. 4q+3 of W [NK][D]: 4 D contiguous floats
#pragma unroll
    for (int f = 0; f < D; ++f) {
        const float4 t = ldg4(W + (size_t)q * 4 * D + 4 * f);
        flat[4 * f] = t.x; flat[4 * f + 1] = t.y; flat[4 * f + 2] = t.z; flat[4 * f + 3] = t.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int d = 0; d < D; ++d) wr[e][d] = flat[e * D + d];
}

constexpr int kC1Chunk = 8;           // samples per reduction round of k_c1_fwd2

template <int D>
__global__ void __launch_bounds__(kC1Threads) k_c1_fwd2(const float* __restrict__ u, const float* __restrict__ W,
                                                        float* __restrict__ v_pub, float* __restrict__ s_save,
                                                        int B, int NK, int spc) {
    constexpr int SC = kC1Chunk, NV = SC * D, NG = (NV + 31) / 32;
    __shared__ float red[kC1Threads / 32][NG * 32];
    __shared__ float tot[NG * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = NK >> 2;
    const long b_begin = (long)blockIdx.x * spc, b_end = min((long)B, b_begin + spc);
    for (long b0 = b_begin; b0 < b_end; b0 += SC) {
        float acc[SC][D];
#pragma unroll
        for (int t = 0; t < SC; ++t)
#pragma unroll
            for (int d = 0; d < D; ++d) acc[t][d] = 0.f;
        for (int q = threadIdx.x; q < nq; q += kC1Threads) {
            float wr[4][D];
            c1_load_w<D>(W, q, wr);
            float4 x[SC];
#pragma unroll
            for (int t = 0; t < SC; ++t)
                x[t] = b0 + t < b_end ? ldg4(u + (size_t)(b0 + t) * NK + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < SC; ++t) {
                const float xe[4] = {x[t].x, x[t].y, x[t].z, x[t].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int d = 0; d < D; ++d) acc[t][d] = fmaf(xe[e], wr[e][d], acc[t][d]);
            }
        }
        // fixed-order reduction: transpose-reduce inside the warp (lane l ends up with the warp's total of value l),
        // then the 8 warps in order
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float vals[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const int f = g * 32 + l;
                vals[l] = f < NV ? acc[f / D][f % D] : 0.f;
            }
            red[warp][g * 32 + lane] = warp_transpose_reduce32(vals, lane);
        }
        __syncthreads();
        if (threadIdx.x < NG * 32) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < kC1Threads / 32; ++w) x += red[w][threadIdx.x];
            tot[threadIdx.x] = x;
        }
        __syncthreads();
        if (threadIdx.x < SC && b0 + threadIdx.x < b_end) {
            float sv[8], vv[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) sv[d] = d < D ? tot[threadIdx.x * D + d] : 0.f;
            squash_vec<8>(sv, vv);                         // reference models.py:64-67 (no epsilon: 0/0 stays NaN)
            const long b = b0 + threadIdx.x;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                s_save[b * 8 + d] = sv[d];
                if (d < D) v_pub[b * D + d] = vv[d];
            }
        }
        // red / tot are rewritten only after the next round's loads and FMAs, behind its first barrier... which the
        // readers of tot above have not necessarily passed: close the round
        __syncthreads();
    }
}

template <int D>
__global__ void __launch_bounds__(kC1Threads) k_c1_bwd2(const float* __restrict__ u, const float* __restrict__ W,
                                                        const float* __restrict__ s_save, const float* __restrict__ grad_v,
                                                        const int64_t* __restrict__ y, float margin_scale,
                                                        const float* __restrict__ loss_grad, float* __restrict__ du,
                                                        float* __restrict__ dW, int B, int NK, int CB) {
    extern __shared__ __align__(16) float c1_smem[];
    float* ds_s = c1_smem;                                 // [CB][D]
    __shared__ float red[kC1Threads / 32][8][4 * D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ql = threadIdx.x & 7, slot = threadIdx.x >> 3;            // quad within the CTA's 8, sample slot 0..31
    const int nq = NK >> 2;
    const int q = blockIdx.x * 8 + ql;
    const bool qvalid = q < nq;
    float wr[4][D], dw[4][D];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int d = 0; d < D; ++d) { wr[e][d] = 0.f; dw[e][d] = 0.f; }
    if (qvalid) c1_load_w<D>(W, q, wr);
    const float lg = loss_grad != nullptr ? __ldg(loss_grad) : 1.f;
    for (long c0 = 0; c0 < B; c0 += CB) {
        const int nb = (int)min((long)CB, (long)B - c0);
        __syncthreads();                                   // the previous chunk's ds has been consumed
        // dv = grad_v (+ margin gradient), ds = squash'(s) dv  -- the arithmetic of k_c1_bwd, for every sample of the chunk
        for (int t = threadIdx.x; t < nb; t += kC1Threads) {
            const long b = c0 + t;
            float sv[8], dv[8], ds[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) { sv[d] = s_save[b * 8 + d]; dv[d] = 0.f; }
            if (grad_v != nullptr) {
#pragma unroll
                for (int d = 0; d < D; ++d) dv[d] = __ldg(grad_v + b * D + d);
            }
            if (y != nullptr) {                            // margin-loss gradient on the single capsule (loss_fns.py:12-17)
                float vv[8];
                squash_vec<8>(sv, vv);
                float m2 = 0.f;
#pragma unroll
                for (int d = 0; d < 8; ++d) m2 = fmaf(vv[d], vv[d], m2);
                const float m = sqrtf(m2);
                const bool hit = (y[b] == 0);
                const float dm = (hit ? -2.f * fmaxf(0.9f - m, 0.f) : fmaxf(m - 0.1f, 0.f)) * (margin_scale * lg);
                const float f = dm / m;
#pragma unroll
                for (int d = 0; d < 8; ++d) dv[d] = fmaf(f, vv[d], dv[d]);
            }
            squash_bwd_vec<8>(sv, dv, ds);
#pragma unroll
            for (int d = 0; d < D; ++d) ds_s[t * D + d] = ds[d];
        }
        __syncthreads();
        if (qvalid) {
            constexpr int UN = 4;                          // samples in flight per thread
            for (int t0 = slot; t0 < nb; t0 += 32 * UN) {
                float4 x[UN];
#pragma unroll
                for (int k = 0; k < UN; ++k) {
                    const int t = t0 + 32 * k;
                    x[k] = t < nb ? ldg4(u + (size_t)(c0 + t) * NK + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < UN; ++k) {
                    const int t = t0 + 32 * k;
                    if (t < nb) {
                        const float xe[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
                        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const float dsd = ds_s[t * D + d];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                g[e] = fmaf(dsd, wr[e][d], g[e]);
                                dw[e][d] = fmaf(xe[e], dsd, dw[e][d]);
                            }
                        }
                        if (du != nullptr) st4(du + (size_t)(c0 + t) * NK + 4 * q, make_float4(g[0], g[1], g[2], g[3]));
                    }
                }
            }
        }
    }
    // dW of the CTA's 32 columns: the 32 sample slots in fixed order -- lanes l, l^8, l^16, l^24 hold the same quad
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float x = dw[e][d];
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if (lane < 8) red[warp][lane][e * D + d] = x;
        }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 8 * 4 * D; idx += kC1Threads) {
        const int qq = idx / (4 * D), f = idx - qq * (4 * D);
        float x = 0.f;
#pragma unroll
        for (int w = 0; w < kC1Threads / 32; ++w) x += red[w][qq][f];
        const int qg = blockIdx.x * 8 + qq;
        if (qg < nq) dW[(size_t)qg * 4 * D + f] = x;
    }
}

}  // namespace

int g_c1_version = 2;    // tuning knob "c1v": 2 = k_c1_fwd2 / k_c1_bwd2 (two launches), 1 = the first generation (three)

// one class capsule, K = 8, D <= 8, and W (N*8*D floats) fits next to nothing else in shared memory
bool c1_supported(int N, int C, int K, int D) { return C == 1 && K == 8 && D <= 8 && (size_t)N * K * D * 4 <= 160 * 1024; }
size_t c1_part_floats(int B, int N, int K, int D) { return (size_t)cdiv(B, kC1Tile) * N * K * D; }

#define CAPS_C1_BY_D(D, ...)                                                         \
    switch (D) {                                                                     \
        case 1: { constexpr int DT = 1; __VA_ARGS__; break; }                               \
        case 2: { constexpr int DT = 2; __VA_ARGS__; break; }                               \
        case 3: { constexpr int DT = 3; __VA_ARGS__; break; }                               \
        case 4: { constexpr int DT = 4; __VA_ARGS__; break; }                               \
        case 5: { constexpr int DT = 5; __VA_ARGS__; break; }                               \
        case 6: { constexpr int DT = 6; __VA_ARGS__; break; }                               \
        case 7: { constexpr int DT = 7; __VA_ARGS__; break; }                               \
        default: { constexpr int DT = 8; __VA_ARGS__; break; }                              \
    }

int launch_c1_forward(const float* u, const float* W, float* v, float* s_save, int B, int N, int K, int D, cudaStream_t st) {
    const int NK = N * K;
    if (g_c1_version == 2) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // whole reduction rounds of kC1Chunk samples per CTA, one wave of CTAs
        const int spc = cdiv(cdiv(B, sms), kC1Chunk) * kC1Chunk;
        CAPS_C1_BY_D(D, k_c1_fwd2<DT><<<cdiv(B, spc), kC1Threads, 0, st>>>(u, W, v, s_save, B, NK, spc));
        LAUNCH_CHECK();
        return 0;
    }
    const size_t smem = (size_t)NK * D * sizeof(float);
    CAPS_SET_SMEM(k_c1_fwd<8>, smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::max(1, std::min(cdiv(B, 2), sms));
    k_c1_fwd<8><<<grid, kC1Threads, smem, st>>>(u, W, v, s_save, B, NK, D);
    LAUNCH_CHECK();
    return 0;
}

int launch_c1_backward(const float* u, const float* W, const float* s_save, const float* grad_v, const int64_t* y,
                       float margin_scale, const float* loss_grad, float* du, float* dW, float* part,
                       int B, int N, int K, int D, cudaStream_t st, int* launches) {
    if (g_c1_version == 2) {
        const int NK2 = N * K;
        const int CB = std::min(B, 4096);                  // samples whose ds sit in shared memory at a time (<= 128 KB)
        const size_t smem = (size_t)CB * D * sizeof(float);
        CAPS_C1_BY_D(D, {
            auto kern = k_c1_bwd2<DT>;
            CAPS_SET_SMEM(kern, smem);
            kern<<<cdiv(NK2 >> 2, 8), kC1Threads, smem, st>>>(u, W, s_save, grad_v, y, margin_scale, loss_grad, du, dW, B, NK2, CB);
        });
        LAUNCH_CHECK();
        if (launches) *launches = 1;
        return 0;
    }
    const int NK = N * K, parts = cdiv(B, kC1Tile);
    k_c1_bwd<8><<<parts, kC1Threads, 0, st>>>(u, W, s_save, grad_v, y, margin_scale, loss_grad, du, part, B, NK, D);
    LAUNCH_CHECK();
    const long n = (long)NK * D;
    k_c1_reduce<<<cdiv(n, 32), 256, 0, st>>>(part, parts, dW, n);
    LAUNCH_CHECK();
    if (launches) *launches = 2;
    return 0;
}

}  // namespace caps
