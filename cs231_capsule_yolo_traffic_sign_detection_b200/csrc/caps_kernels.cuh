// caps_kernels.cuh -- sm_100a kernels of the capsule dynamic-routing path.
//
// What they replace: the caps->caps branch of the reference CapsuleLayer (reference
// models.py:70-79), its autograd backward, and the margin-loss gradient (loss_fns.py:12-17,23).
//
// Design (DESIGN.md has the long form):
//   * u_hat[b,i,j,:] = u[b,i,:] . W[i,j,:,:] is NEVER materialised (3.57 MB per sample at the
//     CapsuleNet shape).  Every pass recomputes it on the fp32 FMA pipe from u (registers) and a
//     W[i, j-group] slab staged in shared memory with cp.async and a multi-stage ring.
//   * lane <-> sample: a warp owns one output capsule j and 32*SPT samples, and walks a range of
//     input capsules i.  The W operand is therefore warp-uniform (shared-memory broadcast) and
//     the per-sample state a thread needs ([D] accumulator or [D] probe vector) lives in
//     registers.  Routing logits use the identity b^r_ij = u_hat_ij . (v^0+..+v^{r-1}).
//   * every routing iteration is: L-pass (logits / dc = u_hat . X_j), a per-(b,i) softmax over j,
//     A-pass (s_j or dv_j = sum_i coef_ij u_hat_ij), a per-(b,j) squash.  The [B,N,C] coefficient
//     arrays are the only per-(b,i,j) data that touch HBM (4 bytes per 2*K*D+... flops).
//   * all cross-CTA reductions go through fixed-order partial buffers: no float atomics anywhere,
//     results are bit-reproducible.
//
// Memory layouts ("lane tiles": 32 consecutive samples are the fastest axis, so that lane <-> b
// accesses are 128-byte coalesced; nbt = ceil(B/32); padded samples carry u = 0):
//   ut    [nbt][N][K/4][32][4]      transposed copy of u
//   coef  [nbt][N][C][32]           logits / couplings c^r / beta^r / dc scratch
//   X     [nbt][C][DP/4][32][4]     per-(b,j) vectors: s^r, v^r, sum of v, ds^r   (DP = D padded to 4)
//   part  [IS][nbt][C][DP/4][32][4] A-pass partial sums, one slot per split of the i range
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace caps {

constexpr int kLanes = 32;
constexpr int kPassIC = 4;        // input capsules per pipeline stage
constexpr int kPassStages = 3;    // cp.async ring depth
constexpr int kGradDuBatch = 4;   // input capsules per du cross-warp reduction round
constexpr int kGradMaxM = 9;      // 2*R-1 terms, R <= 5

enum PassMode { kModeAUniform = 0, kModeA = 1, kModeL = 2 };

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// squash (reference models.py:64-67): same operation order as the reference:
//   scale = n2 / (1 + n2);  v = scale * s / sqrt(n2)
template <int DP>
__device__ __forceinline__ void squash_vec(const float (&s)[DP], float (&v)[DP]) {
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) n2 = fmaf(s[d], s[d], n2);
    const float scale = n2 / (1.f + n2);
    const float rn = sqrtf(n2);
#pragma unroll
    for (int d = 0; d < DP; ++d) v[d] = scale * s[d] / rn;
}
// ds = dv * n/(1+n2) + s * (s.dv) * (1-n2) / (n (1+n2)^2)
template <int DP>
__device__ __forceinline__ void squash_bwd_vec(const float (&s)[DP], const float (&dv)[DP], float (&ds)[DP]) {
    float n2 = 0.f, sdv = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) { n2 = fmaf(s[d], s[d], n2); sdv = fmaf(s[d], dv[d], sdv); }
    const float n = sqrtf(n2);
    const float a = n / (1.f + n2);
    const float b = sdv * (1.f - n2) / (n * (1.f + n2) * (1.f + n2));
#pragma unroll
    for (int d = 0; d < DP; ++d) ds[d] = fmaf(a, dv[d], b * s[d]);
}

// ---------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------
// u [B][N][K] -> ut [nbt][N][K/4][32][4]; samples >= B are zero.
template <int K>
__global__ void k_prep_u(const float* __restrict__ u, float* __restrict__ ut, int B, int N, int nbt) {
    constexpr int K4 = K / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;
    const int i = (int)(ti % N);
    const int bt = (int)(ti / N);
    const long b = (long)bt * kLanes + lane;
#pragma unroll
    for (int kq = 0; kq < K4; ++kq) {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) val = ldg4(u + (b * N + i) * K + kq * 4);
        st4(ut + ((((long)bt * N + i) * K4 + kq) * kLanes + lane) * 4, val);
    }
}

// W [rows][D] -> Wp [rows][DP], zero padded (only used when D % 4 != 0 or D != DP).
static __global__ void k_pad_w(const float* __restrict__ W, float* __restrict__ Wp, long rows, int D, int DP) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * DP) return;
    const int d = (int)(idx % DP);
    const long r = idx / DP;
    Wp[idx] = d < D ? W[r * D + d] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// the pass kernel: one sweep over (b, i, j) with u_hat recomputed on the fly
// ---------------------------------------------------------------------------------------------
struct PassParams {
    const float* ut;     // [nbt][N][K4][32][4]
    const float* W;      // [N][C][K][DP]
    const float* coef;   // kModeA: [nbt][N][C][32]
    const float* X;      // kModeL: [nbt][C][D4][32][4]
    float* out;          // kModeL: [nbt][N][C][32];  kModeA*: part [IS][nbt][C][D4][32][4]
    int N, C, nbt, i_per_split;
};

// grid = (IS, ceil(C/JW), ceil(nbt/SPT));  block = 32*JW threads;  warp w <-> capsule j0+w;
// thread <-> samples {(tg*SPT+s)*32 + lane}.
//   kModeAUniform : part[is][b][j][:] = sum_{i in split} u_hat[b,i,j,:]
//   kModeA        : part[is][b][j][:] = sum_{i in split} coef[b,i,j] * u_hat[b,i,j,:]
//   kModeL        : out[b][i][j]      = u_hat[b,i,j,:] . X[b,j,:]
template <int K, int DP, int SPT, int JW, int MODE>
__global__ void __launch_bounds__(32 * JW) k_pass(PassParams p) {
    constexpr int K4 = K / 4, D4 = DP / 4, ROW = K * DP;
    constexpr int IC = kPassIC, ST = kPassStages, NT = 32 * JW;
    extern __shared__ __align__(16) float smem[];            // [ST][IC][JW][ROW]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.y * JW;
    const int j = j0 + warp;
    const bool jvalid = j < p.C;
    const int njv = min(JW, p.C - j0);
    const int i_begin = blockIdx.x * p.i_per_split;
    const int i_end = min(p.N, i_begin + p.i_per_split);
    const int n_chunks = i_end > i_begin ? (i_end - i_begin + IC - 1) / IC : 0;

    int tile[SPT];
    bool tvalid[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int t = blockIdx.z * SPT + s;
        tvalid[s] = t < p.nbt;
        tile[s] = tvalid[s] ? t : p.nbt - 1;                // clamp: loads stay in bounds, stores are skipped
    }

    auto issue = [&](int c) {
        const int stage = c % ST;
#pragma unroll
        for (int ic = 0; ic < IC; ++ic) {
            const int i = i_begin + c * IC + ic;
            if (i < i_end) {
                const float* src = p.W + ((size_t)i * p.C + j0) * ROW;
                float* dst = smem + (size_t)((stage * IC + ic) * JW) * ROW;
                const int pieces = njv * (ROW / 4);
                for (int t = threadIdx.x; t < pieces; t += NT) cp_async16(dst + 4 * t, src + 4 * t);
            }
        }
    };

    float acc[SPT][DP];     // kModeA*: running sums;  kModeL: the probe vectors X[b,j,:]
#pragma unroll
    for (int s = 0; s < SPT; ++s)
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[s][d] = 0.f;
    if (MODE == kModeL && jvalid) {
#pragma unroll
        for (int s = 0; s < SPT; ++s)
#pragma unroll
            for (int dq = 0; dq < D4; ++dq) {
                const float4 x = ldg4(p.X + ((((size_t)tile[s] * p.C + j) * D4 + dq) * kLanes + lane) * 4);
                acc[s][dq * 4 + 0] = x.x; acc[s][dq * 4 + 1] = x.y; acc[s][dq * 4 + 2] = x.z; acc[s][dq * 4 + 3] = x.w;
            }
    }

    // register double buffer for the per-(b,i) operands
    float4 un[SPT][K4];
    float cn[SPT];
    auto load_i = [&](int i) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
#pragma unroll
            for (int kq = 0; kq < K4; ++kq)
                un[s][kq] = ldg4(p.ut + ((((size_t)tile[s] * p.N + i) * K4 + kq) * kLanes + lane) * 4);
            if (MODE == kModeA) cn[s] = __ldg(p.coef + (((size_t)tile[s] * p.N + i) * p.C + j) * kLanes + lane);
        }
    };

#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < n_chunks) issue(s);
        cp_async_commit();
    }
    if (jvalid && n_chunks > 0) load_i(i_begin);

    for (int c = 0; c < n_chunks; ++c) {
        cp_async_wait<ST - 2>();
        __syncthreads();                                     // chunk c landed; everyone left chunk c-1
        if (c + ST - 1 < n_chunks) issue(c + ST - 1);
        cp_async_commit();
        if (!jvalid) continue;
        const int stage = c % ST;
#pragma unroll 1
        for (int ic = 0; ic < IC; ++ic) {
            const int i = i_begin + c * IC + ic;
            if (i >= i_end) break;
            float a[SPT][K];
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const float f = (MODE == kModeA) ? cn[s] : 1.f;
#pragma unroll
                for (int kq = 0; kq < K4; ++kq) {
                    a[s][kq * 4 + 0] = un[s][kq].x * f; a[s][kq * 4 + 1] = un[s][kq].y * f;
                    a[s][kq * 4 + 2] = un[s][kq].z * f; a[s][kq * 4 + 3] = un[s][kq].w * f;
                }
            }
            if (i + 1 < i_end) load_i(i + 1);                // prefetch next capsule's operands
            const float* wrow = smem + (size_t)((stage * IC + ic) * JW + warp) * ROW;
            if (MODE == kModeL) {
                float uh[SPT][DP];
#pragma unroll
                for (int s = 0; s < SPT; ++s)
#pragma unroll
                    for (int d = 0; d < DP; ++d) uh[s][d] = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int dq = 0; dq < D4; ++dq) {
                        const float4 w = *reinterpret_cast<const float4*>(wrow + k * DP + dq * 4);
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            uh[s][dq * 4 + 0] = fmaf(a[s][k], w.x, uh[s][dq * 4 + 0]);
                            uh[s][dq * 4 + 1] = fmaf(a[s][k], w.y, uh[s][dq * 4 + 1]);
                            uh[s][dq * 4 + 2] = fmaf(a[s][k], w.z, uh[s][dq * 4 + 2]);
                            uh[s][dq * 4 + 3] = fmaf(a[s][k], w.w, uh[s][dq * 4 + 3]);
                        }
                    }
#pragma unroll
                for (int s = 0; s < SPT; ++s) {
                    float dot = 0.f;
#pragma unroll
                    for (int d = 0; d < DP; ++d) dot = fmaf(uh[s][d], acc[s][d], dot);
                    if (tvalid[s]) p.out[(((size_t)tile[s] * p.N + i) * p.C + j) * kLanes + lane] = dot;
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int dq = 0; dq < D4; ++dq) {
                        const float4 w = *reinterpret_cast<const float4*>(wrow + k * DP + dq * 4);
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            acc[s][dq * 4 + 0] = fmaf(a[s][k], w.x, acc[s][dq * 4 + 0]);
                            acc[s][dq * 4 + 1] = fmaf(a[s][k], w.y, acc[s][dq * 4 + 1]);
                            acc[s][dq * 4 + 2] = fmaf(a[s][k], w.z, acc[s][dq * 4 + 2]);
                            acc[s][dq * 4 + 3] = fmaf(a[s][k], w.w, acc[s][dq * 4 + 3]);
                        }
                    }
            }
        }
    }
    cp_async_wait<0>();

    if (MODE != kModeL && jvalid) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            if (!tvalid[s]) continue;
#pragma unroll
            for (int dq = 0; dq < D4; ++dq)
                st4(p.out + (((((size_t)blockIdx.x * p.nbt + tile[s]) * p.C + j) * D4 + dq) * kLanes + lane) * 4,
                    make_float4(acc[s][dq * 4 + 0], acc[s][dq * 4 + 1], acc[s][dq * 4 + 2], acc[s][dq * 4 + 3]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-(b,j) kernels: reduce the A-pass partials, squash / squash backward
// ---------------------------------------------------------------------------------------------
// s = scale * sum_is part[is];  v = squash(s);  Vsum (+)= v;  optional public copy v_pub [B][C][D].
template <int DP>
__global__ void k_squash(const float* __restrict__ part, int IS, size_t xsize, float scale,
                         float* __restrict__ s_out, float* __restrict__ v_out, float* __restrict__ vsum,
                         int accumulate, float* __restrict__ v_pub, int B, int C, int D, int nbt) {
    constexpr int D4 = DP / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * C * kLanes) return;
    const int lane = (int)(idx & 31);
    const long tj = idx >> 5;                               // bt*C + j
    const int j = (int)(tj % C);
    const long b = (tj / C) * kLanes + lane;
    const size_t base = ((size_t)tj * D4) * kLanes * 4 + lane * 4;
    float s[DP], v[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) s[d] = 0.f;
#pragma unroll 4
    for (int is = 0; is < IS; ++is)   // unrolled: the partial loads of four splits are in flight together (same order of additions)
#pragma unroll
        for (int dq = 0; dq < D4; ++dq) {
            const float4 x = ldg4(part + is * xsize + base + (size_t)dq * kLanes * 4);
            s[dq * 4 + 0] += x.x; s[dq * 4 + 1] += x.y; s[dq * 4 + 2] += x.z; s[dq * 4 + 3] += x.w;
        }
#pragma unroll
    for (int d = 0; d < DP; ++d) s[d] *= scale;
    if (b < B) {
        squash_vec<DP>(s, v);
    } else {
#pragma unroll
        for (int d = 0; d < DP; ++d) v[d] = 0.f;            // padded samples: keep everything finite
    }
#pragma unroll
    for (int dq = 0; dq < D4; ++dq) {
        const size_t o = base + (size_t)dq * kLanes * 4;
        st4(s_out + o, make_float4(s[dq * 4], s[dq * 4 + 1], s[dq * 4 + 2], s[dq * 4 + 3]));
        st4(v_out + o, make_float4(v[dq * 4], v[dq * 4 + 1], v[dq * 4 + 2], v[dq * 4 + 3]));
        float4 t = make_float4(v[dq * 4], v[dq * 4 + 1], v[dq * 4 + 2], v[dq * 4 + 3]);
        if (accumulate) {
            const float4 o4 = *reinterpret_cast<const float4*>(vsum + o);
            t.x += o4.x; t.y += o4.y; t.z += o4.z; t.w += o4.w;
        }
        st4(vsum + o, t);
    }
    if (v_pub != nullptr && b < B) {
        float* dst = v_pub + ((size_t)b * C + j) * D;
#pragma unroll
        for (int d = 0; d < DP; ++d)
            if (d < D) dst[d] = v[d];
    }
}

// ds = squash_bwd(s, dv) with
//   dv = sum_is part[is]                                   (inner iterations), or
//   dv = grad_v[b][j][:] + margin_scale * d margin / d v   (top: part == nullptr)
// ds_out = out_scale * ds.
template <int DP>
__global__ void k_dsquash(const float* __restrict__ part, int IS, size_t xsize,
                          const float* __restrict__ grad_v, const int64_t* __restrict__ y, float margin_scale,
                          const float* __restrict__ loss_grad, const float* __restrict__ v_last, const float* __restrict__ s_in,
                          float* __restrict__ ds_out, float out_scale, int B, int C, int D, int nbt) {
    constexpr int D4 = DP / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * C * kLanes) return;
    const int lane = (int)(idx & 31);
    const long tj = idx >> 5;
    const int j = (int)(tj % C);
    const long b = (tj / C) * kLanes + lane;
    const size_t base = ((size_t)tj * D4) * kLanes * 4 + lane * 4;
    float s[DP], dv[DP], ds[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) dv[d] = 0.f;
#pragma unroll
    for (int dq = 0; dq < D4; ++dq) {
        const float4 x = ldg4(s_in + base + (size_t)dq * kLanes * 4);
        s[dq * 4 + 0] = x.x; s[dq * 4 + 1] = x.y; s[dq * 4 + 2] = x.z; s[dq * 4 + 3] = x.w;
    }
    if (part != nullptr) {
#pragma unroll 4
        for (int is = 0; is < IS; ++is)
#pragma unroll
            for (int dq = 0; dq < D4; ++dq) {
                const float4 x = ldg4(part + is * xsize + base + (size_t)dq * kLanes * 4);
                dv[dq * 4 + 0] += x.x; dv[dq * 4 + 1] += x.y; dv[dq * 4 + 2] += x.z; dv[dq * 4 + 3] += x.w;
            }
    } else if (b < B) {
        if (grad_v != nullptr) {
            const float* g = grad_v + ((size_t)b * C + j) * D;
#pragma unroll
            for (int d = 0; d < DP; ++d)
                if (d < D) dv[d] = __ldg(g + d);
        }
        if (y != nullptr) {
            float v[DP];
            float m2 = 0.f;
#pragma unroll
            for (int dq = 0; dq < D4; ++dq) {
                const float4 x = ldg4(v_last + base + (size_t)dq * kLanes * 4);
                v[dq * 4 + 0] = x.x; v[dq * 4 + 1] = x.y; v[dq * 4 + 2] = x.z; v[dq * 4 + 3] = x.w;
            }
#pragma unroll
            for (int d = 0; d < DP; ++d) m2 = fmaf(v[d], v[d], m2);
            const float m = sqrtf(m2);                       // reference models.py:117
            const bool hit = (y[b] == (int64_t)j);
            // reference loss_fns.py:12-17: T*relu(0.9-m)^2 + 0.5*(1-T)*relu(m-0.1)^2
            const float lg = loss_grad != nullptr ? __ldg(loss_grad) : 1.f;
            const float dm = (hit ? -2.f * fmaxf(0.9f - m, 0.f) : fmaxf(m - 0.1f, 0.f)) * (margin_scale * lg);
            const float f = dm / m;
#pragma unroll
            for (int d = 0; d < DP; ++d) dv[d] = fmaf(f, v[d], dv[d]);
        }
    }
    if (b < B) {
        squash_bwd_vec<DP>(s, dv, ds);
#pragma unroll
        for (int d = 0; d < DP; ++d) ds[d] *= out_scale;     // 1, or the uniform coupling 1/C of iteration 0 (see caps_route_backward)
    } else {
#pragma unroll
        for (int d = 0; d < DP; ++d) ds[d] = 0.f;
    }
#pragma unroll
    for (int dq = 0; dq < D4; ++dq)
        st4(ds_out + base + (size_t)dq * kLanes * 4, make_float4(ds[dq * 4], ds[dq * 4 + 1], ds[dq * 4 + 2], ds[dq * 4 + 3]));
}

// ---------------------------------------------------------------------------------------------
// per-(b,i) kernels over the C axis
// ---------------------------------------------------------------------------------------------
// in place: logits -> softmax over j (reference models.py:75; max-subtracted like ATen).
// c_pub (nullable): [B][N][C] public copy.
static __global__ void k_softmax(float* __restrict__ coef, float* __restrict__ c_pub, int B, int N, int C, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;                               // bt*N + i
    float* p = coef + (size_t)ti * C * kLanes + lane;
    float mx = p[0];
    for (int j = 1; j < C; ++j) mx = fmaxf(mx, p[(size_t)j * kLanes]);
    float z = 0.f;
    for (int j = 0; j < C; ++j) z += expf(p[(size_t)j * kLanes] - mx);
    const long b = (ti / N) * kLanes + lane;
    const int i = (int)(ti % N);
    float* pub = (c_pub != nullptr && b < B) ? c_pub + ((size_t)b * N + i) * C : nullptr;
    for (int j = 0; j < C; ++j) {
        const float c = expf(p[(size_t)j * kLanes] - mx) / z;
        p[(size_t)j * kLanes] = c;
        if (pub) pub[j] = c;
    }
}

// beta_out = beta_prev + c * (dc - sum_j c*dc)      (softmax backward + identity carry of db)
static __global__ void k_softmax_bwd(const float* __restrict__ c, const float* __restrict__ dc,
                              const float* __restrict__ beta_prev, float* __restrict__ beta_out,
                              int N, int C, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const size_t o = (size_t)(idx >> 5) * C * kLanes + lane;
    float t = 0.f;
    for (int j = 0; j < C; ++j) t = fmaf(c[o + (size_t)j * kLanes], dc[o + (size_t)j * kLanes], t);
    for (int j = 0; j < C; ++j) {
        const size_t q = o + (size_t)j * kLanes;
        float bnew = c[q] * (dc[q] - t);
        if (beta_prev != nullptr) bnew += beta_prev[q];
        beta_out[q] = bnew;
    }
}

// register-resident variants for C <= CMAX: one read + one write of the [.,C] row per thread
template <int CMAX>
__global__ void k_softmax_reg(float* __restrict__ coef, float* __restrict__ c_pub, int B, int N, int C, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;
    float* p = coef + (size_t)ti * C * kLanes + lane;
    float v[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) v[j] = j < C ? p[(size_t)j * kLanes] : -INFINITY;
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < CMAX; ++j) mx = fmaxf(mx, v[j]);
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) {
        v[j] = j < C ? __expf(v[j] - mx) : 0.f;       // ex2.approx: 2^-22 relative, far inside the 1e-5 budget
        z += v[j];
    }
    const long b = (ti / N) * kLanes + lane;
    const int i = (int)(ti % N);
    float* pub = (c_pub != nullptr && b < B) ? c_pub + ((size_t)b * N + i) * C : nullptr;
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < C) {
            const float c = v[j] / z;
            p[(size_t)j * kLanes] = c;
            if (pub) pub[j] = c;
        }
}

template <int CMAX>
__global__ void k_softmax_bwd_reg(const float* __restrict__ c, const float* __restrict__ dc,
                                  const float* __restrict__ beta_prev, float* __restrict__ beta_out,
                                  int N, int C, int nbt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const size_t o = (size_t)(idx >> 5) * C * kLanes + lane;
    float cv[CMAX], dv[CMAX];
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) {
        cv[j] = j < C ? c[o + (size_t)j * kLanes] : 0.f;
        dv[j] = j < C ? dc[o + (size_t)j * kLanes] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < CMAX; ++j) t = fmaf(cv[j], dv[j], t);
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < C) {
            const size_t q = o + (size_t)j * kLanes;
            float bnew = cv[j] * (dv[j] - t);
            if (beta_prev != nullptr) bnew += beta_prev[q];
            beta_out[q] = bnew;
        }
}

// staged variant for 16 < C <= CMAX: one warp per (lane tile, i) block.  The block's c and dc rows are two
// contiguous C x 128-byte runs, fetched by two cp.async.bulk copies into the warp's shared slice while the lanes
// load the beta carry straight into registers: every byte crosses HBM once (the three-pass kernel above re-reads
// c and dc: 8.5 GB instead of 6.5 per launch) with three arrays in flight at once and no register blow-up.
template <int CMAX>
__global__ void __launch_bounds__(128) k_softmax_bwd_staged(const float* __restrict__ c, const float* __restrict__ dc,
                                                            const float* __restrict__ beta_prev, float* __restrict__ beta_out,
                                                            int C, long nblocks) {
    extern __shared__ __align__(16) float sm_sb[];            // [4 warps][2][C][32]
    __shared__ __align__(8) unsigned long long bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long blk = (long)blockIdx.x * 4 + warp;
    if (blk >= nblocks) return;
    float* sc = sm_sb + (size_t)warp * 2 * C * kLanes;
    float* sd = sc + (size_t)C * kLanes;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
    const size_t o = (size_t)blk * C * kLanes;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)C * kLanes * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sc)), "l"(c + o), "r"(bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sd)), "l"(dc + o), "r"(bytes), "r"(bar) : "memory");
    }
    float bp[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) bp[j] = (beta_prev != nullptr && j < C) ? beta_prev[o + (size_t)j * kLanes + lane] : 0.f;
    __syncwarp();                                              // the barrier is initialised before anyone polls it
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar) : "memory");
    }
    float t = 0.f;
#pragma unroll 8
    for (int j = 0; j < C; ++j) t = fmaf(sc[j * kLanes + lane], sd[j * kLanes + lane], t);
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < C) beta_out[o + (size_t)j * kLanes + lane] = fmaf(sc[j * kLanes + lane], sd[j * kLanes + lane] - t, bp[j]);
}

static __global__ void k_fill(float* __restrict__ p, float val, long n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) p[idx] = val;
}

static __global__ void k_add_inplace(float* __restrict__ acc, const float* __restrict__ x, long n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) acc[idx] += x[idx];
}

// ---------------------------------------------------------------------------------------------
// final backward pass: G_bij = sum_m alpha^m_bij X^m_bj ;  dW_ij = sum_b u_bi (x) G_bij ;
//                      du_bi = sum_j W_ij G_bij
// ---------------------------------------------------------------------------------------------
struct GradParams {
    const float* ut;
    const float* W;                    // [N][C][K][DP]
    float* dW;                         // public [N][C][K][D]
    float* du_part;                    // [JG][nbt][N][K4][32][4]
    const float* coef[kGradMaxM];      // [nbt][N][C][32] or nullptr -> constant cconst[m]
    const float* X[kGradMaxM];         // [nbt][C][D4][32][4]
    float cconst[kGradMaxM];
    int N, C, D, nbt;
    int DP;                            // padded D: row length of W and of the X arrays (k_grad_mma)
    int CS;                            // capsules per coefficient row (C, or C rounded up to 8 when the fused sweep wrote the arrays)
};

// transpose-reduce: every lane holds 32 values; lane l returns sum over lanes of vals[l].
__device__ __forceinline__ float warp_transpose_reduce32(float (&vals)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int t = 0; t < half; ++t) {
            const float send = up ? vals[t] : vals[t + half];
            const float keep = up ? vals[t + half] : vals[t];
            vals[t] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return vals[0];
}

// grid = (ceil(N/IT), ceil(C/JW)); block = 32*JW; the CTA owns dW[i-tile][j-group] completely
// (loops over all samples), so dW needs no atomics; du gets one partial per j-group.
template <int K, int DP, int JW, int IT, int M, bool XREG>
__global__ void __launch_bounds__(32 * JW, 1) k_grad(GradParams p) {
    constexpr int K4 = K / 4, D4 = DP / 4, ROW = K * DP, NT = 32 * JW, DUB = kGradDuBatch;
    static_assert(ROW % 32 == 0, "K*DP must be a multiple of 32");
    static_assert(IT % DUB == 0, "IT must be a multiple of the du batch");
    extern __shared__ __align__(16) float smem[];
    float* Wsm = smem;                                      // [IT][JW][ROW]
    float* dWsm = Wsm + IT * JW * ROW;                      // [IT][JW][ROW]
    float* dusm = dWsm + IT * JW * ROW;                     // [JW][DUB][K4][32][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * IT;
    const int ni = min(IT, p.N - i0);
    const int j0 = blockIdx.y * JW;
    const int j = j0 + warp;
    const bool jvalid = j < p.C;
    const int njv = min(JW, p.C - j0);

    for (int t = threadIdx.x; t < IT * JW * (ROW / 4); t += NT) {
        const int r4 = t % (ROW / 4);
        const int w = (t / (ROW / 4)) % JW;
        const int il = t / ((ROW / 4) * JW);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (il < ni && w < njv) val = ldg4(p.W + ((size_t)(i0 + il) * p.C + j0 + w) * ROW + r4 * 4);
        st4(Wsm + (size_t)(il * JW + w) * ROW + r4 * 4, val);
        st4(dWsm + (size_t)(il * JW + w) * ROW + r4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncthreads();

    for (int tile = 0; tile < p.nbt; ++tile) {
        float xr[XREG ? M : 1][DP];
        if (XREG && jvalid) {
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int dq = 0; dq < D4; ++dq) {
                    const float4 x = ldg4(p.X[m] + ((((size_t)tile * p.C + j) * D4 + dq) * kLanes + lane) * 4);
                    xr[m][dq * 4 + 0] = x.x; xr[m][dq * 4 + 1] = x.y; xr[m][dq * 4 + 2] = x.z; xr[m][dq * 4 + 3] = x.w;
                }
        }
        for (int ib = 0; ib < IT; ib += DUB) {
#pragma unroll 1
            for (int ii = 0; ii < DUB; ++ii) {
                const int il = ib + ii;
                float du[K];
#pragma unroll
                for (int k = 0; k < K; ++k) du[k] = 0.f;
                if (jvalid && il < ni) {
                    const int i = i0 + il;
                    float ua[K];
#pragma unroll
                    for (int kq = 0; kq < K4; ++kq) {
                        const float4 x = ldg4(p.ut + ((((size_t)tile * p.N + i) * K4 + kq) * kLanes + lane) * 4);
                        ua[kq * 4 + 0] = x.x; ua[kq * 4 + 1] = x.y; ua[kq * 4 + 2] = x.z; ua[kq * 4 + 3] = x.w;
                    }
                    float G[DP];
#pragma unroll
                    for (int d = 0; d < DP; ++d) G[d] = 0.f;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float al = p.coef[m] != nullptr
                                             ? __ldg(p.coef[m] + (((size_t)tile * p.N + i) * p.CS + j) * kLanes + lane)
                                             : p.cconst[m];
                        if (XREG) {
#pragma unroll
                            for (int d = 0; d < DP; ++d) G[d] = fmaf(al, xr[m][d], G[d]);
                        } else {
#pragma unroll
                            for (int dq = 0; dq < D4; ++dq) {
                                const float4 x = ldg4(p.X[m] + ((((size_t)tile * p.C + j) * D4 + dq) * kLanes + lane) * 4);
                                G[dq * 4 + 0] = fmaf(al, x.x, G[dq * 4 + 0]); G[dq * 4 + 1] = fmaf(al, x.y, G[dq * 4 + 1]);
                                G[dq * 4 + 2] = fmaf(al, x.z, G[dq * 4 + 2]); G[dq * 4 + 3] = fmaf(al, x.w, G[dq * 4 + 3]);
                            }
                        }
                    }
                    const float* wrow = Wsm + (size_t)(il * JW + warp) * ROW;
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int dq = 0; dq < D4; ++dq) {
                            const float4 w = *reinterpret_cast<const float4*>(wrow + k * DP + dq * 4);
                            du[k] = fmaf(w.x, G[dq * 4 + 0], du[k]); du[k] = fmaf(w.y, G[dq * 4 + 1], du[k]);
                            du[k] = fmaf(w.z, G[dq * 4 + 2], du[k]); du[k] = fmaf(w.w, G[dq * 4 + 3], du[k]);
                        }
                    float* dwrow = dWsm + (size_t)(il * JW + warp) * ROW;
#pragma unroll
                    for (int q = 0; q < ROW / 32; ++q) {
                        float vals[32];
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const int e = q * 32 + t;
                            vals[t] = ua[e / DP] * G[e % DP];
                        }
                        const float r = warp_transpose_reduce32(vals, lane);
                        dwrow[q * 32 + lane] += r;          // lane-private address: no hazard across tiles
                    }
                }
#pragma unroll
                for (int kq = 0; kq < K4; ++kq)
                    st4(dusm + ((size_t)((warp * DUB + ii) * K4 + kq) * kLanes + lane) * 4,
                        make_float4(du[kq * 4], du[kq * 4 + 1], du[kq * 4 + 2], du[kq * 4 + 3]));
            }
            __syncthreads();
            for (int t = threadIdx.x; t < DUB * K4 * kLanes; t += NT) {
                const int l = t & 31;
                const int kq = (t >> 5) % K4;
                const int ii = t / (K4 * kLanes);
                const int il = ib + ii;
                if (il < ni) {
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < JW; ++w) {
                        const float4 x = *reinterpret_cast<const float4*>(dusm + ((size_t)((w * DUB + ii) * K4 + kq) * kLanes + l) * 4);
                        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                    }
                    st4(p.du_part + (((((size_t)blockIdx.y * p.nbt + tile) * p.N + i0 + il) * K4 + kq) * kLanes + l) * 4, sum);
                }
            }
            __syncthreads();
        }
    }

    // dW tile -> public layout [N][C][K][D] (drop the d padding)
    for (int t = threadIdx.x; t < IT * JW * ROW; t += NT) {
        const int e = t % ROW;
        const int w = (t / ROW) % JW;
        const int il = t / (ROW * JW);
        const int k = e / DP, d = e % DP;
        if (il < ni && w < njv && d < p.D)
            p.dW[(((size_t)(i0 + il) * p.C + j0 + w) * K + k) * p.D + d] = dWsm[t];
    }
}

// du [B][N][K] = sum over j-groups of du_part
template <int K>
__global__ void k_reduce_du(const float* __restrict__ du_part, int JG, float* __restrict__ du, int B, int N, int nbt) {
    constexpr int K4 = K / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nbt * N * kLanes) return;
    const int lane = (int)(idx & 31);
    const long ti = idx >> 5;
    const int i = (int)(ti % N);
    const long b = (ti / N) * kLanes + lane;
    if (b >= B) return;
    const size_t usize = (size_t)nbt * N * K * kLanes;
#pragma unroll
    for (int kq = 0; kq < K4; ++kq) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < JG; ++g) {
            const float4 x = ldg4(du_part + g * usize + (((size_t)ti * K4 + kq) * kLanes + lane) * 4);
            sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
        }
        st4(du + ((size_t)b * N + i) * K + kq * 4, sum);
    }
}

// ---------------------------------------------------------------------------------------------
// margin loss value, squash for the primary-capsule branch
// ---------------------------------------------------------------------------------------------
// fixed-order two-level tree: deterministic.  v public [B][C][D].  With gridDim.x > 1 every
// block writes its partial to part[blockIdx.x] and k_margin_loss_final adds them in order.
static __global__ void k_margin_loss(const float* __restrict__ v, const int64_t* __restrict__ y, float scale,
                                     float* __restrict__ loss, float* __restrict__ part,
                                     float* __restrict__ scores, int B, int C, int D) {
    __shared__ float red[1024];              // blockDim.x <= 1024 (a power of two)
    float acc = 0.f;
    const long n = (long)B * C;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float* p = v + e * D;
        float m2 = 0.f;
        for (int d = 0; d < D; ++d) m2 = fmaf(p[d], p[d], m2);
        const float m = sqrtf(m2);                           // reference models.py:117
        if (scores) scores[e] = m;
        const bool hit = (y[e / C] == (int64_t)(e % C));
        const float l = fmaxf(0.9f - m, 0.f), r = fmaxf(m - 0.1f, 0.f);
        acc += hit ? l * l : 0.5f * r * r;                   // reference loss_fns.py:12-17
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) *loss = red[0] * scale;
        else part[blockIdx.x] = red[0];
    }
}

static __global__ void k_margin_loss_final(const float* __restrict__ part, int n, float scale, float* __restrict__ loss) {
    __shared__ float red[1024];              // blockDim.x <= 1024 (a power of two)
    float acc = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) acc += part[e];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = red[0] * scale;
}

// ---------------------------------------------------------------------------------------------
// DarkCapsuleNet loss tail (reference loss_fns.py:187-204 with utils.polar_transform, utils.py:69-85; recon off):
//   v [G*B][5] is the routing layer's output, row q*B + b = cell q of sample b (what models.py:400 produces before its
//   view(g,g,B,5).permute(2,0,1,3)); y [B][G][Y] holds (r, x, y, w, h, classes...) per cell.
//   loss = scale * sum_cells [ y_r relu(0.9 - |v|)^2 + 0.5 (1 - y_r) relu(|v| - 0.1)^2 - v . y_phi ]
//   grad_v (nullable) = d loss / d v in v's own layout, ready to be caps_route_backward's grad_v.
// Fixed-order two-level reduction like k_margin_loss (k_margin_loss_final adds the partials).
// ---------------------------------------------------------------------------------------------
static __global__ void k_dark_loss(const float* __restrict__ v, const float* __restrict__ y, float scale,
                                   float* __restrict__ loss, float* __restrict__ part, float* __restrict__ grad_v,
                                   int B, int G, int Y) {
    __shared__ float red[1024];              // blockDim.x <= 1024 (a power of two)
    float acc = 0.f;
    const long n = (long)B * G;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int q = (int)(e / B);
        const long b = e % B;                                   // e = q*B + b: v's row
        const float* p = v + e * 5;
        const float* t = y + ((size_t)b * G + q) * Y;
        const float pi = 3.14159265358979323846f;
        const float yr = t[0];
        float s1, c1, s2, c2, s3, c3, s4, c4;
        sincosf(t[1] * pi, &s1, &c1);                           // f1 = x pi, f2 = y pi, f3 = h pi, f4 = w 2 pi (utils.py:74)
        sincosf(t[2] * pi, &s2, &c2);
        sincosf(t[4] * pi, &s3, &c3);
        sincosf(t[3] * pi * 2.f, &s4, &c4);
        const float phi[5] = {s1, s1 * c2, s1 * s2 * c3, s1 * s2 * s3 * c4, s1 * s2 * s3 * s4};
        float m2 = 0.f, dot = 0.f;
#pragma unroll
        for (int d = 0; d < 5; ++d) { m2 = fmaf(p[d], p[d], m2); dot = fmaf(p[d], phi[d], dot); }
        const float m = sqrtf(m2);
        const float l = fmaxf(0.9f - m, 0.f), r = fmaxf(m - 0.1f, 0.f);
        acc += yr * l * l + 0.5f * (1.f - yr) * r * r - dot;
        if (grad_v != nullptr) {
            const float dm = (-2.f * yr * l + (1.f - yr) * r) / m;
#pragma unroll
            for (int d = 0; d < 5; ++d) grad_v[e * 5 + d] = scale * (dm * p[d] - phi[d]);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) *loss = red[0] * scale;
        else part[blockIdx.x] = red[0];
    }
}

static __global__ void k_squash_rows(const float* __restrict__ x, float* __restrict__ y, long rows, int D) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = x + r * D;
    float n2 = 0.f;
    for (int d = 0; d < D; ++d) n2 = fmaf(p[d], p[d], n2);
    const float scale = n2 / (1.f + n2), rn = sqrtf(n2);
    for (int d = 0; d < D; ++d) y[r * D + d] = scale * p[d] / rn;
}

static __global__ void k_squash_rows_bwd(const float* __restrict__ x, const float* __restrict__ dy,
                                  float* __restrict__ dx, long rows, int D) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = x + r * D;
    const float* g = dy + r * D;
    float n2 = 0.f, sdv = 0.f;
    for (int d = 0; d < D; ++d) { n2 = fmaf(p[d], p[d], n2); sdv = fmaf(p[d], g[d], sdv); }
    const float n = sqrtf(n2);
    const float a = n / (1.f + n2);
    const float b = sdv * (1.f - n2) / (n * (1.f + n2) * (1.f + n2));
    for (int d = 0; d < D; ++d) dx[r * D + d] = fmaf(a, g[d], b * p[d]);
}


// ---------------------------------------------------------------------------------------------
// primary-capsule tail (reference models.py:81-82): conv [B][K*Cc][HW] (ONE convolution whose output channels
// are the K capsule convolutions' channels, capsule-major) -> u [B][Cc*HW][K] = squash over k.  The reference's
// K x view, cat(dim=-1) and 7-op squash become one pass: thread <-> (b, c, hw) reads K values that are
// coalesced across hw and writes K consecutive floats.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void k_primary_squash(const float* __restrict__ conv, float* __restrict__ u, long total, int K, int Cc, int HW) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;        // (b, c, hw), hw fastest
    if (idx >= total) return;
    const int hw = (int)(idx % HW);
    const long bc = idx / HW;
    const int c = (int)(bc % Cc);
    const long b = bc / Cc;
    const float* src = conv + ((size_t)b * K * Cc + c) * HW + hw;
    float x[KMAX];
    float n2 = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        x[k] = k < K ? __ldg(src + (size_t)k * Cc * HW) : 0.f;
        n2 = fmaf(x[k], x[k], n2);
    }
    const float scale = n2 / (1.f + n2), rn = sqrtf(n2);
    float* dst = u + (size_t)idx * K;
    if (KMAX == 8 && K == 8) {
        st4(dst, make_float4(scale * x[0] / rn, scale * x[1] / rn, scale * x[2] / rn, scale * x[3] / rn));
        st4(dst + 4, make_float4(scale * x[4] / rn, scale * x[5] / rn, scale * x[6] / rn, scale * x[7] / rn));
    } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) dst[k] = scale * x[k] / rn;
    }
}

// dconv [B][K*Cc][HW] = squash'(conv) applied to du [B][Cc*HW][K] (same formula as k_squash_rows_bwd)
template <int KMAX>
__global__ void k_primary_squash_bwd(const float* __restrict__ conv, const float* __restrict__ du, float* __restrict__ dconv,
                                     long total, int K, int Cc, int HW) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int hw = (int)(idx % HW);
    const long bc = idx / HW;
    const int c = (int)(bc % Cc);
    const long b = bc / Cc;
    const size_t o = ((size_t)b * K * Cc + c) * HW + hw;
    const float* g = du + (size_t)idx * K;
    float x[KMAX], gy[KMAX];
    float n2 = 0.f, sdv = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        x[k] = k < K ? __ldg(conv + o + (size_t)k * Cc * HW) : 0.f;
        gy[k] = k < K ? __ldg(g + k) : 0.f;
        n2 = fmaf(x[k], x[k], n2);
        sdv = fmaf(x[k], gy[k], sdv);
    }
    const float n = sqrtf(n2);
    const float a = n / (1.f + n2);
    const float bb = sdv * (1.f - n2) / (n * (1.f + n2) * (1.f + n2));
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < K) dconv[o + (size_t)k * Cc * HW] = fmaf(a, gy[k], bb * x[k]);
}

// ---------------------------------------------------------------------------------------------
// DarkCapsuleNet cell regroup (reference models.py:393-399): feature map x [B][Cch][16*G] (the 28x28 map viewed as
// [4][4*G]) -> u [G*B][2*Cch][8] for the routing layer:
//     u[q*B + b][(a*4 + t)*(Cch/8) + ch/8][ch%8] = x[b][ch][a*4*G + 4*q + t]        q < G, a < 4, t < 4
// i.e. the reference's view + chunk(G) + G x (permute, contiguous, view, unsqueeze) + cat in one pass.
// thread <-> (b, ch/8, s): the 8 reads are coalesced across s, the write is one 32-byte sector.
// BWD: the same index map, gradient flowing from du back into dx.
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void k_dark_regroup(const float* __restrict__ src, float* __restrict__ dst, long total, int B, int Cch, int G) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;        // (b, c8, s), s fastest
    if (idx >= total) return;
    const int S = 16 * G, C8 = Cch >> 3;
    const int sp = (int)(idx % S);
    const long bc = idx / S;
    const int c8 = (int)(bc % C8);
    const long b = bc / C8;
    const int a = sp / (4 * G), r = sp % (4 * G), q = r >> 2, t = r & 3;
    const size_t xo = ((size_t)b * Cch + c8 * 8) * S + sp;                                  // x[b][8 c8 + k][sp], k stride S
    const size_t uo = (((size_t)q * B + b) * (2 * Cch) + (size_t)(a * 4 + t) * C8 + c8) * 8; // u[q B + b][n][0..7]
    if (!BWD) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(src + xo + (size_t)k * S);
        st4(dst + uo, make_float4(v[0], v[1], v[2], v[3]));
        st4(dst + uo + 4, make_float4(v[4], v[5], v[6], v[7]));
    } else {
        const float4 lo = ldg4(src + uo), hi = ldg4(src + uo + 4);
        dst[xo] = lo.x; dst[xo + (size_t)S] = lo.y; dst[xo + (size_t)2 * S] = lo.z; dst[xo + (size_t)3 * S] = lo.w;
        dst[xo + (size_t)4 * S] = hi.x; dst[xo + (size_t)5 * S] = hi.y; dst[xo + (size_t)6 * S] = hi.z; dst[xo + (size_t)7 * S] = hi.w;
    }
}

// fp32 FMA-pipe peak probe: 16 independent FFMA chains per thread (bench.py's FMA roofline
// denominator; MEASURED_PEAKS.json has no fp32 figure).
static __global__ void k_fma_peak(float* __restrict__ sink, int iters, float m0, float c0) {
    // all three operands live in registers (the form the routing kernels issue), not immediates
    float a[16], m[16], c[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        a[t] = 1.0f + 1e-3f * (float)(threadIdx.x + t);
        m[t] = m0 - 1e-6f * (float)(threadIdx.x + t);
        c[t] = c0 + 1e-7f * (float)(threadIdx.x * 3 + t);
    }
    for (int it = 0; it < iters; it += 4) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int t = 0; t < 16; ++t) a[t] = fmaf(a[t], m[t & 3], c[(t >> 2) & 3]);
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += a[t];
    if (s == 12345.678f) sink[threadIdx.x] = s;       // never true; keeps the chains alive
}

}  // namespace caps
