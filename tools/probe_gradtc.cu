// probe_gradtc.cu -- sizing probe for a tcgen05 gradient sweep ("prove tcgen05 cannot win or port it").
//
// One stage = (128 samples, 4 class capsules, one input capsule i).  16 "builder" warps (quadrant q = 32 samples, capsule
// c) each hold G[16] per lane (here: synthetic, built by a few packed FMAs), split it (hi = G, lo = G - trunc(G)) and park it
//   (a) in TMEM, lanes <-> samples, columns (hi/lo, c, d):          the A operand of  du[b][k] = sum_(c,d) G[b][(c,d)] W[(c,d)][k]
//   (b) in shared memory, K-major with the SAMPLES as K:             the A operand of  dW[(c,d)][k] = sum_b G[b][(c,d)] u[b][k]
//       rows m = hl*64 + c*16 + d (hi rows and lo rows of the same MMA), 16-byte chunks = 4 consecutive samples, written with
//       4-byte stores; LBO = 2064, K-step stride 4128 bytes so that the 8 chunks a warp touches fall into 8 different bank groups
// and 4 issuing warps run 16 + 16 tcgen05.mma (kind::tf32, M = 128, N = 16, K = 8) per stage:
//   dW: A from shared memory (tile above), B = u block [n = (hi|lo, k)][samples] K-major;  D_dW[128 rows][16] accumulates over stages
//   du: A from TMEM (hi columns, then lo columns), B = W^T block [n = (hi|lo, k)][d] K-major;  D_du[128 samples][16]
// The probe checks both products against the host for one stage, then times a stream of stages (double-buffered).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_gradtc tools/probe_gradtc.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
                   "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int kThreads = 640;              // warps 0-15 builders (q = w & 3, c = w >> 2), 16-19 issuers
constexpr uint32_t kLboA = 2064, kKsA = 4128, kTileA = 16 * kKsA;     // dW A tile: 66048 bytes per buffer
constexpr uint32_t kOffTile = 0;                                       // 2 buffers
constexpr uint32_t kOffU = 2 * kTileA;                                 // u block: 16 K-steps x 512 bytes (K-major B, N = 16)
constexpr uint32_t kOffW = kOffU + 8192;                               // W^T blocks: 4 capsules x 2 K-steps (8 d each) x 512 bytes
constexpr uint32_t kOffBar = kOffW + 4096;
constexpr uint32_t kSmem = kOffBar + 64;
// TMEM columns: A_du buffers [2][hl 2][c 4][16] = 2 x 128, D_dW 16 at 256, D_du 16 at 272
constexpr uint32_t kColA = 0, kColDW = 256, kColDU = 272;

// G value of (sample b, capsule c, dim d) at stage s: something the host can recompute exactly
__host__ __device__ inline float g_val(int b, int c, int d, int s) { return 0.25f + 0.001f * (float)((b * 7 + c * 13 + d * 3 + s * 5) % 97) - 0.0004f * (float)((b + d) % 11); }
__host__ __device__ inline float u_val(int b, int k) { return 0.5f - 0.003f * (float)((b * 5 + k * 11) % 89); }
__host__ __device__ inline float w_val(int c, int d, int k) { return 0.1f * ((float)((c * 17 + d * 7 + k * 3) % 23) - 11.f) / 11.f; }
__host__ __device__ inline float trunc_tf32(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
#else
    uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x;
#endif
}

__global__ void __launch_bounds__(kThreads, 1) k_probe(int stages, int extra_fma, float* out_dw, float* out_du, long long* cyc, int skip) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = base + kOffBar, bar_empty = bar_full + 16;           // [2] each
    // operand blocks that a real kernel would get by bulk copy
    for (int e = tid; e < 128 * 16; e += kThreads) {       // u: B[n][k = sample]: n < 8: u (hardware truncates), n >= 8: lo
        const int b = e / 16, n = e % 16;
        const float x = u_val(b, n & 7);
        const float v = n < 8 ? x : x - trunc_tf32(x);
        *reinterpret_cast<float*>(smem + kOffU + (b / 8) * 512 + ((b % 8) / 4) * 256 + (n / 8) * 128 + (n % 8) * 16 + (b % 4) * 4) = v;
    }
    for (int e = tid; e < 4 * 16 * 16; e += kThreads) {    // W^T: B[n = (hl, k)][kk = d] per capsule c, K-steps of 8 dims
        const int c = e / 256, d = (e / 16) % 16, n = e % 16;
        const float x = w_val(c, d, n & 7);
        const float v = n < 8 ? x : x - trunc_tf32(x);
        *reinterpret_cast<float*>(smem + kOffW + (c * 2 + d / 8) * 512 + ((d % 8) / 4) * 256 + (n / 8) * 128 + (n % 8) * 16 + (d % 4) * 4) = v;
    }
    if (tid == 0) {
        for (int q = 0; q < 2; ++q) { mbar_init(bar_full + 8 * q, 16); mbar_init(bar_empty + 8 * q, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const long long t0 = clock64();
    if (warp < 16) {
        // ===== builders =====
        const int q = warp & 3, c = warp >> 2, b = q * 32 + lane;
        float acc[16], gbase[16];
        for (int d = 0; d < 16; ++d) { acc[d] = 0.f; gbase[d] = g_val(b, c, d, 0); }
        for (int s = 0; s < stages; ++s) {
            const int buf = s & 1;
            float G[16], lo[16];
            if (stages == 1) { for (int d = 0; d < 16; ++d) G[d] = g_val(b, c, d, s); }        // the checked run
            else { for (int d = 0; d < 16; ++d) G[d] = gbase[d] + (float)s * 1e-6f; }           // timing runs: cheap, still varying
            for (int r = 0; r < extra_fma; ++r)                                  // stands in for the 5-term G build
                for (int d = 0; d < 16; d += 2) {
                    asm volatile("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %4};\n\tfma.rn.f32x2 x, z, y, x;\n\tmov.b64 {%0, %1}, x;\n\t}"
                                 : "+f"(acc[d]), "+f"(acc[d + 1]) : "f"(1e-9f), "f"(G[d]), "f"(G[d + 1]));
                }
            for (int d = 0; d < 16; ++d) { G[d] += acc[d] * 1e-30f; lo[d] = G[d] - trunc_tf32(G[d]); }
            mbar_wait(bar_empty + 8 * buf, ((s >> 1) & 1) ^ 1);                  // MMAs of stage s - 2 have read this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kColA + buf * 128 + c * 16;
            if (!(skip & 8)) { tmem_st16(ta, G); tmem_st16(ta + 64, lo); }
            uint8_t* tile = smem + kOffTile + buf * kTileA + (b / 8) * kKsA + ((b % 8) / 4) * kLboA + (b % 4) * 4;
            if (!(skip & 1)) for (int d = 0; d < 16; ++d) {
                const int m = c * 16 + d;
                *reinterpret_cast<float*>(tile + (m / 8) * 128 + (m % 8) * 16) = G[d];
                *reinterpret_cast<float*>(tile + ((64 + m) / 8) * 128 + (m % 8) * 16) = lo[d];
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (!(skip & 16)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the tensor core reads the tile through the async proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * buf);
        }
    } else if (lane == 0) {
        // ===== issuers: warp 16 + w takes K-steps w, w+4, .. of dW and capsule w of du =====
        const int w = warp - 16;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int s = 0; s < stages; ++s) {
            const int buf = s & 1;
            mbar_wait(bar_full + 8 * buf, (s >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!(skip & 2)) for (int ks = w; ks < 16; ks += 4) {                                 // dW: 16 K-steps of 8 samples
                const uint64_t a = make_desc(base + kOffTile + buf * kTileA + ks * kKsA, kLboA, 128);
                const uint64_t bd = make_desc(base + kOffU + ks * 512, 256, 128);
                mma_ss(tmem_base + kColDW + w * 32, a, bd, idesc, (s > 0 || ks >= 4) ? 1u : 0u);     // one partial accumulator per issuer
            }
            if (!(skip & 4)) for (int hl = 0; hl < 2; ++hl)                                       // du: capsule w, 2 K-steps of 8 dims, hi then lo columns
                for (int kd = 0; kd < 2; ++kd) {
                    const uint64_t bd = make_desc(base + kOffW + (w * 2 + kd) * 512, 256, 128);
                    mma_ts(tmem_base + kColDU + w * 32, tmem_base + kColA + buf * 128 + hl * 64 + w * 16 + kd * 8, bd, idesc, (hl | kd) ? 1u : 0u);
                }
            commit(bar_empty + 8 * buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
    // wait for the last commits, then read the accumulators back (partials of the 4 issuers side by side, 32 columns apart)
    if (warp >= 16 && lane == 0) { const int s = stages - 1; (void)s; }
    mbar_wait(bar_empty + 8 * ((stages - 1) & 1), ((stages - 1) >> 1) & 1);
    if (stages > 1) mbar_wait(bar_empty + 8 * ((stages - 2) & 1), ((stages - 2) >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4 && blockIdx.x == 0) {
        float dw[16], du[16], t[16];
        for (int n = 0; n < 16; ++n) { dw[n] = 0.f; du[n] = 0.f; }
        for (int w = 0; w < 4; ++w) {
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + kColDW + w * 32, t);
            for (int n = 0; n < 16; ++n) dw[n] += t[n];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + kColDU + w * 32, t);
            for (int n = 0; n < 16; ++n) du[n] += t[n];
        }
        for (int n = 0; n < 16; ++n) { out_dw[tid * 16 + n] = dw[n]; out_du[tid * 16 + n] = du[n]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d_dw, *d_du; long long* d_c;
    cudaMalloc(&d_dw, 128 * 16 * 4); cudaMalloc(&d_du, 128 * 16 * 4); cudaMalloc(&d_c, 8 * 1024);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    // ---- numerics: ONE stage
    k_probe<<<1, kThreads, kSmem>>>(1, 0, d_dw, d_du, d_c, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check run: %s\n", cudaGetErrorString(e)); return 1; }
    static float h_dw[128 * 16], h_du[128 * 16];
    cudaMemcpy(h_dw, d_dw, sizeof(h_dw), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_du, d_du, sizeof(h_du), cudaMemcpyDeviceToHost);
    // dW[(c,d)][k] = sum_b G u  (fp64 reference of the exact fp32 inputs): device rows m and 64 + m, columns k and 8 + k add up
    double worst_w = 0, scale_w = 0, worst_u = 0, scale_u = 0;
    for (int m = 0; m < 64; ++m)
        for (int k = 0; k < 8; ++k) {
            double ref = 0;
            for (int b = 0; b < 128; ++b) ref += (double)g_val(b, m / 16, m % 16, 0) * (double)u_val(b, k);
            const double got = (double)h_dw[m * 16 + k] + h_dw[m * 16 + 8 + k] + h_dw[(64 + m) * 16 + k] + h_dw[(64 + m) * 16 + 8 + k];
            worst_w = fmax(worst_w, fabs(ref - got)); scale_w = fmax(scale_w, fabs(ref));
        }
    for (int b = 0; b < 128; ++b)
        for (int k = 0; k < 8; ++k) {
            double ref = 0;
            for (int c = 0; c < 4; ++c) for (int d = 0; d < 16; ++d) ref += (double)g_val(b, c, d, 0) * (double)w_val(c, d, k);
            const double got = (double)h_du[b * 16 + k] + h_du[b * 16 + 8 + k];
            worst_u = fmax(worst_u, fabs(ref - got)); scale_u = fmax(scale_u, fabs(ref));
        }
    printf("dW = G^T u on tcgen05 (A: K-major shared-memory tile, hi and lo rows; fp64 reference): max |err| %.3e of %.3e -> rel %.2e %s\n",
           worst_w, scale_w, worst_w / scale_w, worst_w < 2e-6 * scale_w ? "OK (fp32-grade)" : "DIFFERS");
    printf("du = G W^T on tcgen05 (A: TMEM, hi then lo columns; fp64 reference):            max |err| %.3e of %.3e -> rel %.2e %s\n",
           worst_u, scale_u, worst_u / scale_u, worst_u < 2e-6 * scale_u ? "OK (fp32-grade)" : "DIFFERS");
    // ---- timing: a stream of stages on every SM
    const int skips[] = {0, 1, 2, 4, 8, 16, 1 | 2, 2 | 4, 1 | 8 | 16, 1 | 2 | 4 | 8 | 16};
    for (int extra : {0, 5})
        for (int skip : skips) {
            if (extra && skip) continue;
            const int stages = 4096;
            k_probe<<<sms, kThreads, kSmem>>>(stages, extra, d_dw, d_du, d_c, skip);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("timing run: %s\n", cudaGetErrorString(e)); return 1; }
            static long long hc[1024];
            cudaMemcpy(hc, d_c, sms * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("stage stream, %2d x 8 packed FMAs, skip[%s%s%s%s%s ]: %5.0f SM cycles per stage (128 samples x 4 capsules x i) = %5.0f per 128 x 8-capsule unit (k_grad_mma today: ~4900)\n",
                   extra, skip & 1 ? " tile-stores" : "", skip & 2 ? " dW-MMAs" : "", skip & 4 ? " du-MMAs" : "", skip & 8 ? " tmem-stores" : "", skip & 16 ? " proxy-fence" : "",
                   (double)mx / stages, 2.0 * mx / stages);
        }
    return 0;
}
