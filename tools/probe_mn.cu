// probe_mn.cu -- can the gradient sweep's dW = G^T u run on tcgen05?  The batch is that product's contraction index, so
// both operands would have to be read "transposed" (MN-major) out of tiles that are written sample by sample:
//   A = G^T : M = 128 rows (capsule, d), K = 8 samples per MMA.  Tile as a thread (= sample b) writes it, 16 bytes = 4
//             consecutive (capsule, d) values at a time:   byte(b, m) = (b / 8) * 4096 + (m / 4) * 128 + (b % 8) * 16 + (m % 4) * 4
//             -> MN-major canonical layout, no swizzle: core matrix = 8 samples x 16 bytes, SBO = 128 (next 4 rows of M)
//             (the very same bytes are a K-major operand with M = samples: SBO = 4096, LBO = 128 -- the du product)
//   B = u   : N = 16 columns (k hi | k lo), K = 8 samples.  The sweeps' operand block [hi/lo][k/4][128 samples][4] is
//             MN-major as it stands: core matrix = 8 samples x 16 bytes (4 values of k), SBO = 2048 (next 4 columns)
// The probe fills both tiles with known values, accumulates D[128 x 16] over the 16 K-steps of a 128-sample block with
// tcgen05.mma kind::tf32 (A, B from shared memory, both MN-major), reads D back and compares with the host; then times
// the MMA stream for 1, 2 and 4 issuing warps.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_mn tools/probe_mn.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, no swizzle, version 1: start address, LBO, SBO in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int kSamples = 128, kM = 128, kN = 16;

// grid 1, block 128.  gA: [128 samples][128 m] row-major (G), gB: [128 samples][16 n] row-major (u hi|lo), out: [128 m][16 n]
template <int NW>
__global__ void __launch_bounds__(128, 1) k_probe(const float* gA, const float* gB, float* out, int reps, long long* cyc, int lbo_a, int lbo_b) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                       // 64 KB
    uint8_t* sB = smem + 65536;               // 8 KB: [n/4 = 4][128 samples][4]
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool a_k = lbo_a == -1 || lbo_a == -3, b_k = lbo_a == -1 || lbo_a == -2;     // controls: one or both operands K-major
    const bool kmajor = a_k && b_k;
    if (a_k) {
        // A K-major: element (m, k = sample): 16 K-steps of 8 samples; per K-step: (m/8)*128 + (k/4)*2048 + (m%8)*16 + (k%4)*4, K-steps 4096 apart
        for (int e = tid; e < 128 * 128; e += 128) {
            const int b = e / 128, m = e % 128;
            *reinterpret_cast<float*>(sA + (b / 8) * 4096 + (m / 8) * 128 + ((b % 8) / 4) * 2048 + (m % 8) * 16 + (b % 4) * 4) = gA[(size_t)b * kM + m];
        }
    } else {
        const int b = tid;
        for (int m4 = 0; m4 < kM / 4; ++m4) {
            const float4 v = *reinterpret_cast<const float4*>(gA + (size_t)b * kM + 4 * m4);
            *reinterpret_cast<float4*>(sA + (b / 8) * 4096 + m4 * 128 + (b % 8) * 16) = v;
        }
    }
    if (b_k) {
        // B K-major: element (n, k = sample): per K-step: (n/8)*128 + (k/4)*256 + (n%8)*16 + (k%4)*4, K-steps 512 apart
        for (int e = tid; e < 128 * 16; e += 128) {
            const int b = e / 16, n = e % 16;
            *reinterpret_cast<float*>(sB + (b / 8) * 512 + (n / 8) * 128 + ((b % 8) / 4) * 256 + (n % 8) * 16 + (b % 4) * 4) = gB[(size_t)b * kN + n];
        }
    } else {   // thread = sample b writes its row the way an epilogue thread would
        const int b = tid;
        for (int n4 = 0; n4 < kN / 4; ++n4) {
            const float4 v = *reinterpret_cast<const float4*>(gB + (size_t)b * kN + 4 * n4);
            *reinterpret_cast<float4*>(sB + n4 * 2048 + b * 16) = v;
        }
    }
    if (tid == 0) for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[w])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // instruction descriptor: D = f32 (bit 4), A = B = tf32 (2 << 7, 2 << 10), A MN-major (bit 15), B MN-major (bit 16), N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_k ? 0u : (1u << 15)) | (b_k ? 0u : (1u << 16))) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (warp < NW && lane == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            // this warp's share of the 16 K-steps (8 samples each)
            for (int ks = warp; ks < kSamples / 8; ks += NW) {
                const uint64_t a = a_k ? make_desc(smem_u32(sA) + ks * 4096, 2048, 128) : make_desc(smem_u32(sA) + ks * 4096, lbo_a < 0 ? 4096 : lbo_a, 128);
                const uint64_t b = b_k ? make_desc(smem_u32(sB) + ks * 512, 256, 128) : make_desc(smem_u32(sB) + ks * 128, lbo_a < 0 ? 128 : lbo_b, 2048);
                mma_ss(tmem_base + warp * 16, a, b, idesc, (r > 0 || ks >= NW) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    }
    if (warp < NW) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[warp])) : "memory");
        if (lane == 0) { t1 = clock64(); cyc[warp] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D rows = TMEM lanes: warp w reads lanes 32w..32w+31, the NW partial accumulators side by side
    for (int w = 0; w < NW; ++w) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + w * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 16; ++n) {
            const float v = __uint_as_float(r[n]);
            if (w == 0) out[(size_t)tid * 16 + n] = v; else out[(size_t)tid * 16 + n] += v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

template <int NW>
void run(const float* dA, const float* dB, float* dO, long long* dC, const float* hA, const float* hB, int lbo_a, int lbo_b, bool check) {
    const size_t smem = 65536 + 8192;
    cudaFuncSetAttribute(k_probe<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<NW><<<1, 128, smem>>>(dA, dB, dO, 1, dC, lbo_a, lbo_b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("NW=%d lbo_a=%d lbo_b=%d: %s\n", NW, lbo_a, lbo_b, cudaGetErrorString(e)); exit(1); }
    if (check) {
        static float hO[128 * 16];
        cudaMemcpy(hO, dO, sizeof(hO), cudaMemcpyDeviceToHost);
        double worst = 0, scale = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                double ref = 0;
                for (int b = 0; b < 128; ++b) ref += (double)tf32_trunc(hA[b * 128 + m]) * (double)tf32_trunc(hB[b * 16 + n]);
                worst = fmax(worst, fabs(ref - hO[m * 16 + n]));
                scale = fmax(scale, fabs(ref));
            }
        printf("NW=%d lbo_a=%d lbo_b=%d: D[(capsule,d)][k] = sum_b G[b][(capsule,d)] u[b][k]: max |err| %.3e (max |ref| %.3e) -> %s\n", NW, lbo_a, lbo_b,
               worst, scale, worst < 1e-4 * scale ? "MATCHES" : "DIFFERS");
        if (!(worst < 1e-4 * scale)) {
            for (int m = 0; m < 3; ++m) {
                printf("   D[%d][0..5] =", m);
                for (int n = 0; n < 6; ++n) printf(" %9.5f", hO[m * 16 + n]);
                printf("   ref =");
                for (int n = 0; n < 6; ++n) { double ref = 0; for (int b = 0; b < 128; ++b) ref += (double)tf32_trunc(hA[b * 128 + m]) * tf32_trunc(hB[b * 16 + n]); printf(" %9.5f", ref); }
                printf("\n");
            }
            // which (m', n') of the reference does D[1][2] equal?  (tells a permuted layout from garbage)
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
                double ref = 0; for (int b = 0; b < 128; ++b) ref += (double)tf32_trunc(hA[b * 128 + m]) * tf32_trunc(hB[b * 16 + n]);
                if (fabs(ref - hO[1 * 16 + 2]) < 1e-4) printf("   D[1][2] == ref[%d][%d]\n", m, n);
                if (fabs(ref - hO[5 * 16 + 9]) < 1e-4) printf("   D[5][9] == ref[%d][%d]\n", m, n);
            }
        }
    }
    const int reps = 2048;
    k_probe<NW><<<1, 128, smem>>>(dA, dB, dO, reps, dC, lbo_a, lbo_b);
    cudaDeviceSynchronize();
    long long c[4];
    cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < NW; ++w) mx = c[w] > mx ? c[w] : mx;
    printf("NW=%d: %.1f SM cycles per MMA (M=128, N=16, K=8, A and B MN-major from shared memory), %.0f cycles per 128-sample block of 16 MMAs\n",
           NW, (double)mx / (reps * 16.0), (double)mx / reps);
}

int main() {
    static float hA[128 * 128], hB[128 * 16];
    srand(1);
    for (auto& x : hA) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : hB) x = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dB, *dO; long long* dC;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, 128 * 16 * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    // LBO is the stride between K groups of 8, unused when one MMA covers a single group (K = 8): try two values to show that
    run<1>(dA, dB, dO, dC, hA, hB, -1, 0, true);        // control: K-major operands
    run<1>(dA, dB, dO, dC, hA, hB, -2, 0, true);        // A MN-major, B K-major
    run<1>(dA, dB, dO, dC, hA, hB, -3, 0, true);        // A K-major, B MN-major
    run<1>(dA, dB, dO, dC, hA, hB, 4096, 128, true);
    run<1>(dA, dB, dO, dC, hA, hB, 128, 2048, true);
    run<2>(dA, dB, dO, dC, hA, hB, 4096, 128, true);
    run<4>(dA, dB, dO, dC, hA, hB, 4096, 128, true);
    return 0;
}
