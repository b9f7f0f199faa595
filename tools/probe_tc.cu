// probe_tc.cu -- bring-up probe for the tcgen05 path (run on the B200 via gpurun):
//   1. one CTA: D[128 x N] = A[128 x 8] * B[N x 8]^T with kind::tf32, operands in the canonical
//      no-swizzle K-major shared-memory layout (8-row x 16-byte core matrices, SBO between row
//      groups, LBO between the two 16-byte K chunks), accumulator in TMEM, read back with
//      tcgen05.ld.32x32b -> checks descriptors / instruction descriptor / TMEM lane mapping.
//   2. the 3xTF32 split (hi*hi + hi*lo + lo*hi) against an fp64 host product -> error budget.
//   3. issue-rate of back-to-back MMAs and of tcgen05.ld -> cycle model of the pass kernel.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_tc tools/probe_tc.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    long spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++spins > (1L << 24)) { printf("mbar_wait timeout\n"); __trap(); }
    } while (!ok);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

constexpr int M = 128, NMAX = 256;

// A: [128][8] row-major fp32, B: [N][8] row-major fp32 (B^T of the math).  D: [128][N].
// split=0: plain tf32 (hardware truncates);  split=1: 3xTF32.
// amajor=1: A is stored MN-major: [m/4][k][m%4] (8 k-rows x 16 bytes per 4-row quad; SBO = 128 B between quads)
__global__ void __launch_bounds__(128, 1) k_check(const float* A, const float* B, float* D, int N, int split, int reps, long long* cycles,
                                                  int amajor = 0, int Mrows = 128) {
    __shared__ __align__(1024) float sA[2][2 * M * 4];     // [hi/lo][kq][row][4]
    __shared__ __align__(1024) float sB[2][2 * NMAX * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < M * 8; e += 128) {
        const int r = e / 8, k = e % 8;
        const float x = A[e], hi = split ? tf32_rna(x) : x, lo = split ? tf32_rna(x - hi) : 0.f;
        const int off = amajor ? (r / 4) * 32 + k * 4 + (r % 4) : (k / 4) * M * 4 + r * 4 + (k % 4);
        sA[0][off] = hi;
        sA[1][off] = lo;
    }
    for (int e = tid; e < N * 8; e += 128) {
        const int r = e / 8, k = e % 8;
        const float x = B[e], hi = split ? tf32_rna(x) : x, lo = split ? tf32_rna(x - hi) : 0.f;
        sB[0][(k / 4) * N * 4 + r * 4 + (k % 4)] = hi;
        sB[1][(k / 4) * N * 4 + r * 4 + (k % 4)] = lo;
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (MMA)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amajor << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(Mrows >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t a_hi = amajor ? make_desc(smem_u32(sA[0]), 4096, 128) : make_desc(smem_u32(sA[0]), M * 16, 128);
        const uint64_t a_lo = amajor ? make_desc(smem_u32(sA[1]), 4096, 128) : make_desc(smem_u32(sA[1]), M * 16, 128);
        const uint64_t b_hi = make_desc(smem_u32(sB[0]), N * 16, 128), b_lo = make_desc(smem_u32(sB[1]), N * 16, 128);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (split) {
                mma_tf32(tmem_base, a_lo, b_hi, idesc, 0);
                mma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
                mma_tf32(tmem_base, a_hi, b_hi, idesc, 1);
            } else {
                mma_tf32(tmem_base, a_hi, b_hi, idesc, 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: warp w owns TMEM lanes 32w..32w+31 (= rows of D)
    const int row = warp * 32 + (tid & 31);
    long long l0 = clock64();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
    }
    long long l1 = clock64();
    if (tid == 0) cycles[1] = l1 - l0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

// tcgen05.ld throughput: 4 warps each read `cols` columns repeatedly (no stores)
__global__ void __launch_bounds__(128, 1) k_ldrate(float* sink, int reps, long long* cycles) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (int c0 = 0; c0 < 128; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int i = 0; i < 16; ++i) acc += v[i];
        }
    long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    if (acc == 1234.5f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

int main() {
    std::vector<float> hA(M * 8), hB(NMAX * 8);
    srand(1);
    for (auto& x : hA) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : hB) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
    float *dA, *dB, *dD;
    long long* dC;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, M * NMAX * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    for (int N : {128, 256, 64, 16}) {
        for (int split : {0, 1}) {
            cudaMemset(dD, 0, M * NMAX * 4);
            k_check<<<1, 128>>>(dA, dB, dD, N, split, 1, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d split=%d: CUDA error %s\n", N, split, cudaGetErrorString(e)); return 1; }
            std::vector<float> hD(M * N);
            cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int r = 0; r < M; ++r)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 8; ++k) ref += (double)hA[r * 8 + k] * (double)hB[n * 8 + k];
                    maxerr = fmax(maxerr, fabs(ref - hD[r * N + n]));
                    maxref = fmax(maxref, fabs(ref));
                }
            printf("check N=%3d split=%d: max abs err %.3e (max |ref| %.3e, rel %.3e) %s\n", N, split, maxerr, maxref,
                   maxerr / maxref, maxerr / maxref < (split ? 2e-6 : 3e-3) ? "OK" : "FAIL");
        }
    }
    // MN-major A (the layout the gradient kernel wants for G^T), M = 128 and M = 64
    for (int Mr : {128, 64}) for (int N : {16, 8}) {
        if (Mr == 128 && N == 8) continue;
        cudaMemset(dD, 0, M * NMAX * 4);
        k_check<<<1, 128>>>(dA, dB, dD, N, 1, 1, dC, 1, Mr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("MN-major M=%d N=%d: CUDA error %s\n", Mr, N, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(M * N);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        // M=64: rows 0..63 of D live in TMEM lanes 0..31 and 64..95?  print the row map we observe
        double maxerr = 0, maxref = 0; int bad = 0;
        for (int r = 0; r < Mr; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < 8; ++k) ref += (double)hA[r * 8 + k] * (double)hB[n * 8 + k];
                double err = fabs(ref - hD[r * N + n]);
                if (err > 1e-5) ++bad;
                maxerr = fmax(maxerr, err); maxref = fmax(maxref, fabs(ref));
            }
        printf("check MN-major A, M=%3d N=%2d: max abs err %.3e rel %.3e bad %d %s\n", Mr, N, maxerr, maxerr / maxref, bad,
               maxerr / maxref < 2e-6 ? "OK" : "FAIL(lane map?)");
        if (Mr == 64 && bad) {
            // find which TMEM lane holds math row r (compare column 0)
            for (int r = 0; r < 64; r += 9) {
                double ref = 0; for (int k = 0; k < 8; ++k) ref += (double)hA[r * 8 + k] * (double)hB[0 * 8 + k];
                for (int l = 0; l < 128; ++l) if (fabs(hD[l * N] - ref) < 1e-6) printf("   math row %d found in TMEM lane %d\n", r, l);
            }
        }
    }
    for (int cfg = 0; cfg < 4; ++cfg) {
        const int Ns[4] = {16, 8, 16, 32}, Ms[4] = {128, 64, 64, 128};
        long long hc[2];
        const int reps = 2000;
        k_check<<<1, 128>>>(dA, dB, dD, Ns[cfg], 1, reps, dC, 0, Ms[cfg]);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, dC, 16, cudaMemcpyDeviceToHost);
        printf("issue rate M=%d N=%d: %.1f cycles per MMA\n", Ms[cfg], Ns[cfg], (double)hc[0] / (3.0 * reps));
    }
    for (int N : {128, 256}) {
        long long hc[2];
        const int reps = 2000;
        k_check<<<1, 128>>>(dA, dB, dD, N, 1, reps, dC);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, dC, 16, cudaMemcpyDeviceToHost);
        printf("issue rate N=%d: %.1f cycles per MMA (3 per i) -> %.1f cycles per 3xTF32 step\n", N, (double)hc[0] / (3.0 * reps),
               (double)hc[0] / reps);
        printf("  single-CTA readback of %d columns: %lld cycles (incl. global stores)\n", N, hc[1]);
    }
    {
        long long hc[2];
        const int reps = 1000;
        k_ldrate<<<1, 128>>>(dD, reps, dC);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, dC, 16, cudaMemcpyDeviceToHost);
        printf("tcgen05.ld.32x32b.x16 + wait + 16 FADD: %.1f cycles per x16 load per warp (4 warps concurrently)\n",
               (double)hc[0] / (reps * 8.0));
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
