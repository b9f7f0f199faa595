// probe_tc2.cu -- cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N, operand source and
// accumulator dependence.  Answers VERDICT r1 item 3: is a narrow-N UMMA (the gradient sweep's dW / du products have
// one side only 8 or 16 wide) cheap enough to move k_grad_mma from mma.sync to tcgen05?
//
//   mode SS : A and B from shared memory (no-swizzle K-major core matrices, as in caps_pass_tc.cu)
//   mode TS : A from TMEM (128 lanes x 8 columns), B from shared memory
//   dep     : every MMA accumulates into the SAME TMEM columns (a dependent chain)
//   indep   : MMAs rotate over 4 accumulators
// One CTA, one issuing thread, `reps` back-to-back MMAs, one commit, clock64 around issue..completion.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_tc2 tools/probe_tc2.cu
#include <cstdint>
#include <cstdio>
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

// NW issuing warps (one elected thread each), every warp on its own accumulator columns; the issue loop is
// unrolled 8x with loop-invariant descriptors so that a slow issuing thread does not hide the tensor-pipe cost
// (the first version of this probe measured ~97 cycles per MMA for every N: that was its own issue loop).
template <int TS, int NW>
__global__ void __launch_bounds__(128, 1) k_probe(int N, int indep, int reps, long long* out) {
    __shared__ __align__(1024) float sA[2 * 128 * 4];
    __shared__ __align__(1024) float sB[2 * 256 * 4];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 2 * 128 * 4; e += 128) sA[e] = 1.0f + 0.001f * (e & 15);
    for (int e = tid; e < 2 * 256 * 4; e += 128) sB[e] = 0.5f;
    if (tid == 0) for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[w])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    {   // TS mode: the A tile (TMEM columns 480..487 of this warp's 32 lanes) = 1.0
        const uint32_t one = __float_as_uint(1.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};"
                     ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + 480), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (warp < NW && lane == 0) {
        const uint64_t a = make_desc(smem_u32(sA), 128 * 16, 128), b = make_desc(smem_u32(sB), N * 16, 128);
        // this warp's accumulator region: 480 / NW columns; indep -> 2 or 4 slots inside it
        const int region = (480 / NW) & ~15;
        const int nslot = indep ? ((4 * N <= region) ? 4 : (2 * N <= region) ? 2 : 1) : 1;
        const uint32_t d0 = tmem_base + (uint32_t)(warp * region);
        const uint32_t d1 = d0 + (uint32_t)((nslot > 1 ? 1 : 0) * N), d2 = d0 + (uint32_t)((nslot > 2 ? 2 : 0) * N), d3 = d0 + (uint32_t)((nslot > 2 ? 3 : nslot > 1 ? 1 : 0) * N);
        t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
            if (TS) {
                mma_ts(d0, tmem_base + 480, b, idesc, 1); mma_ts(d1, tmem_base + 480, b, idesc, 1); mma_ts(d2, tmem_base + 480, b, idesc, 1); mma_ts(d3, tmem_base + 480, b, idesc, 1);
                mma_ts(d0, tmem_base + 480, b, idesc, 1); mma_ts(d1, tmem_base + 480, b, idesc, 1); mma_ts(d2, tmem_base + 480, b, idesc, 1); mma_ts(d3, tmem_base + 480, b, idesc, 1);
            } else {
                mma_ss(d0, a, b, idesc, 1); mma_ss(d1, a, b, idesc, 1); mma_ss(d2, a, b, idesc, 1); mma_ss(d3, a, b, idesc, 1);
                mma_ss(d0, a, b, idesc, 1); mma_ss(d1, a, b, idesc, 1); mma_ss(d2, a, b, idesc, 1); mma_ss(d3, a, b, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    }
    if (warp < NW) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[warp])) : "memory");
        if (lane == 0) { t1 = clock64(); out[warp] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int TS, int NW>
double run(int N, int indep, int reps, long long* d) {
    k_probe<TS, NW><<<1, 128>>>(N, indep, 64, d);
    k_probe<TS, NW><<<1, 128>>>(N, indep, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d ts=%d nw=%d: %s\n", N, TS, NW, cudaGetErrorString(e)); exit(1); }
    long long c[4];
    cudaMemcpy(c, d, 32, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < NW; ++w) mx = c[w] > mx ? c[w] : mx;
    return (double)mx / ((double)reps * NW);      // SM cycles per MMA (all issuers together)
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    const int reps = 4096;
    printf("# SM cycles per tcgen05.mma kind::tf32, M=128, K=8: NW issuing warps x %d back-to-back MMAs each (unrolled x8), incl. completion\n", reps);
    printf("# %4s %5s %6s %3s %10s %12s\n", "N", "mode", "chain", "NW", "cyc/MMA", "MAC/clk/SM");
    const int Ns[] = {16, 32, 64, 96, 128, 256};
    for (int N : Ns)
        for (int ts = 0; ts < 2; ++ts)
            for (int indep = 0; indep < 2; ++indep)
                for (int nw = 1; nw <= 4; nw *= 2) {
                    if (N * nw > 480) continue;
                    if (indep && 2 * N * nw > 480) continue;
                    double per = 0;
                    if (ts) per = nw == 1 ? run<1, 1>(N, indep, reps, d) : nw == 2 ? run<1, 2>(N, indep, reps, d) : run<1, 4>(N, indep, reps, d);
                    else per = nw == 1 ? run<0, 1>(N, indep, reps, d) : nw == 2 ? run<0, 2>(N, indep, reps, d) : run<0, 4>(N, indep, reps, d);
                    printf("  %4d %5s %6s %3d %10.1f %12.0f\n", N, ts ? "TS" : "SS", indep ? "indep" : "dep", nw, per, 128.0 * N * 8 / per);
                }
    return 0;
}
