// probe_r2.cu -- round-2 sizing probes for the gradient sweep redesign (run on the B200 via gpurun):
//   * fma.rn.f32x2 (FFMA2) rate per SM        -> does the packed form double the fp32 pipe, or only halve issue slots?
//   * tcgen05.st 32x32b.x32 rate per SM       -> cost of parking a per-sample operand (G) in TMEM for a TS-mode MMA
//   * st.shared.v4 rate per SM                -> cost of parking it in shared memory instead
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_r2 tools/probe_r2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma2(float* out, int iters, float m0, float c0) {
    uint64_t a[8], m[2], c[2];
    for (int t = 0; t < 8; ++t) {
        float lo = 1.f + 1e-3f * (threadIdx.x + t), hi = 1.f + 2e-3f * (threadIdx.x + t);
        asm("mov.b64 %0, {%1, %2};" : "=l"(a[t]) : "f"(lo), "f"(hi));
    }
    for (int t = 0; t < 2; ++t) {
        float x = m0 - 1e-6f * (threadIdx.x + t), y = c0 + 1e-7f * t;
        asm("mov.b64 %0, {%1, %1};" : "=l"(m[t]) : "f"(x));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c[t]) : "f"(y));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int t = 0; t < 8; ++t) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[t]) : "l"(m[t & 1]), "l"(c[(t >> 1) & 1]));
    }
    float s = 0.f;
    for (int t = 0; t < 8; ++t) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[t])); s += lo + hi; }
    if (s == 1234.5f) out[0] = s;
}

__global__ void k_ffma(float* out, int iters, float m0, float c0) {
    float a[16], m[4], c[4];
    for (int t = 0; t < 16; ++t) a[t] = 1.f + 1e-3f * (threadIdx.x + t);
    for (int t = 0; t < 4; ++t) { m[t] = m0 - 1e-6f * (threadIdx.x + t); c[t] = c0 + 1e-7f * t; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int t = 0; t < 16; ++t) a[t] = fmaf(a[t], m[t & 3], c[(t >> 2) & 3]);
    }
    float s = 0.f;
    for (int t = 0; t < 16; ++t) s += a[t];
    if (s == 1234.5f) out[0] = s;
}

// 4 warps of one CTA per SM store 32 columns x 32 lanes each, `iters` x 8 times, then one wait::st
__global__ void __launch_bounds__(256, 1) k_sttm(float* out, int iters, int nwarps) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t r[32];
    for (int t = 0; t < 32; ++t) r[t] = threadIdx.x * 32 + t;
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                    ::"r"(tb + (uint32_t)(q * 32)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
                      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
                      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512));
    if (r[0] == 0xffffffffu) out[0] = 1.f;
}

// every thread stores 16 bytes; a warp covers 512 contiguous bytes (conflict-free: 4 wavefronts)
__global__ void __launch_bounds__(256, 1) k_sts128(float* out, int iters) {
    extern __shared__ float4 sm4[];
    const float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    float4* p = sm4 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(p + q * 256)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    __syncthreads();
    if (sm4[threadIdx.x].x == -1.f) out[0] = 1.f;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 1024);
    const double hz = clk * 1e3;
    printf("SMs %d, clock %.0f MHz (attr)\n", sms, clk / 1e3);
    for (int wps : {8, 16, 32}) {
        const int iters = 5000, threads = 256, blocks = sms * (wps / 8);
        float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters, 0.999f, 1e-4f); });
        double fma = 64.0 * iters * wps * 32;
        printf("FFMA  reg-operand: %2d warps/SM: %.3f ms  %.1f FMA/clk/SM  %.2f warp-instr/clk/SM\n", wps, ms, fma / (ms * 1e-3 * hz), fma / 32 / (ms * 1e-3 * hz));
        ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, iters, 0.999f, 1e-4f); });
        fma = 64.0 * iters * wps * 32;          // 32 FFMA2 per iteration = 64 FMA per thread
        printf("FFMA2 reg-operand: %2d warps/SM: %.3f ms  %.1f FMA/clk/SM  %.2f warp-instr/clk/SM\n", wps, ms, fma / (ms * 1e-3 * hz), fma / 64 / (ms * 1e-3 * hz));
    }
    for (int nw : {4, 8}) {
        const int iters = 4000;
        float ms = time_ms([&] { k_sttm<<<sms, 256>>>(out, iters, nw); });
        double bytes = (double)iters * 8 * 32 * 32 * 4 * nw;       // per SM
        printf("tcgen05.st 32x32b.x32: %d warps/SM: %.3f ms  %.1f B/clk/SM  (%.1f cycles per 64 KB)\n", nw, ms, bytes / (ms * 1e-3 * hz), 65536.0 / (bytes / (ms * 1e-3 * hz)));
    }
    {
        const int iters = 4000;
        cudaFuncSetAttribute(k_sts128, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 256 * 16);
        float ms = time_ms([&] { k_sts128<<<sms, 256, 8 * 256 * 16>>>(out, iters); });
        double bytes = (double)iters * 8 * 256 * 16;
        printf("st.shared.v4 (8 warps/SM): %.3f ms  %.1f B/clk/SM\n", ms, bytes / (ms * 1e-3 * hz));
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
