// probe_pipes.cu -- micro-benchmarks that size the design (run on the B200 via gpurun):
//   * legacy mma.sync TF32 (m16n8k8) issue rate per SM  -> is 3xTF32 on HMMA worth it vs FFMA?
//   * FFMA register-operand rate                        -> fp32 roofline denominator cross-check
//   * LDS.128 warp-broadcast rate                       -> the pass kernel's W operand path
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_pipes tools/probe_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mma_tf32(float* out, int iters) {
    float c[8][4];
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f801000u, 0x3f802000u, 0x3f803000u};
    unsigned b[2] = {0x3f804000u, 0x3f805000u + threadIdx.x};
#pragma unroll
    for (int t = 0; t < 8; ++t) for (int e = 0; e < 4; ++e) c[t][e] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int t = 0; t < 8; ++t) for (int e = 0; e < 4; ++e) s += c[t][e];
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

// every lane reads the same 16 bytes (warp broadcast), 8 independent loads per iteration
__global__ void k_lds128_bcast(float* out, int iters) {
    __shared__ float4 sm[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) sm[t] = make_float4(t, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = (threadIdx.x >> 5);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            float4 v = sm[(idx + t * 8) & 255];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 1;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1234.5f) out[0] = acc.x;
}

// same but 4 FFMA per loaded float4 component pair (the SPT=2 inner loop shape: 8 FFMA per LDS.128)
template <int FPL>
__global__ void k_lds_ffma(float* out, int iters) {
    __shared__ float4 sm[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) sm[t] = make_float4(1e-3f * t, 1e-3f, 2e-3f, 3e-3f);
    __syncthreads();
    float acc[FPL][4];
    float u[FPL];
    for (int s = 0; s < FPL; ++s) { u[s] = 1.f + 1e-3f * (threadIdx.x + s); for (int e = 0; e < 4; ++e) acc[s][e] = 0.f; }
    int idx = (threadIdx.x >> 5);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            float4 v = sm[(idx + t * 8) & 255];
#pragma unroll
            for (int s = 0; s < FPL; ++s) {
                acc[s][0] = fmaf(u[s], v.x, acc[s][0]); acc[s][1] = fmaf(u[s], v.y, acc[s][1]);
                acc[s][2] = fmaf(u[s], v.z, acc[s][2]); acc[s][3] = fmaf(u[s], v.w, acc[s][3]);
            }
        }
        idx += 1;
    }
    float s = 0.f;
    for (int q = 0; q < FPL; ++q) for (int e = 0; e < 4; ++e) s += acc[q][e];
    if (s == 1234.5f) out[0] = s;
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
    for (int wps : {4, 8, 16, 32}) {
        const int iters = 20000, blocks = sms, threads = wps * 32;
        float ms = time_ms([&] { k_mma_tf32<<<blocks, threads>>>(out, iters); });
        double mma = 8.0 * iters * wps;                 // per SM
        double cyc = ms * 1e-3 * hz;
        printf("mma.sync m16n8k8 tf32: %2d warps/SM: %.3f ms  %.2f MMA/clk/SM  -> %.1f dense TFLOP/s tf32 (3xTF32: %.1f)\n",
               wps, ms, mma / cyc, mma * sms * 2048.0 / (ms * 1e-3) / 1e12, mma * sms * 2048.0 / (ms * 1e-3) / 1e12 / 3);
    }
    for (int wps : {8, 16, 32, 64}) {
        const int iters = 5000, threads = 256, blocks = sms * (wps / 8);
        float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters, 0.999f, 1e-4f); });
        double fma = 64.0 * iters * wps * 32;           // per SM
        printf("FFMA reg-operand: %2d warps/SM: %.3f ms  %.1f FMA/clk/SM  -> %.1f TFLOP/s\n", wps, ms,
               fma / (ms * 1e-3 * hz), fma * sms * 2 / (ms * 1e-3) / 1e12);
    }
    for (int wps : {8, 16, 32}) {
        const int iters = 20000, threads = 256, blocks = sms * (wps / 8);
        float ms = time_ms([&] { k_lds128_bcast<<<blocks, threads>>>(out, iters); });
        double lds = 8.0 * iters * wps;
        printf("LDS.128 broadcast: %2d warps/SM: %.3f ms  %.3f LDS.128/clk/SM\n", wps, ms, lds / (ms * 1e-3 * hz));
    }
    for (int wps : {8, 16}) {
        const int iters = 10000, threads = 256, blocks = sms * (wps / 8);
        float m1 = time_ms([&] { k_lds_ffma<1><<<blocks, threads>>>(out, iters); });
        float m2 = time_ms([&] { k_lds_ffma<2><<<blocks, threads>>>(out, iters); });
        float m4 = time_ms([&] { k_lds_ffma<4><<<blocks, threads>>>(out, iters); });
        auto rate = [&](float ms, int f) { return 8.0 * iters * wps * 32 * 4 * f / (ms * 1e-3 * hz); };
        printf("LDS.128+FFMA, %2d warps/SM: 4/8/16 FFMA per LDS: %.1f / %.1f / %.1f FMA/clk/SM\n", wps, rate(m1, 1), rate(m2, 2), rate(m4, 4));
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
