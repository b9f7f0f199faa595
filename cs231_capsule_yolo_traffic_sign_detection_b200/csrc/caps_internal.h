// caps_internal.h -- shared between the translation units of libcaps_routing.so (not installed).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "caps_kernels.cuh"
#include "caps_routing.h"

namespace caps {

int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ::caps::fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define LAUNCH_CHECK() CUDA_TRY(cudaGetLastError())

inline size_t round64(size_t n) { return (n + 63) & ~(size_t)63; }
inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// Everything the launchers need to know about one (dims, tuning) combination; workspace offsets
// are a pure function of the dims so forward and backward agree without any shared state.
struct Plan {
    int B, N, C, K, D, R, Reff, DP, JW, JG, SPT, nbt, ntg, IS, i_per_split, M;
    bool pad_w, with_grad, use_tc;
    size_t xs, cs, us;                 // floats per X / coef / ut array
    // offsets (floats) into the workspace
    size_t o_ua, o_wb, o_ut, o_wp, o_vsum, o_s, o_v, o_part, o_c, o_beta, o_tmp, o_ds, o_dupart, total;
};

int launch_pass(const Plan& pl, int mode, const PassParams& pp, cudaStream_t st);   // caps_pass.cu
int launch_grad(const Plan& pl, const GradParams& gp, cudaStream_t st);             // caps_grad.cu
int launch_grad_mma(const Plan& pl, const GradParams& gp, cudaStream_t st);         // caps_grad_mma.cu (D == 16, C >= 7)
int grad_mma_jw(const Plan& pl);           // output (pseudo-)capsules per CTA it will use (8 or 11)
int grad_mma_parts(const Plan& pl);        // du partials it writes: cdiv(C * D/16, jw)
extern int g_grad_jw;                      // tuning knob "gradjw": 0 = auto
// caps_pass_tc.cu: tcgen05 pass kernel (D == 16 only) and its operand preparation
extern int g_tc_dbg;                    // timing experiments only
extern int g_tc_stages;                 // smem ring depth of the tcgen05 pass kernel (tuning knob "tcstages")
size_t tc_ua_floats(int B, int N);
size_t tc_wb_floats(int N, int C, int D);
int tc_jw(int D);
int launch_prep_u_tc(const Plan& pl, const float* u, float* ua, cudaStream_t st);
int launch_prep_w_tc(const Plan& pl, const float* W, float* wb, cudaStream_t st);
int launch_pass_tc(const Plan& pl, int mode, const PassParams& pp, const float* ua, const float* wb, cudaStream_t st);

}  // namespace caps
