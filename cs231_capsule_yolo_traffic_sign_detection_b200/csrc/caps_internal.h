// caps_internal.h -- shared between the translation units of libcaps_routing.so (not installed).
#pragma once
#include <atomic>
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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: cache what was set per device
// (a process-wide "already set" flag would leave a second GPU of the same process without the opt-in).
constexpr int kMaxDevices = 64;
struct SmemAttrCache { std::atomic<size_t> bytes[kMaxDevices]; };      // static storage: zero-initialised
int ensure_dyn_smem(const void* func, size_t smem, SmemAttrCache& cache);     // caps_api.cu
#define CAPS_SET_SMEM(kern, smem)                                                                 \
    do {                                                                                          \
        static ::caps::SmemAttrCache cache__;                                                     \
        const int rc__ = ::caps::ensure_dyn_smem(reinterpret_cast<const void*>(kern), (smem), cache__); \
        if (rc__) return rc__;                                                                    \
    } while (0)

inline size_t round64(size_t n) { return (n + 63) & ~(size_t)63; }
inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

constexpr int kMaxSplits = 32;     // splits of the i range any sweep may use (sizes the partial-sum slots of the workspace)

// Everything the launchers need to know about one (dims, tuning) combination.  The workspace LAYOUT (offsets, total)
// is a pure function of (dims, with_grad): tuning knobs only pick engines and grid shapes, so a knob changed between
// caps_route_workspace_bytes / forward / backward can never shift an offset.  Which engines the forward actually
// used (and therefore which operand copies the workspace holds) is recorded per workspace and replayed by the
// backward (caps_api.cu: forward records).
struct Plan {
    int B, N, C, K, D, R, Reff, DP, JW, JG, SPT, nbt, ntg, IS, i_per_split, M;
    int CSmax;                         // capsules per coefficient row the layout reserves (C, or C rounded up to 8 where the fused sweep may run)
    bool pad_w, with_grad, use_tc, tc_ok, fused, c1;
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
int launch_prep_u_tc(const Plan& pl, const float* u, float* ua, float* ut, cudaStream_t st);   // ut nullable: also the lane-tile copy
int launch_prep_w_tc(const Plan& pl, const float* W, float* wb, cudaStream_t st);
int launch_pass_tc(const Plan& pl, int mode, const PassParams& pp, const float* ua, const float* wb, cudaStream_t st);
// caps_sweep_fused.cu: one cluster-fused sweep per routing iteration (logits -> softmax -> weighted sum)
extern int g_fs_dbg;                    // timing experiments only
extern int g_fs_ws;                     // tuning knob "fsws": warp-specialised epilogue of the fused sweep (default 1)
bool fused_supported(const Plan& pl);
bool fused_shape_ok(int C, int DP, bool tc_ok);       // depends on the dims only (sizes the coefficient rows)
int fused_cluster_capacity(int jg, bool bwd);
int fused_pick_splits(const Plan& pl, bool bwd, int forced);
int launch_sweep_fused(const Plan& pl, bool bwd, const float* ua, const float* wb, const float* X, const float* coef_in,
                       const float* beta_in, float* coef_out, float* part, int IS, cudaStream_t st);
int launch_coef_public(const Plan& pl, const float* coef, int CS, float* c_pub, cudaStream_t st);
// caps_c1.cu: one class capsule (the DarkCapsuleNet head): the layer is one skinny GEMM + squash
extern int g_c1_version;                   // tuning knob "c1v"
bool c1_supported(int N, int C, int K, int D);
size_t c1_part_floats(int B, int N, int K, int D);
int launch_c1_forward(const float* u, const float* W, float* v, float* s_save, int B, int N, int K, int D, cudaStream_t st);
int launch_c1_backward(const float* u, const float* W, const float* s_save, const float* grad_v, const int64_t* y,
                       float margin_scale, const float* loss_grad, float* du, float* dW, float* part,
                       int B, int N, int K, int D, cudaStream_t st, int* launches);

}  // namespace caps
