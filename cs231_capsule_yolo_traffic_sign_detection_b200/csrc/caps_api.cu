// caps_routing.cu -- C ABI (include/caps_routing.h) over the sm_100a kernels in caps_kernels.cuh.
//
// Host-side orchestration of the routing forward (reference models.py:70-79) and its backward:
//   forward :  prep(u) ; A0 ; squash ; { L ; softmax ; A ; squash } x (R-1)
//   backward:  dsquash(top, +margin grad) ; { L ; softmax_bwd ; A ; dsquash } x (R-1) ; grad ; reduce(du)
// Everything is enqueued on the caller's stream; no host synchronisation, no allocation (the
// caller's workspace is carved deterministically from the dims).
#include "caps_internal.h"

#include <algorithm>
#include <atomic>
#include <mutex>

namespace caps {
thread_local char g_err[512] = "";
int ensure_dyn_smem(const void* func, size_t smem, SmemAttrCache& cache) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) {            // beyond the cache: set unconditionally (cheap, just not free)
        CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }
    if (smem > cache.bytes[dev].load(std::memory_order_relaxed)) {
        CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache.bytes[dev].store(smem, std::memory_order_relaxed);      // racing threads both set it: idempotent
    }
    return 0;
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace caps

namespace {

using namespace caps;

int g_tune_spt = 0;      // 0 = auto
int g_tune_isplit = 0;   // 0 = auto
int g_tune_gradmma = 1;  // 1 = mma.sync gradient kernel where it applies (D == 16, C >= 7)
int g_tune_hostmb = 0;   // caps_route_step_host: 0 = auto (3 micro-batches from B >= 2048), 1 = single batch
int g_tune_sbstaged = 1;  // softmax backward through the staged kernel (16 < C <= 48)
int g_tune_tc = 1;       // 1 = use the tcgen05 pass kernel where it applies, 0 = FFMA kernel only
int g_tune_c1 = 1;       // 1 = dedicated kernels for one class capsule (the DarkCapsuleNet head)
int g_tune_fused = 1;    // 1 = cluster-fused sweep (logits -> softmax -> weighted sum in one kernel) where it applies

// ---- forward records ---------------------------------------------------------------------------
// caps_route_forward notes, per (device, workspace pointer), the dims it ran with and the engines it used (which
// operand copies the workspace now holds); caps_route_backward looks the record up, returns CAPS_E_STATE when there
// is none / the dims differ / the forward ran without with_grad, and replays the recorded engine choices instead of
// reading the tuning knobs again.  Host-side bookkeeping only (mutex-guarded, fixed size, oldest entry recycled).
struct FwdRecord { const void* ws; int dev, B, N, C, K, D, R; bool with_grad, use_tc, fused, c1; unsigned long stamp; };
constexpr int kFwdRecords = 256;
FwdRecord g_records[kFwdRecords];
unsigned long g_record_clock = 0;
std::mutex g_record_mu;

void record_forward(const void* ws, const Plan& pl) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_record_mu);
    int slot = 0;
    for (int e = 0; e < kFwdRecords; ++e) {
        if (g_records[e].ws == ws && g_records[e].dev == dev) { slot = e; break; }
        if (g_records[e].stamp < g_records[slot].stamp) slot = e;
    }
    g_records[slot] = FwdRecord{ws, dev, pl.B, pl.N, pl.C, pl.K, pl.D, pl.R, pl.with_grad, pl.use_tc, pl.fused, pl.c1, ++g_record_clock};
}
bool lookup_forward(const void* ws, FwdRecord& out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_record_mu);
    for (int e = 0; e < kFwdRecords; ++e)
        if (g_records[e].stamp != 0 && g_records[e].ws == ws && g_records[e].dev == dev) { out = g_records[e]; return true; }
    return false;
}

// ---- launch accounting (bench.py: gpu_launches, per-kernel-class CUDA-event times) -------------
enum KClass { kcLayout = 0, kcPassA0, kcPassL, kcPassA, kcSquash, kcSoftmax, kcGrad, kcReduceDu, kcLoss, kcOther, kcFused, kcC1, kcCount };
std::atomic<long> g_launches{0};
int g_prof_on = 0;
constexpr int kProfPool = 8192;
cudaEvent_t g_prof_ev[kProfPool][2];
int g_prof_cls[kProfPool];
int g_prof_n = 0;
bool g_prof_init = false;

struct LaunchScope {          // brackets one kernel launch on `st`
    int slot = -1;
    cudaStream_t st;
    LaunchScope(int cls, cudaStream_t s) : st(s) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (g_prof_on && g_prof_n < kProfPool) {
            if (!g_prof_init) {
                for (int i = 0; i < kProfPool; ++i) { cudaEventCreate(&g_prof_ev[i][0]); cudaEventCreate(&g_prof_ev[i][1]); }
                g_prof_init = true;
            }
            slot = g_prof_n++;
            g_prof_cls[slot] = cls;
            cudaEventRecord(g_prof_ev[slot][0], st);
        }
    }
    ~LaunchScope() { if (slot >= 0) cudaEventRecord(g_prof_ev[slot][1], st); }
};
int pass_class(int mode) { return mode == kModeAUniform ? kcPassA0 : mode == kModeL ? kcPassL : kcPassA; }

int pad_dim(int D) { return D <= 8 ? 8 : D <= 16 ? 16 : D <= 24 ? 24 : D <= 32 ? 32 : D <= 48 ? 48 : 0; }

bool make_plan(Plan& p, int B, int N, int C, int K, int D, int R, int with_grad) {
    if (B < 0 || N <= 0 || C <= 0 || K != 8 || D <= 0 || D > 48 || R < 1 || R > 5 || C > 1024) return false;
    p.B = B; p.N = N; p.C = C; p.K = K; p.D = D; p.R = R;
    p.with_grad = with_grad != 0;
    p.Reff = (C == 1) ? 1 : R;          // one class capsule: softmax == 1, iterations are no-ops
    p.M = 2 * p.Reff - 1;
    p.DP = pad_dim(D);
    p.pad_w = (p.DP != D);
    p.JW = C >= 7 ? 8 : C >= 3 ? 4 : 1;
    if (C == 2) p.JW = 4;
    p.JG = cdiv(C, p.JW);
    p.nbt = B > 0 ? cdiv(B, 32) : 1;
    int spt = g_tune_spt;
    if (spt != 1 && spt != 2 && spt != 4) spt = (p.DP <= 16) ? 2 : 1;
    if (p.DP > 16 && spt > 2) spt = 2;
    if (p.DP > 32) spt = 1;
    while (spt > 1 && p.nbt < spt) spt >>= 1;
    p.SPT = spt;
    p.ntg = cdiv(p.nbt, p.SPT);
    // tc_jw(DP) capsules per tcgen05 CTA; D < DP (21 -> 24, 9..15 -> 16, ...) runs on the zero-padded W copy
    p.tc_ok = p.DP >= 16 && C >= 2 && (p.DP != 16 || C >= 4);      // shape is eligible: sizes the layout
    p.use_tc = g_tune_tc != 0 && p.tc_ok;                          // engine choice: a knob
    p.fused = g_tune_fused != 0 && fused_supported(p);
    p.c1 = g_tune_c1 != 0 && c1_supported(N, C, K, D);
    // split the i range until the grid fills the machine: the FMA kernel wants ~4 CTAs per SM, the tcgen05
    // kernel owns an SM (all of TMEM), so one wave of its CTAs (128-sample quads x 8-capsule groups) is enough
    const long ctas = p.use_tc ? (long)cdiv(C, tc_jw(p.DP)) * cdiv(p.nbt, 4) : (long)p.JG * p.ntg;
    int is = g_tune_isplit > 0 ? g_tune_isplit : cdiv(4 * 148, ctas);
    if (g_tune_isplit <= 0 && p.use_tc) {
        // smallest split count (<= 32) whose grid wastes the least of its last wave of 148 CTAs
        double best = -1.0;
        for (int cand = 1; cand <= kMaxSplits; ++cand) {
            const long grid = ctas * cand;
            const double eff = (double)grid / (double)(((grid + 147) / 148) * 148);
            if (eff > best + 0.05) { best = eff; is = cand; }
        }
    }
    const int max_is = cdiv(N, kPassIC);
    if (is > max_is) is = max_is;
    if (is > kMaxSplits) is = kMaxSplits;
    if (is < 1) is = 1;
    p.i_per_split = cdiv(cdiv(N, is), kPassIC) * kPassIC;
    p.IS = cdiv(N, p.i_per_split);
    p.xs = round64((size_t)p.nbt * C * p.DP * 32);
    p.CSmax = fused_shape_ok(C, p.DP, p.tc_ok) ? cdiv(C, 8) * 8 : C;       // the fused sweep pads coefficient rows to whole capsule groups
    p.cs = round64((size_t)p.nbt * N * p.CSmax * 32);
    p.us = round64((size_t)p.nbt * N * K * 32);
    // ---- layout: depends on (dims, with_grad) only -------------------------------------------------------------
    const int part_slots = std::min(kMaxSplits, max_is);
    size_t o = 0;
    p.o_ua = o; o += p.tc_ok ? round64(tc_ua_floats(B, N)) : 0;
    p.o_wb = o; o += p.tc_ok ? round64(tc_wb_floats(N, C, p.DP)) : 0;
    p.o_ut = o; o += p.us;
    p.o_wp = o; o += p.pad_w ? round64((size_t)N * C * K * p.DP) : 0;
    p.o_vsum = o; o += p.xs;
    p.o_s = o; o += p.xs * p.Reff;
    p.o_v = o; o += p.xs * p.Reff;
    p.o_part = o; o += p.xs * part_slots;
    p.o_c = o; o += p.cs * (p.with_grad ? (p.Reff > 1 ? p.Reff - 1 : 0) : (p.Reff > 1 ? 1 : 0));
    p.o_beta = o; o += p.with_grad ? p.cs * (p.Reff > 1 ? p.Reff - 1 : 0) : 0;
    p.o_tmp = o; o += p.with_grad && p.Reff > 1 ? p.cs : 0;
    p.o_ds = o; o += p.with_grad ? p.xs * p.Reff : 0;
    p.o_dupart = o; o += p.with_grad ? std::max(p.us * std::max(p.JG, cdiv(C * cdiv(p.DP, 16), 8)),
                                                 c1_supported(N, C, K, D) ? round64(c1_part_floats(B, N, K, D)) : (size_t)0) : 0;   // FMA kernel: JG partials; mma kernel: <= cdiv(C * D/16, 8)
    p.total = o;
    return true;
}

#define DISPATCH_DP(pl, CALL)                          \
    switch ((pl).DP) {                                 \
        case 8: { constexpr int DP_ = 8; CALL; } break;   \
        case 16: { constexpr int DP_ = 16; CALL; } break; \
        case 24: { constexpr int DP_ = 24; CALL; } break; \
        case 32: { constexpr int DP_ = 32; CALL; } break; \
        case 48: { constexpr int DP_ = 48; CALL; } break; \
    }

int launch_squash(const Plan& pl, const float* part, int IS, float scale, float* s_out, float* v_out, float* vsum,
                  int accumulate, float* v_pub, cudaStream_t st) {
    const long n = (long)pl.nbt * pl.C * 32;
    LaunchScope ls_(kcSquash, st);
    DISPATCH_DP(pl, (k_squash<DP_><<<cdiv(n, 128), 128, 0, st>>>(part, IS, pl.xs, scale, s_out, v_out, vsum,
                                                                  accumulate, v_pub, pl.B, pl.C, pl.D, pl.nbt)));
    LAUNCH_CHECK();
    return 0;
}

int launch_dsquash(const Plan& pl, const float* part, int IS, const float* grad_v, const int64_t* y, float mscale,
                   const float* lgrad, const float* v_last, const float* s_in, float* ds_out, float out_scale, cudaStream_t st) {
    const long n = (long)pl.nbt * pl.C * 32;
    LaunchScope ls_(kcSquash, st);
    DISPATCH_DP(pl, (k_dsquash<DP_><<<cdiv(n, 128), 128, 0, st>>>(part, IS, pl.xs, grad_v, y, mscale, lgrad, v_last,
                                                                   s_in, ds_out, out_scale, pl.B, pl.C, pl.D, pl.nbt)));
    LAUNCH_CHECK();
    return 0;
}

// one pass on whichever engine the plan selected
int run_pass(const Plan& pl, int mode, const PassParams& pp, const float* ws_base, cudaStream_t st) {
    LaunchScope ls_(pass_class(mode), st);
    if (pl.use_tc) return launch_pass_tc(pl, mode, pp, ws_base + pl.o_ua, ws_base + pl.o_wb, st);
    return launch_pass(pl, mode, pp, st);
}

bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

}  // namespace

// ==============================================================================================
// C ABI
// ==============================================================================================
extern "C" {

int caps_abi_version(void) { return CAPS_ABI_VERSION; }

const char* caps_last_error(void) { return g_err; }

long caps_kernel_launch_count(void) { return g_launches.load(); }

int caps_profile_collect(double* ms_by_class, long* count_by_class, int n_classes) {
    for (int c = 0; c < n_classes; ++c) { if (ms_by_class) ms_by_class[c] = 0.0; if (count_by_class) count_by_class[c] = 0; }
    for (int i = 0; i < g_prof_n; ++i) {
        CUDA_TRY(cudaEventSynchronize(g_prof_ev[i][1]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, g_prof_ev[i][0], g_prof_ev[i][1]));
        const int c = g_prof_cls[i];
        if (c < n_classes) { if (ms_by_class) ms_by_class[c] += ms; if (count_by_class) count_by_class[c] += 1; }
    }
    g_prof_n = 0;
    return 0;
}

int caps_fma_peak(int iters, float* ms_out, double* flops_out, void* stream) {
    if (iters <= 0 || !ms_out || !flops_out) return fail(CAPS_E_BADARG, "caps_fma_peak: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, sizeof(float) * 1024));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    k_fma_peak<<<blocks, threads, 0, st>>>(sink, iters / 8 + 1, 0.999f, 1e-4f);      // warm-up
    CUDA_TRY(cudaEventRecord(e0, st));
    k_fma_peak<<<blocks, threads, 0, st>>>(sink, iters, 0.999f, 1e-4f);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(ms_out, e0, e1));
    *flops_out = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 0;
}

int caps_set_tuning(const char* name, int value) {
    if (name == nullptr) return fail(CAPS_E_BADARG, "caps_set_tuning: null name");
    if (!strcmp(name, "profile")) { g_prof_on = value != 0; if (!g_prof_on) g_prof_n = 0; return 0; }
    if (!strcmp(name, "spt")) {
        if (value != 0 && value != 1 && value != 2 && value != 4) return fail(CAPS_E_BADARG, "spt must be 0,1,2,4");
        g_tune_spt = value;
        return 0;
    }
    if (!strcmp(name, "tcdbg")) { g_tc_dbg = value; return 0; }
    if (!strcmp(name, "fsdbg")) { g_fs_dbg = value; return 0; }
    if (!strcmp(name, "fsws")) { g_fs_ws = value != 0; return 0; }
    if (!strcmp(name, "fsinfo")) {          // prints what the fused sweep would do for C = value (diagnostics)
        fprintf(stderr, "[caps] fused sweep: cluster of %d CTAs, capacity fwd %d / bwd %d clusters\n", cdiv(value, 8),
                fused_cluster_capacity(cdiv(value, 8), false), fused_cluster_capacity(cdiv(value, 8), true));
        return 0;
    }
    if (!strcmp(name, "tcstages")) { if (value < 2 || value > 12) return fail(CAPS_E_BADARG, "tcstages must be in [2,12]"); g_tc_stages = value; return 0; }
    if (!strcmp(name, "gradmma")) { g_tune_gradmma = value != 0; return 0; }
    if (!strcmp(name, "gradjw")) {
        if (value != 0 && value != 8 && value != 11) return fail(CAPS_E_BADARG, "gradjw must be 0, 8 or 11");
        g_grad_jw = value;
        return 0;
    }
    if (!strcmp(name, "sbstaged")) { g_tune_sbstaged = value != 0; return 0; }
    if (!strcmp(name, "hostmb")) { g_tune_hostmb = value; return 0; }
    if (!strcmp(name, "tc")) { g_tune_tc = value != 0; return 0; }
    if (!strcmp(name, "fused")) { g_tune_fused = value != 0; return 0; }
    if (!strcmp(name, "c1")) { g_tune_c1 = value != 0; return 0; }
    if (!strcmp(name, "c1v")) { if (value != 1 && value != 2) return fail(CAPS_E_BADARG, "c1v must be 1 or 2"); g_c1_version = value; return 0; }
    if (!strcmp(name, "isplit")) {
        if (value < 0 || value > kMaxSplits) return fail(CAPS_E_BADARG, "isplit must be in [0,%d]", kMaxSplits);
        g_tune_isplit = value;
        return 0;
    }
    return fail(CAPS_E_BADARG, "caps_set_tuning: unknown knob '%s'", name);
}

size_t caps_route_workspace_bytes(int B, int N, int C, int K, int D, int R, int with_grad) {
    Plan pl;
    if (!make_plan(pl, B, N, C, K, D, R, with_grad)) return 0;
    return pl.total * sizeof(float);
}

int caps_route_forward(const float* u, const float* W, float* v, float* c_out, void* ws, size_t ws_bytes,
                       int B, int N, int C, int K, int D, int R, int with_grad, void* stream) {
    Plan pl;
    if (!make_plan(pl, B, N, C, K, D, R, with_grad))
        return fail(B < 0 || N <= 0 || C <= 0 || D <= 0 || R < 1 ? CAPS_E_BADARG : CAPS_E_UNSUPPORTED,
                    "caps_route_forward: dims B=%d N=%d C=%d K=%d D=%d R=%d not supported (K must be 8, D<=48, R<=5)",
                    B, N, C, K, D, R);
    if (B == 0) return 0;
    if (!u || !W || !v || !ws) return fail(CAPS_E_BADARG, "caps_route_forward: null pointer");
    if (misaligned(u) || misaligned(W) || misaligned(ws))
        return fail(CAPS_E_BADARG, "caps_route_forward: u, W and ws must be 16-byte aligned");
    if (ws_bytes < pl.total * sizeof(float))
        return fail(CAPS_E_WORKSPACE, "caps_route_forward: workspace %zu < %zu bytes", ws_bytes, pl.total * sizeof(float));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* w = static_cast<float*>(ws);
    float* ut = w + pl.o_ut;
    const float* Wp = W;
    int rc;
    if (pl.c1) {
        // one class capsule: softmax == 1, the layer is v = squash(u . W) whatever n_iter is -- one kernel, u read in place
        { LaunchScope ls_(kcC1, st); rc = launch_c1_forward(u, W, v, w + pl.o_s, B, N, K, D, st); }
        if (rc) return rc;
        record_forward(ws, pl);
        if (c_out != nullptr) {
            const long n = (long)B * N * C;
            { LaunchScope ls_(kcOther, st); k_fill<<<cdiv(n, 256), 256, 0, st>>>(c_out, 1.f, n); }
            LAUNCH_CHECK();
        }
        return 0;
    }
    if (pl.pad_w) {
        const long rows = (long)N * C * K;
        { LaunchScope ls_(kcLayout, st); k_pad_w<<<cdiv(rows * pl.DP, 256), 256, 0, st>>>(W, w + pl.o_wp, rows, D, pl.DP); }
        LAUNCH_CHECK();
        Wp = w + pl.o_wp;
    }
    if (pl.use_tc) {
        // one pass over u writes the tcgen05 operand copy (tf32 hi/lo) and, when someone will read it (the gradient
        // kernels), the lane-tile copy
        { LaunchScope ls_(kcLayout, st); rc = launch_prep_u_tc(pl, u, w + pl.o_ua, pl.with_grad ? ut : nullptr, st); }
        if (rc) return rc;
        { LaunchScope ls_(kcLayout, st); rc = launch_prep_w_tc(pl, Wp, w + pl.o_wb, st); }
        if (rc) return rc;
    } else {
        const long n = (long)pl.nbt * N * 32;
        { LaunchScope ls_(kcLayout, st); k_prep_u<8><<<cdiv(n, 256), 256, 0, st>>>(u, ut, B, N, pl.nbt); }
        LAUNCH_CHECK();
    }
    record_forward(ws, pl);
    float* vsum = w + pl.o_vsum;
    float* part = w + pl.o_part;
    const int ISf = pl.fused ? fused_pick_splits(pl, false, g_tune_isplit) : 0;
    for (int r = 0; r < pl.Reff; ++r) {
        float* s_r = w + pl.o_s + pl.xs * r;
        float* v_r = w + pl.o_v + pl.xs * r;
        const bool last = (r == pl.Reff - 1);
        PassParams pp{};
        pp.ut = ut; pp.W = Wp; pp.N = N; pp.C = C; pp.nbt = pl.nbt; pp.i_per_split = pl.i_per_split;
        if (r == 0) {
            pp.out = part;
            if ((rc = run_pass(pl, kModeAUniform, pp, w, st))) return rc;
            if ((rc = launch_squash(pl, part, pl.IS, 1.f / (float)C, s_r, v_r, vsum, 0, last ? v : nullptr, st))) return rc;
            continue;
        }
        float* c_r = w + pl.o_c + (pl.with_grad ? pl.cs * (r - 1) : 0);
        if (pl.fused) {
            // logits -> softmax -> weighted sum in ONE sweep; c^r is written once, only if somebody will read it
            const bool want_c = pl.with_grad || (last && c_out != nullptr);
            {
                LaunchScope ls_(kcFused, st);
                rc = launch_sweep_fused(pl, false, w + pl.o_ua, w + pl.o_wb, vsum, nullptr, nullptr, want_c ? c_r : nullptr, part, ISf, st);
            }
            if (rc) return rc;
            if ((rc = launch_squash(pl, part, ISf, 1.f, s_r, v_r, vsum, 1, last ? v : nullptr, st))) return rc;
            if (last && c_out != nullptr) {
                LaunchScope ls_(kcOther, st);
                if ((rc = launch_coef_public(pl, c_r, pl.CSmax, c_out, st))) return rc;
            }
            continue;
        }
        pp.X = vsum; pp.out = c_r;
        if ((rc = run_pass(pl, kModeL, pp, w, st))) return rc;
        const long n = (long)pl.nbt * N * 32;
        {
            LaunchScope ls_(kcSoftmax, st);
            float* cp = last ? c_out : nullptr;
            if (C <= 16) k_softmax_reg<16><<<cdiv(n, 128), 128, 0, st>>>(c_r, cp, B, N, C, pl.nbt);
            else if (C <= 48) k_softmax_reg<48><<<cdiv(n, 128), 128, 0, st>>>(c_r, cp, B, N, C, pl.nbt);
            else k_softmax<<<cdiv(n, 128), 128, 0, st>>>(c_r, cp, B, N, C, pl.nbt);
        }
        LAUNCH_CHECK();
        pp.X = nullptr; pp.coef = c_r; pp.out = part;
        if ((rc = run_pass(pl, kModeA, pp, w, st))) return rc;
        if ((rc = launch_squash(pl, part, pl.IS, 1.f, s_r, v_r, vsum, 1, last ? v : nullptr, st))) return rc;
    }
    if (c_out != nullptr && pl.Reff == 1) {      // R == 1 or C == 1: the couplings are the constant 1/C
        const long n = (long)B * N * C;
        { LaunchScope ls_(kcOther, st); k_fill<<<cdiv(n, 256), 256, 0, st>>>(c_out, 1.f / (float)C, n); }
        LAUNCH_CHECK();
    }
    return 0;
}

int caps_route_backward(const float* u, const float* W, const float* grad_v, const int64_t* y, float margin_scale,
                        const float* loss_grad_dev, float* du, float* dW, void* ws, size_t ws_bytes,
                        int B, int N, int C, int K, int D, int R, void* stream) {
    return caps_route_backward_ev(u, W, grad_v, y, margin_scale, loss_grad_dev, du, dW, ws, ws_bytes, B, N, C, K, D, R, stream, nullptr);
}

int caps_route_backward_ev(const float* u, const float* W, const float* grad_v, const int64_t* y, float margin_scale,
                           const float* loss_grad_dev, float* du, float* dW, void* ws, size_t ws_bytes,
                           int B, int N, int C, int K, int D, int R, void* stream, void* dw_ready_event) {
    Plan pl;
    if (!make_plan(pl, B, N, C, K, D, R, 1))
        return fail(CAPS_E_UNSUPPORTED, "caps_route_backward: dims B=%d N=%d C=%d K=%d D=%d R=%d not supported", B, N, C, K, D, R);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!W || !dW) return fail(CAPS_E_BADARG, "caps_route_backward: null pointer");
    if (B == 0) {
        CUDA_TRY(cudaMemsetAsync(dW, 0, (size_t)N * C * K * D * sizeof(float), st));
        if (dw_ready_event) CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(dw_ready_event), st));
        return 0;
    }
    if (!ws) return fail(CAPS_E_BADARG, "caps_route_backward: null workspace");
    if (misaligned(W) || misaligned(ws) || misaligned(dW) || (du && misaligned(du)))
        return fail(CAPS_E_BADARG, "caps_route_backward: W, dW, du and ws must be 16-byte aligned");
    if (ws_bytes < pl.total * sizeof(float))
        return fail(CAPS_E_WORKSPACE, "caps_route_backward: workspace %zu < %zu bytes", ws_bytes, pl.total * sizeof(float));
    {
        // the forward that filled this workspace: its dims must be ours, and its engine choices (not the knobs as
        // they stand now) say which operand copies the workspace holds
        FwdRecord rec;
        if (!lookup_forward(ws, rec))
            return fail(CAPS_E_STATE, "caps_route_backward: no caps_route_forward has filled this workspace");
        if (rec.B != B || rec.N != N || rec.C != C || rec.K != K || rec.D != D || rec.R != R)
            return fail(CAPS_E_STATE, "caps_route_backward: workspace was filled for B=%d N=%d C=%d K=%d D=%d R=%d, not B=%d N=%d C=%d K=%d D=%d R=%d",
                        rec.B, rec.N, rec.C, rec.K, rec.D, rec.R, B, N, C, K, D, R);
        if (!rec.with_grad) return fail(CAPS_E_STATE, "caps_route_backward: the forward ran with with_grad = 0 (no saved state)");
        pl.use_tc = rec.use_tc;
        pl.fused = rec.fused;
        pl.c1 = rec.c1;
    }
    if (pl.c1) {
        if (!u) return fail(CAPS_E_BADARG, "caps_route_backward: u is required for a single class capsule");
        float* w1 = static_cast<float*>(ws);
        int nl = 0, rc1;
        { LaunchScope ls_(kcC1, st); rc1 = launch_c1_backward(u, W, w1 + pl.o_s, grad_v, y, margin_scale, loss_grad_dev, du, dW,
                                                              w1 + pl.o_dupart, B, N, K, D, st, &nl); }
        if (nl > 1) g_launches.fetch_add(nl - 1, std::memory_order_relaxed);      // the scope counted one launch
        if (!rc1 && dw_ready_event) CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(dw_ready_event), st));
        return rc1;
    }
    float* w = static_cast<float*>(ws);
    const float* ut = w + pl.o_ut;
    const float* Wp = pl.pad_w ? w + pl.o_wp : W;
    float* part = w + pl.o_part;
    float* tmp = w + pl.o_tmp;
    const int Re = pl.Reff;
    int rc;
    // The mma gradient kernel wants ds^0 pre-multiplied by iteration 0's uniform coupling 1/C (ds^0 has no other
    // reader); the FMA kernel multiplies by cconst[0] itself.  Same fp32 product either way.
    const bool grad_mma = g_tune_gradmma && pl.DP >= 16 && pl.JW == 8;       // D >= 9, C >= 7
    const float ds0_scale = grad_mma ? 1.f / (float)C : 1.f;
    const int ISb = pl.fused ? fused_pick_splits(pl, true, g_tune_isplit) : pl.IS;
    // top: dv = grad_v + margin gradient ; ds^{R-1}
    if ((rc = launch_dsquash(pl, nullptr, 0, grad_v, y, margin_scale, loss_grad_dev, w + pl.o_v + pl.xs * (Re - 1),
                             w + pl.o_s + pl.xs * (Re - 1), w + pl.o_ds + pl.xs * (Re - 1), Re == 1 ? ds0_scale : 1.f, st)))
        return rc;
    for (int r = Re - 1; r >= 1; --r) {
        const float* c_r = w + pl.o_c + pl.cs * (r - 1);
        float* beta_r = w + pl.o_beta + pl.cs * (r - 1);
        const float* beta_next = (r == Re - 1) ? nullptr : w + pl.o_beta + pl.cs * r;
        if (pl.fused) {
            // dc = u_hat . ds^r ; beta^r = beta^{r+1} + c^r (dc - sum_j c^r dc) ; dv^{r-1} = sum_i beta^r u_hat : one sweep
            {
                LaunchScope ls_(kcFused, st);
                rc = launch_sweep_fused(pl, true, w + pl.o_ua, w + pl.o_wb, w + pl.o_ds + pl.xs * r, c_r, beta_next, beta_r, part, ISb, st);
            }
            if (rc) return rc;
        } else {
            PassParams pp{};
            pp.ut = ut; pp.W = Wp; pp.N = N; pp.C = C; pp.nbt = pl.nbt; pp.i_per_split = pl.i_per_split;
            pp.X = w + pl.o_ds + pl.xs * r; pp.out = tmp;                       // dc = u_hat . ds^r
            if ((rc = run_pass(pl, kModeL, pp, w, st))) return rc;
            const long n = (long)pl.nbt * N * 32;
            {
                LaunchScope ls_(kcSoftmax, st);
                // the register-resident variant (114 registers) loses to the three-pass one here: the second and
                // third passes hit L1, and occupancy matters more than the re-reads (measured 1.9 vs 1.0 ms)
                if (C <= 16) k_softmax_bwd_reg<16><<<cdiv(n, 128), 128, 0, st>>>(c_r, tmp, beta_next, beta_r, N, C, pl.nbt);
                else if (C <= 48 && g_tune_sbstaged) {
                    const size_t smem = (size_t)4 * 2 * C * 32 * sizeof(float);
                    CAPS_SET_SMEM(k_softmax_bwd_staged<48>, smem);
                    const long nblocks = (long)pl.nbt * N;
                    k_softmax_bwd_staged<48><<<cdiv(nblocks, 4), 128, smem, st>>>(c_r, tmp, beta_next, beta_r, C, nblocks);
                } else k_softmax_bwd<<<cdiv(n, 128), 128, 0, st>>>(c_r, tmp, beta_next, beta_r, N, C, pl.nbt);
            }
            LAUNCH_CHECK();
            pp.X = nullptr; pp.coef = beta_r; pp.out = part;                    // dv^{r-1} = sum_i beta u_hat
            if ((rc = run_pass(pl, kModeA, pp, w, st))) return rc;
        }
        if ((rc = launch_dsquash(pl, part, ISb, nullptr, nullptr, 0.f, nullptr, nullptr, w + pl.o_s + pl.xs * (r - 1),
                                 w + pl.o_ds + pl.xs * (r - 1), r == 1 ? ds0_scale : 1.f, st)))
            return rc;
    }
    GradParams gp{};
    gp.ut = ut; gp.W = Wp; gp.dW = dW; gp.du_part = w + pl.o_dupart;
    gp.N = N; gp.C = C; gp.D = D; gp.DP = pl.DP; gp.nbt = pl.nbt;
    gp.CS = pl.fused ? pl.CSmax : C;
    int m = 0;
    for (int r = 0; r < Re; ++r) {                                          // c^r (x) ds^r
        gp.coef[m] = (r == 0) ? nullptr : w + pl.o_c + pl.cs * (r - 1);
        gp.cconst[m] = (r == 0 && grad_mma) ? 1.f : 1.f / (float)C;
        gp.X[m] = w + pl.o_ds + pl.xs * r;
        ++m;
    }
    for (int r = 1; r < Re; ++r) {                                          // beta^r (x) v^{r-1}
        gp.coef[m] = w + pl.o_beta + pl.cs * (r - 1);
        gp.cconst[m] = 0.f;
        gp.X[m] = w + pl.o_v + pl.xs * (r - 1);
        ++m;
    }
    {
        LaunchScope ls_(kcGrad, st);
        rc = grad_mma ? launch_grad_mma(pl, gp, st) : launch_grad(pl, gp, st);
    }
    if (rc) return rc;
    // dW is complete here (the du reduction below does not touch it): a data-parallel caller starts its all-reduce now
    if (dw_ready_event) CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(dw_ready_event), st));
    if (du != nullptr) {
        const long n = (long)pl.nbt * N * 32;
        const int parts = grad_mma ? grad_mma_parts(pl) : pl.JG;
        { LaunchScope ls_(kcReduceDu, st); k_reduce_du<8><<<cdiv(n, 128), 128, 0, st>>>(gp.du_part, parts, du, B, N, pl.nbt); }
        LAUNCH_CHECK();
    }
    return 0;
}

int caps_margin_loss(const float* v, const int64_t* y, float scale, float* loss, float* scores_out,
                     float* scratch, int B, int C, int D, void* stream) {
    if (!v || !y || !loss || B < 0 || C <= 0 || D <= 0) return fail(CAPS_E_BADARG, "caps_margin_loss: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long n = (long)B * C;
    int blocks = scratch ? (int)((n + 1023) / 1024) : 1;        // ~4 (b,j) rows per thread
    if (blocks > CAPS_MARGIN_SCRATCH_FLOATS) blocks = CAPS_MARGIN_SCRATCH_FLOATS;
    if (blocks < 1) blocks = 1;
    // small problems (the DarkCapsuleNet head: 1568 rows): ONE block of 1024 threads, no second launch -- the kernel is
    // a latency chain of a few dependent loads per thread, so fewer rows per thread and one launch less is all that counts
    const int threads = n <= 4096 ? 1024 : 256;
    if (n <= 4096) blocks = 1;
    { LaunchScope ls_(kcLoss, st); k_margin_loss<<<blocks, threads, 0, st>>>(v, y, scale, loss, scratch, scores_out, B, C, D); }
    LAUNCH_CHECK();
    if (blocks > 1) {
        { LaunchScope ls_(kcLoss, st); k_margin_loss_final<<<1, 256, 0, st>>>(scratch, blocks, scale, loss); }
        LAUNCH_CHECK();
    }
    return 0;
}

int caps_squash(const float* x, float* y, long rows, int D, void* stream) {
    if (!x || !y || rows < 0 || D <= 0) return fail(CAPS_E_BADARG, "caps_squash: bad argument");
    if (rows == 0) return 0;
    { LaunchScope ls_(kcOther, static_cast<cudaStream_t>(stream)); k_squash_rows<<<cdiv(rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rows, D); }
    LAUNCH_CHECK();
    return 0;
}

int caps_squash_backward(const float* x, const float* dy, float* dx, long rows, int D, void* stream) {
    if (!x || !dy || !dx || rows < 0 || D <= 0) return fail(CAPS_E_BADARG, "caps_squash_backward: bad argument");
    if (rows == 0) return 0;
    { LaunchScope ls_(kcOther, static_cast<cudaStream_t>(stream)); k_squash_rows_bwd<<<cdiv(rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dx, rows, D); }
    LAUNCH_CHECK();
    return 0;
}

int caps_primary_squash(const float* conv, float* u, int B, int n_caps, int Cc, int HW, void* stream) {
    if (!conv || !u || B < 0 || n_caps <= 0 || n_caps > 16 || Cc <= 0 || HW <= 0)
        return fail(CAPS_E_BADARG, "caps_primary_squash: bad argument (n_caps must be in 1..16)");
    if (misaligned(u)) return fail(CAPS_E_BADARG, "caps_primary_squash: u must be 16-byte aligned");
    const long total = (long)B * Cc * HW;
    if (total == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        LaunchScope ls_(kcLayout, st);
        if (n_caps <= 8) k_primary_squash<8><<<cdiv(total, 256), 256, 0, st>>>(conv, u, total, n_caps, Cc, HW);
        else k_primary_squash<16><<<cdiv(total, 256), 256, 0, st>>>(conv, u, total, n_caps, Cc, HW);
    }
    LAUNCH_CHECK();
    return 0;
}

int caps_primary_squash_backward(const float* conv, const float* du, float* dconv, int B, int n_caps, int Cc, int HW, void* stream) {
    if (!conv || !du || !dconv || B < 0 || n_caps <= 0 || n_caps > 16 || Cc <= 0 || HW <= 0)
        return fail(CAPS_E_BADARG, "caps_primary_squash_backward: bad argument (n_caps must be in 1..16)");
    const long total = (long)B * Cc * HW;
    if (total == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        LaunchScope ls_(kcLayout, st);
        if (n_caps <= 8) k_primary_squash_bwd<8><<<cdiv(total, 256), 256, 0, st>>>(conv, du, dconv, total, n_caps, Cc, HW);
        else k_primary_squash_bwd<16><<<cdiv(total, 256), 256, 0, st>>>(conv, du, dconv, total, n_caps, Cc, HW);
    }
    LAUNCH_CHECK();
    return 0;
}

int caps_dark_regroup(const float* x, float* u, int B, int Cch, int G, void* stream) {
    if (!x || !u || B < 0 || Cch <= 0 || (Cch & 7) || G <= 0) return fail(CAPS_E_BADARG, "caps_dark_regroup: bad argument (Cch must be a multiple of 8)");
    if (misaligned(u)) return fail(CAPS_E_BADARG, "caps_dark_regroup: u must be 16-byte aligned");
    const long total = (long)B * (Cch / 8) * 16 * G;
    if (total == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { LaunchScope ls_(kcLayout, st); k_dark_regroup<false><<<cdiv(total, 256), 256, 0, st>>>(x, u, total, B, Cch, G); }
    LAUNCH_CHECK();
    return 0;
}

int caps_dark_regroup_backward(const float* du, float* dx, int B, int Cch, int G, void* stream) {
    if (!du || !dx || B < 0 || Cch <= 0 || (Cch & 7) || G <= 0) return fail(CAPS_E_BADARG, "caps_dark_regroup_backward: bad argument (Cch must be a multiple of 8)");
    if (misaligned(du)) return fail(CAPS_E_BADARG, "caps_dark_regroup_backward: du must be 16-byte aligned");
    const long total = (long)B * (Cch / 8) * 16 * G;
    if (total == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { LaunchScope ls_(kcLayout, st); k_dark_regroup<true><<<cdiv(total, 256), 256, 0, st>>>(du, dx, total, B, Cch, G); }
    LAUNCH_CHECK();
    return 0;
}

int caps_dark_loss(const float* v, const float* y, float scale, float* loss, float* grad_v, float* scratch,
                   int B, int G, int Y, void* stream) {
    if (!v || !y || !loss || B < 0 || G <= 0 || Y < 5) return fail(CAPS_E_BADARG, "caps_dark_loss: bad argument (Y >= 5)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long n = (long)B * G;
    const int blocks = (scratch && n > 4096) ? (int)std::min<long>(std::max<long>(cdiv(n, 256), 1), CAPS_MARGIN_SCRATCH_FLOATS) : 1;
    const int threads = n <= 4096 ? 1024 : 256;            // one block, one launch for the reference's batch (32 x 49 cells)
    { LaunchScope ls_(kcLoss, st); k_dark_loss<<<blocks, threads, 0, st>>>(v, y, scale, loss, scratch, grad_v, B, G, Y); }
    LAUNCH_CHECK();
    if (blocks > 1) {
        { LaunchScope ls_(kcLoss, st); k_margin_loss_final<<<1, 256, 0, st>>>(scratch, blocks, scale, loss); }
        LAUNCH_CHECK();
    }
    return 0;
}

// ---- host-buffer step ------------------------------------------------------------------------
// The batch is cut into up to three micro-batches (B/8, 3B/8, B/2, multiples of 128) so that the
// host->device copy of micro-batch m+1 (on an internal copy stream) runs under the kernels of
// micro-batch m; the first one is small so the pipeline fills quickly.  dW is summed over the
// micro-batches in a fixed order (bit-reproducible); loss terms are summed on the host.
namespace {
constexpr int kHostMaxMb = 3;
struct HostStepLayout { size_t o_u, o_y, o_v, o_du, o_loss, o_lscr, o_dwt, o_ws, total; int nmb; int mb[kHostMaxMb]; };
bool host_step_layout(HostStepLayout& L, int B, int N, int C, int K, int D, int R) {
    L.nmb = 1; L.mb[0] = B;
    if (g_tune_hostmb != 1 && B >= 2048) {
        const int c0 = std::max(128, (B / 8) / 128 * 128), c1 = std::max(128, (3 * (B / 8)) / 128 * 128);
        L.nmb = 3; L.mb[0] = c0; L.mb[1] = c1; L.mb[2] = B - c0 - c1;
    }
    size_t wsb = 0;
    for (int m = 0; m < L.nmb; ++m) {
        const size_t w = caps_route_workspace_bytes(L.mb[m], N, C, K, D, R, 1);
        if (w == 0) return false;
        wsb = std::max(wsb, w);
    }
    auto r256 = [](size_t n) { return (n + 255) & ~(size_t)255; };
    size_t o = 0;
    L.o_u = o; o += r256((size_t)B * N * K * 4);
    L.o_y = o; o += r256((size_t)B * 8);
    L.o_v = o; o += r256((size_t)B * C * D * 4);
    L.o_du = o; o += r256((size_t)B * N * K * 4);
    L.o_loss = o; o += 256;
    L.o_lscr = o; o += r256(CAPS_MARGIN_SCRATCH_FLOATS * 4);
    L.o_dwt = o; o += L.nmb > 1 ? r256((size_t)N * C * K * D * 4) : 0;
    L.o_ws = o; o += r256(wsb);
    L.total = o;
    return true;
}
// one internal copy stream per device, created on first use (mutex); the events of a call are its own
cudaStream_t g_copy_stream[kMaxDevices];
std::mutex g_copy_mu;
int copy_stream_for_current_device(cudaStream_t* out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(CAPS_E_UNSUPPORTED, "device ordinal %d beyond %d", dev, kMaxDevices);
    std::lock_guard<std::mutex> lk(g_copy_mu);
    if (!g_copy_stream[dev]) CUDA_TRY(cudaStreamCreateWithFlags(&g_copy_stream[dev], cudaStreamNonBlocking));
    *out = g_copy_stream[dev];
    return 0;
}
struct CallEvents {          // per call: two host threads (or two devices) never share an event
    cudaEvent_t ev[kHostMaxMb + 1] = {};
    int n = 0;
    int create(int count) {
        for (n = 0; n < count; ++n) CUDA_TRY(cudaEventCreateWithFlags(&ev[n], cudaEventDisableTiming));
        return 0;
    }
    ~CallEvents() { for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]); }
};
}  // namespace

size_t caps_route_step_host_scratch_bytes(int B, int N, int C, int K, int D, int R) {
    HostStepLayout L;
    return host_step_layout(L, B, N, C, K, D, R) ? L.total : 0;
}

int caps_route_step_host(const float* u_host, const int64_t* y_host, const float* W_dev, float* loss_host,
                         float* v_host, float* du_host, float* dW_dev, void* dev_scratch, size_t scratch_bytes,
                         int B, int N, int C, int K, int D, int R, void* stream) {
    HostStepLayout L;
    if (!host_step_layout(L, B, N, C, K, D, R)) return fail(CAPS_E_UNSUPPORTED, "caps_route_step_host: dims not supported");
    if (!u_host || !y_host || !W_dev || !loss_host || !dW_dev || !dev_scratch || B <= 0)
        return fail(CAPS_E_BADARG, "caps_route_step_host: bad argument");
    if (scratch_bytes < L.total) return fail(CAPS_E_WORKSPACE, "caps_route_step_host: scratch %zu < %zu", scratch_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(dev_scratch);
    float* u_d = reinterpret_cast<float*>(base + L.o_u);
    int64_t* y_d = reinterpret_cast<int64_t*>(base + L.o_y);
    float* v_d = reinterpret_cast<float*>(base + L.o_v);
    float* du_d = reinterpret_cast<float*>(base + L.o_du);
    float* loss_d = reinterpret_cast<float*>(base + L.o_loss);
    float* dwt = reinterpret_cast<float*>(base + L.o_dwt);
    void* ws = base + L.o_ws;
    const size_t wsb = L.total - L.o_ws;
    const size_t row_u = (size_t)N * K, row_v = (size_t)C * D;
    const float scale = 1.f / (float)B;
    cudaStream_t cs = st;
    CallEvents evs;
    int rc;
    if (L.nmb > 1) {
        if ((rc = copy_stream_for_current_device(&cs))) return rc;
        if ((rc = evs.create(kHostMaxMb + 1))) return rc;
        // the copy stream may only overwrite u/y once everything already queued on `st` is done
        CUDA_TRY(cudaEventRecord(evs.ev[kHostMaxMb], st));
        CUDA_TRY(cudaStreamWaitEvent(cs, evs.ev[kHostMaxMb], 0));
    }
    CUDA_TRY(cudaMemcpyAsync(y_d, y_host, (size_t)B * 8, cudaMemcpyHostToDevice, cs));
    for (int m = 0, b0 = 0; m < L.nmb; b0 += L.mb[m], ++m) {
        CUDA_TRY(cudaMemcpyAsync(u_d + b0 * row_u, u_host + b0 * row_u, L.mb[m] * row_u * 4, cudaMemcpyHostToDevice, cs));
        if (L.nmb > 1) CUDA_TRY(cudaEventRecord(evs.ev[m], cs));
    }
    for (int m = 0, b0 = 0; m < L.nmb; b0 += L.mb[m], ++m) {
        const int Bm = L.mb[m];
        if (L.nmb > 1) CUDA_TRY(cudaStreamWaitEvent(st, evs.ev[m], 0));
        float* dW_m = m == 0 ? dW_dev : dwt;
        if ((rc = caps_route_forward(u_d + b0 * row_u, W_dev, v_d + b0 * row_v, nullptr, ws, wsb, Bm, N, C, K, D, R, 1, stream))) return rc;
        if ((rc = caps_margin_loss(v_d + b0 * row_v, y_d + b0, scale, loss_d + m, nullptr, reinterpret_cast<float*>(base + L.o_lscr), Bm, C, D, stream))) return rc;
        if ((rc = caps_route_backward(u_d + b0 * row_u, W_dev, nullptr, y_d + b0, scale, nullptr, du_host ? du_d + b0 * row_u : nullptr, dW_m,
                                      ws, wsb, Bm, N, C, K, D, R, stream)))
            return rc;
        if (m > 0) {
            const long n = (long)N * C * K * D;
            LaunchScope ls_(kcOther, st);
            k_add_inplace<<<cdiv(n, 256), 256, 0, st>>>(dW_dev, dwt, n);
            LAUNCH_CHECK();
        }
    }
    float loss_parts[kHostMaxMb] = {0.f, 0.f, 0.f};
    CUDA_TRY(cudaMemcpyAsync(L.nmb > 1 ? loss_parts : loss_host, loss_d, 4 * L.nmb, cudaMemcpyDeviceToHost, st));
    if (v_host) CUDA_TRY(cudaMemcpyAsync(v_host, v_d, (size_t)B * C * D * 4, cudaMemcpyDeviceToHost, st));
    if (du_host) CUDA_TRY(cudaMemcpyAsync(du_host, du_d, (size_t)B * N * K * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (L.nmb > 1) *loss_host = (loss_parts[0] + loss_parts[1]) + loss_parts[2];
    return 0;
}

// ---- host pipeline: the H2D copy of step n+1 runs under the kernels of step n ---------------------------------
namespace {
struct HostPipe {
    int B, N, C, K, D, R, dev;
    char* base;
    size_t o_u[2], o_y[2], o_v, o_du, o_loss, o_lscr, o_ws, ws_bytes, total;
    cudaStream_t copy_stream;
    cudaEvent_t copied[2];        // slot s: its H2D copies have landed
    cudaEvent_t consumed[2];      // slot s: the step that read it has finished with u / y
    long submitted, stepped;
};
bool host_pipe_layout(HostPipe& P) {
    auto r256 = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const size_t wsb = caps_route_workspace_bytes(P.B, P.N, P.C, P.K, P.D, P.R, 1);
    if (wsb == 0 || P.B <= 0) return false;
    size_t o = 0;
    for (int s = 0; s < 2; ++s) { P.o_u[s] = o; o += r256((size_t)P.B * P.N * P.K * 4); P.o_y[s] = o; o += r256((size_t)P.B * 8); }
    P.o_v = o; o += r256((size_t)P.B * P.C * P.D * 4);
    P.o_du = o; o += r256((size_t)P.B * P.N * P.K * 4);          // the step computes du like any backward (it stays on the device)
    P.o_loss = o; o += 256;
    P.o_lscr = o; o += r256(CAPS_MARGIN_SCRATCH_FLOATS * 4);
    P.o_ws = o; o += r256(wsb);
    P.ws_bytes = wsb;
    P.total = o;
    return true;
}
}  // namespace

size_t caps_host_pipe_scratch_bytes(int B, int N, int C, int K, int D, int R) {
    HostPipe P{};
    P.B = B; P.N = N; P.C = C; P.K = K; P.D = D; P.R = R;
    return host_pipe_layout(P) ? P.total : 0;
}

int caps_host_pipe_create(void** pipe_out, void* dev_scratch, size_t scratch_bytes, int B, int N, int C, int K, int D, int R) {
    if (!pipe_out || !dev_scratch) return fail(CAPS_E_BADARG, "caps_host_pipe_create: null pointer");
    HostPipe* P = new HostPipe();
    P->B = B; P->N = N; P->C = C; P->K = K; P->D = D; P->R = R;
    if (!host_pipe_layout(*P)) { delete P; return fail(CAPS_E_UNSUPPORTED, "caps_host_pipe_create: dims not supported"); }
    if (scratch_bytes < P->total) { const size_t need = P->total; delete P; return fail(CAPS_E_WORKSPACE, "caps_host_pipe_create: scratch %zu < %zu", scratch_bytes, need); }
    P->base = static_cast<char*>(dev_scratch);
    cudaError_t e = cudaGetDevice(&P->dev);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P->copy_stream, cudaStreamNonBlocking);
    for (int s = 0; s < 2 && e == cudaSuccess; ++s) {
        e = cudaEventCreateWithFlags(&P->copied[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->consumed[s], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { delete P; return fail((int)e, "caps_host_pipe_create: %s", cudaGetErrorString(e)); }
    P->submitted = P->stepped = 0;
    *pipe_out = P;
    return 0;
}

int caps_host_pipe_destroy(void* pipe) {
    HostPipe* P = static_cast<HostPipe*>(pipe);
    if (!P) return 0;
    cudaStreamSynchronize(P->copy_stream);
    for (int s = 0; s < 2; ++s) { cudaEventDestroy(P->copied[s]); cudaEventDestroy(P->consumed[s]); }
    cudaStreamDestroy(P->copy_stream);
    delete P;
    return 0;
}

int caps_host_pipe_submit(void* pipe, const float* u_host, const int64_t* y_host) {
    HostPipe* P = static_cast<HostPipe*>(pipe);
    if (!P || !u_host || !y_host) return fail(CAPS_E_BADARG, "caps_host_pipe_submit: null pointer");
    if (P->submitted - P->stepped >= 2) return fail(CAPS_E_STATE, "caps_host_pipe_submit: both slots hold batches that have not been stepped");
    const int s = (int)(P->submitted & 1);
    // the slot's previous tenant (two submissions ago) must have been consumed by its step
    if (P->submitted >= 2) CUDA_TRY(cudaStreamWaitEvent(P->copy_stream, P->consumed[s], 0));
    CUDA_TRY(cudaMemcpyAsync(P->base + P->o_y[s], y_host, (size_t)P->B * 8, cudaMemcpyHostToDevice, P->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(P->base + P->o_u[s], u_host, (size_t)P->B * P->N * P->K * 4, cudaMemcpyHostToDevice, P->copy_stream));
    CUDA_TRY(cudaEventRecord(P->copied[s], P->copy_stream));
    ++P->submitted;
    return 0;
}

int caps_host_pipe_step(void* pipe, const float* W_dev, float* dW_dev, float* loss_host, float* v_host, void* stream,
                        void* dw_ready_event) {
    HostPipe* P = static_cast<HostPipe*>(pipe);
    if (!P || !W_dev || !dW_dev || !loss_host) return fail(CAPS_E_BADARG, "caps_host_pipe_step: null pointer");
    if (P->stepped >= P->submitted) return fail(CAPS_E_STATE, "caps_host_pipe_step: no submitted batch to step on");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int s = (int)(P->stepped & 1);
    const float* u_d = reinterpret_cast<const float*>(P->base + P->o_u[s]);
    const int64_t* y_d = reinterpret_cast<const int64_t*>(P->base + P->o_y[s]);
    float* v_d = reinterpret_cast<float*>(P->base + P->o_v);
    float* loss_d = reinterpret_cast<float*>(P->base + P->o_loss);
    void* ws = P->base + P->o_ws;
    const int B = P->B, N = P->N, C = P->C, K = P->K, D = P->D, R = P->R;
    int rc;
    CUDA_TRY(cudaStreamWaitEvent(st, P->copied[s], 0));
    if ((rc = caps_route_forward(u_d, W_dev, v_d, nullptr, ws, P->ws_bytes, B, N, C, K, D, R, 1, stream))) return rc;
    if ((rc = caps_margin_loss(v_d, y_d, 1.f / (float)B, loss_d, nullptr, reinterpret_cast<float*>(P->base + P->o_lscr), B, C, D, stream))) return rc;
    if ((rc = caps_route_backward_ev(u_d, W_dev, nullptr, y_d, 1.f / (float)B, nullptr, reinterpret_cast<float*>(P->base + P->o_du), dW_dev, ws, P->ws_bytes,
                                     B, N, C, K, D, R, stream, dw_ready_event)))
        return rc;
    CUDA_TRY(cudaEventRecord(P->consumed[s], st));
    CUDA_TRY(cudaMemcpyAsync(loss_host, loss_d, 4, cudaMemcpyDeviceToHost, st));
    if (v_host) CUDA_TRY(cudaMemcpyAsync(v_host, v_d, (size_t)B * C * D * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ++P->stepped;
    return 0;
}

}  // extern "C"
