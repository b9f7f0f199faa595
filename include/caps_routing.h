/* caps_routing.h -- C ABI of the B200-native capsule dynamic-routing path.
 *
 * This is the drop-in boundary for ONE path of Cranial-XIX/cs231-capsule-yolo-traffic-sign-detection:
 * the caps->caps branch of `CapsuleLayer` (reference models.py:46-83) plus the margin-loss
 * gradient that feeds it (reference loss_fns.py:11-23, models.py:117); since ABI version 2 also the
 * steps directly before and after that branch (caps_primary_squash, caps_dark_regroup,
 * caps_dark_loss: reference models.py:81-82, :393-399, loss_fns.py:187-204).  The reference has no
 * FFI of its own -- its boundary is the Python class -- so these entry points are what a ctypes
 * binding behind that class calls (the in-repo binding is
 * cs231_capsule_yolo_traffic_sign_detection_b200/_cabi.py).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.  All `float*` are fp32, row-major,
 *     contiguous, 16-byte aligned.  Unless a function says HOST, pointers are DEVICE pointers on
 *     the current CUDA device and work is enqueued on `stream` (a cudaStream_t passed as void*);
 *     nothing synchronises the host.
 *   - dims: B batch, N = n_nodes (input capsules), C = n_caps (output capsules), K = in_C,
 *     D = out_C, R = n_iter                                   (reference models.py:47-58)
 *   - tensors:  u [B,N,K]   W [N,C,K,D] (= route_weights[0])   v [B,C,D] (= output[:,0,:,0,:])
 *               c [B,N,C] (last-iteration coupling coefficients)   y [B] int64 labels
 *   - return value: 0 on success; CAPS_E_* (<0) for argument errors; a positive cudaError_t if a
 *     CUDA call failed.  caps_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every buffer (inputs, outputs, workspace) for the duration of the enqueued work; launchers are
 *     re-entrant across host threads and devices.  Process-wide state is limited to: the caps_set_tuning() knobs;
 *     per-device caches of kernel attributes and one internal copy stream per device (created on first use, mutex-
 *     guarded); a fixed-size, mutex-guarded table that remembers, per workspace pointer, the dims and engines of the
 *     last caps_route_forward so that caps_route_backward can validate its workspace (CAPS_E_STATE) and replay them;
 *     and the profiling counters of the measurement helpers (single-threaded use).
 *   - results are bit-reproducible run to run for identical (dims, B, tuning, device model): every cross-CTA sum has
 *     a fixed order, but how the input-capsule range is split across CTAs depends on the batch size and on how many
 *     thread-block clusters the device can co-schedule, so the same sample can differ in its last bits between two
 *     batch sizes (caps_set_tuning("isplit", n) pins the split).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CAPS_ROUTING_H_
#define CAPS_ROUTING_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPS_ABI_VERSION 3

#define CAPS_E_BADARG     (-1)   /* null pointer / non-positive dim / misaligned pointer        */
#define CAPS_E_UNSUPPORTED (-2)  /* K != 8, D > 48, R > 5, C > 1024: shapes with no kernel       */
#define CAPS_E_WORKSPACE  (-3)   /* workspace smaller than caps_route_workspace_bytes() says     */
#define CAPS_E_STATE      (-4)   /* backward on a workspace no (matching, with_grad) forward filled; pipe misuse */

/* ABI version of the loaded library (== CAPS_ABI_VERSION of the header it was built from). */
int caps_abi_version(void);

/* Thread-local description of the last non-zero return on this thread ("" if none). */
const char* caps_last_error(void);

/* Bytes of device workspace caps_route_forward/backward need for these dims.  with_grad != 0
 * sizes it to also hold the forward state the backward pass reads (per-iteration capsule sums and
 * coupling coefficients) and the backward scratch.  Returns 0 for unsupported dims. */
size_t caps_route_workspace_bytes(int B, int N, int C, int K, int D, int R, int with_grad);

/* Forward of the routing branch: replaces reference models.py:71-79 (+ :83).
 *   u_hat = u.W ; R x { c = softmax_j(b) ; s = sum_i c u_hat ; v = squash(s) ; b += u_hat.v }
 *   v      out [B,C,D]   class capsules of the last iteration
 *   c_out  out [B,N,C]   last-iteration couplings, or NULL (only tests want it)
 *   ws     workspace of >= caps_route_workspace_bytes(..., with_grad) bytes, 256-byte aligned.
 *          With with_grad != 0 it holds the state caps_route_backward needs; keep it (and u, W)
 *          unchanged until that call. */
int caps_route_forward(const float* u, const float* W, float* v, float* c_out,
                       void* ws, size_t ws_bytes,
                       int B, int N, int C, int K, int D, int R, int with_grad, void* stream);

/* Backward through ALL R iterations (the reference never detaches; autograd of models.py:71-79),
 * with the margin-loss gradient optionally fused:
 *   g = grad_v (if non-NULL)  +  d/dv [ margin_loss(|v|, y) * margin_scale * lg ] (if y non-NULL)
 *       margin term: reference loss_fns.py:12-17,23 on scores = |v| (models.py:117);
 *       margin_scale is the reference's 1/y.size(0); lg = *loss_grad_dev if that DEVICE scalar
 *       is non-NULL (the upstream d/d loss autograd hands over; avoids a host sync), else 1.
 *   du  out [B,N,K]     d loss / d u      (or NULL to skip writing it)
 *   dW  out [N,C,K,D]   d loss / d W, summed over the batch; OVERWRITTEN, not accumulated
 * `ws` must be the workspace a with_grad forward with the same dims, u and W filled.
 * Summation orders are fixed: results are bit-reproducible run to run. */
int caps_route_backward(const float* u, const float* W, const float* grad_v, const int64_t* y,
                        float margin_scale, const float* loss_grad_dev, float* du, float* dW,
                        void* ws, size_t ws_bytes,
                        int B, int N, int C, int K, int D, int R, void* stream);

/* caps_route_backward plus a hook for data-parallel callers: `dw_ready_event` (a cudaEvent_t passed as void*, or
 * NULL) is recorded on `stream` right behind the kernel that completes dW -- before the du reduction -- so that an
 * all-reduce of dW on another stream can start while this stream finishes du (the reference has no
 * counterpart: it is single-device, main.py:231). */
int caps_route_backward_ev(const float* u, const float* W, const float* grad_v, const int64_t* y,
                           float margin_scale, const float* loss_grad_dev, float* du, float* dW,
                           void* ws, size_t ws_bytes,
                           int B, int N, int C, int K, int D, int R, void* stream, void* dw_ready_event);

/* Margin loss value: sum_{b,j} [ T relu(0.9-m)^2 + 0.5 (1-T) relu(m-0.1)^2 ] * scale with
 * m = |v[b,j,:]|, T = (y[b]==j).  Replaces reference models.py:117 + loss_fns.py:12-17,23
 * (recon term off).  loss: one float (device).  scores_out: [B,C] or NULL.  scratch: device
 * buffer of CAPS_MARGIN_SCRATCH_FLOATS floats for the two-level fixed-order reduction, or NULL
 * (then one thread block does all the work: slow for large B, same result contract). */
#define CAPS_MARGIN_SCRATCH_FLOATS 2048
int caps_margin_loss(const float* v, const int64_t* y, float scale, float* loss,
                     float* scores_out, float* scratch, int B, int C, int D, void* stream);

/* squash over the last dim of a [rows, D] array (reference models.py:64-67), any D >= 1.
 * Used by the primary-capsule branch (models.py:82).  y may alias x. */
int caps_squash(const float* x, float* y, long rows, int D, void* stream);
/* its backward: dx = d squash(x)/dx applied to dy. */
int caps_squash_backward(const float* x, const float* dy, float* dx, long rows, int D, void* stream);

/* Primary-capsule tail, the step directly before the routing layer (SURVEY.md section 8(f) row 1;
 * reference models.py:81-82: `[cap(x).view(B,-1,1) for cap in self.capsules]`, `torch.cat(dim=-1)`,
 * `squash`).  `conv` [B][n_caps*Cc][HW] is the output of ONE convolution whose weight is the n_caps
 * capsule convolutions' weights concatenated along the output-channel axis (channel = k*Cc + c);
 * `u` [B][Cc*HW][n_caps] is the routing layer's input: u[b][c*HW + hw][k] = squash over k of
 * conv[b][k*Cc + c][hw] (reference models.py:64-67, no epsilon).  The K views, the cat and the
 * seven elementwise ops of squash are one pass over the data.  Backward: dconv from du.
 * n_caps <= 16; u 16-byte aligned. */
int caps_primary_squash(const float* conv, float* u, int B, int n_caps, int Cc, int HW, void* stream);
int caps_primary_squash_backward(const float* conv, const float* du, float* dconv,
                                 int B, int n_caps, int Cc, int HW, void* stream);

/* DarkCapsuleNet cell regroup, the step directly before the routing layer in that model (SURVEY.md
 * section 8(f) row 2; reference models.py:393-399: `x.view(B,256,4,4*g*g)`, `torch.chunk(.., g*g, 3)`,
 * then per chunk `permute(0,2,3,1).contiguous().view(B,-1,8).unsqueeze(0)`, `torch.cat(.., 0)`,
 * `.view(-1,512,8)`).  x [B][Cch][16*G] (G = g*g cells), u [G*B][2*Cch][8]:
 *   u[q*B + b][(a*4 + t)*(Cch/8) + ch/8][ch%8] = x[b][ch][a*4*G + 4*q + t],  q < G, a < 4, t < 4.
 * One pass instead of ~2 G + 2 kernels and two full copies; the backward is the inverse map.
 * Cch % 8 == 0; u / du 16-byte aligned. */
int caps_dark_regroup(const float* x, float* u, int B, int Cch, int G, void* stream);
int caps_dark_regroup_backward(const float* du, float* dx, int B, int Cch, int G, void* stream);

/* DarkCapsuleNet loss tail, the step directly after the routing layer in that model (SURVEY.md
 * section 8(f) row 3; reference loss_fns.py:187-204 `darkcapsule_loss` with `utils.polar_transform`,
 * utils.py:69-85; reconstruction term off).  v [G*B][5]: the routing layer's output, row q*B + b =
 * cell q of sample b (what models.py:400 returns before its view/permute); y [B][G][Y], Y >= 5, the
 * label tensor's (r, x, y, w, h) in columns 0..4.
 *   loss = scale * sum over cells of  y_r relu(0.9-|v|)^2 + 0.5 (1-y_r) relu(|v|-0.1)^2 - v . y_phi
 * (scale = 1/B is the reference's `/ y.size(0)`).  grad_v (nullable) [G*B][5] = d loss / d v, in the
 * layout caps_route_backward takes as grad_v.  scratch: CAPS_MARGIN_SCRATCH_FLOATS floats or NULL
 * (single block).  One kernel for the norm, both relu branches, the four sin/cos pairs, the
 * products, the sum -- and the gradient autograd would need ~25 more launches for. */
int caps_dark_loss(const float* v, const float* y, float scale, float* loss, float* grad_v,
                   float* scratch, int B, int G, int Y, void* stream);

/* HOST-buffer step, the end-to-end call: copies u (and y) host->device, runs forward, margin
 * loss, fused backward, and copies loss (and, if non-NULL, v / du / dW) device->host, all on
 * `stream`, then synchronises that stream.  u_host/y_host/..._host are HOST pointers (pinned for
 * speed).  W_dev/dW_dev stay on the device (weights live there).  dev_scratch is a device
 * buffer of >= caps_route_step_host_scratch_bytes(...) bytes.  From B >= 2048 the batch is cut
 * into three micro-batches (B/8, 3B/8, B/2) whose copies run on an internal stream under the
 * previous micro-batch's kernels; dW is their sum in a fixed order (caps_set_tuning("hostmb", 1)
 * forces a single batch; the scratch size depends on the setting at sizing time). */
size_t caps_route_step_host_scratch_bytes(int B, int N, int C, int K, int D, int R);
int caps_route_step_host(const float* u_host, const int64_t* y_host, const float* W_dev,
                         float* loss_host, float* v_host, float* du_host, float* dW_dev,
                         void* dev_scratch, size_t scratch_bytes,
                         int B, int N, int C, int K, int D, int R, void* stream);

/* HOST pipeline: like caps_route_step_host, but the host->device copy of the NEXT step's inputs runs on an
 * internal stream while the current step computes (two device input slots).  Replaces the per-batch
 * `torch.from_numpy(x).to(device)` of the reference's train loop (main.py:57-59) on the routing-alone path.
 *   scratch   caps_host_pipe_scratch_bytes(...) bytes of device memory, owned by the caller for the pipe's lifetime
 *   submit    enqueues the H2D copy of one batch (u_host [B,N,K], y_host [B] int64; HOST, pinned for overlap) into
 *             the free slot; at most two batches may be submitted and not yet stepped (CAPS_E_STATE otherwise)
 *   step      waits for the oldest submitted batch, runs forward + margin loss + fused backward on `stream`
 *             (du is computed like in any backward and stays in the pipe's device scratch), copies the loss (and v if non-NULL) to the
 *             host and synchronises `stream`.  dw_ready_event as in caps_route_backward_ev.
 * Typical loop:  submit(b0); for n: { submit(b[n+1]); step(...); }   A pipe belongs to one host thread at a time. */
size_t caps_host_pipe_scratch_bytes(int B, int N, int C, int K, int D, int R);
int caps_host_pipe_create(void** pipe_out, void* dev_scratch, size_t scratch_bytes, int B, int N, int C, int K, int D, int R);
int caps_host_pipe_submit(void* pipe, const float* u_host, const int64_t* y_host);
int caps_host_pipe_step(void* pipe, const float* W_dev, float* dW_dev, float* loss_host, float* v_host, void* stream,
                        void* dw_ready_event);
int caps_host_pipe_destroy(void* pipe);

/* Measurement helpers (bench.py).  caps_kernel_launch_count: kernels this library has launched in
 * this process.  With caps_set_tuning("profile", 1) every launch is bracketed by CUDA events on
 * its own stream; caps_profile_collect synchronises them, sums milliseconds / launch counts per
 * kernel class (0 layout, 1 pass-A0, 2 pass-L, 3 pass-A, 4 squash, 5 softmax, 6 grad, 7 du-reduce,
 * 8 loss, 9 other, 10 fused sweep, 11 single-capsule kernels) and resets.  Profiling state is process-global: single-threaded use only.
 * caps_fma_peak: times `iters` x 16 dependent-chain FFMAs per thread on a full grid and returns
 * milliseconds and the flop count (the fp32-FMA roofline denominator). */
long caps_kernel_launch_count(void);
int caps_profile_collect(double* ms_by_class, long* count_by_class, int n_classes);
int caps_fma_peak(int iters, float* ms_out, double* flops_out, void* stream);

/* Tuning knobs (process-wide, read at call time; defaults are chosen per shape).
 *   name = "spt"  samples per thread in the pass kernels (1, 2 or 4; 0 = auto)
 *   name = "isplit" forced number of splits of the N range (0 = auto)
 *   name = "tc"   1 (default): tcgen05 tensor-core pass kernel where it applies (9 <= D <= 16 and C >= 4, or
 *                 17 <= D <= 48 and C >= 2; D is zero-padded to 16 / 24 / 32 / 48);
 *                 0: fp32-FMA pass kernel everywhere
 *   name = "fused" 1 (default): one cluster-fused sweep per routing iteration (logits -> softmax -> weighted sum, and
 *                 its backward counterpart) where it applies (9 <= D <= 16, 4 <= C <= 64); 0: three kernels per iteration
 *                 with max-subtracted softmax (the fused sweep exponentiates logits directly: |logit| must stay < 80)
 *   name = "fsws" 1 (default): warp-specialised epilogue of the fused sweep (8 logit warps + 8 accumulate warps per CTA);
 *                 0: 8 epilogue warps doing both halves (the round-2 first version; same arithmetic, same order)
 *   name = "c1"   1 (default): dedicated GEMM + squash kernels for a single class capsule (C == 1, D <= 8); 0: general kernels
 *   name = "c1v"  which generation of those kernels: 2 (default) two launches per forward + backward, no dW partials;
 *                 1 the first generation (three launches), kept as a cross-check
 *   name = "tcstages" shared-memory ring depth of the tcgen05 pass kernel, 2..12 (default 10)
 *   name = "gradmma" 1 (default): tensor-core (mma.sync 3xTF32) gradient kernel where it applies
 *                 (D >= 9, C >= 7); 0: fp32-FMA gradient kernel everywhere
 *   name = "gradjw" output capsules per CTA of that kernel: 0 (default) auto, 8 or 11
 *   name = "hostmb" caps_route_step_host micro-batching: 0 (default) auto, 1 single batch
 *   name = "profile" 1: bracket every launch with CUDA events (see caps_profile_collect)
 * Returns 0, or CAPS_E_BADARG for an unknown name/value.  Meant for benchmarks and tests. */
int caps_set_tuning(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* CAPS_ROUTING_H_ */
