#!/usr/bin/env python
"""bench.py -- capsule-routing samples/sec, fwd+bwd, on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--scaling weak|strong] [--graph]
                    [--workload cfg2|cfg1|cfg3|cfg3_head]

A "step" is one pass of the hot path over one batch of synthetic input: routing forward
(reference models.py:70-79), margin loss (loss_fns.py:12-17,23), backward to du and dW
(autograd of the same), on BASELINE.json configs[1]: 1152 primary capsules (8D) -> 43 class
capsules (16D), 3 routing iterations.  Per-GPU batch is fixed (weak scaling); ranks shard the
batch with no data-path collective and average dW with one NCCL all-reduce per step.

Prints ONE JSON line (rank 0).  `value` times K steps with inputs resident in HBM (CUDA events,
max over ranks); `e2e` times the same K steps through the host-buffer C-ABI call
(caps_route_step_host: H2D of u,y from pinned memory and D2H of the loss inside the timed region);
`roofline` is for the dominant kernel class (the gradient sweep, or the fused sweeps) from CUDA events recorded
around every launch in a second run of the same K steps; `cpu_baseline` / `--impl reference` time the UNMODIFIED reference layer (baseline/_ref, brought along by
baseline/install_reference.py; `kind: "reference"`) on this box's host cores over a bounded sample -- or the
oracle's op-for-op torch port (`kind: "port"`) when the reference is not installed.  `eager_b200` is the same
unmodified reference run eagerly on the B200 (informational second baseline).  After the timed region a 64-sample
slice of the benchmarked batch is checked against the fp64 C oracle (`parity_checked`).

Other workloads (not the driver's line; numbers go to profiles/): `--workload cfg1` the reference's CapsNet train
step at batch 16, `--workload cfg3` its DarkCapsuleNet step at batch 32 x 224^2 (both: reference model code with
`models.CapsuleLayer` replaced by the B200-native layer, next to the unmodified reference on the CPU and eagerly on
the GPU), `--workload cfg3_head` the DarkCapsuleNet head's routing shape alone.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES, N_CAPS, IN_C, OUT_C, N_ITER = 1152, 43, 8, 16, 3
METRIC = 'capsule-routing samples/sec fwd+bwd'
UNIT = 'samples/s'
KCLASS = ['layout', 'pass_A0', 'pass_L', 'pass_A', 'squash', 'softmax', 'grad', 'du_reduce', 'loss', 'other', 'fused_sweep', 'c1_kernels']


def workload_name(batch):
    return 'configs[1]: routing layer alone, %d primary caps (8D) -> %d class caps (16D), %d iters, batch %d per GPU' % (
        N_NODES, N_CAPS, N_ITER, batch)


def flops_per_sample(N=N_NODES, C=N_CAPS, K=IN_C, D=OUT_C, R=N_ITER):
    """Algorithmic fwd+bwd flops: 6 NCKD + (6R-4) 2 NCD."""
    return 6.0 * N * C * K * D + (6 * R - 4) * 2.0 * N * C * D


def hbm_bytes_per_step(B, N=N_NODES, C=N_CAPS, K=IN_C, D=OUT_C, R=N_ITER):
    """Algorithmic HBM bytes: (12 NK + (8+8R) CD) per sample + 12 NCKD per batch."""
    return B * (12.0 * N * K + (8 + 8 * R) * C * D) + 12.0 * N * C * K * D


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(',')[1]) for _, l in self.lines[-3:] if len(l.split(',')) > 2] or [0.0]
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the ONLY places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(sample):
    """-> (step(), threads, kind): one fwd + margin loss + bwd of a `sample`-sized micro-batch on the host cores.
    kind 'reference': the unmodified reference CapsuleLayer + capsule_loss from baseline/_ref;
    kind 'port': oracle/routing_torch.py (same ATen ops, op for op) when the reference is not installed."""
    import torch
    from oracle import routing_np as onp
    torch.set_num_threads(os.cpu_count() or 1)
    u, W, y = onp.make_inputs(sample, N_NODES, N_CAPS, IN_C, OUT_C, seed=0)
    ut, yt = torch.from_numpy(u), torch.from_numpy(y)
    try:
        from baseline import refload, workloads
        ref_ok = refload.load() is not None
    except Exception:
        ref_ok = False
    if ref_ok:
        ref_step = workloads.make_routing_reference_step(torch.device('cpu'), N_NODES, N_CAPS, IN_C, OUT_C, N_ITER)
        Wt = torch.from_numpy(W)

        def step():
            return ref_step(ut, Wt, yt)
        return step, torch.get_num_threads(), 'reference'
    from oracle import routing_torch as ot
    W5 = torch.from_numpy(W)[None]

    def step():
        return ot.routing_step_t(ut, W5, yt, N_ITER)
    return step, torch.get_num_threads(), 'port'


def run_cpu_baseline(sample=64, reps=2):
    """Times the torch op-for-op port of the reference path (oracle/routing_torch.py) on the host."""
    step, threads, kind = cpu_reference_step_fn(sample)
    step()                                   # warm-up (allocator, thread pool)
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {'value': sample / dt, 'unit': UNIT, 'cores': threads, 'kind': kind,
            'sample': '%d steps of a %d-sample micro-batch of the same workload (reference needs ~65 MB/sample; '
                      'its samples/s is flat in batch)' % (reps, sample),
            'ms_per_microbatch': dt * 1e3}


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    sample = args.cpu_sample
    step, threads, kind = cpu_reference_step_fn(sample)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.batch),
                   'note': ('reference CPU path: the UNMODIFIED reference CapsuleLayer + capsule_loss (baseline/_ref)'
                            if kind == 'reference' else
                            'reference CPU path: torch op-for-op port (oracle/routing_torch.py; baseline/_ref not installed)')
                           + '; each step is a %d-sample micro-batch' % sample},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': kind,
                         'sample': '%d-sample micro-batch per step' % sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def main_gpu(args, rank, world, device):
    import torch
    import torch.distributed as dist
    import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi
    L = _cabi.lib()                       # raises if the CUDA library is missing (no fallback)
    B, N, C, K, D, R = args.batch, N_NODES, N_CAPS, IN_C, OUT_C, N_ITER
    if args.scaling == 'strong':
        B = max(128, (args.batch // world) // 128 * 128)      # fixed GLOBAL batch, sharded
    if args.spt:
        _cabi.set_tuning('spt', args.spt)
    if args.isplit:
        _cabi.set_tuning('isplit', args.isplit)
    for kv in filter(None, args.tune.split(',')):
        k, val = kv.split('=')
        _cabi.set_tuning(k, int(val))
    torch.manual_seed(1234 + rank)
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)

    # synthetic inputs of the reference's shape/value range: u = squash(N(0,1)), W = 0.1 N(0,1)
    x = torch.randn(B, N, K, generator=g)
    sq = (x ** 2).sum(-1, keepdim=True)
    u_host = ((sq / (1 + sq)) * x / sq.sqrt()).contiguous().pin_memory()
    y_host = torch.randint(0, C, (B,), generator=g).pin_memory()
    gw = torch.Generator(device='cpu').manual_seed(99)           # same weights on every rank
    W = (0.1 * torch.randn(N, C, K, D, generator=gw)).to(device)
    u = u_host.to(device)
    y = y_host.to(device)
    v = torch.empty(B, C, D, device=device)
    du = torch.empty(B, N, K, device=device)
    # the product's data-parallel plumbing: dW lives in the flat gradient bucket the collective reduces (parallel.GradBucket)
    Wparam = torch.nn.Parameter(W)
    bucket = pkg.GradBucket([Wparam])
    dW = Wparam.grad
    dw_ready = torch.cuda.Event()
    dw_ready.record()                      # materialises the cudaEvent_t handle
    loss = torch.empty((), device=device)
    lscr = torch.empty(_cabi.MARGIN_SCRATCH_FLOATS, device=device)
    nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, R, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()

    def step_local():
        # one rank's share of a step: forward, margin loss, backward (no collective: also what the parity check runs)
        _cabi.check(L.caps_route_forward(P(u), P(W), P(v), None, P(ws), nbytes, B, N, C, K, D, R, 1, stream), 'fwd')
        _cabi.check(L.caps_margin_loss(P(v), P(y), 1.0 / B, P(loss), None, P(lscr), B, C, D, stream), 'loss')
        bucket.wait()                      # the previous step's all-reduce must be done before dW is rewritten (stream-level)
        _cabi.check(L.caps_route_backward_ev(P(u), P(W), None, P(y), 1.0 / B, None, P(du), P(dW), P(ws), nbytes,
                                             B, N, C, K, D, R, stream, dw_ready.cuda_event), 'bwd')

    def step_device():
        step_local()
        if world > 1:
            # gradient AVERAGE across ranks on the bucket's comm stream, gated by the event the backward records right
            # behind the kernel that completes dW: overlaps the du reduction and the next step's forward
            bucket.allreduce_async(dw_ready, average=True)

    def barrier():
        bucket.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        bucket.wait()                      # the last step's all-reduce belongs to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t1 = time.time()
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, t0, t1

    # ---- kernel-resident throughput -----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(device.index if device.index is not None else 0)
    clocks.start()
    time.sleep(0.25)
    launches0 = L.caps_kernel_launch_count()
    ms, t0, t1 = timed(step_device, args.steps)                 # THE timed region (no per-kernel events)
    launches = L.caps_kernel_launch_count() - launches0
    clk = clocks.stop(t0, t1)
    # same K steps again with CUDA events around every launch (on the launching stream): per-kernel-class
    # durations for the roofline.  Kept out of the headline region: 2 event records per launch cost
    # ~1 ms per step of host time, which would dominate at small batch.
    _cabi.set_tuning('profile', 1)
    ms_prof, _, _ = timed(step_device, args.steps)
    ms_cls = (ctypes.c_double * len(KCLASS))()
    n_cls = (ctypes.c_long * len(KCLASS))()
    _cabi.check(L.caps_profile_collect(ms_cls, n_cls, len(KCLASS)), 'profile')
    _cabi.set_tuning('profile', 0)
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    loss_val = float(loss)

    # ---- the same step replayed from a CUDA graph (one launch instead of 17): matters where the step is launch-bound -----
    graphed = None
    if args.graph and world == 1:
        try:
            gs = pkg.GraphedStep(B, N, C, K, D, R, W)
            gs.u.copy_(u); gs.y.copy_(y)
            for _ in range(3):
                gs.replay()
            ms_g, _, _ = timed(gs.replay, args.steps)
            graphed = {'ms_per_step': ms_g / args.steps, 'value': B / (ms_g / args.steps * 1e-3), 'unit': UNIT, 'loss': float(gs.loss),
                       'what': 'forward + margin loss + backward captured once in a CUDA graph (GraphedStep), replayed per step'}
            del gs
            torch.cuda.empty_cache()
        except Exception as e:
            graphed = {'unavailable': repr(e)[:300]}

    # ---- end to end through the host-buffer C-ABI call ----------------------------------------
    # caps_host_pipe_*: every step copies ITS inputs host -> device (pinned memory) and its loss device -> host; the copy
    # of step n+1 is submitted before step n runs, so it overlaps the kernels (two device input slots)
    host = pkg.HostPipe(B, N, C, K, D, R, device=device)
    host.submit(u_host, y_host)

    def step_host():
        host.submit(u_host, y_host)        # next step's inputs: H2D on the pipe's copy stream
        bucket.wait()
        host.step(W, dW, dw_ready_event=dw_ready)
        if world > 1:
            bucket.allreduce_async(dw_ready, average=True)
    for _ in range(max(1, min(args.warmup, 3))):
        step_host()
    ms_e2e, _, _ = timed(step_host, args.steps)
    e2e_val = world * B / (ms_e2e / args.steps * 1e-3)

    barrier()                              # every collective of this run is complete on every rank
    if rank != 0:
        return 0

    # ---- parity of THIS run: a 64-sample slice of the benchmarked batch against the fp64 C oracle ------------------
    # (samples are independent: v and du of a sample do not depend on its batch-mates; 1/B is the loss scale)
    parity = {'parity_checked': False}
    try:
        import numpy as np
        from oracle import routing_c as oc
        step_local()                       # rank 0 only from here on: nothing below may enter a collective
        torch.cuda.synchronize()
        ns = min(64, B)
        ref = oc.routing_step(u_host[:ns].numpy().astype(np.float64), W.cpu().numpy().astype(np.float64),
                              y_host[:ns].numpy(), R, inv_batch=1.0 / B, want_c=False)

        def rel(a, b):
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        ev, eu = rel(v[:ns].cpu().numpy(), ref['v']), rel(du[:ns].cpu().numpy(), ref['du'])
        parity = {'parity_checked': bool(ev < 1e-5 and eu < 1e-4), 'parity': {'samples': ns, 'rel_err_v': ev, 'rel_err_du': eu,
                  'tolerance': 'v 1e-5, du 1e-4 (max-norm relative, vs the fp64 C oracle); dW is covered by tests/ at B=512 and 2176'}}
    except Exception as e:       # the checker must never take the measurement down
        parity = {'parity_checked': False, 'parity': {'error': repr(e)[:200]}}

    # ---- second, informational baseline: the UNMODIFIED reference layer run eagerly on this B200 -------------------
    eager = None
    if not args.no_eager:
        try:
            from baseline import workloads
            mb = min(args.eager_batch, B)
            ref_step = workloads.make_routing_reference_step(device, N, C, K, D, R)
            ue, ye = u[:mb].contiguous(), y[:mb].contiguous()
            sec = workloads.time_steps(lambda: ref_step(ue, W, ye), 3, 2, True)
            eager = {'value': mb / sec, 'unit': UNIT, 'ms_per_microbatch': sec * 1e3, 'micro_batch': mb,
                     'what': 'reference models.CapsuleLayer + loss_fns.capsule_loss, unmodified (baseline/_ref), eager PyTorch on the same GPU; '
                             'its autograd state is ~65 MB per sample, hence the micro-batch'}
            del ref_step, ue, ye
            torch.cuda.empty_cache()
        except Exception as e:
            eager = {'unavailable': repr(e)[:200]}

    # ---- rooflines --------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    fma_ms, fma_fl = ctypes.c_float(), ctypes.c_double()
    _cabi.check(L.caps_fma_peak(20000, ctypes.byref(fma_ms), ctypes.byref(fma_fl), stream), 'fma_peak')
    fma_peak = fma_fl.value / (fma_ms.value * 1e-3) / 1e12

    per_class = {KCLASS[i]: {'ms_per_step': ms_cls[i] / args.steps, 'launches_per_step': n_cls[i] / args.steps}
                 for i in range(len(KCLASS)) if n_cls[i]}
    # --- per-kernel-class algorithmic work of ONE launch ------------------------------------
    #   uniform / L / A sweep : flops B (2 NCKD + 2 NCD)        bytes B (4 NK + 4 NC) + 4 NCKD
    #   fused sweep           : flops B (2 NCKD + 4 NCD)        bytes B (4 NK + a 4 NC) + 4 NCKD, a = [B,N,C] arrays it
    #                           touches: 1 (forward: c written), 2 (top backward: c read, beta written), 3 (inner backward)
    #   gradient sweep        : flops B (4 NCKD + 2 (2R-1) NCD) bytes B (4 (2R-2) NC + 8 NK) + 8 NCKD
    #   single-capsule kernels: bytes B 12 NK + 12 NKD (u read twice, du written; W read twice, dW written)
    Re = 1 if C == 1 else R
    cls_bytes, cls_flops = {}, {}
    sweep_flops = B * (2.0 * N * C * K * D + 2.0 * N * C * D)
    for c in (1, 2, 3):
        cls_bytes[c] = B * (4.0 * N * K + 4.0 * N * C) + 4.0 * N * C * K * D
        cls_flops[c] = sweep_flops
    arrays_fused = (Re - 1) * 1 + 2 + 3 * max(Re - 2, 0)                        # per step, over its 2 (Re - 1) launches
    cls_bytes[10] = (2 * (Re - 1) * (B * 4.0 * N * K + 4.0 * N * C * K * D) + arrays_fused * B * 4.0 * N * C) / max(2 * (Re - 1), 1)
    cls_flops[10] = B * (2.0 * N * C * K * D + 4.0 * N * C * D)
    cls_bytes[6] = B * (4.0 * (2 * Re - 2) * N * C + 8.0 * N * K) + 8.0 * N * C * K * D
    cls_flops[6] = B * (4.0 * N * C * K * D + 2.0 * (2 * Re - 1) * N * C * D)
    cls_bytes[11] = (B * 12.0 * N * K + 12.0 * N * K * D) / 3.0                    # three launches per step share it
    cls_flops[11] = B * 6.0 * N * K * D / 3.0
    names = {1: 'k_pass_tc<uniform> (u_hat sweep, tcgen05 3xTF32, TMEM-accumulated)', 2: 'k_pass_tc<L>', 3: 'k_pass_tc<A>',
             6: 'k_grad_mma (dW/du sweep, mma.sync 3xTF32)' if C >= 7 and D >= 9 else 'k_grad (dW/du sweep, fp32 FMA)',
             10: 'k_sweep_fused (logits -> softmax -> weighted sum in one sweep; tcgen05 3xTF32 + DSMEM exchange)',
             11: 'k_c1_fwd / k_c1_bwd / k_c1_reduce (single class capsule: skinny GEMM + squash)'}
    traffic_by_class = {}
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')))
        if tj.get('batch') == B and tj.get('n_nodes') == N and tj.get('n_caps') == C:
            traffic_by_class = {int(k): v for k, v in tj.get('dram_bytes_per_launch_by_class', {}).items()}
    except Exception:
        pass

    def roof(c):
        if not n_cls[c] or c not in cls_bytes:
            return None
        ms1 = ms_cls[c] / n_cls[c]
        gbs = cls_bytes[c] / (ms1 * 1e-3) / 1e9
        return {'bound': 'hbm', 'kernel': '%s; %d launches/step' % (names.get(c, KCLASS[c]), n_cls[c] // args.steps),
                'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                'traffic': traffic_by_class.get(c), 'peak_source': hbm_src, 'share_of_step': ms_cls[c] / ms_prof,
                'ms_per_launch': ms1, 'algorithmic_bytes_per_launch': cls_bytes[c],
                'algorithmic_tflops': cls_flops[c] / (ms1 * 1e-3) / 1e12}
    cand = [c for c in cls_bytes if n_cls[c]]
    dom = max(cand, key=lambda c: ms_cls[c])
    roofline = roof(dom)
    binds = {6: 'instruction issue latency, not HBM and not the tensor pipe: 11 + 1 warps per SM is what the register file holds '
                '(80 registers of per-sample state per thread), every warp is a stream of short dependent steps (~200 '
                'instructions per 32-sample unit, a third of them the tf32 hi/lo splits of the 3xTF32 products); ncu: issue 45 %, '
                'shared-memory pipe 56-63 %, legacy HMMA pipe 30 %; timing probes with all of G\'s shared-memory traffic off (-2 %), '
                'shorter HMMA chains and a software-pipelined G build (slower) rule the other candidates out '
                '(profiles/r2_analysis.md); tools/probe_tc2.cu + DESIGN.md 3.3 say why tcgen05 does not help',
             10: 'the instruction stream of the 8 logit + 8 accumulate warps per CTA (1966 warp instructions per 128-sample x '
                 '8-capsule stage, ~950 cycles); ncu: issue slots ~55 % busy, FMA pipe 42 % of cycles, tensor pipe 21-24 %, DRAM '
                 '20-36 %; a scratch build without the cluster exchange is no faster, so the per-stage cluster synchronisation '
                 'is not on the critical path (profiles/r2_ncu_summary.md, r2_analysis.md)',
             11: 'HBM (u is read twice and du written once; nothing else is large)'}
    roofline['binds'] = binds.get(dom, 'see DESIGN.md section 5')
    roofline_other = {KCLASS[c]: roof(c) for c in cand if c != dom}
    # tensor-pipe view of the sweeps: 3xTF32 issues 3 MMAs per algorithmic one; tf32 dense peak = half the measured bf16 peak
    tf32_peak = float(peaks.get('bf16_tflops', 1590.0)) / 2.0
    for c in (1, 10):
        r_ = roofline if c == dom else roofline_other.get(KCLASS[c])
        if r_:
            mma_tf = 3.0 * B * 2.0 * N * C * K * D * (1 if c == 1 else 1) / (r_['ms_per_launch'] * 1e-3) / 1e12
            r_['tensor'] = {'issued_tflops_3xtf32': mma_tf, 'peak': tf32_peak, 'frac': mma_tf / tf32_peak,
                            'peak_source': 'MEASURED_PEAKS.json bf16_tflops / 2 (tf32 runs at half the bf16 rate)'}
    step_tf = flops_per_sample() * B / (ms_per_step * 1e-3) / 1e12
    step_gbs = hbm_bytes_per_step(B) / (ms_per_step * 1e-3) / 1e9
    roofline_step = {'algorithmic_tflops': step_tf, 'frac_of_fp32_peak': step_tf / fma_peak,
                     'algorithmic_gbs': step_gbs, 'frac_of_hbm_peak': step_gbs / hbm_peak}

    cpu = run_cpu_baseline(args.cpu_sample, 2) if not args.no_cpu else None

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(B), 'n_nodes': N, 'n_caps': C, 'in_C': K, 'out_C': D, 'n_iter': R,
                   'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': 'dp%d' % world,
                   'l2': 'inputs larger than L2 (u = %.0f MB per step; saved coupling arrays %.1f GB)' % (
                       B * N * K * 4 / 1e6, 4.0 * B * N * C * 4 / 1e9),
                   'loss': loss_val},
        'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': host.h2d_bytes,
                'd2h_bytes_per_step': host.d2h_bytes, 'ms_per_step': ms_e2e / args.steps,
                'api': 'caps_host_pipe_submit/step (pinned host u,y -> device every step, prefetched one step ahead; loss -> host)'},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': roofline, 'roofline_other_kernels': roofline_other, 'roofline_step': roofline_step,
        'kernel_ms_per_step': per_class, 'profiled_ms_per_step': ms_prof / args.steps,
        'cpu_baseline': cpu,
        'eager_b200': eager,
        'cuda_graph': graphed,
    }
    line.update(parity)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# full-model workloads: BASELINE.json configs[0] (cfg1), configs[2] (cfg3) and configs[3] (cfg3 with --gpus N)
# ------------------------------------------------------------------------------------------------
def main_model(args, rank, world, device):
    """The reference's own train step (main.py:55-77: H2D of the numpy batch, forward, loss, .cpu() of the prediction,
    zero_grad / backward / Adam step, loss.item()) on the reference's own model code, with `models.CapsuleLayer`
    replaced by the B200-native layer; next to it the unmodified reference eagerly on the same GPU and on the host
    cores.  With --gpus N (cfg3 = BASELINE.json configs[3]) every rank steps on its own shard and ALL gradients
    (backbone + routing weights) are averaged through one flat NCCL bucket per step."""
    import torch
    import torch.distributed as dist
    from baseline import refload, workloads
    if refload.load() is None:
        if rank == 0:
            print(json.dumps({'workload': args.workload, 'unavailable': 'reference not installed under baseline/_ref (run __graft_entry__.build() where /root/reference is mounted)'}))
        return 0
    cfg = args.workload
    B = args.batch if args.batch != 8192 else (16 if cfg == 'cfg1' else 32)
    torch.backends.cudnn.allow_tf32 = False            # fp32 convolutions on every arm, like the CPU reference
    torch.backends.cuda.matmul.allow_tf32 = False
    x, y = (workloads.synth_cfg1 if cfg == 'cfg1' else workloads.synth_cfg3)(B, seed=rank)

    def build(dev, dropin, **kw):
        if cfg == 'cfg1':
            return workloads.make_cfg1_step(dev, dropin=dropin)[0]
        return workloads.make_cfg3_step(dev, dropin=dropin, **kw)[0]

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step(x, y)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step(x, y)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        if world > 1:
            t = torch.tensor([sec], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t)
        return sec, loss

    out = {}
    variants = [('dropin', dict(dropin=True))]
    if cfg == 'cfg3':
        variants.append(('dropin_fused_tail', dict(dropin=True, fused_tail=True)))
    if world == 1:
        variants.append(('reference_eager_b200', dict(dropin=False)))
    for name, kw in variants:
        if cfg == 'cfg3':
            kw = dict(kw, world=world)
        step = build(device, **kw)
        sec, loss = timed(step, args.steps, args.warmup)
        out[name] = {'samples_per_s': world * B / sec, 'ms_per_step': sec * 1e3, 'loss': loss}
        del step
        torch.cuda.empty_cache()
    if cfg == 'cfg1' and world == 1:
        # the same step through the sync-free side runner: an epoch of `steps` batches per call;
        # graphed: every step replayed from one CUDA graph (forward, loss, backward, Adam)
        import numpy as np
        for name, graph in (('dropin_syncfree_runner', False), ('dropin_graphed_runner', True)):
            epoch = workloads.make_cfg1_epoch(device, B, graph=graph)
            epoch(np.concatenate([x] * max(args.warmup, 3)), np.concatenate([y] * max(args.warmup, 3)))
            xe, ye = np.concatenate([x] * args.steps), np.concatenate([y] * args.steps)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            avg = epoch(xe, ye)
            torch.cuda.synchronize()
            sec = (time.perf_counter() - t0) / args.steps
            out[name] = {'samples_per_s': B / sec, 'ms_per_step': sec * 1e3, 'loss': avg,
                         'what': 'runner.train: main.py:38-95 with staged copies and one host synchronisation per epoch instead of three per step'
                                 + ('; every step replayed from one CUDA graph' if graph else '')}
            del epoch
            torch.cuda.empty_cache()
    if rank != 0:
        return 0
    cpu = None
    if not args.no_cpu and world == 1:
        threads = workloads.cpu_threads()
        Bc = B if cfg == 'cfg1' else min(B, 4)           # a DarkCapsuleNet step is ~90 GFLOP per sample: bounded sample
        xc, yc = x[:Bc], y[:Bc]
        step = build(torch.device('cpu'), False)
        step(xc, yc)
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            step(xc, yc)
        sec = (time.perf_counter() - t0) / n
        cpu = {'value': Bc / sec, 'unit': 'samples/s', 'cores': threads, 'kind': 'reference',
               'sample': '%d steps of batch %d of the same workload' % (n, Bc), 'ms_per_step': sec * 1e3}
    main_v = out['dropin']
    line = {
        'metric': 'train-step samples/sec (reference main.py:55-77 step on the reference model, B200-native CapsuleLayer dropped in)',
        'value': main_v['samples_per_s'], 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': main_v['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': ('configs[0]: CapsNet (conv1 -> primary caps -> routing 1296->43x16 -> decoder), recon on, Adam, batch %d' % B)
                   if cfg == 'cfg1' else
                   ('configs[%d]: DarkCapsuleNet (5-conv backbone -> regroup -> routing 512->1x5 per cell) at 224x224, --recon --no_metric, Adam, batch %d per GPU'
                    % (3 if world > 1 else 2, B)),
                   'parallelism': 'dp%d' % world,
                   'allreduce': 'one flat NCCL bucket per step over backbone + routing gradients (parallel.GradBucket)' if world > 1 else None},
        'e2e': {'value': main_v['samples_per_s'], 'unit': 'samples/s', 'h2d_bytes_per_step': int(x.nbytes + y.nbytes),
                'd2h_bytes_per_step': int(4 + (B * 43 * 4 if cfg == 'cfg1' else B * 49 * 5 * 4)),
                'note': 'the step IS end to end: numpy batch -> device, prediction and loss -> host, every step'},
        'variants': out, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=8192, help='per-GPU batch (weak scaling)')
    ap.add_argument('--cpu-sample', type=int, default=64, help='micro-batch of the CPU baseline')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak (default): --batch per GPU; strong: --batch is the global batch, sharded over the GPUs')
    ap.add_argument('--graph', action='store_true', help='also time the step replayed from a CUDA graph (N = 1)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-eager', action='store_true', help='skip the eager-reference-on-GPU leg')
    ap.add_argument('--eager-batch', type=int, default=256, help='micro-batch of the eager reference on the GPU')
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg1', 'cfg3', 'cfg3_head'],
                    help='cfg2 (default, the line the driver reads): routing layer alone; cfg1: CapsNet train step B=16; '
                         'cfg3: DarkCapsuleNet train step B=32 @224; cfg3_head: the DarkCapsuleNet head routing shape alone')
    ap.add_argument('--spt', type=int, default=0)
    ap.add_argument('--isplit', type=int, default=0)
    ap.add_argument('--tune', default='', help='comma list name=value passed to caps_set_tuning')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return main_reference(args, rank, world)
    from cs231_capsule_yolo_traffic_sign_detection_b200.parallel import init_from_env
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    rank, world, device = init_from_env()
    if args.workload == 'cfg3_head':
        # the DarkCapsuleNet head (reference models.py:368-370, :398): routing batch 32 * 49 cells, 512 -> 1 x 5
        global N_NODES, N_CAPS, OUT_C
        N_NODES, N_CAPS, OUT_C = 512, 1, 5
        if args.batch == 8192:
            args.batch = 1568
    try:
        if args.workload in ('cfg1', 'cfg3'):
            return main_model(args, rank, world, device)
        return main_gpu(args, rank, world, device)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
