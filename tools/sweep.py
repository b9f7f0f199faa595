#!/usr/bin/env python
"""Shape / batch sweep of the routing step (fwd + margin loss + bwd) on one B200.

Covers BASELINE.json configs 2 (batch sweep 64..8192 at N=1152 and N=1296), 3 (the DarkCapsuleNet
head: 512 -> 1 x 5D at batch 32*49) and 5 (iterations 1..5, N up to 4x, D 16..32).  Prints a
markdown table; `python tools/sweep.py > profiles/rNN_sweep.md` under gpurun.
Times K steps with CUDA events after W warm-up steps, inputs resident in HBM."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi  # noqa: E402


def time_step(B, N, C, D, R, steps=5, warmup=3, K=8):
    L = _cabi.lib()
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(0)
    x = torch.randn(B, N, K, generator=g)
    sq = (x ** 2).sum(-1, keepdim=True)
    u = ((sq / (1 + sq)) * x / sq.sqrt()).to(dev)
    W = (0.1 * torch.randn(N, C, K, D, generator=g)).to(dev)
    y = torch.randint(0, C, (B,), generator=g).to(dev)
    v = torch.empty(B, C, D, device=dev); du = torch.empty_like(u); dW = torch.empty_like(W)
    loss = torch.empty((), device=dev); lscr = torch.empty(_cabi.MARGIN_SCRATCH_FLOATS, device=dev)
    nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, R, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()

    def step():
        _cabi.check(L.caps_route_forward(P(u), P(W), P(v), None, P(ws), nbytes, B, N, C, K, D, R, 1, st), 'fwd')
        _cabi.check(L.caps_margin_loss(P(v), P(y), 1.0 / B, P(loss), None, P(lscr), B, C, D, st), 'loss')
        _cabi.check(L.caps_route_backward(P(u), P(W), None, P(y), 1.0 / B, None, P(du), P(dW), P(ws), nbytes,
                                          B, N, C, K, D, R, st), 'bwd')
    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = B * (6.0 * N * C * K * D + (6 * R - 4) * 2.0 * N * C * D)
    return ms, B / ms * 1e3, flops / ms / 1e9, nbytes / 2 ** 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    a = ap.parse_args()
    rows = []
    for N in (1152, 1296):
        for B in ((64, 1024, 8192) if a.quick else (64, 128, 256, 512, 1024, 2048, 4096, 8192)):
            rows.append(('cfg2 batch sweep', B, N, 43, 16, 3))
    rows.append(('cfg3 DarkCapsuleNet head', 32 * 49, 512, 1, 5, 3))
    for R in (1, 2, 3, 4, 5):
        rows.append(('cfg5 iterations', 1024, 1296, 43, 16, R))
    for N in (2592, 5184):
        rows.append(('cfg5 primary caps x2/x4', 1024, N, 43, 16, 3))
    for D in (24, 32):
        rows.append(('cfg5 class-capsule dim', 1024, 1296, 43, D, 3))
    rows.append(('DarkCapsuleNet3 head shape', 1024, 512, 43, 21, 3))
    rows.append(('DarkCapsuleNet2 head shape', 256, 784, 49, 48, 3))
    print('| case | B | N | C | D | R | engine | ms/step | samples/s | algorithmic TFLOP/s | workspace GiB |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for name, B, N, C, D, R in rows:
        ms, sps, tf, gib = time_step(B, N, C, D, R)
        eng = ('fused cluster sweep (tcgen05) + mma.sync' if (D > 8 and D <= 16 and C >= 7 and C <= 64 and R > 1) else 'tcgen05 + mma.sync' if (D > 8 and C >= 7) else 'tcgen05 passes + fp32 FMA gradient sweep' if (D > 8 and C >= 2)
               else 'single-capsule kernels' if C == 1 else 'fp32 FMA')
        print('| %s | %d | %d | %d | %d | %d | %s | %.3f | %.0f | %.2f | %.2f |' % (name, B, N, C, D, R, eng, ms, sps, tf, gib))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
