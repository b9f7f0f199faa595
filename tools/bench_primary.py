"""Times the primary-capsule branch (reference models.py:59-62, 81-82) on one GPU:
the reference's formulation (8 convolutions, 8 views, cat, 7-op squash; stock PyTorch ops on the GPU) against
this repo's (one cuDNN convolution over the concatenated weights + caps_primary_squash), forward + backward,
CapsuleNet shapes: x [B,256,24,24] -> 8 x Conv2d(256->16, 8x8, stride 2) -> u [B,1296,8].

    python tools/bench_primary.py [--batches 256,1024] [--steps 10]
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cs231_capsule_yolo_traffic_sign_detection_b200 as capsb   # noqa: E402


def reference_formulation(x, convs):
    outs = [cap(x).view(x.size(0), -1, 1) for cap in convs]          # reference models.py:81
    v = torch.cat(outs, dim=-1)
    sq = (v ** 2).sum(dim=-1, keepdim=True)                           # reference models.py:64-67
    return (sq / (1 + sq)) * v / torch.sqrt(sq)


def reference_regroup(x, g):
    B = x.size(0)
    xs = torch.chunk(x.view(B, 256, 4, 4 * g ** 2), g ** 2, 3)                                   # reference models.py:395
    xs = [xx.permute(0, 2, 3, 1).contiguous().view(B, -1, 8).unsqueeze(0) for xx in xs]          # :397
    return torch.cat(xs, 0).view(-1, 512, 8)                                                     # :398, :400


def timed(fn, x, gu, steps):
    for _ in range(3):
        x.grad = None
        fn(x).backward(gu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        x.grad = None
        fn(x).backward(gu)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='256,1024')
    ap.add_argument('--steps', type=int, default=10)
    args = ap.parse_args()
    dev = torch.device('cuda')
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    layer = capsb.CapsuleLayer(None, n_caps=8, n_nodes=-1, in_C=256, out_C=16, kernel=8, stride=2).to(dev)
    print('| B | reference formulation (8 convs + views + cat + squash) ms | one conv + caps_primary_squash ms | speed-up | max rel diff u | max rel diff dx |')
    print('|---|---|---|---|---|---|')
    for B in [int(b) for b in args.batches.split(',')]:
        x = torch.randn(B, 256, 24, 24, device=dev, requires_grad=True)
        gu = torch.randn(B, 1296, 8, device=dev)
        x.grad = None
        u_ref = reference_formulation(x, layer.capsules)
        u_ref.backward(gu)
        dx_ref = x.grad.clone()
        x.grad = None
        u = layer(x)
        u.backward(gu)
        du_rel = float((u.detach() - u_ref.detach()).abs().max() / u_ref.detach().abs().max())
        dx_rel = float((x.grad - dx_ref).abs().max() / dx_ref.abs().max())
        t_ref = timed(lambda t: reference_formulation(t, layer.capsules), x, gu, args.steps)
        t_new = timed(layer, x, gu, args.steps)
        print('| %d | %.3f | %.3f | %.2fx | %.1e | %.1e |' % (B, t_ref, t_new, t_ref / t_new, du_rel, dx_rel))


def dark():
    dev = torch.device('cuda')
    print()
    print('DarkCapsuleNet cell regroup (reference models.py:393-399), x [B,256,28,28] -> u [49 B,512,8], fwd + bwd')
    print('| B | reference formulation (view, chunk, 49 x permute/contiguous/view, cat) ms | caps_dark_regroup ms | speed-up | identical |')
    print('|---|---|---|---|---|')
    for B in (32, 256):
        x = torch.randn(B, 256, 28, 28, device=dev, requires_grad=True)
        gu = torch.randn(49 * B, 512, 8, device=dev)
        same = bool(torch.equal(reference_regroup(x, 7), capsb.dark_regroup(x, 7)))
        t_ref = timed(lambda t: reference_regroup(t, 7), x, gu, 10)
        t_new = timed(lambda t: capsb.dark_regroup(t, 7), x, gu, 10)
        print('| %d | %.3f | %.3f | %.1fx | %s |' % (B, t_ref, t_new, t_ref / t_new, same))


if __name__ == '__main__':
    main()
    dark()
