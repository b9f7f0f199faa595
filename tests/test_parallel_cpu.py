"""World-size-2 gloo test of the data-parallel plumbing (host logic only; CPU tensors)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cs231_capsule_yolo_traffic_sign_detection_b200.parallel import GradBucket, shard_bounds


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 64, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    W = torch.nn.Parameter(torch.randn(4, 3))
    b = torch.nn.Parameter(torch.randn(3))
    bucket = GradBucket([W, b])
    x = torch.arange(8 * 4, dtype=torch.float32).view(8, 4) / 10
    lo, hi = shard_bounds(8, rank, world)
    bucket.zero()
    loss = ((x[lo:hi] @ W + b) ** 2).sum() / (hi - lo)       # per-rank 1/B_local like loss_fns.py:23
    loss.backward()
    bucket.allreduce(average=True)
    first = bucket.flat.clone()
    # a second step the way the reference's loop does it (main.py:70): optimizer.zero_grad() with set_to_none=True
    # drops the views into the bucket; the bucket must notice, re-bind and still reduce the fresh gradients
    torch.optim.SGD([W, b], lr=0.0).zero_grad()
    assert W.grad is None
    loss = ((x[lo:hi] @ W + b) ** 2).sum() / (hi - lo)
    loss.backward()
    assert W.grad.data_ptr() != bucket._views[0].data_ptr()
    bucket.allreduce_async(None, average=True)               # CPU tensors: same result through the async entry point
    bucket.wait()
    assert W.grad.data_ptr() == bucket._views[0].data_ptr()
    assert torch.allclose(bucket.flat, first, rtol=1e-6, atol=1e-6)
    if rank == 0:
        torch.save(bucket.flat.clone(), out)
    dist.destroy_process_group()


def test_bucket_allreduce_matches_single_process(tmp_path):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    W = torch.nn.Parameter(torch.randn(4, 3))
    b = torch.nn.Parameter(torch.randn(3))
    x = torch.arange(8 * 4, dtype=torch.float32).view(8, 4) / 10
    (((x @ W + b) ** 2).sum() / 8).backward()
    want = torch.cat([W.grad.reshape(-1), b.grad.reshape(-1)])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
