"""Hardware multi-rank parity (SURVEY section 4 tier 4): two ranks, one process per GPU over NCCL, each on its half of
the batch; the gradient bucket after the (asynchronous, event-gated) all-reduce must equal the single-GPU dW of the
concatenated batch.  Needs >= 2 GPUs: skipped on the one-GPU boxes, run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi
    from oracle import routing_np as onp
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    rank, world, dev = pkg.init_from_env()
    L = _cabi.lib()
    B, N, C, K, D, R = 384, 160, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=31)
    lo, hi = pkg.shard_bounds(B, rank, world)
    Bl = hi - lo
    ut, yt = torch.from_numpy(u[lo:hi]).to(dev), torch.from_numpy(y[lo:hi]).to(dev)
    Wp = torch.nn.Parameter(torch.from_numpy(W).to(dev))
    bucket = pkg.GradBucket([Wp])
    v = torch.empty(Bl, C, D, device=dev)
    du = torch.empty(Bl, N, K, device=dev)
    nbytes = L.caps_route_workspace_bytes(Bl, N, C, K, D, R, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ev = torch.cuda.Event()
    ev.record()
    P = lambda t: t.data_ptr()
    for _ in range(2):                                       # twice: the second step must wait for the first all-reduce
        bucket.wait()
        _cabi.check(L.caps_route_forward(P(ut), P(Wp), P(v), None, P(ws), nbytes, Bl, N, C, K, D, R, 1, st), 'fwd')
        _cabi.check(L.caps_route_backward_ev(P(ut), P(Wp), None, P(yt), 1.0 / Bl, None, P(du), P(Wp.grad), P(ws), nbytes,
                                             Bl, N, C, K, D, R, st, ev.cuda_event), 'bwd')
        bucket.allreduce_async(ev, average=True)
    bucket.wait()
    torch.cuda.synchronize()
    got = Wp.grad.cpu().numpy()
    if rank == 0:
        # single-GPU run on the concatenated batch (1/B scale): the all-reduced AVERAGE of the per-rank 1/B_local gradients
        uf, yf = torch.from_numpy(u).to(dev), torch.from_numpy(y).to(dev)
        vf, duf, dWf = torch.empty(B, C, D, device=dev), torch.empty(B, N, K, device=dev), torch.empty_like(Wp.data)
        nb = L.caps_route_workspace_bytes(B, N, C, K, D, R, 1)
        wsf = torch.empty(nb, dtype=torch.uint8, device=dev)
        _cabi.check(L.caps_route_forward(P(uf), P(Wp), P(vf), None, P(wsf), nb, B, N, C, K, D, R, 1, st), 'fwd')
        _cabi.check(L.caps_route_backward(P(uf), P(Wp), None, P(yf), 1.0 / B, None, P(duf), P(dWf), P(wsf), nb, B, N, C, K, D, R, st), 'bwd')
        torch.cuda.synchronize()
        np.savez(out, got=got, want=dWf.cpu().numpy(), du_half=du.cpu().numpy(), du_full=duf[lo:hi].cpu().numpy() * (B / Bl))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_two_rank_allreduce_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'dw.npz')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    rel = np.abs(r['got'] - r['want']).max() / np.abs(r['want']).max()
    assert rel < 1e-5, rel
    # per-sample gradients do not depend on the sharding (up to the 1/B_local vs 1/B loss scale)
    assert np.abs(r['du_half'] - r['du_full']).max() / np.abs(r['du_full']).max() < 1e-5
