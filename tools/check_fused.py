"""Bring-up check of the cluster-fused sweep against the fp64 C oracle (one shape per process: a device trap kills
the context).  python tools/check_fused.py B N C D R [fused]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cs231_capsule_yolo_traffic_sign_detection_b200 as m
from oracle import routing_c as oc
from oracle import routing_np as onp

B, N, C, D, R = [int(x) for x in sys.argv[1:6]]
fused = int(sys.argv[6]) if len(sys.argv) > 6 else 1
m._cabi.set_tuning('fused', fused)
u, W, y = onp.make_inputs(B, N, C, 8, D, seed=11)
ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R)
dev = torch.device('cuda')
ut = torch.from_numpy(u).to(dev).requires_grad_(True)
Wt = torch.from_numpy(W)[None].to(dev).requires_grad_(True)
yt = torch.from_numpy(y).to(dev)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


t0 = time.time()
with torch.no_grad():
    v, c = m.dynamic_routing(ut, Wt, R, return_couplings=True)
torch.cuda.synchronize()
vn, cn = v.cpu().numpy(), c.cpu().numpy()
print('dims', (B, N, C, D, R), 'fused', fused, 'fwd ok %.2fs' % (time.time() - t0), 'v', rel(vn, ref['v']), 'c', rel(cn, ref['c']), flush=True)
ev = np.abs(vn - ref['v']).reshape(B, -1).max(1)
worst = np.argsort(-ev)[:8]
print('   worst samples (b, err, |v|max):', [(int(b), float('%.2e' % ev[b]), float('%.3f' % np.abs(ref['v'][b]).max())) for b in worst], 'median err %.2e' % np.median(ev), flush=True)
ec = np.abs(cn - ref['c'])
bi = np.unravel_index(np.argmax(ec), ec.shape)
print('   worst c at (b,i,j)', bi, 'err %.2e' % ec[bi], 'ref %.4e' % ref['c'][bi], 'sum_j c -1: %.2e' % np.abs(cn.sum(-1) - 1).max(), flush=True)
v, loss = m.routing_margin_loss(ut, Wt, yt, R)
loss.backward()
torch.cuda.synchronize()
print('   loss', abs(float(loss) - ref['loss']), 'du', rel(ut.grad.cpu().numpy(), ref['du']), 'dW', rel(Wt.grad[0].cpu().numpy(), ref['dW']), flush=True)
