"""Times the routing forward+backward per kernel class at the benchmark shape for a list of tuning settings.
python tools/time_fused.py [B] [name=value,name=value ...]   (each argument after B is one setting to time)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi

KCLASS = ['layout', 'pass_A0', 'pass_L', 'pass_A', 'squash', 'softmax', 'grad', 'du_reduce', 'loss', 'other', 'fused']
L = _cabi.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N, C, K, D, R = [int(x) for x in os.environ.get('CAPS_DIMS', '1152,43,8,16,3').split(',')]
settings = sys.argv[2:] or ['fused=1']
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
x = torch.randn(B, N, K, generator=g)
sq = (x ** 2).sum(-1, keepdim=True)
u = ((sq / (1 + sq)) * x / sq.sqrt()).to(dev)
W = (0.1 * torch.randn(N, C, K, D, generator=g)).to(dev)
y = torch.randint(0, C, (B,), generator=g).to(dev)
v = torch.empty(B, C, D, device=dev); du = torch.empty(B, N, K, device=dev); dW = torch.empty_like(W)
nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, R, 1)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
_cabi.set_tuning('fsinfo', C)


def step():
    _cabi.check(L.caps_route_forward(P(u), P(W), P(v), None, P(ws), nbytes, B, N, C, K, D, R, 1, st), 'fwd')
    _cabi.check(L.caps_route_backward(P(u), P(W), None, P(y), 1.0 / B, None, P(du), P(dW), P(ws), nbytes, B, N, C, K, D, R, st), 'bwd')


ref = None
for setting in settings:
    for kv in setting.split(','):
        k, val = kv.split('=')
        _cabi.set_tuning(k, int(val))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    NREP = int(os.environ.get('CAPS_REPS', '3'))
    for _ in range(NREP):
        step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / NREP
    if ref is None:
        ref = (v.clone(), du.clone(), dW.clone())
        dev_s = ''
    else:       # deviation from the first setting's results (max-norm relative): 0 = same bits
        dev_s = ' | vs first: ' + ' '.join('%s %.1e' % (nm, float((a - b).abs().max() / b.abs().max())) for nm, a, b in zip(('v', 'du', 'dW'), (v, du, dW), ref))
    _cabi.set_tuning('profile', 1)
    step()
    ms = (ctypes.c_double * len(KCLASS))(); n = (ctypes.c_long * len(KCLASS))()
    _cabi.check(L.caps_profile_collect(ms, n, len(KCLASS)), 'profile')
    _cabi.set_tuning('profile', 0)
    print('%-28s step %.3f ms | ' % (setting, total) + ' '.join('%s %.3f/%d' % (KCLASS[i], ms[i], n[i]) for i in range(len(KCLASS)) if n[i]) + dev_s, flush=True)
