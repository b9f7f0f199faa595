"""Randomised shape fuzz of the CUDA path against the fp64 C oracle (test infrastructure: `python tools/fuzz_shapes.py [n] [seed]`
on a GPU box).  Shapes are drawn across every engine the library picks between -- one class capsule, fp32-FMA kernels
(D <= 8 or few capsules), tensor-core sweeps with D padded to 16 / 24 / 32 / 48, the cluster-fused sweep (9 <= D <= 16,
4 <= C <= 64) -- with ragged batches / tiles / capsule groups, R in 1..5, with and without an external grad_v.
Tolerances are the parity suite's: rel 1e-5 on v, c, loss; rel 1e-4 on du, dW."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cs231_capsule_yolo_traffic_sign_detection_b200 as capsb   # noqa: E402
from conftest import rel_err   # noqa: E402
from oracle import routing_c as oc   # noqa: E402
from oracle import routing_np as onp   # noqa: E402
from test_routing_gpu import cuda_step   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
capsb._cabi.lib()
worst = {'v': 0.0, 'c': 0.0, 'loss': 0.0, 'du': 0.0, 'dW': 0.0}
bad = 0
for it in range(n):
    kind = it % 4
    if kind == 0:       # single class capsule
        C, D = 1, int(rng.integers(1, 9))
    elif kind == 1:     # FMA kernels
        C, D = int(rng.integers(2, 70)), int(rng.integers(1, 9))
    elif kind == 2:     # fused sweep range
        C, D = int(rng.integers(4, 65)), int(rng.integers(9, 17))
    else:               # padded tensor-core range
        C, D = int(rng.integers(2, 40)), int(rng.integers(17, 49))
    B = int(rng.choice([1, 2, 31, 32, 33, 100, 129, 257, int(rng.integers(1, 400))]))
    N = int(rng.choice([1, 3, 8, 33, 64, 100, int(rng.integers(1, 300))]))
    R = int(rng.integers(1, 6))
    u, W, y = onp.make_inputs(B, N, C, 8, D, seed=1000 + it)
    gext = (rng.standard_normal((B, C, D)) * 0.05).astype(np.float32) if it % 3 == 0 else None
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R,
                          grad_v_extra=None if gext is None else gext.astype(np.float64))
    got = cuda_step(capsb, u, W, y, R, grad_v_extra=gext)
    errs = {'v': rel_err(got['v'], ref['v']), 'c': rel_err(got['c'], ref['c']),
            'loss': abs(got['loss'] - ref['loss']) / max(1.0, abs(ref['loss'])),
            'du': rel_err(got['du'], ref['du']), 'dW': rel_err(got['dW'], ref['dW'])}
    ok = errs['v'] < 1e-5 and errs['c'] < 1e-5 and errs['loss'] < 1e-5 and errs['du'] < 1e-4 and errs['dW'] < 1e-4
    bad += not ok
    for k in worst:
        worst[k] = max(worst[k], errs[k])
    print('%s B=%-4d N=%-4d C=%-3d D=%-3d R=%d ext=%d  ' % ('ok  ' if ok else 'FAIL', B, N, C, D, R, gext is not None)
          + ' '.join('%s %.1e' % kv for kv in errs.items()), flush=True)
print('%d shapes, %d outside tolerance; worst ' % (n, bad) + ' '.join('%s %.1e' % kv for kv in worst.items()))
sys.exit(1 if bad else 0)
