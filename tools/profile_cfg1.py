"""Kernel-time breakdown of the configs[0] CapsNet train step (batch 16) with the drop-in layer: torch.profiler over a few
eager steps, CUDA kernel self times summed per kernel name.   python tools/profile_cfg1.py [batch]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
x, y = workloads.synth_cfg1(B, seed=0)
step = workloads.make_cfg1_step(dev, dropin=True)[0]
for _ in range(5):
    step(x, y)
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step(x, y)
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / N, e.count / N) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('device time per step: %.1f us in %.0f kernels / copies' % (tot, sum(r[2] for r in rows)))
for k, t, n in rows[:40]:
    print('%8.1f us  %5.1f x  %s' % (t, n, k[:150]))
