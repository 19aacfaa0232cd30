#!/usr/bin/env python
"""Times the two fused MSDeformAttn kernels at the bench shape (720p pyramid, B frames per launch).
  python tools/msda_time.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpvsg_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shapes = [(23, 40), (46, 80), (92, 160)]
n = sum(h * w for h, w in shapes)
g = torch.Generator().manual_seed(0)
dev = torch.device('cuda')
value = torch.randn(B, n, 256, generator=g).to(dev)
proj = torch.cat([torch.randn(B, n, 192, generator=g) * 1.7, torch.randn(B, n, 96, generator=g)], -1).to(dev)
refs = []
for h, w in shapes:
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
ref = torch.cat(refs, 0).to(dev)
alg_bytes = B * n * (1024 + 1152 + 1024)
for impl in ('tile', 'group'):
    if impl == 'group':
        os.environ['PVSG_MSDA_IMPL'] = 'group'
    else:
        os.environ.pop('PVSG_MSDA_IMPL', None)
    for _ in range(3):
        out = ops.msda_fused_forward(value, shapes, proj, ref, out_mode='split')
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.msda_fused_forward(value, shapes, proj, ref, out_mode='split')
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f'{impl}: {1e3 * ms:.1f} us per launch ({B} frames) = {1e3 * ms / B:.2f} us per frame-layer, '
          f'{alg_bytes / ms / 1e6:.0f} GB/s algorithmic = {alg_bytes / ms / 1e6 / 6538.6:.3f} of HBM peak')
