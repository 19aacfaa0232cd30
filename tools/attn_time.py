#!/usr/bin/env python
"""Times the tcgen05 (t5) and mma.sync attention kernels on the three shapes of the path:
decoder cross-attention (20 frames x 8 heads, 100 x Lk x 32, masked), relation ObjectEncoder (128 x 8, 200 x 200 x 32),
TemporalTransformer (100 x 4, 128 x 128 x 128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpvsg_b200 import lib, ops  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)


def planes(x):
    return ops.Split(*ops.split_bf16(x.contiguous()))


def bench(name, B, H, Lq, Lk, D, masked):
    E = H * D
    q = torch.randn(B, Lq, E, generator=g).to(dev)
    k, v = planes(torch.randn(B, Lk, E, generator=g).to(dev)), planes(torch.randn(B, Lk, E, generator=g).to(dev))
    mask = row_open = None
    if masked:
        mask = (torch.rand(B, Lq, Lk, generator=g) < 0.7).to(torch.uint8).to(dev)
        row_open = (mask == 0).sum(-1).to(torch.int32)
    res = {}
    for impl in ('t5', 'mma'):
        lib.ATTN_IMPL[0] = impl
        for _ in range(3):
            out = ops.attention(q, k, v, H, mask=mask, row_open=row_open)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.attention(q, k, v, H, mask=mask, row_open=row_open)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[impl] = (sorted(ts)[len(ts) // 2], out)
    err = (res['t5'][1] - res['mma'][1]).abs().max().item()
    fl = 4.0 * B * H * Lq * Lk * D
    print(f'{name}: t5 {1e3 * res["t5"][0]:.1f} us ({fl / res["t5"][0] / 1e9:.1f} TF)   mma {1e3 * res["mma"][0]:.1f} us '
          f'({fl / res["mma"][0] / 1e9:.1f} TF)   max |t5 - mma| {err:.2e}')


bench('decoder cross-attn s8 ', 20, 8, 100, 14720, 32, True)
bench('decoder cross-attn s16', 20, 8, 100, 3680, 32, True)
bench('decoder cross-attn s32', 20, 8, 100, 920, 32, True)
bench('relation ObjectEncoder', 128, 8, 200, 200, 32, False)
bench('TemporalTransformer   ', 100, 4, 128, 128, 128, False)
