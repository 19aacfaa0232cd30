import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openpvsg_b200 import ops
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, H, W, C) in ((32, 96, 120, 64), (32, 96, 120, 256), (32, 48, 60, 128), (32, 24, 30, 1024)):
    x = torch.randn(B, H, W, C, device='cuda')
    T = B * H * W
    Tp = (T + 63) // 64 * 64
    hi, lo = ops.alloc_planes(C, T, Tp, x.device)
    us = t(lambda: ops.transpose_split(x.view(T, C), Tp, hi=hi, lo=lo))
    gb = T * C * 8 / 1e9
    print(f'contig [{T},{C}]: {us:.1f} us  {gb / us * 1e6:.0f} GB/s')
    xp = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))
    us = t(lambda: ops.transpose_split(xp[:, 1:1 + H, 2:2 + W], Tp, hi=hi, lo=lo))
    print(f'  tap view: {us:.1f} us  {gb / us * 1e6:.0f} GB/s')
    us = t(lambda: ops.split_bf16(x))
    print(f'  split_bf16 (no transpose): {us:.1f} us  {gb / us * 1e6:.0f} GB/s')
