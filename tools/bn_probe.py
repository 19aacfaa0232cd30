#!/usr/bin/env python
"""tcgen05 engine alone (operands already split): time per call for the frame's GEMM / conv shapes.
Run once per tile width:  PVSG_TC_BN=128 python tools/bn_probe.py ; PVSG_TC_BN=256 python tools/bn_probe.py"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpvsg_b200 import ops


def graph_time(fn, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


B = int(os.environ.get('PROBE_BATCH', '8'))
bn = os.environ.get('PVSG_TC_BN', 'auto')
for (M, N, K) in [(19320, 256, 256), (19320, 1024, 256), (19320, 256, 1024), (14720, 256, 256), (58880, 256, 64),
                  (58880, 64, 256), (3680, 2048, 512), (920, 2048, 1024)]:
    M *= B
    x = ops.Split(*ops.split_bf16(torch.randn(M, K, device='cuda')))
    w = torch.randn(N, K, device='cuda'); b = torch.randn(N, device='cuda')
    out = torch.empty(M, N, device='cuda')
    us = graph_time(lambda: ops.linear(x, w, b, out=out), reps=10)
    print(json.dumps(dict(bn=bn, kind='linear', M=M, N=N, K=K, us=round(us, 1), TF=round(2 * M * N * K / us / 1e6, 1))),
          flush=True)
for (H, W, Cin, Cout) in [(184, 320, 64, 64), (92, 160, 128, 128), (46, 80, 256, 256), (23, 40, 512, 512),
                          (92, 160, 256, 256)]:
    x = torch.randn(B, H, W, Cin, device='cuda')
    xs = ops.Split(*ops.split_bf16(x))
    w = torch.randn(Cout, 3, 3, Cin, device='cuda')
    us = graph_time(lambda: ops.conv2d_nhwc(xs, w, None, stride=1, pad=1, act=ops.ACT_RELU), reps=10)
    fl = 2 * B * H * W * Cout * Cin * 9
    print(json.dumps(dict(bn=bn, kind='conv3x3', B=B, H=H, W=W, Cin=Cin, Cout=Cout, us=round(us, 1),
                          TF=round(fl / us / 1e6, 1))), flush=True)
