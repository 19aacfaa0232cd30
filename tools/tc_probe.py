#!/usr/bin/env python
"""GPU-side cost per launch of small GEMMs: N back-to-back calls captured in one CUDA graph."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpvsg_b200 import ops


def graph_time(fn, reps=50):
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


for eng in ('tc', 'simt'):
    ops.set_engine(eng)
    for (M, N, K) in [(100, 256, 256), (100, 2048, 256), (100, 256, 2048), (100, 512, 256), (920, 256, 256),
                      (19320, 256, 256), (19320, 1024, 256), (19320, 256, 1024), (58880, 256, 64), (58880, 64, 256)]:
        x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); b = torch.randn(N, device='cuda')
        us = graph_time(lambda: ops.linear(x, w, b))
        print(json.dumps(dict(engine=eng, M=M, N=N, K=K, us_per_call=round(us, 2),
                              TF=round(2 * M * N * K / us / 1e6, 1))), flush=True)
x = torch.randn(19320, 256, device='cuda')
print('split 19320x256 us', graph_time(lambda: ops.split_bf16(x)))
x = torch.randn(100, 256, device='cuda')
print('split 100x256 us', graph_time(lambda: ops.split_bf16(x)))
g = torch.randn(256, device='cuda')
print('layernorm 100x256 us', graph_time(lambda: ops.layernorm(x, g, g)))
