#!/usr/bin/env python
"""Per-kernel timing at the BASELINE 720p shapes (CUDA events, L2 flushed between runs).

Prints one line per kernel: microseconds, achieved GB/s or TFLOP/s on ALGORITHMIC bytes /
flops.  Used to fill profiles/ and to steer optimisation; the
official numbers come from bench.py.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpvsg_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--match', default='', help='only shapes whose name contains this')
    ap.add_argument('--batch', type=int, default=1, help='frames per launch (bench.py runs 8)')
    args = ap.parse_args()
    dev = 'cuda'
    NB = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    rows = []

    def rec(name, us, gbytes=None, gflop=None):
        r = dict(kernel=name, us=round(us, 2))
        if gbytes is not None:
            r['GBps'] = round(gbytes / (us * 1e-6), 1)
        if gflop is not None:
            r['TFLOPs'] = round(gflop / (us * 1e-6) / 1e3, 2)
        rows.append(r)
        print(json.dumps(r), flush=True)

    def want(n):
        return not args.only or args.only in n

    shapes = [(23, 40), (46, 80), (92, 160)]
    N = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(0)
    if want('msda'):
        value = torch.randn(NB, N, 256, device=dev, generator=g)
        proj = torch.cat([torch.randn(NB, N, 192, device=dev, generator=g) * 2, torch.randn(NB, N, 96, device=dev, generator=g)], -1)
        refs = []
        for h, w in shapes:
            ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
            refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
        ref = torch.cat(refs).to(dev)
        us = timeit(lambda: ops.msda_fused_forward(value, shapes, proj, ref), args.iters, flush=flush)
        rec(f'msda_fused_720p_b{NB}', us, gbytes=NB * N * (1024 + 1152 + 1024) / 1e9)
        loc = torch.rand(NB, N, 8, 3, 4, 2, device=dev, generator=g)
        aw = torch.softmax(torch.randn(NB, N, 8, 12, device=dev, generator=g), -1).view(NB, N, 8, 3, 4)
        v4 = value.view(NB, N, 8, 32)
        us = timeit(lambda: ops.msda_forward(v4, shapes, loc, aw), args.iters, flush=flush)
        rec(f'msda_unfused_720p_b{NB}', us, gbytes=NB * N * (1024 + 768 + 384 + 1024) / 1e9)
    if want('gemm'):
        for (M, Nn, K) in [(19320, 256, 256), (19320, 1024, 256), (19320, 256, 1024), (58880, 256, 256),
                           (100, 256, 256), (100, 2048, 256), (100, 256, 2048), (14720, 256, 256)]:
            M = M * NB
            if args.match and args.match not in f'linear_{M}x{Nn}x{K}':
                continue
            x = torch.randn(M, K, device=dev, generator=g)
            w = torch.randn(Nn, K, device=dev, generator=g)
            b = torch.randn(Nn, device=dev, generator=g)
            us = timeit(lambda: ops.linear(x, w, b), args.iters, flush=flush)
            rec(f'linear_{M}x{Nn}x{K}', us, gbytes=4 * (M * K + Nn * K + M * Nn) / 1e9, gflop=2 * M * Nn * K / 1e9)
    if want('mask'):
        embed = torch.randn(1, 100, 256, device=dev, generator=g)
        feat = torch.randn(1, 58880, 256, device=dev, generator=g)
        us = timeit(lambda: ops.mask_logits(embed, feat, True, False), args.iters, flush=flush)
        rec('mask_logits_full_720p', us, gbytes=((256 + 100) * 58880 * 4 + 100 * 256 * 4) / 1e9, gflop=2 * 100 * 256 * 58880 / 1e9)
        pooled = torch.randn(1, 14720, 256, device=dev, generator=g)
        us = timeit(lambda: ops.mask_logits(embed, pooled, False, True), args.iters, flush=flush)
        rec('mask_logits_attnmask_s8', us, gflop=2 * 100 * 256 * 14720 / 1e9)
    if want('conv'):
        for (cin, cout, k, s, hw) in [(3, 64, 7, 2, (736, 1280)), (64, 64, 3, 1, (184, 320)), (128, 128, 3, 1, (92, 160)),
                                      (256, 256, 3, 1, (46, 80)), (512, 512, 3, 1, (23, 40)), (256, 256, 3, 1, (184, 320)),
                                      (64, 256, 1, 1, (184, 320)), (1024, 256, 1, 1, (46, 80)), (512, 2048, 1, 1, (23, 40))]:
            if args.match and args.match not in f'conv{k}x{k}s{s}_{cin}->{cout}@{hw[0]}x{hw[1]}_b{NB}':
                continue
            x = torch.randn(NB, hw[0], hw[1], cin, device=dev, generator=g)
            w = torch.randn(cout, k, k, cin, device=dev, generator=g)
            b = torch.randn(cout, device=dev, generator=g)
            pad = k // 2
            oh, ow = (hw[0] + 2 * pad - k) // s + 1, (hw[1] + 2 * pad - k) // s + 1
            us = timeit(lambda: ops.conv2d_nhwc(x, w, b, stride=s, pad=pad, act=1), args.iters, flush=flush)
            rec(f'conv{k}x{k}s{s}_{cin}->{cout}@{hw[0]}x{hw[1]}_b{NB}', us, gflop=2 * NB * oh * ow * cout * cin * k * k / 1e9)
    if want('attn'):
        for (B, H, Lq, Lk, E) in [(1, 8, 100, 920, 256), (1, 8, 100, 3680, 256), (1, 8, 100, 14720, 256),
                                  (1, 8, 100, 100, 256), (128, 8, 200, 200, 256), (100, 4, 128, 128, 512)]:
            if Lq == 100:
                B = B * NB
            if args.match and args.match not in f'attention_B{B}_H{H}_{Lq}x{Lk}_E{E}':
                continue
            q = torch.randn(B, Lq, E, device=dev, generator=g)
            k = torch.randn(B, Lk, E, device=dev, generator=g)
            v = torch.randn(B, Lk, E, device=dev, generator=g)
            mask = (torch.rand(B, Lq, Lk, device=dev, generator=g) < 0.5).to(torch.uint8) if Lk > 200 else None
            us = timeit(lambda: ops.attention(q, k, v, H, mask=mask), args.iters, flush=flush)
            rec(f'attention_B{B}_H{H}_{Lq}x{Lk}_E{E}', us, gflop=4 * B * Lq * Lk * E / 1e9)
    if want('norm'):
        x = torch.randn(N, 256, device=dev, generator=g)
        gm = torch.randn(256, device=dev, generator=g)
        us = timeit(lambda: ops.layernorm(x, gm, gm), args.iters, flush=flush)
        rec('layernorm_19320x256', us, gbytes=2 * N * 256 * 4 / 1e9)
        x = torch.randn(1, 184, 320, 256, device=dev, generator=g)
        us = timeit(lambda: ops.groupnorm_nhwc(x, gm, gm, 32, act=1), args.iters, flush=flush)
        rec('groupnorm_184x320x256', us, gbytes=3 * 58880 * 256 * 4 / 1e9)
    if want('pan'):
        cls = torch.randn(100, 127, device=dev, generator=g)
        cls[::3, 5] += 15
        mp = torch.randn(100, 184, 320, device=dev, generator=g)
        us = timeit(lambda: ops.panoptic_fuse(cls, mp, (736, 1280), (720, 1280), (720, 1280), 115, 126), args.iters, flush=flush)
        rec('panoptic_fuse_720p_34kept', us, gbytes=(100 * 184 * 320 * 4 + 720 * 1280 * 4) / 1e9)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'kbench.json'), 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
