#!/usr/bin/env python
"""Where the GEMM engine's time goes in one batch of 720p frames: every top-level ops.linear /
conv2d_nhwc / mask_logits call of a batched forward is recorded, grouped by shape signature, and one
representative per group is replayed in a CUDA graph.  Prints time x count, TFLOP/s and the
algorithmic HBM bytes (operand planes + outputs + residual) per group."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpvsg_b200 as pv
from openpvsg_b200 import configs, ops, synthetic as syn

H, W = 720, 1280
B = int(os.environ.get('PROBE_BATCH', '8'))
dev = torch.device('cuda:0')
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.to(dev)
meta = syn.frame_meta(H, W)
x = syn.synthetic_frame(0, H, W)[None].to(dev).expand(B, -1, -1, -1).contiguous()
calls, depth = [], [0]


def nbytes(v):
    if isinstance(v, ops.Split):
        return v.hi.numel() * 4
    if torch.is_tensor(v):
        return v.numel() * v.element_size()
    if isinstance(v, (tuple, list)):
        return sum(nbytes(u) for u in v)
    return 0


def sig(name, a, k):
    def s(v):
        if isinstance(v, ops.Split):
            return 'S' + str(tuple(v.shape))
        return str(tuple(v.shape)) if torch.is_tensor(v) else None
    parts = [name] + [t for t in (s(v) for v in a[:2]) if t]
    parts += [f'{kk}={vv if not torch.is_tensor(vv) else "T"}' for kk, vv in sorted(k.items())
              if kk in ('stride', 'out_mode', 'residual', 'add_input', 'act', 'want_logits', 'want_mask') and vv is not None]
    return ' '.join(parts)


def flops(name, a, k):
    if name == 'linear':
        return 2.0 * int(np.prod(a[0].shape[:-1])) * a[1].shape[0] * a[1].shape[1]
    if name == 'conv2d_nhwc':
        s = k.get('stride', 1)
        xx, w = a[0], a[1]
        return 2.0 * xx.shape[0] * (xx.shape[1] // s) * (xx.shape[2] // s) * w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3]
    e, f = a[0], a[1]
    return 2.0 * e.shape[0] * e.shape[1] * e.shape[2] * f.shape[1]


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **k):
        top = depth[0] == 0
        depth[0] += 1
        try:
            out = orig(*a, **k)
        finally:
            depth[0] -= 1
        if top:
            io = nbytes(a[0]) + nbytes(out) + nbytes(k.get('residual')) + nbytes(k.get('add_input'))
            calls.append((sig(name, a, k), orig, a, k, flops(name, a, k), io))
        return out
    setattr(ops, name, f)


for n in ('linear', 'conv2d_nhwc', 'mask_logits'):
    wrap(n)
with torch.no_grad():
    det.panoptic_head.simple_test_with_query(det.extract_feat(x), [[meta]] * B, upsample=False)
    calls.clear()
    det.panoptic_head.simple_test_with_query(det.extract_feat(x), [[meta]] * B, upsample=False)
torch.cuda.synchronize()

groups = {}
for c in calls:
    g = groups.setdefault(c[0], dict(n=0, call=c))
    g['n'] += 1
rows = []
REPS = 5
for s, g in groups.items():
    _, fn, a, k, fl, io = g['call']
    st = torch.cuda.Stream()
    with torch.cuda.stream(st), torch.no_grad():
        fn(*a, **k)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr), torch.no_grad():
        for _ in range(REPS):
            fn(*a, **k)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    rows.append((us * g['n'], g['n'], us, fl / us / 1e6, io / us / 1e3, s))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f'batch {B}: {len(calls)} calls, {len(rows)} shapes, total {tot:.0f} us = {tot / B:.0f} us/frame')
for t, n, us, tf, gbs, s in rows[:40]:
    print(f'{t:8.0f} us  x{n:<3d} {us:7.1f} us/call {tf:6.0f} TF {gbs:6.0f} GB/s  {s}')
if len(sys.argv) > 1:
    json.dump(dict(source=f'tools/gemm_calls.py, PROBE_BATCH={B}: every distinct dense call of one {B}-frame 720p forward replayed alone in '
                          'a CUDA graph (warm L2); TF = algorithmic 2MNK flops / time (split-bf16 ceiling 469 TF), GB/s = operand '
                          'planes + outputs + residual bytes / time (HBM peak 6539 GB/s)',
                   total_us=round(tot, 1), us_per_frame=round(tot / B, 1),
                   shapes=[dict(total_us=round(t, 1), calls=n, us_per_call=round(us, 1), tflops=round(tf, 1), gbps=round(gbs, 1), shape=s)
                           for t, n, us, tf, gbs, s in rows]), open(sys.argv[1], 'w'), indent=1)
