#!/usr/bin/env python
"""Per-call device time of every ops.* call of one 720p frame (GPU parked behind a spin kernel so
that event intervals exclude host launch latency).  Prints the calls sorted by time."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpvsg_b200 as pv
from openpvsg_b200 import configs, ops, synthetic as syn

H, W = 720, 1280
dev = torch.device('cuda:0')
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.to(dev)
meta = syn.frame_meta(H, W)
x = syn.synthetic_frame(0, H, W)[None].to(dev)
events = []


def shp(a):
    if isinstance(a, ops.Split):
        return 'S' + str(tuple(a.shape))
    if torch.is_tensor(a):
        return str(tuple(a.shape))
    return None


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(*a, **k)
        e1.record()
        desc = name + ' ' + ' '.join(s for s in (shp(v) for v in a[:2]) if s) + \
            ''.join(f' {kk}={vv}' for kk, vv in k.items() if kk in ('stride', 'out_mode'))
        events.append((desc, e0, e1))
        return out
    setattr(ops, name, f)


for n in ('linear', 'conv2d_nhwc', 'mask_logits', 'msda_fused_forward', 'attention', 'layernorm', 'groupnorm_nhwc',
          'bilinear_resize_nhwc', 'maxpool3x3s2_nhwc', 'add_rowvec', 'split_bf16', 'panoptic_fuse', 'instance_masks'):
    wrap(n)


def frame():
    cls, mlr, _ = det.panoptic_head.simple_test_with_query(det.extract_feat(x), [[meta]], upsample=False)
    fh = det.panoptic_fusion_head
    fh._panoptic(cls[0], mlr[0, 0], (736, 1280), (H, W), (H, W))
    fh._instance_device(cls[0], mlr[0, 0], (736, 1280), (H, W), (H, W), True)


frame(); frame()
torch.cuda.synchronize()
events.clear()
torch.cuda._sleep(int(0.25 * 1.9e9))
frame()
torch.cuda.synchronize()
rows = [(d, e0.elapsed_time(e1) * 1e3) for d, e0, e1 in events]
# nested calls (linear -> split_bf16) are reported separately; outer time includes inner
tot = sum(t for d, t in rows if not d.startswith('split_bf16'))
print('calls', len(rows), 'sum_us(excl nested split)', round(tot, 1))
agg = {}
for d, t in rows:
    a = agg.setdefault(d, [0, 0.0])
    a[0] += 1
    a[1] += t
for d, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{t:9.1f} us  x{n:<3d} {d}')
