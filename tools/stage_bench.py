#!/usr/bin/env python
"""Stage-level GPU time of one 720p frame: nested CUDA graphs (backbone | + pixel decoder |
+ decoder | + post-processing), each replayed and timed with CUDA events."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpvsg_b200 as pv
from openpvsg_b200 import configs, synthetic as syn, lib

H, W = 720, 1280
dev = torch.device('cuda:0')
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.to(dev)
meta = syn.frame_meta(H, W)
x = syn.synthetic_frame(0, H, W)[None].to(dev)
head, fh = det.panoptic_head, det.panoptic_fusion_head
in_hw, img_hw = (736, 1280), (720, 1280)


def s_backbone():
    return det.extract_feat(x)


def s_pixdec():
    return head.pixel_decoder(det.extract_feat(x))


def s_head():
    return head.simple_test_with_query(det.extract_feat(x), [[meta]], upsample=False)


def s_pan():
    cls, m, q = s_head()
    return fh._panoptic(cls[0], m[0, 0], in_hw, img_hw, img_hw)


def s_full():
    cls, m, q = s_head()
    a = fh._panoptic(cls[0], m[0, 0], in_hw, img_hw, img_hw)
    b = fh._instance_device(cls[0], m[0, 0], in_hw, img_hw, img_hw, True)
    return a, b


def graph_ms(fn, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.synchronize()
    n0 = lib.launch_count[0]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    n = lib.launch_count[0] - n0
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], n


prev, out = 0.0, {}
for name, fn in (('backbone', s_backbone), ('pixel_decoder', s_pixdec), ('decoder', s_head), ('panoptic', s_pan),
                 ('instance', s_full)):
    ms, n = graph_ms(fn)
    out[name] = dict(ms=round(ms - prev, 3), cumulative_ms=round(ms, 3), launches_cumulative=n)
    prev = ms
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'stage_bench.json'), 'w'), indent=1)
