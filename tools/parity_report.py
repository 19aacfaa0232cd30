#!/usr/bin/env python
"""Per-stage parity table: B200 backend vs CPU oracle on one synthetic frame.

Shows where end-to-end differences come from: fp32 re-association noise per stage, and the
discrete attention-mask sign flips (mask2former_head.py:391) that noise can trigger when a
down-sampled mask logit is within rounding distance of zero.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpvsg_b200 as pv  # noqa: E402
from openpvsg_b200 import configs, synthetic as syn  # noqa: E402
from oracle import m2f as om  # noqa: E402


def err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b).abs()
    return dict(max=float(d.max()), mean=float(d.mean()), ref_absmax=float(b.abs().max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--seed', type=int, default=17)
    ap.add_argument('--out', default='gpurun_out/parity_report.json')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    sd = syn.mask2former_state_dict(seed=3)
    det = pv.build_detector(configs.mask2former_r50(True))
    det.load_state_dict(sd)
    det.cuda()
    img = syn.synthetic_frame(args.seed, args.height, args.width)[None]
    meta = syn.frame_meta(args.height, args.width)
    rep = {}
    with torch.no_grad():
        rf = om.resnet50(sd, img)
        gf = det.extract_feat(img.cuda())
        for n, a, b in zip(('C2', 'C3', 'C4', 'C5'), gf, rf):
            rep[n] = err(a, b)
        rmf, rmem, rinter = om.pixel_decoder(sd, rf, return_intermediate=True)
        gmf, gmem = det.panoptic_head.pixel_decoder([f.cuda() for f in rf])
        rep['mask_feature(oracle feats)'] = err(gmf, rmf)
        for n, a, b in zip(('m32', 'm16', 'm8'), gmem, rmem):
            rep[n + '(oracle feats)'] = err(a, b)
        rc, rm, rq, ex = om.head_forward(sd, rf, video=True, num_frames=1, return_all=True)
        gc, gm, gq = det.panoptic_head.forward([f.cuda() for f in rf], [[meta]], return_query=True)
        shapes = [m.shape[-2:] for m in ex['memories']]
        for i in range(10):
            rep[f'cls[{i}]'] = err(gc[i], rc[i])
            rep[f'mask[{i}]'] = err(gm[i], rm[i])
            if i < 9:
                tgt = shapes[i % 3] if i > 0 else shapes[0]
                tgt = shapes[(i) % 3] if i == 0 else shapes[i % 3]
                # attention mask used by layer i comes from prediction i, target level i % 3
                dr = F.interpolate(rm[i].flatten(0, 1), tuple(shapes[i % 3]), mode='bilinear', align_corners=False)
                dg = F.interpolate(gm[i].flatten(0, 1).cpu(), tuple(shapes[i % 3]), mode='bilinear', align_corners=False)
                rep[f'attn_mask_flips[{i}]'] = dict(flips=int(((dr < 0) != (dg < 0)).sum()), total=dr.numel(),
                                                    near_zero=int((dr.abs() < 1e-4).sum()))
        rep['query'] = err(gq, rq)
    for k, v in rep.items():
        print(f'{k:32s} {json.dumps(v)}')
    os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, args.out), 'w'), indent=1)


if __name__ == '__main__':
    main()
