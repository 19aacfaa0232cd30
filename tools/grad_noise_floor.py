"""Noise floor of the training gradients: the CPU oracle against ITSELF in fp64 vs fp32 (same targets, points and
attention-mask decisions).  ReLU kinks and max-pooling ties make the backbone gradients of any fp32 implementation differ
from the exact ones by up to ~7e-3 of the tensor maximum; the head tensors agree to 3e-6.  tests/test_training_slice.py
uses 2e-2 (backbone) / 2e-3 (head) accordingly.  CPU only: python tools/grad_noise_floor.py"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import m2f as om, losses as ol
from openpvsg_b200 import synthetic as syn
torch.set_num_threads(8)
sd = syn.mask2former_state_dict(seed=3)
H, W, T = 96, 160, 2
frames = torch.stack([syn.synthetic_frame(5 + t, H, W) for t in range(T)])
import torch.nn.functional as F
def _ps(input, points, align_corners=False):
    return F.grid_sample(input.to(points.dtype), (2.0 * points - 1.0).unsqueeze(2), align_corners=align_corners).squeeze(3)
ol.point_sample = _ps
def run(dtype, ref_masks=None, gt=None):
    torch.set_default_dtype(dtype)
    osd = {k: v.clone().to(dtype) if v.is_floating_point() else v.clone() for k, v in sd.items()}
    keys = [k for k in osd if (k.startswith('panoptic_head.') or 'conv' in k or 'downsample.0' in k) and osd[k].is_floating_point() and 'bn' not in k and 'downsample.1' not in k and 'num_batches' not in k]
    for k in keys: osd[k].requires_grad_(True)
    feats = om.resnet50(osd, frames.to(dtype))
    ocls, omask, _, ex = om.head_forward(osd, feats, video=True, num_frames=T, return_all=True, tie_masks=ref_masks)
    if gt is None:
        mp = omask[-1].detach()
        picks = [7, 31, 64]
        gtm = torch.stack([(mp[0, :, q] > mp[0, :, q].median()).float() for q in picks])
        gt = (gtm, torch.tensor([5, 120, 60]))
    g = torch.Generator().manual_seed(9)
    K = 400
    tot = 0
    for c, m in zip(ocls, omask):
        a = torch.rand(1, K, 2, generator=g, dtype=torch.float32).to(dtype); l = torch.rand(3, K, 2, generator=g, dtype=torch.float32).to(dtype)
        wc, wm, wd, _, pos = ol.loss_single(c, m, [gt[1]], [gt[0].to(dtype)], a, lambda n: l[:n])
        tot = tot + wc + wm + wd
    tot.backward()
    return {k: osd[k].grad for k in keys}, [m.clone() for m in ex['raw_attn_masks']], gt, float(tot)
g64, masks, gt, l64 = run(torch.float64)
g32, _, _, l32 = run(torch.float32, ref_masks=masks, gt=gt)
print('loss', l64, l32)
rel = {k: float((g32[k].double() - g64[k]).abs().max() / max(float(g64[k].abs().max()), 1e-3)) for k in g64 if g64[k] is not None}
top = sorted(rel.items(), key=lambda kv: -kv[1])[:12]
for k, v in top: print(f'{v:.2e} {k}')
import statistics
print('median', statistics.median(rel.values()), 'backbone max', max(v for k, v in rel.items() if k.startswith('backbone')), 'head max', max(v for k, v in rel.items() if k.startswith('panoptic')))
