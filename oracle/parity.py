"""Tie-aware parity helpers for the GPU tests (test infrastructure).

north_star: outputs within 1e-3 fp32 of the reference CPU path, argmax mask ids bit-exact.  Two discrete decisions
sit on fp32 values and need a TIE rule to be comparable between two correct fp32 implementations:

* the decoder's attention mask is ``sigmoid(logit) < 0.5`` (mask2former_head.py:391): a logit within re-association
  noise of zero may land on either side, and self-attention then spreads the difference to every query;
* the panoptic id of a pixel is an arg-max over ``score * sigmoid(mask)`` plus a ``>= 0.5`` test
  (mask2former_fusion_head.py:128-150): near-equal candidates may swap.

The helpers make both explicit instead of allowing a blanket error budget: the oracle is re-run with the product's
mask decisions adopted ONLY where its own logit is within ``TIE_EPS`` of the threshold (every other flipped bit is
counted and must be zero), after which everything continuous must agree to 1e-3 and every differing pixel must be a
provable tie of the oracle's own scores.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import m2f as om

TOL = 1e-3
TIE_EPS = 1e-3          # attention-mask logits: |logit| below this may flip (the product's logits are within TOL)
PIX_EPS = 2e-3          # panoptic arg-max: score gap / mask-logit margin below which a pixel is a tie


def oracle_frame(sd, img, meta, gpu_masks, video=True, backbone=None, rescale=True):
    """Oracle forward of one frame with the product's attention-mask decisions adopted at ties.
    img [3,H,W] (CPU); gpu_masks: list over decoder layers of uint8 [Q,hw] / [1,Q,hw] arrays.
    Returns dict(result, cls, masks (full-resolution logits [Q,H,W]), query, tie_stats)."""
    masks = [torch.as_tensor(np.asarray(m)).reshape(1, *np.asarray(m).shape[-2:]) for m in gpu_masks]
    stats = []
    with torch.no_grad():
        if video:
            res, raw = om.vps_simple_test(sd, img[None, None], [[meta]], rescale=rescale, instance_on=False, return_raw=True,
                                          backbone=backbone, tie_masks=[masks], tie_eps=TIE_EPS, tie_stats=stats)
            res, cls, mp, q = res[0][0], raw['cls'][0], raw['masks'][0, 0], raw['embds'][0]
        else:
            res, raw = om.ips_simple_test(sd, img[None], [meta], rescale=rescale, instance_on=False, return_raw=True,
                                          backbone=backbone, tie_masks=masks, tie_eps=TIE_EPS, tie_stats=stats)
            res, cls, mp, q = res[0], raw['cls'][0], raw['masks'][0], raw['embds'][0, :, 0]
    return dict(result=res, cls=cls, masks=mp, query=q, tie_stats=stats)


def assert_no_real_flips(tie_stats, what=''):
    bad = [s for s in tie_stats if s['flipped_non_ties']]
    assert not bad, f'{what}: attention-mask bits differ away from the threshold: {bad}'
    return sum(s['flipped_ties'] for s in tie_stats)


def tie_pixels(cls, masks, meta, num_classes=126, object_mask_thr=0.8, rescale=True):
    """Pixels whose panoptic id is not decided by a clear margin in the ORACLE's own scores: the top-2 candidates of
    ``score * sigmoid(mask)`` are within PIX_EPS (and at least one of them would be written), or the winner's mask
    logit is within PIX_EPS of 0 (the >= 0.5 test of filter_low_score).  cls [Q,NC+1], masks [Q,Hp,Wp] full-resolution
    logits.  Returns (tie mask, distance of every pixel's decision from a tie)."""
    ih, iw = meta['img_shape'][:2]
    mp = masks[:, :ih, :iw]
    if rescale:
        mp = F.interpolate(mp[:, None], size=tuple(meta['ori_shape'][:2]), mode='bilinear', align_corners=False)[:, 0]
    scores, labels = F.softmax(cls, dim=-1).max(-1)
    keep = labels.ne(num_classes) & (scores > object_mask_thr)
    if int(keep.sum()) == 0:
        return torch.zeros(mp.shape[-2:], dtype=torch.bool).numpy(), torch.full(mp.shape[-2:], 1e9).numpy()
    logits = mp[keep]
    prob = scores[keep].view(-1, 1, 1) * logits.sigmoid()
    if prob.shape[0] > 1:
        top = prob.topk(2, dim=0)
        gap = top.values[0] - top.values[1]
        win, second = top.indices[0], top.indices[1]
        l2 = logits.gather(0, second[None])[0]
    else:
        gap = torch.full(prob.shape[1:], 1.0)
        win = torch.zeros(prob.shape[1:], dtype=torch.long)
        l2 = torch.full(prob.shape[1:], -1e9)
    l1 = logits.gather(0, win[None])[0]
    # a swap of the two best candidates only shows if one of them passes the >= 0.5 test (filter_low_score)
    visible = (l1 > -PIX_EPS) | (l2 > -PIX_EPS)
    closeness = torch.minimum(torch.where(visible, gap, torch.full_like(gap, 1e9)), l1.abs())   # distance from a tie
    return (closeness < PIX_EPS).numpy(), closeness.numpy()


def assert_pan_tie_aware(pan, ref, what=''):
    """pan: product's int32 map; ref: dict from ``oracle_frame``.  Every differing pixel must be a tie pixel of the
    oracle.  Returns (mismatching pixels, tie pixels)."""
    ref_pan = np.asarray(ref['result']['pan_results'])
    assert pan.shape == ref_pan.shape and pan.dtype == np.int32, (what, pan.shape, pan.dtype)
    diff = pan != ref_pan
    n = int(diff.sum())
    ties = ref.get('_tie_pixels')
    if isinstance(ties, tuple):
        ties, closeness = ties
        ref['_max_tie_distance_of_mismatch'] = float(closeness[diff].max()) if n else 0.0
    if n:
        assert ties is not None, f'{what}: {n} panoptic ids differ'
        assert not (diff & ~ties).any(), f'{what}: {int((diff & ~ties).sum())} panoptic ids differ away from any tie'
    return n, (int(ties.sum()) if ties is not None else None)


def check_frame(res, sd, img, meta, gpu_masks, what='', video=True, backbone=None, gpu_cls=None, feat_tol=TOL):
    """Full tie-aware parity of one product result dict against the oracle.  Returns a stats dict."""
    ref = oracle_frame(sd, img, meta, gpu_masks, video=video, backbone=backbone)
    flips = assert_no_real_flips(ref['tie_stats'], what)
    ref['_tie_pixels'] = tie_pixels(ref['cls'], ref['masks'], meta)
    n_diff, n_tie = assert_pan_tie_aware(np.asarray(res['pan_results']), ref, what)
    ref_q = ref['result']['query_feats']
    assert sorted(res['query_feats']) == sorted(ref_q), (what, sorted(res['query_feats']), sorted(ref_q))
    err = 0.0
    for k in ref_q:
        a = torch.as_tensor(np.asarray(torch.as_tensor(res['query_feats'][k][0]).cpu())).float().flatten()
        b = torch.as_tensor(np.asarray(ref_q[k][0])).float().flatten()
        err = max(err, (a - b).abs().max().item())
    assert err <= feat_tol, f'{what}: query features differ by {err:.3e}'
    cls_err = None
    if gpu_cls is not None:
        cls_err = (torch.as_tensor(np.asarray(gpu_cls)).float().reshape(ref['cls'].shape) - ref['cls']).abs().max().item()
        assert cls_err <= TOL, f'{what}: class logits differ by {cls_err:.3e}'
    return dict(what=what, adopted_tie_bits=flips, near_threshold_bits=sum(s['near_threshold'] for s in ref['tie_stats']),
                max_abs_logit_of_adopted_bits=max([s['max_abs_logit_of_flips'] for s in ref['tie_stats']] + [0.0]),
                pan_mismatch_pixels=n_diff, pan_tie_pixels=n_tie,
                max_tie_distance_of_mismatching_pixels=ref.get('_max_tie_distance_of_mismatch', 0.0),
                pixels=int(np.asarray(res['pan_results']).size), query_feat_max_err=err, cls_max_err=cls_err)
