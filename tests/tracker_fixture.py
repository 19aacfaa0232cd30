"""Synthetic detections for the IPS tracker path (shared by tests/golden/make_golden_tracker.py and the tests)."""
import copy

import numpy as np
import torch

# configs/unitrack/imagenet_resnet50_s3_womotion_timecycle.py:5-43 (the keys the tracker reads)
TRACKER_CFG = dict(common=dict(device='cpu', down_factor=8),
                   mots=dict(track_buffer=300, conf_thres=0.5, max_mask_area=300, dup_iou_thres=0.15, confirm_iou_thres=0.7,
                             feat_size=[4, 10], use_kalman=True, asso_with_motion=False, motion_lambda=1, motion_gated=False))
D = 32          # appearance feature width


def tracker_cfg():
    return copy.deepcopy(TRACKER_CFG)


def embedding_sets(seed=1, ntrk=5, ndet=6):
    """Mask-pooled embeddings [1, D, n_pix] with ragged pixel counts: detections are noisy copies of some tracks."""
    g = torch.Generator().manual_seed(seed)
    protos = torch.randn(max(ntrk, ndet), D, generator=g)
    trk = [protos[i][None, :, None] + 0.3 * torch.randn(1, D, int(n), generator=g)
           for i, n in zip(range(ntrk), torch.randint(3, 12, (ntrk,), generator=g))]
    det = [protos[(i + 1) % max(ntrk, ndet)][None, :, None] + 0.3 * torch.randn(1, D, int(n), generator=g)
           for i, n in zip(range(ndet), torch.randint(3, 12, (ndet,), generator=g))]
    return trk, det


def boxes(seed, n):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, 60, (n, 2))
    wh = rng.uniform(10, 40, (n, 2))
    return np.concatenate([xy, xy + wh], 1)


def clip(seed=7, frames=14, objects=6):
    """Per frame: (obs = [(tlwh, feat [1,D,n_pix], cls)], query_feats = [dict(cls_id, query_feat)]).  Objects move
    linearly; some vanish for a few frames (lost -> re-found), two appear late, one frame is empty, two objects of
    different classes share a location (the class gate must keep them apart)."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    protos = torch.randn(objects, D, generator=g)
    pos = rng.uniform(10, 120, (objects, 2))
    pos[1] = pos[0] + 2.0                                  # overlapping pair ...
    vel = rng.uniform(-4, 4, (objects, 2))
    vel[1] = vel[0]
    size = rng.uniform(16, 40, (objects, 2))
    size[1] = size[0]
    cls = np.array([3, 7, 3, 12, 7, 3])[:objects]          # ... of different classes
    born = np.array([0, 0, 0, 0, 5, 8])[:objects]
    gone = {2: range(4, 7), 3: range(9, 11)}               # object -> frames where it is not detected
    out = []
    for f in range(frames):
        obs, qf = [], []
        if f != 6:                                          # frame 6 has no detections at all
            for o in rng.permutation(objects):
                if f < born[o] or f in gone.get(int(o), ()):
                    continue
                p = pos[o] + vel[o] * f + rng.normal(0, 0.3, 2)
                s = size[o] * (1 + rng.normal(0, 0.01, 2))
                n_pix = int(rng.integers(4, 10))
                feat = protos[o][None, :, None] + 0.25 * torch.randn(1, D, n_pix, generator=g)
                obs.append((np.array([p[0], p[1], s[0], s[1]]), feat, int(cls[o])))
                qf.append(dict(cls_id=int(cls[o]) + 1000 * (len(qf) + 1), query_feat=np.full(4, float(o * 100 + f), np.float32)))
        out.append((obs, qf))
    return out
