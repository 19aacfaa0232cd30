#!/usr/bin/env python
"""Golden vectors for the IPS tracker path from the REFERENCE's own code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_tracker.py
Writes tests/golden/tracker.json.

What runs is the reference, unmodified: ``models/unitrack/multitracker.py::AssociationTracker.update`` (the whole
association state machine), ``core/association/matching.py`` (``reconsdot_distance``, ``linear_assignment``,
``iou_distance``, ``fuse_motion``), ``core/motion/kalman_filter.py``, ``basetrack.py`` (``STrack`` and the list
helpers), ``data/query_feat_tracklet.py``.  Three things are absent from this container and are stubbed:

* ``lap`` (pinned nowhere in the reference; ``lap.lapjv(cost, extend_cost=True, cost_limit=thresh)`` of lap 0.4): restated
  from its published algorithm -- the rectangular problem is extended to a square one of size n + m whose extra
  entries cost ``cost_limit / 2`` (0 in the extra-extra block) and solved exactly (here with scipy's
  ``linear_sum_assignment``, which returns the same optimum for generic costs); rows / columns assigned to an
  extension entry are unmatched (-1);
* ``cython_bbox.bbox_overlaps`` (py-faster-rcnn's IoU with the +1 pixel convention), restated in numpy;
* the appearance network (``models/unitrack/model``, needs ``checkpoints/UniTrack/timecycle.pth``): detections carry
  seeded synthetic mask-pooled embeddings instead -- ``prepare_obs`` is the one overridden method, exactly the hook the
  reference's own subclasses (``mask.py``, ``box.py``) override.
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import tracker_fixture as fx  # noqa: E402


def install_stubs():
    np.float = float      # the reference predates numpy 1.24
    np.int = int
    from oracle import tracker as otr
    lap = types.ModuleType('lap')
    lap.lapjv = otr.lapjv
    sys.modules['lap'] = lap
    cb = types.ModuleType('cython_bbox')
    cb.bbox_overlaps = otr.bbox_overlaps
    sys.modules['cython_bbox'] = cb
    pm = types.ModuleType('pycocotools')
    pmm = types.ModuleType('pycocotools.mask')
    pm.mask = pmm
    sys.modules['pycocotools'], sys.modules['pycocotools.mask'] = pm, pmm
    # `models` / `models.unitrack` (/ `.utils`: its __init__ imports the visualisation stack) as path-only packages:
    # the reference's models/__init__.py imports mmcv / mmdet
    for name, path in (('models', f'{REF}/models'), ('models.unitrack', f'{REF}/models/unitrack'),
                       ('models.unitrack.utils', f'{REF}/models/unitrack/utils'),
                       ('models.unitrack.data', f'{REF}/models/unitrack/data')):   # data/__init__ imports mmdet
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    model = types.ModuleType('models.unitrack.model')      # appearance network: not used (prepare_obs is overridden)
    model.AppearanceModel = object
    model.partial_load = lambda *a, **k: None
    sys.modules['models.unitrack.model'] = model
    prop = types.ModuleType('models.unitrack.core.propagation')
    prop.propagate = lambda *a, **k: None
    sys.modules['models.unitrack.core.propagation'] = prop


def main():
    install_stubs()
    mt = importlib.import_module('models.unitrack.multitracker')
    matching = importlib.import_module('models.unitrack.core.association.matching')
    bt = importlib.import_module('models.unitrack.basetrack')
    from openpvsg_b200.registry import to_cfg
    cfg = to_cfg(fx.TRACKER_CFG)

    class Tracker(mt.AssociationTracker):
        def __init__(self, tracker_cfg):                      # the reference __init__ minus the appearance network
            self.tracker_cfg = tracker_cfg
            self.tracked_stracks, self.lost_stracks, self.removed_stracks = [], [], []
            self.query_feat_tubes = []
            self.frame_id = 0
            self.det_thresh = tracker_cfg.mots.conf_thres
            self.buffer_size = tracker_cfg.mots.track_buffer
            self.max_time_lost = self.buffer_size
            self.kalman_filter = mt.KalmanFilter()
            if not self.tracker_cfg.mots.asso_with_motion:
                self.tracker_cfg.mots.motion_lambda = 1
                self.tracker_cfg.mots.motion_gated = False

        def prepare_obs(self, img, img0, obs, embs=None):
            return [bt.STrack(tlwh, 1, f, self.buffer_size, None, ac=True) for tlwh, f in obs]

    out = {}
    # ---- 1. stand-alone functions on fixed inputs
    trk, det = fx.embedding_sets(seed=1)
    T = [types.SimpleNamespace(curr_feat=f) for f in trk]
    D = [types.SimpleNamespace(curr_feat=f) for f in det]
    cost, _ = matching.reconsdot_distance(T, D)
    out['reconsdot'] = dict(cost=np.asarray(cost, np.float64).tolist())
    rng = np.random.default_rng(3)
    laps = []
    for n, m, thresh in ((5, 7, 0.9), (8, 8, 0.5), (9, 4, 0.7), (1, 6, 0.9), (12, 12, 2.0)):
        c = rng.random((n, m))
        c[rng.random((n, m)) < 0.15] = np.inf
        matches, ua, ub = matching.linear_assignment(c.copy(), thresh)
        laps.append(dict(cost=np.where(np.isinf(c), -1.0, c).tolist(), thresh=thresh, matches=np.asarray(matches).tolist(),
                         unmatched_a=np.asarray(ua).tolist(), unmatched_b=np.asarray(ub).tolist()))
    out['lap'] = laps
    boxes_a, boxes_b = fx.boxes(seed=4, n=6), fx.boxes(seed=5, n=5)
    out['iou_distance'] = matching.iou_distance(list(boxes_a), list(boxes_b)).tolist()
    kf = mt.KalmanFilter()
    mean, cov = kf.initiate(np.array([50., 40., 0.5, 80.]))
    steps = []
    for z in ([52., 41., 0.5, 81.], [55., 43., 0.52, 80.], [59., 44., 0.5, 79.]):
        mean, cov = kf.predict(mean, cov)
        gd = kf.gating_distance(mean, cov, np.array([z, [0., 0., 1., 10.]]), metric='maha')
        mean, cov = kf.update(mean, cov, np.array(z))
        steps.append(dict(z=z, mean=mean.tolist(), cov=cov.tolist(), gating=gd.tolist()))
    out['kalman'] = steps
    # ---- 2. the whole tracker over a synthetic clip (use_kalman on and off, as the two branches of update())
    for use_kalman in (True, False):
        bt.BaseTrack.reset_count()
        c = to_cfg(fx.TRACKER_CFG)
        c.mots.use_kalman = use_kalman
        tracker = Tracker(c)
        frames = []
        for obs, query_feats in fx.clip(seed=7):
            dets = [(tlwh, f) for tlwh, f, _ in obs]
            if not dets:
                frames.append(dict(ids=[], tlwh=[], num_tubes=len(tracker.query_feat_tubes)))
                continue
            online, n_tubes = tracker.update(None, None, _Obs(dets), query_feats, 0)
            frames.append(dict(ids=[int(t.track_id) for t in online], tlwh=[np.asarray(t.tlwh).tolist() for t in online],
                               num_tubes=int(n_tubes),
                               lost=[int(t.track_id) for t in tracker.lost_stracks],
                               removed=[int(t.track_id) for t in tracker.removed_stracks]))
        tubes = [dict(track_id=int(q.track_id), start=int(q.start_frame_id), end=int(q.end_frame_id),
                      present=[None if e is None else int(e['cls_id']) for e in q.qf_tube]) for q in tracker.query_feat_tubes]
        out['clip_kalman' if use_kalman else 'clip_nokalman'] = dict(frames=frames, tubes=tubes)
    json.dump(out, open(os.path.join(HERE, 'tracker.json'), 'w'))
    print('tracker.json', {k: (len(v) if isinstance(v, list) else list(v)) for k, v in out.items()})


class _Obs(list):
    """``obs`` of update(): only ``obs.shape[1]`` is inspected there (6 columns = category gating)."""
    @property
    def shape(self):
        return (len(self), 5)


if __name__ == '__main__':
    main()
