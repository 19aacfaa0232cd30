"""Seeded synthetic clip for the relation-set builder tests: GT instance-id maps, predicted panoptic maps
(+ seg_info rows as pvsg_panoptic_fuse writes them), per-frame query features, a pvsg.json-style annotation.
Shared by tests/golden/make_golden_relset.py (reference side) and tests/test_relation_set_*.py."""
import numpy as np

CLASSES = dict(thing=['adult', 'ball', 'dog'], stuff=['floor', 'wall'])
RELATIONS = ['holding', 'next to', 'on', 'chasing']
VID = '0001_synthetic'          # 4-digit head => data source 'vidor' (utils/relation_matching.py:212-217)
Q = 12                          # seg_info capacity per frame
INSTANCE_OFFSET = 1000


def _boxes(t):
    """object id -> (class index, y0, x0, h, w) at frame t; objects drift so IoUs cross 0.5 over time."""
    return {
        1: (0, 4 + t // 6, 3 + t // 2, 18, 12),        # adult walking right
        2: (0, 20, 40 - t // 3, 16, 10),               # second adult (same class: several candidate tubes)
        3: (1, 8, 30 + (t % 7), 6, 6),                 # ball, jittering
        4: (2, 28 - t // 8, 14, 8, 14),                # dog
    }


def make_clip(T=40, H=48, W=64, seed=0, variant='base'):
    """variant 'base': every object predicted in (almost) every frame.  variant 'gaps': the edge cases of the matching /
    compaction rules -- frames without any kept segment, an object whose prediction comes and goes in stretches
    separated by more than 5 frames (find_ranges splits), a second tube for the same object (list-of-ranges form), a tube
    matched on fewer than 5 frames (dropped), and a GT object that is never predicted."""
    rng = np.random.default_rng(seed)
    gaps = variant == 'gaps'
    gt = np.zeros((T, H, W), np.int32)
    pan = np.full((T, H, W), 126, np.int32)            # void label of the panoptic head
    seg_info = np.zeros((T, 1 + 4 * Q), np.int32)
    seg_ids, feats = [], []
    for t in range(T):
        gt[t, H // 2:] = 5                             # stuff object 'floor' (id 5)
        pan[t, H // 2 + (t % 3 == 0):] = 3             # stuff class 3 ('floor'), a row off every third frame
        if gaps and t in (10, 11, 12):
            pan[t] = 126
        rows = [(0, 3, 3)]                             # (query, class, segment id)
        empty = gaps and t in (10, 11, 12)              # the detector kept nothing in these frames
        for oid, (cls, y0, x0, h, w) in _boxes(t).items():
            gt[t, y0:y0 + h, x0:x0 + w] = oid
            if empty or (gaps and oid == 2 and not (t < 7 or 14 <= t < 21 or 28 <= t < 35)) \
                    or (gaps and oid == 4 and not (3 <= t < 7)) or (gaps and oid == 3):
                continue                                 # 2: three stretches; 4: four frames only; 3: never predicted
            # prediction: shifted / shrunk copy; object 1 is lost for frames 14..21 and comes back as a NEW
            # instance id (two tubes for one GT object); the ball is badly localised on odd frames
            if oid == 1 and 14 <= t < 22:
                continue
            inst = oid + (10 if (oid == 1 and t >= 22) else 0)
            dy, dx = (3, 4) if (oid == 3 and t % 2) else (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
            shrink = 1 if oid != 4 else 0
            pan[t, y0 + dy:y0 + dy + h - shrink, x0 + dx:x0 + dx + w - shrink] = cls + inst * INSTANCE_OFFSET
            rows.append((inst, cls, cls + inst * INSTANCE_OFFSET))
        if t % 5 == 0:      # short-lived distractor balls: one new tube each (tube ids reach two digits)
            inst = 20 + t // 5
            pan[t, 0:3, W - 4:W - 1] = 1 + inst * INSTANCE_OFFSET
            rows.append((inst, 1, 1 + inst * INSTANCE_OFFSET))
        # keep only the segments that survived the painting order (as the fusion head does)
        rows = [r for r in rows if (pan[t] == r[2]).any()]
        seg_info[t, 0] = len(rows)
        for k, (q, cls, seg) in enumerate(rows):
            seg_info[t, 1 + 4 * k:5 + 4 * k] = (q, cls, seg, int((pan[t] == seg).sum()))
        seg_ids.append([r[2] for r in rows])
        feats.append(rng.standard_normal((len(rows), 256)).astype(np.float32))
    return dict(gt=gt, pan=pan, seg_info=seg_info, seg_ids=seg_ids, feats=feats, T=T, H=H, W=W)


def make_anno():
    objects = [dict(object_id=1, category='adult', is_thing=True), dict(object_id=2, category='adult', is_thing=True),
               dict(object_id=3, category='ball', is_thing=True), dict(object_id=4, category='dog', is_thing=True),
               dict(object_id=5, category='floor', is_thing=False)]
    relations = [[1, 3, 'holding', [[0, 12], [24, 38]]], [4, 1, 'chasing', [[2, 40]]], [2, 5, 'on', [[0, 40]]],
                 [1, 2, 'next to', [[5, 30]]], [3, 5, 'on', [[10, 13]]], [4, 2, 'looking at', [[0, 40]]],
                 [1, 5, 'on', [[0, 9], [30, 40]]]]
    return dict(split=dict(vidor=dict(train=[VID], val=[]), epic_kitchen=dict(train=[], val=[]),
                           ego4d=dict(train=[], val=[])),
                objects=CLASSES, relations=RELATIONS,
                data=[dict(video_id=VID, objects=objects, relations=relations)])


def link(clip):
    """TubeLinker over the clip (masks.txt rows from the panoptic maps)."""
    from openpvsg_b200 import tubes
    linker = tubes.TubeLinker()
    for t in range(clip['T']):
        linker.add_frame(clip['seg_ids'][t], clip['feats'][t], pan=clip['pan'][t])
    return linker


def frame_tube_ids(clip, linker):
    """slot -> tube id per frame (slots = order of first appearance of a segment id in seg_info)."""
    from openpvsg_b200 import tubes
    return [[linker.object_list.index(s) + 1 for s in tubes.slot_ids(clip['seg_info'][t])] for t in range(clip['T'])]
