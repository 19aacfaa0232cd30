"""Golden vectors for the training-target preparation (build container only: needs /root/reference).

Runs the reference's OWN ``preprocess_video_panoptic_gt`` (models/mask2former_vps/utils.py:94-140) on seeded inputs and
stores inputs + outputs in tests/golden/train_gt.json.  The function is executed from the reference file where it lies
(its module imports cv2 / pycocotools / the tracker at module scope, so only this function's source is compiled); the one
third-party type it touches, mmdet's BitmapMasks, is replaced by a stand-in with the two methods it calls (pad, to_tensor).

    python tests/golden/make_golden_train_gt.py
"""
import ast
import json
import os

import numpy as np
import torch

REF = '/root/reference/models/mask2former_vps/utils.py'
HERE = os.path.dirname(os.path.abspath(__file__))


class Masks:
    """BitmapMasks stand-in: masks [n,h,w] uint8 numpy."""

    def __init__(self, m):
        self.m = np.asarray(m, np.uint8)

    def pad(self, shape, pad_val=0):
        n, h, w = self.m.shape
        out = np.full((n, shape[0], shape[1]), pad_val, np.uint8)
        out[:, :h, :w] = self.m
        return Masks(out)

    def to_tensor(self, dtype, device):
        return torch.as_tensor(self.m, dtype=dtype, device=device)


def load_reference_function():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'preprocess_video_panoptic_gt')
    ns = {'torch': torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, 'exec'), ns)
    return ns['preprocess_video_panoptic_gt']


def cases():
    rng = np.random.default_rng(7)
    out = []
    # (frames, instances per frame as lists of (id, label)), image size, pad size
    specs = [
        ([[(11, 3), (12, 40)], [(12, 40), (11, 3)]], (6, 8), (8, 8)),                      # same instances, different order
        ([[(5, 7)], [(5, 7), (9, 120)], [(9, 120)]], (5, 7), (5, 7)),                      # instances missing in some frames
        ([[(1, 0), (2, 0), (3, 125)], [(3, 125)]], (4, 4), (8, 12)),                       # two instances of one class
    ]
    for frames, (h, w), pad in specs:
        masks = [rng.integers(0, 2, (len(f), h, w)).astype(np.uint8) for f in frames]
        labels = [[t, lab] for t, f in enumerate(frames) for (_, lab) in f]
        ids = [[t, iid] for t, f in enumerate(frames) for (iid, _) in f]
        out.append(dict(masks=[m.tolist() for m in masks], labels=labels, ids=ids, pad=list(pad)))
    return out


if __name__ == '__main__':
    ref = load_reference_function()
    golden = []
    for c in cases():
        metas = [dict(pad_shape=(c['pad'][0], c['pad'][1], 3)) for _ in c['masks']]
        lab, msk = ref(torch.tensor(c['labels']), [Masks(m) for m in c['masks']], None, torch.tensor(c['ids']), 115, 11, metas)
        golden.append(dict(c, out_labels=lab.tolist(), out_masks=msk.tolist()))
    json.dump(golden, open(os.path.join(HERE, 'train_gt.json'), 'w'))
    print('wrote', len(golden), 'cases', [np.asarray(g['out_masks']).shape for g in golden])


# ---------------------------------------------------------------------------------------------------------------------
# frame unpacking of the IPS tracker driver: LoadOutputsFromMask2Former._get_binary_masks_and_query_feats /
# _unify_query_feat_dim (models/unitrack/data/single_video.py:49-85), executed from the reference file (two methods only: the
# module imports cv2 / torchvision at module scope; `np.int` no longer exists in numpy 2, so the namespace maps it to int)
def load_single_video_methods():
    src = '/root/reference/models/unitrack/data/single_video.py'
    tree = ast.parse(open(src).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'LoadOutputsFromMask2Former')
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('_get_binary_masks_and_query_feats', '_unify_query_feat_dim')]

    class NP:
        int = int

        def __getattr__(self, k):
            return getattr(np, k)

    ns = {'np': NP(), 'INSTANCE_OFFSET': 1000}
    exec(compile(ast.Module(body=fns, type_ignores=[]), src, 'exec'), ns)
    return type('Loader', (), {'num_classes': 126, '_get_binary_masks_and_query_feats': ns['_get_binary_masks_and_query_feats'],
                               '_unify_query_feat_dim': ns['_unify_query_feat_dim']})()


def frame_cases():
    rng = np.random.default_rng(3)
    out = []
    for ids in ([126, 1005, 2005, 120], [126], [3007, 121, 1003], [120, 121, 122]):
        pan = rng.choice(ids, size=(6, 9)).astype(np.int32)
        present = [i for i in np.unique(pan).tolist() if i != 126]
        qf = {i: [rng.standard_normal((1, 8)).astype(np.float32) for _ in range(1 if i >= 1000 else int(rng.integers(1, 4)))] for i in present}
        out.append((pan, qf))
    return out


if __name__ == '__main__':
    loader = load_single_video_methods()
    golden = []
    for pan, qf in frame_cases():
        masks, feats = loader._get_binary_masks_and_query_feats(pan, qf)
        golden.append(dict(pan=pan.tolist(), query_feats={str(k): [x.tolist() for x in v] for k, v in qf.items()},
                           masks=np.asarray(masks).tolist(), feats=[dict(query_feat=np.asarray(f['query_feat']).tolist(), cls_id=int(f['cls_id'])) for f in feats]))
    json.dump(golden, open(os.path.join(HERE, 'tracker_frames.json'), 'w'))
    print('wrote', len(golden), 'frame cases', [np.asarray(g['masks']).shape for g in golden])
