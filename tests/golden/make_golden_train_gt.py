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
