#!/usr/bin/env python
"""Golden vectors of the relation-set builder from the REFERENCE's own code (build container only).

``utils/relation_matching.py`` and ``datasets/datasets/pvsg_relation.py`` are imported by file path,
unmodified; their only missing import, ``pycocotools.mask``, is stubbed with a stand-alone RLE decoder
written here (pycocotools is third party and not installable offline).  The reference then runs its
file-based flow -- masks.txt -> decoded tubes -> per-frame IoU matching against PNG ground truth ->
compaction -> relation translation -> relations.pickle -> PVSGRelationDataset -- on the synthetic clip of
tests/relset_fixture.py, and the intermediate and final structures are stored as JSON
(tests/golden/relset.json).

    python tests/golden/make_golden_relset.py
"""
import importlib.util
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import relset_fixture as fx  # noqa: E402
from openpvsg_b200 import relation_set as rs  # noqa: E402  (only SimpleTracker / query_feat_tubes containers)


def _decode(rle):
    """Stand-alone restatement of pycocotools rleFrString + rleDecode (column-major)."""
    h, w = rle['size']
    s = rle['counts']
    cnts, p = [], 0
    while p < len(s):
        x, k = 0, 0
        while True:
            c = ord(s[p]) - 48
            p += 1
            x |= (c & 0x1f) << (5 * k)
            k += 1
            if not (c & 0x20):
                if c & 0x10:
                    x |= -1 << (5 * k)
                break
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    flat = np.zeros(h * w, np.uint8)
    pos = 0
    for i, c in enumerate(cnts):
        if i % 2:
            flat[pos:pos + c] = 1
        pos += c
    return flat.reshape(w, h).T.copy()


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def ordered(d):
    """dict -> list of [key, value] pairs (JSON objects lose int keys; insertion ORDER is part of the contract)."""
    return [[k, v] for k, v in d.items()]


def setup_reference_modules():
    """pycocotools stub + the reference's utils/relation_matching.py and datasets/datasets/pvsg_relation.py by path."""
    pm = types.ModuleType('pycocotools')
    pm.mask = types.ModuleType('pycocotools.mask')
    pm.mask.decode = _decode
    sys.modules['pycocotools'] = pm
    sys.modules['pycocotools.mask'] = pm.mask
    rm = _load('utils.relation_matching', f'{REF}/utils/relation_matching.py')
    utils_pkg = types.ModuleType('utils')
    utils_pkg.relation_matching = rm
    sys.modules['utils'] = utils_pkg
    ds = _load('ref_pvsg_relation', f'{REF}/datasets/datasets/pvsg_relation.py')
    return rm, ds


def write_clip_files(tmp, clip, linker):
    """The files the reference flow reads: pvsg.json, GT PNGs, masks.txt, query_feats.pickle."""
    from PIL import Image
    data_dir, work_dir = os.path.join(tmp, 'data'), os.path.join(tmp, 'work')
    os.makedirs(os.path.join(data_dir, 'vidor', 'masks', fx.VID))
    os.makedirs(os.path.join(work_dir, fx.VID, 'quantitive'))
    json.dump(fx.make_anno(), open(os.path.join(data_dir, 'pvsg.json'), 'w'))
    for t in range(clip['T']):
        Image.fromarray(clip['gt'][t].astype(np.uint8)).save(
            os.path.join(data_dir, 'vidor', 'masks', fx.VID, f'{t:04d}.png'))
    open(os.path.join(work_dir, fx.VID, 'quantitive', 'masks.txt'), 'w').write(linker.masks_txt())
    pickle.dump(rs.query_feat_tubes(linker), open(os.path.join(work_dir, fx.VID, 'query_feats.pickle'), 'wb'))
    return data_dir, work_dir


def main(variant='base', out_name='relset.json'):
    rm, ds = setup_reference_modules()
    clip = fx.make_clip(variant=variant)
    linker = fx.link(clip)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        data_dir, work_dir = write_clip_files(tmp, clip, linker)

        # ---- tools/prepare_rel_set.py:24-52, the reference's own calls
        pvsg_dataset = rm.PVSGRelationAnnotation(os.path.join(data_dir, 'pvsg.json'), 'train')
        query_feats = rm.load_pickle(os.path.join(work_dir, fx.VID, 'query_feats.pickle'))
        pred_mask_tubes = rm.get_pred_mask_tubes_one_video(fx.VID, work_dir)
        matching = rm.match_and_process_gt_tubes(fx.VID, pvsg_dataset, pred_mask_tubes, data_dir=data_dir)
        compact = rm.compact_matching_dict(matching)
        gt_relations = pvsg_dataset[fx.VID]['relations']
        pred_relations = rm.translate_gt_relations(compact, gt_relations)
        pred_feat_tubes = {q.track_id: q.qf_tube for q in query_feats}
        relation_dict = rm.process_feats_and_relations(pred_relations, pred_feat_tubes)
        rm.save_pickle(os.path.join(work_dir, fx.VID, 'relations.pickle'), relation_dict)
        relations_full = rm.process_relations(pred_relations, pred_feat_tubes)
        sample = ds.PVSGRelationDataset(os.path.join(data_dir, 'pvsg.json'), 'train', work_dir, return_mask=True)[0]

        # GT tubes + the alternative matcher (match_tubes) for the IoU helper
        gt_tubes = rm.get_gt_mask_tubes_one_video(fx.VID, pvsg_dataset, data_dir)
        match_tubes = rm.match_tubes(gt_tubes, pred_mask_tubes)

        out['pred_mask_tubes'] = [[tid, t['cid'], [list(m.keys())[0] for m in t['mask']],
                                   [int(list(m.values())[0].sum()) for m in t['mask']]]
                                  for tid, t in pred_mask_tubes.items()]
        out['annotation'] = jsonable(pvsg_dataset[fx.VID])
        out['matching'] = [[k, ordered(v)] for k, v in matching.items()]
        out['match_tubes'] = [[k, ordered(v)] for k, v in match_tubes.items()]
        out['compact'] = [[k, ordered(v)] for k, v in compact.items()]
        out['pred_relations'] = jsonable(pred_relations)
        out['pairs'] = rm.process_pairs(pred_relations)
        out['relation_dict'] = dict(
            feat_keys=[int(k) for k in relation_dict['feats']],
            feat_sums=[float(np.abs(v).sum()) for v in relation_dict['feats'].values()],
            feat_dtype=str(next(iter(relation_dict['feats'].values())).dtype),
            relations=jsonable(relation_dict['relations']))
        out['relations_full'] = [dict(relation=r['relation'], span=r['relation_span'].tolist(),
                                      s_sum=float(np.abs(r['tube_s']).sum()), o_sum=float(np.abs(r['tube_o']).sum()))
                                 for r in relations_full]
        out['sample'] = dict(vid=sample['vid'], feats_shape=list(sample['feats'].shape),
                             feats_sum=float(np.abs(sample['feats']).sum()), pairs=sample['pairs'],
                             relations=jsonable(sample['relations']),
                             idx2key=[[int(k), int(v)] for k, v in sample['idx2key'].items()],
                             mask_frames=[[list(m.keys())[0] for m in tube.get('mask', [])] for tube in sample['masks']])
        out['convert_to_ranges'] = [[fr, rm.convert_to_ranges(fr)] for fr in
                                    ([0, 1, 2, 3, 4], [0, 1, 2, 9, 10, 11, 12, 13, 20], [5], [3, 1, 2, 0, 4, 8, 12, 16, 17])]
        out['find_ranges'] = [[fr, rm.find_ranges(fr)] for fr in ([0, 1, 2, 3, 4], [0, 5, 11, 12, 30], [7])]
        out['checksum'] = float(np.abs(np.concatenate([f.ravel() for f in clip['feats']])).sum()
                                + clip['gt'].sum() + clip['pan'].sum())
    json.dump(out, open(os.path.join(HERE, out_name), 'w'), indent=0)
    print('matching:', out['matching'])
    print('compact:', out['compact'])
    print('pred_relations:', out['pred_relations'])
    print('n relations kept:', len(out['relation_dict']['relations']))


if __name__ == '__main__':
    main()
    main('gaps', 'relset_gaps.json')
