#!/usr/bin/env python
"""Golden R@K metrics from the REFERENCE's own evaluation loop (build container only).

``tools/rel_test.py`` is imported by file path, unmodified, and its ``evaluate`` (:16-110) is run on CPU with the
reference's relation-head classes (models/relation_head/*.py, torch only), the reference's
``utils/rel_metrics.py`` / ``utils/show_log.py`` and the reference's ``PVSGRelationDataset`` behind a torch
``DataLoader(batch_size=1)`` -- exactly the objects ``tools/rel_test.py:113-180`` wires together.  Only the package
scaffolding is stubbed (``datasets`` / ``models`` / ``utils`` package objects pointing at the reference files, the
pycocotools decoder), none of the executed code.

Data: the synthetic clip of tests/relset_fixture.py (14 tubes x 40 frames); seeded relation weights
(openpvsg_b200.synthetic.relation_state_dicts(seed=1)).  The ground-truth relations are drawn from the reference
model's own ranked predictions ("synthetic GT relations drawn from the CPU-oracle's own tubes so
R@K is well-defined"): ranks on both sides of K = 20 / 50 / 100, some with the predicted span (hits), some with a
disjoint span (weak hits only), some never predicted (misses).  Output: tests/golden/releval.json.

    python tests/golden/make_golden_releval.py
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_relset as mg  # noqa: E402  (sets sys.path, shares the stubs / file writers)

fx, rs, REF = mg.fx, mg.rs, mg.REF
from openpvsg_b200 import synthetic as syn  # noqa: E402

NUM_RELATIONS = 57
RELATION_LIST = [f'relation_{i}' for i in range(NUM_RELATIONS)]
GT_RANKS = (2, 11, 19, 20, 34, 49, 50, 77, 99)       # ranks of the reference's pairwise results used as ground truth
DISJOINT = (11, 50)                                   # of those: GT span made disjoint from the prediction (weak hit only)


def load_reference_stack():
    rm, ds = mg.setup_reference_modules()
    utils_pkg = sys.modules['utils']
    utils_pkg.rel_metrics = mg._load('utils.rel_metrics', f'{REF}/utils/rel_metrics.py')
    utils_pkg.show_log = mg._load('utils.show_log', f'{REF}/utils/show_log.py')
    datasets_pkg = types.ModuleType('datasets')
    datasets_pkg.PVSGRelationDataset = ds.PVSGRelationDataset
    sys.modules['datasets'] = datasets_pkg
    models_pkg, rel_pkg = types.ModuleType('models'), types.ModuleType('models.relation_head')
    models_pkg.relation_head = rel_pkg
    sys.modules['models'], sys.modules['models.relation_head'] = models_pkg, rel_pkg
    for name in ('base', 'convolution', 'transformer', 'train_utils', 'test_utils'):
        setattr(rel_pkg, name, mg._load(f'models.relation_head.{name}', f'{REF}/models/relation_head/{name}.py'))
    rel_test = mg._load('ref_tools_rel_test', f'{REF}/tools/rel_test.py')
    return rm, ds, rel_test


def reference_models(rel_test):
    sds = syn.relation_state_dicts(seed=1)
    mods = [rel_test.ObjectEncoder(feature_dim=256), rel_test.ObjectEncoder(feature_dim=256),
            rel_test.PairProposalNetwork(256, 1024), rel_test.TemporalTransformer(512, NUM_RELATIONS)]
    for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
        m.load_state_dict(sds[k])
        m.eval()
    return mods


def main():
    torch.set_num_threads(8)
    rm, ds, rel_test = load_reference_stack()
    clip = fx.make_clip()
    linker = fx.link(clip)
    mods = reference_models(rel_test)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        data_dir, work_dir = mg.write_clip_files(tmp, clip, linker)
        feat_tubes = {q.track_id: q.qf_tube for q in rs.query_feat_tubes(linker)}
        feats = rm.process_feats(feat_tubes)                      # {tube id: [T,256]}
        keys = list(feats)
        # pass 1: the reference forward on the sample (tools/rel_test.py:35-67) to draw the ground truth from
        with torch.no_grad():
            x = torch.as_tensor(np.array([feats[k] for k in keys])).float()
            sub, obj = mods[0](x), mods[1](x)
            pairs = rel_test.pick_top_pairs_eval(mods[2](sub, obj), 100)
            span_pred, prob = mods[3](rel_test.concatenate_sub_obj(sub, obj, pairs))
            results = rel_test.generate_pairwise_results(span_pred, prob, pairs)
        T = clip['T']
        gt = []
        for r in GT_RANKS:
            res = results[r]
            span = res['relation_span'].copy()
            if r in DISJOINT or span.sum() == 0:
                span = 1.0 - span if r in DISJOINT else np.ones(T)
            gt.append(dict(subject_index=keys[res['subject_index']], object_index=keys[res['object_index']],
                           relation=int(res['relation']), relation_span=span))
        predicted = {(r['subject_index'], r['object_index'], r['relation']) for r in results}
        for s, o, rel in ((0, 1, 5), (3, 2, 40), (7, 7, 1)):      # never predicted with that relation: misses
            assert (s, o, rel) not in predicted
            gt.append(dict(subject_index=keys[s], object_index=keys[o], relation=rel, relation_span=np.ones(T)))
        rm.save_pickle(os.path.join(work_dir, fx.VID, 'relations.pickle'), dict(feats=feats, relations=gt))

        # pass 2: the reference's evaluate() over its own dataset + DataLoader
        from torch.utils.data import DataLoader
        dataset = ds.PVSGRelationDataset(os.path.join(data_dir, 'pvsg.json'), 'train', work_dir)
        loader = DataLoader(dataset, batch_size=1, shuffle=False)
        captured = {}

        def capture(final_metrics, pair_recall_list, K_values, csv_file_path, model_name):
            captured.update(final_metrics=final_metrics, pair_recall_list=pair_recall_list, K_values=K_values)

        rel_test.save_metrics_to_csv = capture
        rel_test.evaluate(mods[0], mods[1], mods[2], mods[3], loader, 100, RELATION_LIST, torch.device('cpu'),
                          os.path.join(tmp, 'm.csv'), 'golden')
        out['K_values'] = list(captured['K_values'])
        out['final_metrics'] = {str(K): {k: float(v) for k, v in m.items()} for K, m in captured['final_metrics'].items()}
        out['pair_recall_list'] = [float(v) for v in captured['pair_recall_list']]
        out['gt_relations'] = [dict(subject_index=int(g['subject_index']), object_index=int(g['object_index']),
                                    relation=g['relation'], relation_span=g['relation_span'].tolist()) for g in gt]
        out['tube_keys'] = [int(k) for k in keys]
        out['top_results'] = [[r['subject_index'], r['object_index'], r['relation'], int(r['relation_span'].sum())]
                              for r in results]
        mp = torch.max(prob, dim=1).values.sort(descending=True).values
        out['min_rank_gap'] = float((mp[:-1] - mp[1:]).min())
        out['rank_gap_20'], out['rank_gap_50'] = float(mp[19] - mp[20]), float(mp[49] - mp[50])
        flat = mods[2](sub, obj).detach().clone()
        flat.fill_diagonal_(-float('inf'))
        top = flat.flatten().sort(descending=True).values[:101]
        out['min_pair_gap'] = float((top[:-1] - top[1:]).min())
        # the only places where a re-ordered fp32 forward can change a metric: the top-20 / top-100 pair boundaries
        out['pair_gap_20'] = float(top[19] - top[20])
        out['pair_gap_100'] = float(top[99] - top[100])
        # metric helpers on fixed inputs
        rmx = sys.modules['utils.rel_metrics']
        a, b = np.array([1, 1, 0, 0, 1.]), np.array([0, 1, 1, 0, 1.])
        out['helpers'] = dict(iou=float(rmx.calculate_iou(a, b)), iou_empty=float(rmx.calculate_iou(a * 0, b * 0)),
                              pair_recall=float(rmx.calculate_pair_recall_at_k([[0, 1], [2, 3], [4, 5]], [[2, 3], [9, 9], [2, 3]], 2)))
    json.dump(out, open(os.path.join(HERE, 'releval.json'), 'w'), indent=0)
    print('final_metrics', out['final_metrics'])
    print('pair recall', out['pair_recall_list'], 'min rank gap', out['min_rank_gap'], 'min pair gap', out['min_pair_gap'],
          'gaps at 20 / 100', out['pair_gap_20'], out['pair_gap_100'], 'rank gaps at 20 / 50', out['rank_gap_20'], out['rank_gap_50'])


if __name__ == '__main__':
    main()
