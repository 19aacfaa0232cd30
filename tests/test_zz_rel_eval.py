"""R@K evaluation (SURVEY 8c parity gate "R@K identical"; BASELINE configs[4] "R@K vs CPU reference").

tests/golden/releval.json comes from the reference's own ``tools/rel_test.py::evaluate`` run on CPU with the
reference's relation-head classes, metrics, dataset and a torch DataLoader (tests/golden/make_golden_releval.py).
CPU: ``rel_eval.evaluate`` with the oracle forward reproduces every number.  GPU: the same call through the CUDA
relation head (passes on a B200: profiles/r01u_gpu_test_logs.txt).  File named to sort last: it was the round's last
addition."""
import json
import os

import numpy as np
import pytest
import torch

import relset_fixture as fx
from openpvsg_b200 import rel_eval, relation_set as rs, synthetic as syn
from oracle import relation as orel

HERE = os.path.dirname(os.path.abspath(__file__))
RELATION_LIST = [f'relation_{i}' for i in range(57)]


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(HERE, 'golden', 'releval.json')))


@pytest.fixture(scope='module')
def dataset(golden):
    """The relation set of the golden run, served from memory by the product's PVSGRelationDataset."""
    clip = fx.make_clip()
    linker = fx.link(clip)
    feats = rs.process_feats({t.track_id: t.qf_tube for t in rs.query_feat_tubes(linker)})
    assert [int(k) for k in feats] == golden['tube_keys']
    rels = [dict(subject_index=g['subject_index'], object_index=g['object_index'], relation=g['relation'],
                 relation_span=np.array(g['relation_span'])) for g in golden['gt_relations']]
    return rs.PVSGRelationDataset(fx.make_anno(), 'train', memory={fx.VID: dict(feats=feats, relations=rels)})


def _check(golden, final, pair_recalls):
    assert pair_recalls == pytest.approx(golden['pair_recall_list'], abs=1e-12)
    for K in golden['K_values']:
        for name, want in golden['final_metrics'][str(K)].items():
            assert final[K][name] == pytest.approx(want, abs=1e-12), (K, name)


def test_metric_helpers_match_reference(golden):
    a, b = np.array([1, 1, 0, 0, 1.]), np.array([0, 1, 1, 0, 1.])
    h = golden['helpers']
    assert rel_eval.calculate_iou(a, b) == pytest.approx(h['iou'], abs=1e-15)
    assert rel_eval.calculate_iou(a * 0, b * 0) == h['iou_empty'] == 0
    assert rel_eval.calculate_pair_recall_at_k([[0, 1], [2, 3], [4, 5]], [[2, 3], [9, 9], [2, 3]], 2) == h['pair_recall']
    assert rel_eval.calculate_pair_recall_at_k([[0, 1]], [], 20) == 0


def test_evaluate_matches_reference_cpu(golden, dataset):
    """Oracle forward (CPU) + the product's accounting == the reference's evaluate()."""
    sds = syn.relation_state_dicts(seed=1)
    seen = {}

    def forward(feats):
        out = orel.relation_forward(sds, feats, 100)
        seen.update(out)
        return out

    final, pair_recalls, recall = rel_eval.evaluate(None, [dataset[0]], RELATION_LIST, 100, forward_fn=forward)
    # the oracle's ranked results are the reference's (so the metrics below are compared on the same predictions)
    from openpvsg_b200 import relation_head as rh
    res = rh.generate_pairwise_results(seen['span_pred'], seen['prob'], seen['pairs'])
    assert [[r['subject_index'], r['object_index'], r['relation'], int(r['relation_span'].sum())] for r in res] \
        == golden['top_results']
    _check(golden, final, pair_recalls)
    assert sum(c['total'] for c in recall[20].values()) == len(golden['gt_relations'])
    # the non-pairwise strategy (generate_results) runs through the same accounting
    final2, _, _ = rel_eval.evaluate(None, [dataset[0]], RELATION_LIST, 100, forward_fn=forward, pairwise=False)
    assert set(final2) == {20, 50, 100} and 0 <= final2[100]['weak_recall'] <= 1


def test_accumulate_first_match_and_rank_rule():
    """A GT relation is matched by the FIRST equal triplet only; idx < K decides the K buckets; t-IoU >= 0.5 a hit."""
    span = np.array([1, 1, 1, 1, 0, 0.])
    results = [dict(subject_index=0, object_index=1, relation=3, relation_span=np.array([0, 0, 0, 0, 1, 1.]))] + \
              [dict(subject_index=5, object_index=5, relation=0, relation_span=span)] * 24 + \
              [dict(subject_index=0, object_index=1, relation=3, relation_span=span),
               dict(subject_index=2, object_index=1, relation=4, relation_span=np.array([1, 1, 0, 0, 0, 0.]))]
    rd = rel_eval.new_recall_dict([f'r{i}' for i in range(6)])
    gts = [dict(subject_index=0, object_index=1, relation=3, relation_span=span),        # first match at rank 0: weak only
           dict(subject_index=2, object_index=1, relation=4, relation_span=span[None]),   # rank 26, t-IoU 0.5: hit for K >= 50
           dict(subject_index=4, object_index=4, relation=5, relation_span=span)]         # never predicted
    rel_eval.accumulate(rd, results, gts)
    assert (rd[20][3]['weak_hit'], rd[20][3]['hit'], rd[20][3]['total']) == (1, 0, 1)
    assert (rd[20][4]['weak_hit'], rd[50][4]['weak_hit'], rd[50][4]['hit'], rd[100][4]['hit']) == (0, 1, 1, 1)
    assert rd[100][5] == dict(name='r5', total=1, hit=0, weak_hit=0)
    m = rel_eval.calculate_final_metrics(rd, [20, 50, 100])
    assert m[20]['weak_recall'] == pytest.approx(1 / 3) and m[50]['recall'] == pytest.approx(1 / 3)
    assert m[100]['mean_recall'] == pytest.approx((0 + 1 + 0) / 3)


@pytest.mark.gpu
def test_evaluate_matches_reference_gpu(golden, dataset):
    """The CUDA relation head through rel_eval.evaluate: R@K / mR@K / weak recalls / pair recall identical to the
    reference's CPU run.  The decision margins of this fixture (golden: score gap at the top-20 / top-100 pair
    boundaries 3.9e-2 / 3.1e-4, rank gaps at K = 20 / 50 of 5.8e-3 / 1.5e-3) are far above the fp32 re-association
    error of the device path (~1e-5)."""
    from openpvsg_b200 import relation_head as rh
    assert min(golden['pair_gap_20'], golden['pair_gap_100'], golden['rank_gap_20'], golden['rank_gap_50']) > 1e-4
    dev = torch.device('cuda')
    sds = syn.relation_state_dicts(seed=1)
    mods = [rh.ObjectEncoder(256), rh.ObjectEncoder(256), rh.PairProposalNetwork(256, 1024), rh.TemporalTransformer(512, 57)]
    for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
        m.load_state_dict(sds[k])
        m.to(dev)
    final, pair_recalls, _ = rel_eval.evaluate(mods, [dataset[0]], RELATION_LIST, 100, device=dev)
    _check(golden, final, pair_recalls)
