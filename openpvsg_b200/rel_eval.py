"""Relation-head evaluation: Recall@K / mean Recall@K / weak recalls and pair recall (SURVEY.md 8c parity gate "R@K
identical", BASELINE configs[4] "R@K vs CPU reference").

Reference: ``tools/rel_test.py:16-110`` (``evaluate``: per video -- forward, top pairs, pair recall@20, pairwise
results, hit / weak-hit accounting per GT relation) and ``utils/rel_metrics.py:6-56`` (``calculate_iou``,
``calculate_pair_recall_at_k``, ``calculate_final_metrics``).  Same function names and return values; the forward
section is ``relation_head.relation_forward`` (device-resident pair selection), everything after it is small host
bookkeeping on the result lists.  Golden vectors: the reference's own ``evaluate`` run on the synthetic relation
set of tests/relset_fixture.py (``tests/golden/make_golden_releval.py``).
"""
import numpy as np
import torch

from . import relation_head as rh

K_VALUES = (20, 50, 100)


def calculate_iou(span1, span2):
    """Temporal IoU of two 0/1 spans (utils/rel_metrics.py:6-9)."""
    span1, span2 = np.asarray(span1, dtype=np.float64), np.asarray(span2, dtype=np.float64)
    inter = (span1 * span2).sum()
    union = span1.sum() + span2.sum() - inter
    return inter / union if union > 0 else 0


def calculate_pair_recall_at_k(selected_pairs, gt_pairs, k=20):
    """Share of the distinct GT (subject, object) pairs among the first k selected pairs (utils/rel_metrics.py:21-26)."""
    gt = {tuple(p) for p in gt_pairs}
    top = {tuple(p) for p in selected_pairs[:k]}
    return len(top & gt) / len(gt) if gt else 0


def new_recall_dict(relation_list, K_values=K_VALUES):
    """tools/rel_test.py:22-23."""
    return {K: {idx: dict(name=name, total=0, hit=0, weak_hit=0) for idx, name in enumerate(relation_list)} for K in K_values}


def accumulate(relation_recall_dict, results, gt_relations, K_values=K_VALUES):
    """tools/rel_test.py:69-92: every GT relation counts once per K; the FIRST result with the same (subject, object,
    relation) is a weak hit for every K above its rank and a hit if its span has temporal IoU >= 0.5."""
    for gt in gt_relations:
        key = (int(gt['subject_index']), int(gt['object_index']), int(gt['relation']))
        for K in K_values:
            relation_recall_dict[K][key[2]]['total'] += 1
        for idx, res in enumerate(results):
            if (res['subject_index'], res['object_index'], res['relation']) == key:
                t_iou = calculate_iou(np.asarray(gt['relation_span']).reshape(-1), res['relation_span'])
                for K in K_values:
                    if idx < K:
                        relation_recall_dict[K][key[2]]['weak_hit'] += 1
                        if t_iou >= 0.5:
                            relation_recall_dict[K][key[2]]['hit'] += 1
                break


def calculate_final_metrics(relation_recall_dict, K_values):
    """utils/rel_metrics.py:29-56: micro recall, mean-over-relations recall, and their weak (span-agnostic) forms."""
    cells0 = relation_recall_dict[K_values[0]].values()
    num_valid = len([c for c in cells0 if c['total'] != 0])
    out = {}
    for K in K_values:
        cells = list(relation_recall_dict[K].values())
        total = sum(c['total'] for c in cells)
        valid = [c for c in cells if c['total'] != 0]
        out[K] = dict(recall=sum(c['hit'] for c in cells) / total if total > 0 else 0,
                      mean_recall=sum(c['hit'] / c['total'] for c in valid) / num_valid,
                      weak_recall=sum(c['weak_hit'] for c in cells) / total if total > 0 else 0,
                      weak_mean_recall=sum(c['weak_hit'] / c['total'] for c in valid) / num_valid)
    return out


@torch.no_grad()
def evaluate(models, samples, relation_list, num_top_pairs=100, K_values=K_VALUES, device='cuda', pairwise=True,
             forward_fn=None):
    """``tools/rel_test.py::evaluate`` over an iterable of ``PVSGRelationDataset`` samples (dicts with ``feats``
    [N,T,256] and ``relations``).  models: (subject_encoder, object_encoder, pair_proposal_model, relation_model)
    of ``openpvsg_b200.relation_head``.  Returns (final_metrics, pair_recall_list, relation_recall_dict).
    forward_fn(feats) -> dict(pairs, span_pred, prob) replaces the device forward (used by the CPU tests)."""
    recall = new_recall_dict(relation_list, K_values)
    pair_recalls = []
    for sample in samples:
        feats = torch.as_tensor(np.asarray(sample['feats'])).float()
        gt_relations = sample['relations']
        if forward_fn is None:
            out = rh.relation_forward(*models, feats.to(device), num_top_pairs)
        else:
            out = forward_fn(feats)
        pairs = out['pairs'].cpu().tolist() if torch.is_tensor(out['pairs']) else [list(p) for p in out['pairs']]
        gt_pairs = [[int(r['subject_index']), int(r['object_index'])] for r in gt_relations]
        pair_recalls.append(calculate_pair_recall_at_k(pairs, gt_pairs, 20))
        gen = rh.generate_pairwise_results if pairwise else rh.generate_results
        accumulate(recall, gen(out['span_pred'], out['prob'], pairs), gt_relations, K_values)
    return calculate_final_metrics(recall, list(K_values)), pair_recalls, recall
