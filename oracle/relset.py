"""TEST INFRASTRUCTURE (oracle): CPU statement of the per-frame overlap counts behind the relation-set builder.

Follows ``calculate_iou`` / ``match_and_process_gt_tubes`` (reference utils/relation_matching.py:156-165,
205-260): the reference evaluates |gt & pred| and |gt | pred| per (frame, GT object, tube) on decoded masks; both
sides partition the frame, so the same numbers are the cells / row sums / column sums of a joint histogram.
Pinned against the reference's own functions through tests/golden/relset.json
(tests/golden/make_golden_relset.py).  Imported by tests/ only; the product path is ``pvsg_tube_overlap``.
"""
import numpy as np


def slot_ids(seg_info_row):
    """Kept segment ids in order of first appearance (seg_info row = [n, (query, class, segment id, area) * n])."""
    n = int(seg_info_row[0])
    ids = []
    for seg in np.asarray(seg_info_row[1:1 + 4 * n]).reshape(n, 4)[:, 2].tolist():
        if seg >= 0 and seg not in ids:
            ids.append(int(seg))
    return ids


def joint_histogram(gt, pan, seg_info, num_gt):
    """counts int32 [T, num_gt+1, Q+1]: counts[t,g,s] = #{gt == g and pan == id of slot s}; row num_gt = GT ids
    outside [0, num_gt), column Q = pixels of no kept segment."""
    T = gt.shape[0]
    Qn = (seg_info.shape[1] - 1) // 4
    out = np.zeros((T, num_gt + 1, Qn + 1), np.int32)
    for t in range(T):
        slot = np.full(pan[t].shape, Qn, np.int64)
        for k, s in enumerate(slot_ids(seg_info[t])):
            slot[pan[t] == s] = k
        row = np.where((gt[t] >= 0) & (gt[t] < num_gt), gt[t], num_gt).astype(np.int64)
        np.add.at(out[t], (row.ravel(), slot.ravel()), 1)
    return out


def iou_from_masks(gt_mask, pred_mask):
    """calculate_iou, utils/relation_matching.py:156-165."""
    union = np.logical_or(gt_mask, pred_mask).sum()
    return 0 if union == 0 else np.logical_and(gt_mask, pred_mask).sum() / union
