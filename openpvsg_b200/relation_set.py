"""Relation-set builder: predicted tubes <-> ground-truth objects -> the relation head's training /
evaluation samples, handed over in memory (second half).

Reference data flow (``tools/prepare_rel_set.py:24-52``): ``masks.txt`` + ``query_feats.pickle`` are read
back, every RLE row is decoded to a full-frame mask, every (frame, GT object, same-class tube) triple gets a
two-pass numpy IoU (``utils/relation_matching.py:205-260``), the matches are compacted into frame ranges,
the GT relations are translated onto predicted tube ids and ``relations.pickle`` is written for
``datasets/datasets/pvsg_relation.py:42-79``.

Here the IoUs of a frame come from ONE device pass over the GT id map and the panoptic map
(``ops.tube_overlap`` -> ``pvsg_tube_overlap``, a joint histogram -- both sides are partitions of the frame),
the host keeps only the small dictionary logic, and the result goes straight into ``PVSGRelationDataset``
without touching the disk.  The file formats stay readable / writable for tools that want them.

Function names, argument meaning and return structures follow ``utils/relation_matching.py`` so the parity
tests read like calls of the reference (golden vectors: ``tests/golden/make_golden_relset.py``).
"""
import copy
import json
import os
import pickle
from collections import Counter

import numpy as np

from . import tubes

__all__ = ['PVSGRelationAnnotation', 'PVSGRelationDataset', 'SimpleTracker', 'get_pred_mask_tubes_one_video',
           'pred_mask_tubes_from_rows', 'calculate_iou', 'convert_to_ranges', 'find_ranges', 'match_from_counts',
           'match_and_process_gt_tubes', 'compact_matching_dict', 'translate_gt_relations', 'process_relations',
           'process_feats', 'process_pairs', 'process_feats_and_relations', 'query_feat_tubes',
           'build_relation_dict', 'label_maps_from_tubes', 'overlap_counts', 'gather_counts', 'assemble_sharded', 'load_pickle', 'save_pickle']

_SOURCES = ('vidor', 'epic_kitchen', 'ego4d')


def load_pickle(filepath):
    with open(filepath, 'rb') as f:
        return pickle.load(f)


def save_pickle(filepath, data):
    with open(filepath, 'wb') as f:
        pickle.dump(data, f)


class SimpleTracker:
    """One tube of ``query_feats.pickle`` (models/mask2former_vps/utils.py:75-89): ``qf_tube[t]`` is None or
    ``{'query_feat': float32 [256], 'cls_id': int}``."""

    def __init__(self, track_id, qf_tube):
        self.track_id = track_id
        self.qf_tube = qf_tube


def query_feat_tubes(linker):
    """TubeLinker -> the list ``concat_seq`` pickles (models/mask2former_vps/utils.py:75-89)."""
    return [SimpleTracker(tid, [frames.get(t) for t in range(linker.num_frames)])
            for tid, frames in linker.feat_tubes.items()]


# --------------------------------------------------------------------- annotation ---------
def _split_ids(anno, split):
    return [v for src in _SOURCES for v in anno['split'][src][split]]


class PVSGRelationAnnotation:
    """pvsg.json accessor (utils/relation_matching.py:15-51); ``anno_file`` may be a path or the loaded dict."""

    def __init__(self, anno_file, split='train'):
        anno = anno_file if isinstance(anno_file, dict) else json.load(open(anno_file, 'r'))
        self.video_ids = _split_ids(anno, split)
        self.classes = anno['objects']['thing'] + anno['objects']['stuff']
        self.relations = anno['relations']
        self.videos = {v['video_id']: v for v in anno['data']}

    def __getitem__(self, vid):
        assert vid in self.videos
        info = copy.deepcopy(self.videos[vid])
        objects = []
        for obj in info['objects']:
            obj['category'] = self.classes.index(obj['category'])
            objects.append(obj)
        rels = []
        for rel in info['relations']:
            if rel[2] in self.relations:
                rel[2] = self.relations.index(rel[2])
                rels.append(rel)
        return dict(video_id=vid, objects=objects, relations=rels, relation_str=self.videos[vid]['relations'])


# --------------------------------------------------------------------- masks.txt ----------
def pred_mask_tubes_from_rows(rows, decode=True):
    """rows: iterable of (frame (1-based), tube id, class id, h, w, rle string) -- the fields of a
    ``masks.txt`` line (models/unitrack/utils/io.py:14-37).  Returns ``{tube id: {'cid': str, 'mask':
    [{frame: uint8 [h,w]}, ...]}}`` exactly as utils/relation_matching.py:65-105: tubes ordered by the STRING
    tube id, class = most frequent class string of the tube.  ``decode=False`` keeps ``(h, w, rle)`` instead of
    the decoded mask (the device path never needs the per-tube masks)."""
    recs = [tuple(str(x) for x in r) for r in rows]
    order = sorted(range(len(recs)), key=lambda i: recs[i][1])      # stable, by tid string
    out = {}
    for i in order:
        fid, tid, cid, h, w, m = recs[i]
        t = out.setdefault(int(tid), dict(cid=[], mask=[]))
        t['cid'].append(cid)
        t['mask'].append({int(fid) - 1: tubes.rle_decode(m, int(h), int(w)) if decode else (int(h), int(w), m)})
    for t in out.values():
        t['cid'] = Counter(t['cid']).most_common(1)[0][0]
    return out


def get_pred_mask_tubes_one_video(vid, work_dir, decode=True):
    with open(f'{work_dir}/{vid}/quantitive/masks.txt', 'r') as f:
        return pred_mask_tubes_from_rows([line.strip().split() for line in f], decode=decode)


# --------------------------------------------------------------------- small helpers ------
def calculate_iou(gt_mask, pred_mask):
    union = np.logical_or(gt_mask, pred_mask).sum()
    return 0 if union == 0 else np.logical_and(gt_mask, pred_mask).sum() / union


def convert_to_ranges(frames):
    """utils/relation_matching.py:140-153: maximal stretches whose neighbours are <= 3 apart and that span >= 4."""
    fr = sorted(frames)
    out, start = [], fr[0]
    for prev, cur in zip(fr, fr[1:]):
        if cur - prev > 3:
            if prev - start >= 4:
                out.append([start, prev])
            start = cur
    if fr[-1] - start >= 4:
        out.append([start, fr[-1]])
    return out


def find_ranges(num_list):
    """'a-b' strings of the stretches of a sorted list whose gaps are <= 5 (utils/relation_matching.py:263-273)."""
    cuts = [0] + [i for i in range(1, len(num_list)) if num_list[i] > num_list[i - 1] + 5] + [len(num_list)]
    return [f'{num_list[a]}-{num_list[b - 1]}' for a, b in zip(cuts, cuts[1:])]


# --------------------------------------------------------------------- matching -----------
def match_from_counts(counts, frame_tube_ids, tube_cids, object_list, frame_offset=0, matching_dict=None):
    """The matching of ``match_and_process_gt_tubes`` (utils/relation_matching.py:205-260) from overlap counts.

    counts: int [T, G+1, S+1] from ``ops.tube_overlap`` (row g = GT object id g, column s = slot s);
    frame_tube_ids[t][s]: predicted tube id of slot s in frame t; tube_cids: {tube id: class (str / int)} in the
    reference's tube order (``pred_mask_tubes`` order -- string-sorted ids); object_list: GT objects with
    ``object_id`` and integer ``category``.  IoU > 0.5  <=>  2 * inter > union (integers, exact).
    Returns ``{object id: {tube id: [frames]}}`` with the reference's insertion orders."""
    md = {} if matching_dict is None else matching_dict
    counts = np.asarray(counts, dtype=np.int64)
    tube_order = {tid: i for i, tid in enumerate(tube_cids)}
    tube_cls = {tid: int(c) for tid, c in tube_cids.items()}
    row_area = counts.sum(2)
    col_area = counts.sum(1)
    for t in range(counts.shape[0]):
        slots = frame_tube_ids[t]
        if not len(slots):
            continue
        by_tube = sorted(range(len(slots)), key=lambda s: tube_order[slots[s]])
        for obj in object_list:
            g, cid = int(obj['object_id']), int(obj['category'])
            if g < 0 or g >= counts.shape[1] - 1:
                continue                      # an id the GT map cannot hold: empty mask, IoU 0
            for s in by_tube:
                tid = slots[s]
                inter = counts[t, g, s]
                if tube_cls[tid] == cid and 2 * inter > row_area[t, g] + col_area[t, s] - inter:
                    md.setdefault(g, {}).setdefault(tid, []).append(t + frame_offset)
    return md


def label_maps_from_tubes(pred_mask_tubes, num_frames, hw, max_slots=None):
    """Paint decoded tube masks (disjoint: they were cut from one panoptic map) back into per-frame label maps so
    the file-based entry point shares the device path.  Returns (pan int32 [T,H,W] holding tube ids, 0 = none;
    seg_info int32 [T, 1+4Q]; frame_tube_ids)."""
    H, W = hw
    per_frame = [[] for _ in range(num_frames)]
    pan = np.zeros((num_frames, H, W), np.int32)
    for tid, tube in pred_mask_tubes.items():
        for entry in tube['mask']:
            (f, m), = entry.items()
            if 0 <= f < num_frames and tid not in per_frame[f]:
                pan[f][np.asarray(m, bool)] = tid
                per_frame[f].append(tid)
    Q = max(1, max_slots or max(len(p) for p in per_frame))
    seg_info = np.zeros((num_frames, 1 + 4 * Q), np.int32)
    for f, tids in enumerate(per_frame):
        seg_info[f, 0] = len(tids)
        for k, tid in enumerate(tids):
            seg_info[f, 1 + 4 * k:5 + 4 * k] = (k, 0, tid, 0)
    return pan, seg_info, per_frame


def overlap_counts(gt_maps, pan_maps, seg_info, num_gt, device='cuda', chunk=64):
    """counts [T, num_gt+1, Q+1] for a clip: GT id maps and panoptic maps (numpy or torch, host or device) go
    through ``ops.tube_overlap`` in chunks of frames.  There is no CPU path."""
    import torch
    from . import ops
    T = len(gt_maps)
    out = []
    for lo in range(0, T, chunk):
        def dev(x):
            x = x[lo:lo + chunk]
            x = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
            return x.to(device=device, dtype=torch.int32).contiguous()
        out.append(ops.tube_overlap(dev(gt_maps), dev(pan_maps), dev(seg_info), num_gt).cpu())
    return torch.cat(out, 0).numpy()


def gather_counts(local_counts, num_frames, device='cpu'):
    """Multi-GPU form: ranks own contiguous frame blocks (``tubes.shard_frames``) and compute the overlap counts of
    their own frames; the small count tensors ([frames, G+1, Q+1] int32, ~13 KB per frame at G = 32) are all-gathered
    (NCCL for device tensors, gloo on CPU) so that every rank can run the matching over the whole clip.  Maps and
    masks are never exchanged.  Single process: returns the input."""
    import torch
    import torch.distributed as dist
    counts = torch.as_tensor(np.asarray(local_counts)).to(device=device, dtype=torch.int32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts.cpu().numpy()
    ws = dist.get_world_size()
    per = (num_frames + ws - 1) // ws
    pad = torch.zeros((per,) + tuple(counts.shape[1:]), dtype=torch.int32, device=device)
    pad[:counts.shape[0]] = counts
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat(parts, 0)[:num_frames].cpu().numpy()


def assemble_sharded(frame_entries, local_counts, num_frames, object_list, gt_relations, device='cpu', max_segments=100):
    """The N > 1 tail of ``end2end.relation_set_clip``: every rank passes the kept (panoptic ids, query features) of
    its own contiguous frame block and the overlap counts of those frames.  Two small all-gathers (entries:
    ``tubes.gather_and_link``; counts: ``gather_counts``) give every rank the clip-wide tubes and matching; panoptic
    maps, GT maps and masks stay on the rank that owns the frame.  Returns dict(linker, counts, frame_tube_ids,
    relation_dict), identical on every rank and identical to the single-process result."""
    linker = tubes.gather_and_link(frame_entries, num_frames, max_segments=max_segments, device=device)
    counts = gather_counts(local_counts, num_frames, device=device)
    slot_tubes = linker.frame_tube_ids()
    rd = build_relation_dict(linker, counts, slot_tubes, object_list, gt_relations)
    return dict(linker=linker, counts=counts, frame_tube_ids=slot_tubes, relation_dict=rd)


def match_and_process_gt_tubes(vid, pvsg_dataset, pred_mask_tubes, data_dir='./data', gt_maps=None, device='cuda'):
    """Reference signature (utils/relation_matching.py:205-260).  GT id maps are read from
    ``{data_dir}/{source}/masks/{vid}/*.png`` unless passed as ``gt_maps`` [T,H,W]."""
    if gt_maps is None:
        from pathlib import Path
        from PIL import Image
        head = vid.split('_')[0]
        source = 'epic_kitchen' if vid.startswith('P') else 'vidor' if head.isdigit() and len(head) == 4 else 'ego4d'
        paths = sorted(Path(os.path.join(data_dir, source, 'masks', vid)).rglob('*.png'))
        gt_maps = np.stack([np.array(Image.open(p)) for p in paths]).astype(np.int32)
    gt_maps = np.asarray(gt_maps)
    object_list = pvsg_dataset[vid]['objects']
    T, H, W = gt_maps.shape
    pan, seg_info, per_frame = label_maps_from_tubes(pred_mask_tubes, T, (H, W))
    num_gt = max([int(o['object_id']) for o in object_list] + [0]) + 1
    counts = overlap_counts(gt_maps, pan, seg_info, num_gt, device=device)
    return match_from_counts(counts, per_frame, {t: v['cid'] for t, v in pred_mask_tubes.items()}, object_list)


def compact_matching_dict(matching_dict):
    """utils/relation_matching.py:276-299: drop tube matches shorter than 5 frames; a GT object matched by a
    single tube keeps one 'min-max' string, otherwise each tube keeps its list of gap-split range strings."""
    out = {}
    for obj_id, per_tube in matching_dict.items():
        kept = {}
        for tid, frames in per_tube.items():
            if len(frames) < 5:
                continue
            kept[tid] = f'{min(frames)}-{max(frames)}' if len(per_tube) == 1 else find_ranges(sorted(frames))
        if kept:
            out[obj_id] = kept
    return out


def _ranges(spec):
    for r in ([spec] if isinstance(spec, str) else spec):
        a, b = r.split('-')
        yield int(a), int(b) + 1          # inclusive frame range -> half-open


def translate_gt_relations(matching_dict, gt_relations):
    """GT relations [subject object id, object object id, label, [[start, end), ...]] -> relations between
    predicted tubes, clipped to the frames where both matches hold (utils/relation_matching.py:302-370).
    Returns ``[tube_s, tube_o, label, [[start, end], ...]]`` merged per (tube_s, tube_o, label)."""
    merged = {}
    for sub, obj, label, spans in gt_relations:
        if sub not in matching_dict or obj not in matching_dict:
            continue
        for lo, hi in spans:
            for tid_s, spec_s in matching_dict[sub].items():
                for a1, b1 in _ranges(spec_s):
                    for tid_o, spec_o in matching_dict[obj].items():
                        for a2, b2 in _ranges(spec_o):
                            start, end = max(lo, a1, a2), min(hi, b1, b2)
                            if start < end:
                                merged.setdefault((tid_s, tid_o, label), []).append([start, end])
    return [[s, o, l, v] for (s, o, l), v in merged.items()]


# --------------------------------------------------------------------- samples ------------
def _video_length(pred_feat_tubes):
    return len(next(iter(pred_feat_tubes.values())))


def _present(qf_tube):
    return np.fromiter((e is not None for e in qf_tube), bool, len(qf_tube))


def _dense(qf_tube, d):
    out = np.zeros([len(qf_tube), d])
    for t, e in enumerate(qf_tube):
        if e is not None:
            out[t] = e['query_feat']
    return out


def _span(time_span, length, *tubes_):
    span = np.zeros(length)
    for a, b in time_span:
        span[a:b] = 1        # an out-of-range start raises in the reference; spans come from matched frames
    for tube in tubes_:
        span[~_present(tube)] = 0
    return span


def process_feats(pred_feat_tubes, d=256):
    """{tube id: float64 [T,d]}, zero rows where the tube is absent (utils/relation_matching.py:431-442)."""
    return {tid: _dense(tube, d) for tid, tube in pred_feat_tubes.items()}


def process_pairs(pred_relations):
    return [[r[0], r[1]] for r in pred_relations]


def process_relations(pred_relations, pred_feat_tubes, d=256):
    """Per relation: dense subject / object features + 0/1 span; spans shorter than 3 frames are dropped
    (utils/relation_matching.py:373-428)."""
    T = _video_length(pred_feat_tubes)
    out = []
    for s, o, relation, time_span in pred_relations:
        span = _span(time_span, T, pred_feat_tubes[s], pred_feat_tubes[o])
        if span.sum() >= 3:
            out.append(dict(relation=relation, tube_s=_dense(pred_feat_tubes[s], d), tube_o=_dense(pred_feat_tubes[o], d),
                            relation_span=span))
    return out


def process_feats_and_relations(pred_relations, pred_feat_tubes, d=256):
    """The ``relations.pickle`` payload (utils/relation_matching.py:452-486)."""
    T = _video_length(pred_feat_tubes)
    rels = []
    for s, o, relation, time_span in pred_relations:
        span = _span(time_span, T, pred_feat_tubes[s], pred_feat_tubes[o])
        if span.sum() >= 3:
            rels.append(dict(subject_index=s, object_index=o, relation=relation, relation_span=span))
    return dict(feats=process_feats(pred_feat_tubes, d), relations=rels)


def build_relation_dict(linker, counts, frame_tube_ids, object_list, gt_relations):
    """In-memory ``tools/prepare_rel_set.py:24-52`` for one video: TubeLinker (tubes + features) and the clip's
    overlap counts -> the ``relations.pickle`` dictionary."""
    order = sorted(linker.feat_tubes, key=str)               # the reference's pred_mask_tubes order
    cids = {}
    for tid in order:
        c = Counter(str(e['cls_id']) for _, e in sorted(linker.feat_tubes[tid].items()))
        cids[tid] = c.most_common(1)[0][0]
    matching = compact_matching_dict(match_from_counts(counts, frame_tube_ids, cids, object_list))
    pred_relations = translate_gt_relations(matching, gt_relations)
    feat_tubes = {t.track_id: t.qf_tube for t in query_feat_tubes(linker)}
    return process_feats_and_relations(pred_relations, feat_tubes)


# --------------------------------------------------------------------- dataset ------------
class PVSGRelationDataset:
    """datasets/datasets/pvsg_relation.py:15-79.  ``memory`` = {video id: relations dict} serves samples without
    the ``relations.pickle`` round trip; otherwise the files under ``work_dir`` are read as the reference does."""

    def __init__(self, anno_file, split='train', work_dir='./work_dirs/train_save_qf_1106', return_mask=False,
                 memory=None):
        anno = anno_file if isinstance(anno_file, dict) else json.load(open(anno_file, 'r'))
        self.video_ids = _split_ids(anno, split)
        self.work_dir = work_dir
        self.split = split
        self.classes = anno['objects']['thing'] + anno['objects']['stuff']
        self.relations = anno['relations']
        self.return_mask = return_mask
        self.videos = {v['video_id']: v for v in anno['data']}
        self.memory = memory

    def __len__(self):
        return len(self.video_ids)

    def __getitem__(self, index):
        vid = self.video_ids[index]
        if self.memory is not None:
            sample = copy.deepcopy(self.memory[vid])
        else:
            sample = load_pickle(os.path.join(self.work_dir, vid, 'relations.pickle'))
        sample['vid'] = vid
        keys = list(sample['feats'])
        index_of = {k: i for i, k in enumerate(keys)}
        sample['feats'] = np.array([sample['feats'][k] for k in keys])
        for rel in sample['relations']:
            rel['subject_index'] = index_of[rel['subject_index']]
            rel['object_index'] = index_of[rel['object_index']]
        sample['pairs'] = [[r['subject_index'], r['object_index']] for r in sample['relations']]
        if self.return_mask:
            sample['idx2key'] = dict(enumerate(keys))
            tubes_ = get_pred_mask_tubes_one_video(vid, self.work_dir)
            sample['masks'] = [tubes_.get(k, {}) for k in keys]
        return sample
