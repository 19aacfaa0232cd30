"""End-to-end PVSG inference on one clip: VPS forward -> tube linking -> relation head.

The reference splits this over three tools with files in between
(tools/prepare_query_tube_vps.py -> work_dirs/.../{quantitive/masks.txt, query_feats.pickle};
utils/relation_matching.py:431-442 / datasets/datasets/pvsg_relation.py:47-53 build the zero-filled
[N_tubes, T, 256] tube features; tools/rel_test.py:35-67 runs the relation head) and leaves
tools/end2end_inference.py empty.  Here the stages hand over in memory:

  frames --FrameRunner (8 frames per graph replay, device RLE events)--> per-frame results
         --TubeLinker (concat_seq semantics: tube id = first appearance of a panoptic id)-->
  tube features [N, T, 256] --relation_forward--> top pairs, span / relation scores --> triplets

Multi-GPU: ranks own contiguous frame blocks (tubes.shard_frames); the only exchange is the
all-gather of the kept (segment id, query feature) entries (tubes.gather_and_link); the relation
stage then runs on every rank (it is ~2 ms).
"""
import numpy as np
import torch

from . import engine, relation_head as rh, relation_set, tubes


@torch.no_grad()
def vps_clip(detector, frames, meta, batch=8, rle=True, consume=None, debug_masks=False):
    """frames: list of [3,H,W] tensors (pinned host or device).  Returns the per-frame result dicts
    (reference format, plus 'rle' strings) of this rank's frames, pipelined through the runner.
    consume(result): called per frame as soon as its batch is collected; the results are then views
    of the runner's pinned ring (no 13 MB-per-frame host copies) and nothing is retained."""
    if getattr(detector, '_runners', None) is None:
        engine.enable_cuda_graph(detector)
    runner = engine.get_runner(detector, meta, True, batch=batch, rle=rle, debug_masks=debug_masks)
    results, pend = [], None

    def drain(p):
        for r in runner.collect(p, copy=consume is None):
            if consume is None:
                results.append(r)
            else:
                consume(r)

    for i in range(0, len(frames), batch):
        nxt = runner.submit(frames[i:i + batch])
        if pend is not None:
            drain(pend)
        pend = nxt
    if pend is not None:
        drain(pend)
    return results


def link_tubes(results, num_frames=None, device='cpu'):
    """concat_seq over the clip.  Single process: local linking with masks.txt rows.  Under
    torch.distributed every rank passes the results of its own frame block; the kept entries are
    all-gathered and every rank gets the same tubes (mask rows stay with the owning rank)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        entries = []
        for r in results:
            ids = list(r['query_feats'].keys())
            feats = np.stack([np.asarray(torch.as_tensor(r['query_feats'][k][0]).cpu()) for k in ids]) if ids \
                else np.zeros((0, 256), np.float32)
            entries.append((ids, feats))
        return tubes.gather_and_link(entries, num_frames, device=device)
    return tubes.concat_seq([[r] for r in results])


@torch.no_grad()
def relations(linker, models, num_top_pairs=100, device='cuda'):
    """models: (subject_encoder, object_encoder, pair_proposal_model, relation_model).
    Returns (results list as generate_results / test_utils.py:59-84, raw forward dict)."""
    feats = torch.as_tensor(linker.tube_features()).to(device)
    if feats.shape[0] < 2:
        return [], None
    out = rh.relation_forward(*models, feats, num_top_pairs)
    res = rh.generate_results(out['span_pred'], out['prob'], out['pairs'].cpu().tolist())
    return res, out


def _add_result(linker, r):
    ids = list(r['query_feats'].keys())
    feats = [np.asarray(torch.as_tensor(r['query_feats'][k][0]).cpu()) for k in ids]
    if 'rle' in r:
        linker.add_frame(ids, feats, rle=r['rle'], hw=r['pan_results'].shape)
    else:
        linker.add_frame(ids, feats, r.get('pan_results'))


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


@torch.no_grad()
def run_clip(detector, models, frames, meta, batch=8, num_top_pairs=100, keep_results=False, num_frames=None,
             debug_masks=False):
    """The whole path for one clip.  Returns dict(results, linker, relations, raw).
    With keep_results=False (default) frames are linked as their batch completes and the per-frame
    panoptic maps are not retained (the tube wire format carries the masks as RLE rows).
    Under torch.distributed (world size > 1) ``frames`` is this rank's contiguous block of the clip
    (``tubes.shard_frames``) and ``num_frames`` the clip length: the kept (segment id, query feature) entries are
    all-gathered once (``tubes.gather_and_link``), every rank links the same clip-wide tubes and runs the ~1 ms
    relation stage on them; ``linker.rows`` (masks.txt) holds only the rows of this rank's frames."""
    device = next(detector.parameters()).device
    if _world() > 1:
        if num_frames is None:
            raise ValueError('run_clip: num_frames (clip length) is required when the clip is sharded over ranks')
        entries, rows = [], []

        def consume(r):
            ids = [int(k) for k in r['query_feats'].keys()]
            feats = np.stack([np.asarray(torch.as_tensor(r['query_feats'][k][0]).cpu()) for k in ids]) if ids \
                else np.zeros((0, 256), np.float32)
            entries.append((ids, feats))
            if 'rle' in r:
                rows.append((ids, r['rle'], r['pan_results'].shape))

        results = None
        if keep_results:
            results = vps_clip(detector, frames, meta, batch, debug_masks=debug_masks)
            for r in results:
                consume(r)
        else:
            vps_clip(detector, frames, meta, batch, consume=consume)
        linker = tubes.gather_and_link(entries, num_frames, device=device)
        import torch.distributed as dist
        lo, _ = tubes.shard_frames(num_frames, dist.get_world_size(), dist.get_rank())
        for f, (ids, rle, hw) in enumerate(rows):       # masks.txt rows of the frames this rank owns, clip-wide ids
            linker.rows.extend((lo + f + 1, linker._tube_of[i], i % 1000, hw[0], hw[1], rle[i]) for i in ids)
    elif keep_results:
        results = vps_clip(detector, frames, meta, batch, debug_masks=debug_masks)
        linker = link_tubes(results, len(frames))
    else:
        results, linker = None, tubes.TubeLinker()
        vps_clip(detector, frames, meta, batch, consume=lambda r: _add_result(linker, r))
    rel, raw = relations(linker, models, num_top_pairs, device=device)
    return dict(results=results, linker=linker, relations=rel, raw=raw)


@torch.no_grad()
def relation_set_clip(detector, frames, meta, gt_maps, object_list, gt_relations, batch=8, max_segments=100, num_frames=None):
    """tools/prepare_query_tube_vps.py:170-240 + tools/prepare_rel_set.py:24-52 for one video, in memory: VPS
    forward -> tube linking -> overlap of every frame's panoptic map with its ground-truth instance-id map
    (``ops.tube_overlap``, one device pass per batch of frames) -> matched tubes -> the ``relations.pickle``
    dictionary that ``relation_set.PVSGRelationDataset(memory=...)`` serves.  gt_maps: [T,H,W] integer maps at the
    frames' ``ori_shape``; object_list / gt_relations as ``PVSGRelationAnnotation.__getitem__`` returns them.
    Under torch.distributed (world size > 1) ``frames`` / ``gt_maps`` are this rank's contiguous block
    (``tubes.shard_frames``) and ``num_frames`` the clip length: the kept entries and the overlap counts are
    all-gathered (``relation_set.assemble_sharded``) and every rank returns the clip-wide result.
    Returns dict(linker, counts [T,G+1,Q+1], frame_tube_ids, relation_dict)."""
    from . import ops
    device = next(detector.parameters()).device
    linker = tubes.TubeLinker()
    num_gt = max([int(o['object_id']) for o in object_list] + [0]) + 1
    counts, slot_tubes, held = [], [], []

    def flush():
        if not held:
            return
        lo = len(counts)
        pan = torch.stack([h[0] for h in held])
        seg_info = torch.zeros(len(held), 1 + 4 * max_segments, dtype=torch.int32)
        for b, (_, ids) in enumerate(held):
            seg_info[b, 0] = len(ids)
            for k, seg in enumerate(ids):
                seg_info[b, 3 + 4 * k] = seg
        gt = torch.as_tensor(np.ascontiguousarray(gt_maps[lo:lo + len(held)])).to(device=device, dtype=torch.int32)
        counts.extend(ops.tube_overlap(gt, pan, seg_info.to(device), num_gt).cpu().numpy())
        held.clear()

    def consume(r):
        _add_result(linker, r)
        ids = [int(k) for k in r['query_feats'].keys()]
        slot_tubes.append([linker._tube_of[i] for i in ids])
        held.append((torch.as_tensor(r['pan_results']).to(device=device, dtype=torch.int32), ids))
        if len(held) == batch:
            flush()

    vps_clip(detector, frames, meta, batch, consume=consume)
    flush()
    counts = np.stack(counts) if counts else np.zeros((0, num_gt + 1, max_segments + 1), np.int32)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        entries = [(ids, np.stack([linker.feat_tubes[linker._tube_of[i]][t]['query_feat'] for i in ids])
                    if ids else np.zeros((0, 256), np.float32)) for t, ids in enumerate(linker.frame_seg_ids)]
        return relation_set.assemble_sharded(entries, counts, num_frames, object_list, gt_relations, device=device,
                                             max_segments=max_segments)
    rd = relation_set.build_relation_dict(linker, counts, slot_tubes, object_list, gt_relations)
    return dict(linker=linker, counts=counts, frame_tube_ids=slot_tubes, relation_dict=rd)
