"""Relation-set builder: host logic against the reference's own outputs
(tests/golden/relset.json, produced by utils/relation_matching.py + datasets/datasets/pvsg_relation.py run
unmodified -- tests/golden/make_golden_relset.py) and the device overlap kernel against the oracle."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

import relset_fixture as fx
from openpvsg_b200 import relation_set as rs
from openpvsg_b200 import tubes
from oracle import relset as orl

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_GT = 6


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(HERE, 'golden', 'relset.json')))


@pytest.fixture(scope='module')
def clip(golden):
    c = fx.make_clip()
    cs = float(np.abs(np.concatenate([f.ravel() for f in c['feats']])).sum() + c['gt'].sum() + c['pan'].sum())
    assert cs == pytest.approx(golden['checksum'], rel=1e-9), 'fixture RNG drifted: regenerate the golden file'
    c['linker'] = fx.link(c)
    return c


def _pairs(d):
    return [[k, [[kk, vv] for kk, vv in v.items()]] for k, v in d.items()]


def _check_flow(golden, clip, counts):
    """counts (from the oracle on CPU, from pvsg_tube_overlap on the GPU) -> every structure of the reference flow."""
    linker = clip['linker']
    anno = rs.PVSGRelationAnnotation(fx.make_anno(), 'train')
    info = anno[fx.VID]
    assert json.loads(json.dumps(info)) == golden['annotation']
    cids = {tid: t['cid'] for tid, t in rs.pred_mask_tubes_from_rows(linker.rows, decode=False).items()}
    matching = rs.match_from_counts(counts, fx.frame_tube_ids(clip, linker), cids, info['objects'])
    assert _pairs(matching) == golden['matching']
    compact = rs.compact_matching_dict(matching)
    assert _pairs(compact) == golden['compact']
    pred_relations = rs.translate_gt_relations(compact, info['relations'])
    assert pred_relations == golden['pred_relations']
    assert rs.process_pairs(pred_relations) == golden['pairs']
    feat_tubes = {t.track_id: t.qf_tube for t in rs.query_feat_tubes(linker)}
    rd = rs.process_feats_and_relations(pred_relations, feat_tubes)
    g = golden['relation_dict']
    assert [int(k) for k in rd['feats']] == g['feat_keys']
    assert str(next(iter(rd['feats'].values())).dtype) == g['feat_dtype']
    assert [float(np.abs(v).sum()) for v in rd['feats'].values()] == pytest.approx(g['feat_sums'], rel=1e-12)
    assert len(rd['relations']) == len(g['relations'])
    for a, b in zip(rd['relations'], g['relations']):
        assert (a['subject_index'], a['object_index'], a['relation']) == (b['subject_index'], b['object_index'], b['relation'])
        assert a['relation_span'].tolist() == b['relation_span']
    full = rs.process_relations(pred_relations, feat_tubes)
    assert len(full) == len(golden['relations_full'])
    for a, b in zip(full, golden['relations_full']):
        assert a['relation'] == b['relation'] and a['relation_span'].tolist() == b['span']
        assert float(np.abs(a['tube_s']).sum()) == pytest.approx(b['s_sum'], rel=1e-12)
        assert float(np.abs(a['tube_o']).sum()) == pytest.approx(b['o_sum'], rel=1e-12)
    # the one-call form and the in-memory dataset
    rd2 = rs.build_relation_dict(linker, counts, fx.frame_tube_ids(clip, linker), info['objects'], info['relations'])
    ds = rs.PVSGRelationDataset(fx.make_anno(), 'train', memory={fx.VID: rd2})
    assert len(ds) == 1
    sample = ds[0]
    gs = golden['sample']
    assert sample['vid'] == gs['vid'] and list(sample['feats'].shape) == gs['feats_shape']
    assert float(np.abs(sample['feats']).sum()) == pytest.approx(gs['feats_sum'], rel=1e-12)
    assert sample['pairs'] == gs['pairs']
    for a, b in zip(sample['relations'], gs['relations']):
        assert (a['subject_index'], a['object_index'], a['relation']) == (b['subject_index'], b['object_index'], b['relation'])
        assert a['relation_span'].tolist() == b['relation_span']
    ds[0]   # memory samples are not consumed by a read


def test_flow_matches_reference_cpu(golden, clip):
    counts = orl.joint_histogram(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT)
    _check_flow(golden, clip, counts)


def test_flow_matches_reference_cpu_gaps_variant():
    """Edge cases of the matching / compaction rules (tests/relset_fixture.py variant 'gaps'): frames without any
    kept segment, a prediction that comes and goes, a tube matched on fewer than 5 frames (dropped), a GT object
    that is never predicted, relations whose span shrinks below 3 frames (dropped) -- against the reference's outputs."""
    golden = json.load(open(os.path.join(HERE, 'golden', 'relset_gaps.json')))
    c = fx.make_clip(variant='gaps')
    cs = float(np.abs(np.concatenate([f.ravel() for f in c['feats']])).sum() + c['gt'].sum() + c['pan'].sum())
    assert cs == pytest.approx(golden['checksum'], rel=1e-9)
    assert any(len(ids) == 0 for ids in c['seg_ids'])
    c['linker'] = fx.link(c)
    counts = orl.joint_histogram(c['gt'], c['pan'], c['seg_info'], NUM_GT)
    _check_flow(golden, c, counts)
    tubes_ = rs.pred_mask_tubes_from_rows(c['linker'].rows)
    got = [[tid, v['cid'], [list(m.keys())[0] for m in v['mask']], [int(list(m.values())[0].sum()) for m in v['mask']]]
           for tid, v in tubes_.items()]
    assert got == golden['pred_mask_tubes']
    assert len(golden['relation_dict']['relations']) < len(golden['pred_relations'])      # the >= 3 frames rule fired
    assert len(golden['compact']) < len(golden['matching'])                              # the >= 5 frames rule fired


def test_oracle_counts_are_the_reference_ious(clip):
    """Pins the oracle's histogram to calculate_iou on decoded masks (what the reference evaluates)."""
    counts = orl.joint_histogram(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT).astype(np.int64)
    for t in (0, 7, 15, 22, 39):
        ids = orl.slot_ids(clip['seg_info'][t])
        for g in range(1, NUM_GT):
            for s, seg in enumerate(ids):
                gm, pm = clip['gt'][t] == g, clip['pan'][t] == seg
                inter = counts[t, g, s]
                union = counts[t, g].sum() + counts[t, :, s].sum() - inter
                assert inter == np.logical_and(gm, pm).sum() and union == np.logical_or(gm, pm).sum()
                assert (2 * inter > union) == (orl.iou_from_masks(gm, pm) > 0.5) == (rs.calculate_iou(gm, pm) > 0.5)


def test_masks_txt_reader_and_files(golden, clip, tmp_path):
    linker = clip['linker']
    work = tmp_path / 'work'
    (work / fx.VID / 'quantitive').mkdir(parents=True)
    (work / fx.VID / 'quantitive' / 'masks.txt').write_text(linker.masks_txt())
    t = rs.get_pred_mask_tubes_one_video(fx.VID, str(work))
    got = [[tid, v['cid'], [list(m.keys())[0] for m in v['mask']], [int(list(m.values())[0].sum()) for m in v['mask']]]
           for tid, v in t.items()]
    assert got == golden['pred_mask_tubes']
    # painting the decoded tubes back gives label maps whose overlap counts reproduce the matching
    pan, seg_info, per_frame = rs.label_maps_from_tubes(t, clip['T'], (clip['H'], clip['W']))
    counts = orl.joint_histogram(clip['gt'], pan, seg_info, NUM_GT)
    md = rs.match_from_counts(counts, per_frame, {k: v['cid'] for k, v in t.items()},
                              rs.PVSGRelationAnnotation(fx.make_anno())[fx.VID]['objects'])
    assert _pairs(md) == golden['matching']
    # file-based dataset (relations.pickle + return_mask) as the reference reads it
    counts0 = orl.joint_histogram(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT)
    info = rs.PVSGRelationAnnotation(fx.make_anno())[fx.VID]
    rd = rs.build_relation_dict(linker, counts0, fx.frame_tube_ids(clip, linker), info['objects'], info['relations'])
    rs.save_pickle(str(work / fx.VID / 'relations.pickle'), rd)
    with open(work / fx.VID / 'query_feats.pickle', 'wb') as f:
        pickle.dump(rs.query_feat_tubes(linker), f)
    sample = rs.PVSGRelationDataset(fx.make_anno(), 'train', str(work), return_mask=True)[0]
    gs = golden['sample']
    assert [[int(k), int(v)] for k, v in sample['idx2key'].items()] == gs['idx2key']
    assert [[list(m.keys())[0] for m in tube.get('mask', [])] for tube in sample['masks']] == gs['mask_frames']
    assert sample['pairs'] == gs['pairs']


def test_range_helpers(golden):
    for frames, want in golden['convert_to_ranges']:
        assert rs.convert_to_ranges(frames) == want
    for frames, want in golden['find_ranges']:
        assert rs.find_ranges(frames) == want


def test_no_cpu_path_for_counts(clip):
    from openpvsg_b200.lib import PvsgError
    with pytest.raises(PvsgError):
        rs.overlap_counts(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT, device='cpu')


class _FakeDetector:
    def parameters(self):
        yield torch.zeros(1)


def _mock_device_side(monkeypatch_setattr, c, lo, hi):
    """end2end.relation_set_clip with the device side replaced: frames come from the fixture, the overlap counts from
    the oracle -- what remains is exactly the host logic of the function (hand-over, slot order, batching, gathers)."""
    from openpvsg_b200 import end2end, ops

    def fake_vps_clip(det, frames, meta, batch, consume=None, rle=True):
        for t in range(lo, hi):
            consume(dict(pan_results=c['pan'][t],
                         query_feats={int(s): [torch.as_tensor(c['feats'][t][k])] for k, s in enumerate(c['seg_ids'][t])}))

    monkeypatch_setattr(end2end, 'vps_clip', fake_vps_clip)
    monkeypatch_setattr(ops, 'tube_overlap',
                        lambda gt, pan, si, n: torch.as_tensor(orl.joint_histogram(gt.numpy(), pan.numpy(), si.numpy(), n)))
    return end2end


def _relations(rd):
    return [(r['subject_index'], r['object_index'], r['relation'], np.asarray(r['relation_span']).tolist()) for r in rd['relations']]


def test_relation_set_clip_host_logic(golden, clip, monkeypatch):
    e2e = _mock_device_side(monkeypatch.setattr, clip, 0, clip['T'])
    info = rs.PVSGRelationAnnotation(fx.make_anno())[fx.VID]
    out = e2e.relation_set_clip(_FakeDetector(), None, None, clip['gt'], info['objects'], info['relations'], batch=7,
                                max_segments=fx.Q)
    assert _relations(out['relation_dict']) == _relations(golden['relation_dict'])
    assert out['counts'].shape == (clip['T'], NUM_GT + 1, fx.Q + 1)
    assert out['frame_tube_ids'] == out['linker'].frame_tube_ids() == fx.frame_tube_ids(clip, clip['linker'])


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    c = fx.make_clip()
    lo, hi = tubes.shard_frames(c['T'], world, rank)
    local = orl.joint_histogram(c['gt'][lo:hi], c['pan'][lo:hi], c['seg_info'][lo:hi], NUM_GT)   # this rank's frames only
    entries = [(c['seg_ids'][t], c['feats'][t]) for t in range(lo, hi)]                          # and its kept entries
    info = rs.PVSGRelationAnnotation(fx.make_anno())[fx.VID]
    out = rs.assemble_sharded(entries, local, c['T'], info['objects'], info['relations'], max_segments=fx.Q)
    counts, rd = out['counts'], out['relation_dict']
    assert out['frame_tube_ids'] == fx.frame_tube_ids(c, fx.link(c))
    # the same through end2end.relation_set_clip's distributed branch (device side mocked)
    e2e = _mock_device_side(setattr, c, lo, hi)
    out2 = e2e.relation_set_clip(_FakeDetector(), None, None, c['gt'][lo:hi], info['objects'], info['relations'], batch=7,
                                 max_segments=fx.Q, num_frames=c['T'])
    assert _relations(out2['relation_dict']) == _relations(rd) and np.array_equal(out2['counts'], counts)
    q.put((rank, counts.shape, [(r['subject_index'], r['object_index'], r['relation'], r['relation_span'].tolist())
                                for r in rd['relations']]))
    dist.destroy_process_group()


def test_sharded_counts_world2_gloo(golden):
    """N > 1 (relation_set.assemble_sharded, the tail of end2end.relation_set_clip): every rank passes the kept entries
    and the overlap counts of its own frame block; two all-gathers (entries, counts -- never maps) give every rank the
    reference's relation set."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want = [(r['subject_index'], r['object_index'], r['relation'], r['relation_span'])
            for r in golden['relation_dict']['relations']]
    for rank, shape, rels in got:
        assert tuple(shape) == (40, NUM_GT + 1, fx.Q + 1)
        assert rels == want


# ------------------------------------------------------------------------------ GPU ----------
@pytest.mark.gpu
def test_tube_overlap_kernel_bit_exact(golden, clip):
    from openpvsg_b200 import ops
    dev = 'cuda'
    want = orl.joint_histogram(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT)
    got = ops.tube_overlap(torch.as_tensor(clip['gt']).to(dev), torch.as_tensor(clip['pan']).to(dev),
                           torch.as_tensor(clip['seg_info']).to(dev), NUM_GT).cpu().numpy()
    assert np.array_equal(got, want)
    _check_flow(golden, clip, rs.overlap_counts(clip['gt'], clip['pan'], clip['seg_info'], NUM_GT, chunk=16))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1, 1, 1), (3, 7, 13), (2, 33, 130), (5, 61, 67)])
def test_tube_overlap_ragged_shapes(shape):
    """Odd sizes (scalar path, partial strips), GT ids out of range, empty seg_info, duplicate stuff rows."""
    from openpvsg_b200 import ops
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H)
    Qn, G = 9, 5
    gt = rng.integers(-1, G + 2, (B, H, W)).astype(np.int32)
    pan = rng.choice([126, 3, 3, 1005, 2005, 40], (B, H, W)).astype(np.int32)
    seg_info = np.zeros((B, 1 + 4 * Qn), np.int32)
    for b in range(B):
        rows = [] if b == 1 else [(0, 3, 3, 1), (1, 5, 1005, 1), (2, 3, 3, 1), (3, 7, -1, 0), (4, 5, 2005, 1)]
        seg_info[b, 0] = len(rows)
        for k, r in enumerate(rows):
            seg_info[b, 1 + 4 * k:5 + 4 * k] = r
    want = orl.joint_histogram(gt, pan, seg_info, G)
    got = ops.tube_overlap(torch.as_tensor(gt).cuda(), torch.as_tensor(pan).cuda(), torch.as_tensor(seg_info).cuda(), G)
    assert np.array_equal(got.cpu().numpy(), want)
    assert got.sum().item() == B * H * W


@pytest.mark.gpu
def test_tube_overlap_full_size_properties():
    """720p x 20 frames, 100 segment slots, 255 GT ids: totals, marginals and one frame against the oracle."""
    from openpvsg_b200 import ops
    B, H, W, Qn, G = 20, 720, 1280, 100, 255
    g = torch.Generator().manual_seed(3)
    # blocky label maps (spatially coherent, like real segmentations)
    gt = torch.randint(0, G, (B, H // 16, W // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2).int()
    slot = torch.randint(0, 60, (B, H // 8, W // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    ids = (torch.arange(100) % 127 + 1000 * (torch.arange(100) // 3)).int()
    pan = ids[slot].int()
    seg_info = torch.zeros(B, 1 + 4 * Qn, dtype=torch.int32)
    seg_info[:, 0] = 50                                       # slots 50..59 are painted but not kept
    for k in range(50):
        seg_info[:, 1 + 4 * k + 2] = ids[k]
    got = ops.tube_overlap(gt.cuda(), pan.cuda(), seg_info.cuda(), G).cpu()
    assert got.sum(dim=(1, 2)).tolist() == [H * W] * B
    for b in (0, 19):
        assert torch.equal(got[b].sum(1)[:G], torch.bincount(gt[b].flatten(), minlength=G).int())
        col = torch.bincount(slot[b].flatten(), minlength=60)
        assert torch.equal(got[b].sum(0)[:50], col[:50].int()) and got[b].sum(0)[Qn].item() == col[50:].sum().item()
    want = orl.joint_histogram(gt[:1].numpy(), pan[:1].numpy(), seg_info[:1].numpy(), G)
    assert np.array_equal(got[:1].numpy(), want)


@pytest.mark.gpu
def test_match_and_process_gt_tubes_reference_signature(golden, clip, tmp_path):
    """The reference call (PNG ground truth on disk + decoded masks.txt tubes) through the device path."""
    from PIL import Image
    d = tmp_path / 'data' / 'vidor' / 'masks' / fx.VID
    d.mkdir(parents=True)
    for t in range(clip['T']):
        Image.fromarray(clip['gt'][t].astype(np.uint8)).save(d / f'{t:04d}.png')
    pred = rs.pred_mask_tubes_from_rows(clip['linker'].rows)
    md = rs.match_and_process_gt_tubes(fx.VID, rs.PVSGRelationAnnotation(fx.make_anno()), pred,
                                       data_dir=str(tmp_path / 'data'))
    assert _pairs(md) == golden['matching']
