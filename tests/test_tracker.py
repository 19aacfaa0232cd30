"""IPS tracker path: golden vectors from the reference's own tracker code
(tests/golden/make_golden_tracker.py -> tracker.json), the oracle's numerics, the product's host state machine and
(GPU) the device kernels pvsg_reconsdot / pvsg_lap_assign."""
import json
import os
import types

import numpy as np
import pytest
import torch

import tracker_fixture as fx
from openpvsg_b200 import tracker as trk
from openpvsg_b200.registry import to_cfg
from oracle import tracker as otr

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(HERE, 'golden', 'tracker.json')))


def _inf(c):
    c = np.asarray(c, np.float64)
    return np.where(c < 0, np.inf, c)


def test_oracle_numerics_match_reference(golden):
    t, d = fx.embedding_sets(seed=1)
    np.testing.assert_allclose(otr.reconsdot_distance(t, d), np.array(golden['reconsdot']['cost']), atol=1e-6)
    for case in golden['lap']:
        m, ua, ub = otr.linear_assignment(_inf(case['cost']), case['thresh'])
        assert np.asarray(m).tolist() == case['matches'] and list(ua) == case['unmatched_a'] and list(ub) == case['unmatched_b']
    a, b = fx.boxes(seed=4, n=6), fx.boxes(seed=5, n=5)
    np.testing.assert_allclose(otr.iou_distance(a, b), np.array(golden['iou_distance']), atol=1e-12)
    np.testing.assert_allclose(trk.iou_distance(list(a), list(b)), np.array(golden['iou_distance']), atol=1e-12)


def test_kalman_filter_matches_reference(golden):
    kf = trk.KalmanFilter()
    mean, cov = kf.initiate(np.array([50., 40., 0.5, 80.]))
    for step in golden['kalman']:
        mean, cov = kf.predict(mean, cov)
        gd = kf.gating_distance(mean, cov, np.array([step['z'], [0., 0., 1., 10.]]), metric='maha')
        mean, cov = kf.update(mean, cov, np.array(step['z']))
        np.testing.assert_allclose(gd, step['gating'], rtol=1e-10)
        np.testing.assert_allclose(mean, step['mean'], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(cov, step['cov'], rtol=1e-9, atol=1e-12)
    # multi_predict == predict per track
    m2 = np.stack([mean, mean * 1.1])
    c2 = np.stack([cov, cov * 1.2])
    mm, cc = kf.multi_predict(m2, c2)
    for i in range(2):
        a, b = kf.predict(m2[i], c2[i])
        np.testing.assert_allclose(mm[i], a, rtol=1e-12)
        np.testing.assert_allclose(cc[i], b, rtol=1e-12)


class _Tracker(trk.AssociationTracker):
    def prepare_obs(self, img, img0, obs, embs=None):
        return [trk.STrack(tlwh, 1, f, self.buffer_size, None, ac=True) for tlwh, f in obs]


class _Obs(list):
    @property
    def shape(self):
        return (len(self), 5)


def _run_clip(use_kalman):
    trk.BaseTrack.reset_count()
    cfg = to_cfg(fx.tracker_cfg())
    cfg.mots.use_kalman = use_kalman
    tracker = _Tracker(cfg)
    frames = []
    for obs, query_feats in fx.clip(seed=7):
        dets = [(tlwh, f) for tlwh, f, _ in obs]
        if not dets:
            frames.append(dict(ids=[], tlwh=[], num_tubes=len(tracker.query_feat_tubes)))
            continue
        online, n_tubes = tracker.update(None, None, _Obs(dets), query_feats, 0)
        frames.append(dict(ids=[int(t.track_id) for t in online], tlwh=[np.asarray(t.tlwh).tolist() for t in online],
                           num_tubes=int(n_tubes), lost=[int(t.track_id) for t in tracker.lost_stracks],
                           removed=[int(t.track_id) for t in tracker.removed_stracks]))
    tubes = [dict(track_id=int(q.track_id), start=int(q.start_frame_id), end=int(q.end_frame_id),
                  present=[None if e is None else int(e['cls_id']) for e in q.qf_tube]) for q in tracker.query_feat_tubes]
    return frames, tubes


def _check_clip(golden, use_kalman):
    frames, tubes = _run_clip(use_kalman)
    want = golden['clip_kalman' if use_kalman else 'clip_nokalman']
    assert tubes == want['tubes']
    assert len(frames) == len(want['frames'])
    for a, b in zip(frames, want['frames']):
        assert a['ids'] == b['ids'] and a['num_tubes'] == b['num_tubes']
        assert a.get('lost') == b.get('lost') and a.get('removed') == b.get('removed')
        np.testing.assert_allclose(np.array(a['tlwh']).reshape(-1, 4), np.array(b['tlwh']).reshape(-1, 4), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('use_kalman', [True, False])
def test_state_machine_matches_reference_with_oracle_numerics(golden, monkeypatch, use_kalman):
    """Host logic of AssociationTracker.update / STrack / QueryFeatTube against the reference's run.  No GPU in the CPU
    suite and no CPU fallback in the product, so the two device calls are replaced by the oracle (TEST DOUBLE)."""
    from openpvsg_b200 import ops

    def fake_reconsdot(t, d, tmp=100.0):     # inputs arrive position-major [n, positions, d] zero padded
        tl = [x.t()[None][:, :, :int((x.abs().sum(1) > 0).sum())] for x in t]
        dl = [x.t()[None][:, :, :int((x.abs().sum(1) > 0).sum())] for x in d]
        return torch.as_tensor(otr.reconsdot_distance(tl, dl, tmp)).float()

    def fake_lap(cost, limit):
        _, x, y = otr.lapjv(cost.double().numpy(), extend_cost=True, cost_limit=limit)
        return torch.as_tensor(x).int(), torch.as_tensor(y).int()

    monkeypatch.setattr(ops, 'reconsdot', fake_reconsdot)
    monkeypatch.setattr(ops, 'lap_assign', fake_lap)
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)
    monkeypatch.setattr(torch.Tensor, 'to', lambda self, *a, **k: self)
    _check_clip(golden, use_kalman)


def test_box_helpers():
    m = torch.zeros(2, 1, 12, 16)
    m[0, 0, 3:7, 4:10] = 1
    boxes = trk.mask2box(m)
    assert boxes[1].tolist() == [-1, -1, 10, 10]
    rows, cols = np.arange(3, 7).repeat(6), np.tile(np.arange(4, 10), 4)
    cx, cy = cols.mean(), rows.mean()
    dx, dy = max(np.abs(cols - cx).mean(), 1), max(np.abs(rows - cy).mean(), 1)
    np.testing.assert_allclose(boxes[0], [cx - 2 * dx, cy - 2 * dy, cx + 2 * dx, cy + 2 * dy], rtol=1e-6)
    b = np.array([[0, 0, 10, 10], [1, 1, 10, 10], [20, 20, 30, 30], [-1, -1, 10, 10]], np.float64)
    assert trk.remove_duplicated_box(b, iou_th=0.7).tolist() == [0, 2]
    assert trk.remove_duplicated_box(b, iou_th=0.9).tolist() == [0, 1, 2]


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_reconsdot_kernel_vs_oracle(golden):
    from openpvsg_b200 import ops
    t, d = fx.embedding_sets(seed=1)
    T = [types.SimpleNamespace(curr_feat=f) for f in t]
    D = [types.SimpleNamespace(curr_feat=f) for f in d]
    cost, _ = trk.reconsdot_distance(T, D)
    np.testing.assert_allclose(cost, np.array(golden['reconsdot']['cost']), atol=2e-4)
    g = torch.Generator().manual_seed(9)
    for ntrk, ndet, dim in ((1, 1, 16), (7, 3, 64), (20, 31, 128), (3, 12, 40)):
        tl = [torch.randn(1, dim, int(n), generator=g) for n in torch.randint(1, 40, (ntrk,), generator=g)]
        dl = [torch.randn(1, dim, int(n), generator=g) for n in torch.randint(1, 40, (ndet,), generator=g)]
        for i in range(min(ntrk, ndet)):       # some detections resemble some tracks
            n = dl[i].shape[2]
            dl[i] = tl[i].repeat(1, 1, 40)[:, :, :n] + 0.1 * torch.randn(1, dim, n, generator=g)
        want = otr.reconsdot_distance(tl, dl)
        got, _ = trk.reconsdot_distance([types.SimpleNamespace(curr_feat=f) for f in tl], [types.SimpleNamespace(curr_feat=f) for f in dl])
        np.testing.assert_allclose(got, want, atol=3e-4)
        assert ops is not None


@pytest.mark.gpu
def test_lap_kernel_vs_oracle(golden):
    for case in golden['lap']:
        m, ua, ub = trk.linear_assignment(_inf(case['cost']), case['thresh'])
        assert np.asarray(m).tolist() == case['matches'] and list(ua) == case['unmatched_a'] and list(ub) == case['unmatched_b']
    rng = np.random.default_rng(0)
    for trial in range(60):
        n, m = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        c = rng.random((n, m)).astype(np.float32).astype(np.float64)
        c[rng.random((n, m)) < 0.2] = np.inf
        thresh = float(rng.choice([0.3, 0.5, 0.9, 5.0]))
        got, ga, gb = trk.linear_assignment(c, thresh)
        want, wa, wb = otr.linear_assignment(c, thresh)
        cost = lambda mm: sum(c[i, j] for i, j in np.asarray(mm).reshape(-1, 2))  # noqa: E731
        assert len(got) == len(want) and abs(cost(got) - cost(want)) < 1e-9, (trial, n, m)
        assert np.asarray(got).tolist() == np.asarray(want).tolist()      # generic costs: the optimum is unique
        assert list(ga) == list(wa) and list(gb) == list(wb)
    empty = trk.linear_assignment(np.zeros((0, 4)), 0.9)
    assert empty[0].shape == (0, 2) and empty[2] == (0, 1, 2, 3)


@pytest.mark.gpu
def test_minvis_chain_kernels_vs_scipy():
    """All T-1 MinVIS matchings at once (pvsg_cosine_chain_cost + pvsg_lap_square_batched + pvsg_perm_chain) against the
    reference's sequential order of operations (mask2former_min_vis.py:176-181, 244-258) with scipy on the host."""
    from scipy.optimize import linear_sum_assignment
    from openpvsg_b200 import ops, tubes
    g = torch.Generator().manual_seed(5)
    for T, Q, C in ((9, 100, 256), (2, 100, 256), (5, 7, 16), (3, 256, 32)):
        base = torch.randn(Q, C, generator=g)
        embeds = torch.stack([base[torch.randperm(Q, generator=g)] * (1 + 0.3 * torch.rand(Q, 1, generator=g))
                              + 0.3 * torch.randn(Q, C, generator=g) for _ in range(T)])
        out, want = [embeds[0]], [torch.arange(Q)]
        for t in range(1, T):
            cur = embeds[t] / embeds[t].norm(dim=1)[:, None]
            tgt = out[-1] / out[-1].norm(dim=1)[:, None]
            idx = torch.as_tensor(linear_sum_assignment((1 - cur @ tgt.T).T.numpy())[1])
            want.append(idx)
            out.append(embeds[t][idx])
        dev = embeds.cuda()
        sigma = ops.minvis_chain(dev)
        assert sigma.shape == (T - 1, Q) and all(sorted(r.tolist()) == list(range(Q)) for r in sigma)
        perms = ops.perm_chain(sigma, Q)
        assert torch.equal(perms.cpu().long(), torch.stack(want)), (T, Q, C)
        assert torch.equal(tubes.minvis_link_sharded(dev, T).cpu(), torch.stack(want))
    one = ops.perm_chain(torch.empty(0, 11, dtype=torch.int32, device='cuda'), 11)
    assert one.cpu().tolist() == [list(range(11))]


@pytest.mark.gpu
@pytest.mark.parametrize('use_kalman', [True, False])
def test_tracker_on_device_matches_reference(golden, use_kalman):
    _check_clip(golden, use_kalman)


@pytest.mark.gpu
def test_mask_tracker_extract_emb_vs_oracle():
    """MaskAssociationTracker.extract_emb / prepare_obs (mask.py:21-60) with a fixed appearance map: small masks (plain
    gather), a large mask (bilinear rescale with the caller's scale factor) and an empty one."""
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(1, 32, 45, 80, generator=g)
    H, W = 360, 640
    obs = np.zeros((4, H, W), np.float32)
    obs[0, 40:120, 100:220] = 1          # 10 x 15 = 150 feature positions
    obs[1, 0:300, 0:500] = 1             # large: > max_mask_area positions
    obs[3, 200:260, 300:420] = 1
    cfg = to_cfg(fx.tracker_cfg())
    tracker = trk.MaskAssociationTracker(cfg, app_model=lambda img: feat.cuda())
    masks, embs = tracker.extract_emb(torch.zeros(3, H, W), obs)
    want_masks, want = otr.extract_emb(feat, obs, cfg.mots.max_mask_area, cfg.mots.feat_size)
    assert torch.equal(masks.cpu(), want_masks)
    for k in (0, 1, 3):
        assert embs[k].shape == want[k].shape, (k, embs[k].shape, want[k].shape)
        assert (embs[k] - want[k]).abs().max().item() < 1e-5
    assert embs[2].shape == (32, 40)      # empty mask: random template, as the reference
    dets = tracker.prepare_obs(torch.zeros(3, H, W), None, obs)
    assert len(dets) == 3 and all(d.curr_feat.shape[1] == 32 for d in dets)


def test_frame_observations_match_reference_golden():
    """tracker.frame_observations against the reference's own LoadOutputsFromMask2Former._get_binary_masks_and_query_feats
    (golden vectors from tests/golden/make_golden_train_gt.py): id order, void dropped, binary masks, class ids, the mean
    query feature of merged stuff segments, the empty frame."""
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tracker_frames.json')))
    for c in golden:
        qf = {int(k): [np.asarray(x, np.float32) for x in v] for k, v in c['query_feats'].items()}
        masks, feats = trk.frame_observations(np.asarray(c['pan'], np.int32), qf, 126)
        assert np.asarray(masks).tolist() == c['masks']
        assert [f['cls_id'] for f in feats] == [f['cls_id'] for f in c['feats']]
        for a, b in zip(feats, c['feats']):
            assert np.allclose(a['query_feat'], np.asarray(b['query_feat'], np.float32), atol=1e-7)


@pytest.mark.gpu
def test_track_clip_links_moving_segments():
    """tracker.track_clip (eval_seq of test_mots_from_mask2former.py:29-95) on a synthetic IPS clip: three segments drift
    across 8 frames (one frame is all void, one segment is missing for two frames).  Every segment keeps ONE track id, the
    masks.txt rows decode back to the panoptic segments, and the query-feature tubes are complete with None where the
    track was not seen (positions follow the reference's quirk for void frames, see the end of the test)."""
    from openpvsg_b200 import tubes
    H, W, T = 96, 160, 8
    rng = np.random.default_rng(0)
    protos = rng.standard_normal((3, 32)).astype(np.float32)
    ids = [1005, 2005, 120]                                   # two instances of thing class 5, one stuff segment

    def boxes(t):
        return [(10 + 4 * t, 10, 30, 30), (60, 20 + 3 * t, 36, 28), (100 - 3 * t, 50, 40, 36)]

    outputs, frames, feats_per_frame = [], [], []
    for t in range(T):
        pan = np.full((H, W), 126, np.int32)
        qf = {}
        for k, (x, y, w, h) in enumerate(boxes(t)):
            if t == 4 or (k == 1 and t in (2, 3)):
                continue
            pan[y:y + h, x:x + w] = ids[k]
            qf[ids[k]] = [np.full((1, 256), float(k + 1), np.float32)]
        qf = {i: v for i, v in qf.items() if (pan == i).any()}
        # appearance map consistent with the panoptic map: the cell whose top-left pixel belongs to segment k carries proto k
        cell = pan[::8, ::8]
        fmap = np.zeros((32, H // 8, W // 8), np.float32)
        for k, i in enumerate(ids):
            fmap[:, cell == i] = protos[k][:, None]
        outputs.append(dict(pan_results=pan, query_feats=qf))
        frames.append(torch.zeros(3, H, W))
        feats_per_frame.append(torch.from_numpy(fmap)[None])
    it = iter(f for t, f in enumerate(feats_per_frame) if t != 4)      # the void frame never reaches the appearance network
    app = lambda img: next(it).cuda()                         # noqa: E731  the appearance map of the frame being tracked
    cfg = to_cfg(fx.tracker_cfg())
    results, qtubes = trk.track_clip(outputs, frames, cfg, 126, app)
    assert len(results) == T and results[4] == (5, [], [], [])
    seen = {}
    for fid, tlwhs, masks, tids in results:
        pan = outputs[fid - 1]['pan_results']
        for m, tid in zip(masks, tids):
            dec = tubes.rle_decode(m['counts'], H, W).astype(bool)
            seg = [i for i in ids if np.array_equal(dec, pan == i)]
            assert len(seg) == 1, (fid, tid)
            assert m['class_id'] == seg[0] % 1000
            seen.setdefault(seg[0], set()).add(tid)
    assert all(len(v) == 1 for v in seen.values()) and len({next(iter(v)) for v in seen.values()}) == 3, seen
    rows = trk.mots_rows(results)
    assert len(rows) == sum(len(r[3]) for r in results) and rows[0].split()[0] == '1' and rows[0].split()[3:5] == [str(H), str(W)]
    assert len(qtubes) == 3 and all(len(q.qf_tube) == T for q in qtubes)
    missing = {tuple(i for i, x in enumerate(q.qf_tube) if x is None) for q in qtubes}
    # as in the reference, an all-void frame never reaches tracker.update, so the tubes index TRACKER frames: the void
    # frame shows up as one missing entry at the end (complete_empty_postfix), not at position 4
    assert missing == {(7,), (2, 3, 7)}


def test_track_clip_all_void_clip():
    """No segment in any frame: the tracker is never called, every result row is empty, there are no tubes (CPU: the
    appearance network is never reached)."""
    outputs = [dict(pan_results=np.full((8, 12), 126, np.int32), query_feats={}) for _ in range(3)]

    def never(img):
        raise AssertionError('appearance network called on a void clip')

    results, qtubes = trk.track_clip(outputs, [None] * 3, to_cfg(fx.tracker_cfg()), 126, never)
    assert results == [(1, [], [], []), (2, [], [], []), (3, [], [], [])] and qtubes == [] and trk.mots_rows(results) == []
