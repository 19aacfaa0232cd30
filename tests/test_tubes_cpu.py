"""CPU: tube linking / wire formats / frame sharding, incl. the N > 1 path over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from openpvsg_b200 import tubes
from oracle import m2f as om


def _fake_outputs(num_frames=7, seed=0, h=12, w=9):
    rng = np.random.default_rng(seed)
    outs = []
    for f in range(num_frames):
        pan = np.full((h, w), 126, np.int32)
        ids = []
        for sid in (1005, 117, 2005, 3040):
            if rng.random() < 0.6:
                y, x = rng.integers(0, h - 3), rng.integers(0, w - 3)
                pan[y:y + 3, x:x + 3] = sid
                ids.append(sid)
        ids = [i for i in ids if (pan == i).any()]
        outs.append([dict(pan_results=pan, query_feats={i: [torch.full((256,), float(i + f))] for i in ids})])
    return outs


def test_rle_roundtrip_and_oracle():
    rng = np.random.default_rng(1)
    for shape in ((1, 1), (5, 7), (33, 21), (64, 64)):
        for p in (0.0, 0.1, 0.5, 1.0):
            m = (rng.random(shape) < p).astype(np.uint8)
            counts = tubes.rle_counts(m)
            assert counts == om.rle_encode(m)
            assert sum(counts) == m.size
            s = tubes.rle_string(counts)
            assert s == om.rle_to_string(counts)
            assert np.array_equal(tubes.rle_decode(s, *shape), m)
    # long runs need multi-character codes
    m = np.zeros((300, 300), np.uint8)
    m[100:200, 50:250] = 1
    assert np.array_equal(tubes.rle_decode(tubes.rle_string(tubes.rle_counts(m)), 300, 300), m)


def test_concat_seq_matches_oracle():
    outs = _fake_outputs()
    linker = tubes.concat_seq(outs)
    rows, feat_tubes = om.concat_seq(outs)
    assert linker.rows == rows
    assert sorted(linker.feat_tubes) == sorted(feat_tubes)
    for tid in feat_tubes:
        assert sorted(linker.feat_tubes[tid]) == sorted(feat_tubes[tid])
        for f in feat_tubes[tid]:
            assert np.array_equal(linker.feat_tubes[tid][f]['query_feat'], feat_tubes[tid][f]['query_feat'])
            assert linker.feat_tubes[tid][f]['cls_id'] == feat_tubes[tid][f]['cls_id']
    txt = linker.masks_txt().splitlines()
    assert len(txt) == len(rows) and txt[0].split(' ')[:5] == [str(v) for v in rows[0][:5]]
    feats = linker.tube_features()
    assert feats.shape == (len(linker.object_list), len(outs), 256)
    # empty clip / frames with nothing detected
    assert tubes.concat_seq([[dict(pan_results=np.full((4, 4), 126, np.int32), query_feats={})]]).rows == []


def test_bulk_linking_equals_incremental():
    """TubeLinker.add_frames_bulk (the all-gathered compact form) == frame-by-frame add_frame, incl. empty frames,
    a stuff id kept in several frames and a block appended after incremental frames."""
    outs = _fake_outputs(num_frames=11, seed=5)
    ref = tubes.concat_seq(outs)
    entries = []
    for o in outs:
        ids = list(o[0]['query_feats'].keys())
        entries.append((ids, [o[0]['query_feats'][k][0].numpy() for k in ids]))
    counts, ids, feats = tubes.pack_frames(entries)
    assert counts.tolist() == [len(e[0]) for e in entries] and feats.shape == (len(ids), 256)
    lk = tubes.TubeLinker()
    for e in entries[:4]:
        lk.add_frame(*e)
    c2, i2, f2 = tubes.pack_frames(entries[4:])
    lk.add_frames_bulk(c2, i2, f2)
    assert lk.object_list == ref.object_list and lk.num_frames == ref.num_frames
    assert lk.frame_seg_ids == ref.frame_seg_ids and lk.frame_tube_ids() == ref.frame_tube_ids()
    assert np.array_equal(lk.tube_features(), ref.tube_features())
    for tid in ref.feat_tubes:
        assert {f: d['cls_id'] for f, d in lk.feat_tubes[tid].items()} == {f: d['cls_id'] for f, d in ref.feat_tubes[tid].items()}
    with pytest.raises(ValueError):
        tubes.TubeLinker().add_frames_bulk([2], [5], np.zeros((1, 256), np.float32))
    # single process gather_and_link = the bulk path
    one = tubes.gather_and_link(entries, len(entries))
    assert one.object_list == ref.object_list and np.array_equal(one.tube_features(), ref.tube_features())


def test_shard_frames():
    for T, G in ((100, 1), (300, 8), (380, 8), (5, 8), (0, 2)):
        blocks = [tubes.shard_frames(T, G, r) for r in range(G)]
        assert blocks[0][0] == 0 and blocks[-1][1] == T
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(b[1] - b[0] for b in blocks) == (T + G - 1) // G if T else True


def _worker(rank, world, port, outs, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = tubes.shard_frames(len(outs), world, rank)
    entries = []
    for o in outs[lo:hi]:
        ids = list(o[0]['query_feats'].keys())
        entries.append((ids, [o[0]['query_feats'][k][0].numpy() for k in ids]))
    linker = tubes.gather_and_link(entries, len(outs))
    q.put((rank, linker.object_list, linker.tube_features()))
    dist.destroy_process_group()


def test_gather_and_link_world2_gloo():
    outs = _fake_outputs(num_frames=9, seed=3)
    outs[6][0]['query_feats'] = {}          # a frame that keeps nothing
    ref = tubes.concat_seq(outs)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, outs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, object_list, feats in got:
        assert object_list == ref.object_list
        assert np.array_equal(feats, ref.tube_features())


def _host_solve(embeds):
    """scipy stand-in for ops.minvis_chain (host logic tests only)."""
    from scipy.optimize import linear_sum_assignment
    e = embeds.numpy().astype(np.float64)
    e = e / np.linalg.norm(e, axis=2, keepdims=True)
    return torch.from_numpy(np.stack([linear_sum_assignment(1.0 - e[t] @ e[t + 1].T)[1] for t in range(len(e) - 1)]).astype(np.int32))


def _host_compose(sigma, Q):
    perms = [torch.arange(Q)]
    for s in sigma.long():
        perms.append(s[perms[-1]])
    return torch.stack(perms)


def _sequential_minvis(embeds):
    """The reference's order of operations: frame t is matched against the RE-ORDERED queries of frame t-1."""
    from scipy.optimize import linear_sum_assignment
    out = [embeds[0]]
    perms = [torch.arange(embeds.shape[1])]
    for t in range(1, embeds.shape[0]):
        cur = embeds[t] / embeds[t].norm(dim=1)[:, None]
        tgt = out[-1] / out[-1].norm(dim=1)[:, None]
        C = 1 - cur @ tgt.T
        idx = torch.as_tensor(linear_sum_assignment(C.T.numpy())[1])
        perms.append(idx)
        out.append(embeds[t][idx])
    return torch.stack(perms)


def _minvis_embeds(T, Q=12, C=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(Q, C, generator=g)
    return torch.stack([base[torch.randperm(Q, generator=g)] + 0.05 * torch.randn(Q, C, generator=g) for _ in range(T)])


def test_minvis_parallel_linking_equals_sequential_chain():
    for T in (1, 2, 7):
        e = _minvis_embeds(T, seed=T)
        got = tubes.minvis_link_sharded(e, T, _host_solve, _host_compose)
        assert torch.equal(got, _sequential_minvis(e))


def test_batch_schedule_ramps_and_covers_every_frame():
    """engine.batch_schedule: half-sized first / last batch, every batch within the runner capacity, all frames covered."""
    from openpvsg_b200 import engine
    assert engine.batch_schedule(100, 20) == [10, 20, 20, 20, 20, 10]
    assert engine.batch_schedule(100, 20, ramp=False) == [20] * 5
    assert engine.batch_schedule(39, 20) == [20, 19] and engine.batch_schedule(7, 20) == [7] and engine.batch_schedule(0, 20) == []
    for n in range(1, 130):
        for b in (1, 2, 8, 20):
            s = engine.batch_schedule(n, b)
            assert sum(s) == n and all(0 < x <= b for x in s), (n, b, s)
            if n >= 2 * b and b >= 2:
                assert s[0] == b // 2


def test_splitk_plan_covers_the_reduction():
    """ops.splitk_plan (weight-gradient GEMMs): chunk length a multiple of the 64-column k-block, chunks x length covers the
    token count with less than one chunk of padding, enough chunks to fill the SMs when the output has few tiles."""
    from openpvsg_b200 import ops
    for T in (1, 30, 100, 960, 3840, 23040, 368640, 2949120):
        for M, N in ((64, 64), (256, 576), (2048, 512), (100, 256), (64, 147)):
            S, Kc, Tp = ops.splitk_plan(T, M, N)
            assert Kc % 64 == 0 and Tp == S * Kc and Tp >= T and Tp - T < Kc + 64, (T, M, N, S, Kc)
            tiles = ((M + 127) // 128) * ((N + 127) // 128)
            if T >= 256 * 296:
                assert S * tiles >= 148, (T, M, N, S)
    assert ops.splitk_plan(100, 256, 256) == (1, 128, 128)


def test_single_process_collectives_are_noops():
    """world size 1 (no process group): the gradient exchange and the parameter broadcast do nothing, the MinVIS linking of a
    one-frame clip is the identity."""
    from openpvsg_b200 import dist_train
    model = torch.nn.Linear(3, 2)
    model(torch.ones(1, 3)).sum().backward()
    before = [p.grad.clone() for p in model.parameters()]
    assert dist_train.allreduce_gradients(list(model.parameters())) is None and dist_train.world() == 1
    dist_train.broadcast_parameters(model)
    assert all(torch.equal(a, p.grad) for a, p in zip(before, model.parameters()))
    e = _minvis_embeds(1, Q=5)
    assert tubes.minvis_link_sharded(e, 1, _host_solve, _host_compose).tolist() == [[0, 1, 2, 3, 4]]


def _minvis_worker(rank, world, port, embeds, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = tubes.shard_frames(len(embeds), world, rank)
    perms = tubes.minvis_link_sharded(embeds[lo:hi], len(embeds), _host_solve, _host_compose)
    q.put((rank, perms.numpy()))
    dist.destroy_process_group()


def test_minvis_link_sharded_world2_gloo():
    for T in (7, 2):                        # 2 frames: rank 1 owns the only pair, through the halo
        embeds = _minvis_embeds(T, seed=11 + T)
        ref = _sequential_minvis(embeds).numpy()
        s = socket.socket()
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
        s.close()
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_minvis_worker, args=(r, 2, port, embeds, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        for rank, perms in got:
            assert np.array_equal(perms, ref), rank


def _grad_worker(rank, world, port, q):
    import torch.distributed as dist
    from openpvsg_b200 import dist_train
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    model[1].bias.requires_grad_(False)
    dist_train.broadcast_parameters(model)                         # every rank now holds rank 0's weights
    start = [p.detach().clone() for p in model.parameters()]
    x = torch.full((4, 5), float(rank + 1))
    if rank == 0:
        model(x).sum().backward()
    else:                                                          # this rank's batch does not touch layer 1
        model[0](x).sum().backward()
    bucket = dist_train.allreduce_gradients(list(model.parameters()))
    bucket2 = dist_train.allreduce_gradients(list(model.parameters()), bucket)     # reusable buffer, idempotent on equal grads
    q.put((rank, [t.numpy() for t in start], [None if p.grad is None else p.grad.numpy() for p in model.parameters()],
           bucket2.data_ptr() == bucket.data_ptr()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    """dist_train: weights broadcast from rank 0, gradients averaged over the ranks through one flat bucket; a parameter
    without a gradient on one rank contributes zeros, frozen parameters are left alone."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (_, w0, g0, same0), (_, w1, g1, same1) = got
    assert same0 and same1
    for a, b in zip(w0, w1):
        assert np.array_equal(a, b)
    assert g0[3] is None and g1[3] is None                         # frozen bias: untouched
    for a, b in zip(g0[:3], g1[:3]):
        assert np.array_equal(a, b)                                # both ranks hold the same averaged gradients
    # reference: the mean of the two ranks' gradients computed in one process
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    with torch.no_grad():
        for p, w in zip(model.parameters(), w0):
            p.copy_(torch.from_numpy(w))
    model(torch.full((4, 5), 1.0)).sum().backward()
    ga = [p.grad.clone() for p in list(model.parameters())[:3]]
    model.zero_grad()
    model[0](torch.full((4, 5), 2.0)).sum().backward()
    gb = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in list(model.parameters())[:3]]
    for a, b, got_g in zip(ga, gb, g0[:3]):
        assert np.allclose((a + b).numpy() / 2, got_g, atol=1e-6)


def test_rle_events_host_side_matches_scalar_encoder():
    """tubes.rle_from_events / rle_string_np (host side of the device RLE encoder, pvsg_rle_events)
    against the scalar pycocotools-style encoder, including a segment that owns pixel 0, an absent
    segment (empty mask) and a stuff class kept twice."""
    import numpy as np
    from openpvsg_b200 import tubes
    rng = np.random.default_rng(0)
    H, W = 37, 53
    pan = np.full((H, W), 126, np.int32)
    for sid in (5, 1003, 2007, 120):
        y0, x0 = rng.integers(0, H - 8), rng.integers(0, W - 8)
        pan[y0:y0 + rng.integers(3, 20), x0:x0 + rng.integers(3, 25)] = sid
    pan[0, 0] = 5
    ids = [5, 1003, 2007, 120, 77]
    flat = pan.reshape(-1, order='F')
    ev_pos, ev_slot, prev = [], [], None
    for p, cur in enumerate(flat.tolist()):          # the walk pvsg_rle_events performs
        if prev is None or cur != prev:
            if prev is not None and prev in ids:
                ev_pos.append(p), ev_slot.append(ids.index(prev))
            if cur in ids:
                ev_pos.append(p), ev_slot.append(ids.index(cur))
            prev = cur
    out = tubes.rle_from_events(np.array(ev_pos, np.uint32).view(np.int32), np.array(ev_slot, np.int16), len(ev_pos),
                                ids, H, W)                       # C++ host routine of the library
    assert out == tubes.rle_from_events(np.array(ev_pos, np.uint32).view(np.int32), np.array(ev_slot, np.int16),
                                        len(ev_pos), ids, H, W, native=False)   # numpy cross-check
    for sid in ids:
        assert out[sid] == tubes.rle_string(tubes.rle_counts(pan == sid))
        assert np.array_equal(tubes.rle_decode(out[sid], H, W), (pan == sid).astype(np.uint8))
    for _ in range(20):
        c = rng.integers(0, 200000, size=rng.integers(1, 50))
        assert tubes.rle_string_np(c) == tubes.rle_string(c.tolist())
    info = np.array([3, 0, 5, 120, 10, 1, 7, -1, 0, 2, 5, 120, 4] + [0] * 8, np.int32)
    assert tubes.slot_ids(info) == [120]
