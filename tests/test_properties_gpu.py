"""Size-independent properties at the BASELINE sizes (720p frames, 100 queries, the reference's training batch), where the
CPU oracle would take minutes: linearity of the deformable attention and of the mask contraction, optimality / validity of
the assignment solver, RLE round trip of a full panoptic map, and the analytic training gradient against a central finite
difference of the loss.  Complements the oracle-vs-kernel parity tests, which run at sizes the oracle finishes in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pyramid(H, W):
    return [((H + 31) // 32, (W + 31) // 32), ((H + 15) // 16, (W + 15) // 16), ((H + 7) // 8, (W + 7) // 8)]


def test_msda_is_linear_in_value_at_720p():
    """out(a v1 + b v2) = a out(v1) + b out(v2) for fixed projections, both kernels (TMA-staged and lane-group)."""
    import os
    from openpvsg_b200 import ops
    shapes = _pyramid(736, 1280)
    n = sum(h * w for h, w in shapes)
    g = torch.Generator(device='cuda').manual_seed(0)
    v1, v2 = (torch.randn(2, n, 256, device='cuda', generator=g) for _ in range(2))
    proj = torch.randn(2, n, 288, device='cuda', generator=g) * 3
    ref = torch.cat([torch.stack(((torch.arange(w).float().repeat(h) + 0.5) / w, (torch.arange(h).float().repeat_interleave(w) + 0.5) / h), -1)
                     for h, w in shapes]).cuda()
    for impl in ('tile', 'group'):
        os.environ['PVSG_MSDA_IMPL'] = impl
        try:
            o1, o2 = (ops.msda_fused_forward(v, shapes, proj, ref) for v in (v1, v2))
            mix = ops.msda_fused_forward(0.75 * v1 - 1.5 * v2, shapes, proj, ref)
        finally:
            os.environ.pop('PVSG_MSDA_IMPL', None)
        err = float((mix - (0.75 * o1 - 1.5 * o2)).abs().max())
        assert err <= 2e-5 * float(o1.abs().max()), (impl, err)
    assert float(o1.abs().max()) > 0.1


def test_mask_contraction_is_bilinear_at_720p():
    """logits(E1 + E2, F) = logits(E1, F) + logits(E2, F) and logits(E, 2F) = 2 logits(E, F) on the 184 x 320 mask features."""
    from openpvsg_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(1)
    e1, e2 = (torch.randn(2, 100, 256, device='cuda', generator=g) for _ in range(2))
    f = torch.randn(2, 184 * 320, 256, device='cuda', generator=g)
    l1, l2, l12 = (ops.mask_logits(e, f)[0] for e in (e1, e2, e1 + e2))
    scale = float(l12.abs().max())
    assert float((l12 - (l1 + l2)).abs().max()) <= 3e-5 * scale
    assert float((ops.mask_logits(e1, 2 * f)[0] - 2 * l1).abs().max()) <= 1e-6 * scale
    # the sign masks the decoder consumes are the signs of the same numbers
    _, mask, row_open = ops.mask_logits(e1, f, False, True)
    assert torch.equal(mask.bool(), l1 < 0) and torch.equal(row_open, (l1 >= 0).sum(-1).int())


def test_assignment_is_an_optimal_permutation_at_maximum_size():
    """pvsg_lap_square_batched on 6 problems of the largest supported size (256): every row a permutation, total cost equal
    to scipy's optimum and not above any of 2000 random permutations; pvsg_lap_assign with a cost limit never matches a pair
    above the limit and leaves exactly the unmatched rows / columns out."""
    from scipy.optimize import linear_sum_assignment
    from openpvsg_b200 import ops, tracker as trk
    rng = np.random.default_rng(0)
    e = torch.from_numpy(rng.standard_normal((7, 256, 48)).astype(np.float32)).cuda()
    sigma = ops.minvis_chain(e).cpu().numpy()
    en = e.cpu().double().numpy()
    en /= np.linalg.norm(en, axis=2, keepdims=True)
    for t in range(6):
        cost = 1.0 - en[t] @ en[t + 1].T
        assert sorted(sigma[t].tolist()) == list(range(256))
        got = cost[np.arange(256), sigma[t]].sum()
        r, c = linear_sum_assignment(cost)
        assert abs(got - cost[r, c].sum()) < 1e-4
        assert all(got <= cost[np.arange(256), rng.permutation(256)].sum() for _ in range(2000))
    c = rng.random((120, 100)).astype(np.float32).astype(np.float64)
    m, ua, ub = trk.linear_assignment(c, 0.05)
    m = np.asarray(m).reshape(-1, 2)
    assert all(c[i, j] <= 0.05 for i, j in m)
    assert sorted(list(m[:, 0]) + list(ua)) == list(range(120)) and sorted(list(m[:, 1]) + list(ub)) == list(range(100))


def test_rle_round_trip_of_a_full_panoptic_map():
    """device RLE events -> COCO RLE strings -> decode reproduces every segment mask of a 720 x 1280 map with 40 segments."""
    from openpvsg_b200 import ops, tubes
    rng = np.random.default_rng(3)
    H, W, S = 720, 1280, 40
    yy, xx = np.mgrid[0:H, 0:W]
    centers = rng.uniform(0, 1, (S, 2)) * (H, W)
    pan = np.argmin(((yy[..., None] - centers[:, 0]) ** 2 + (xx[..., None] - centers[:, 1]) ** 2), -1).astype(np.int32)   # Voronoi cells
    ids = (np.arange(S) % 126) + 1000 * (np.arange(S) // 3 + 1)
    pan_ids = ids[pan].astype(np.int32)
    seg_info = np.full(1 + 4 * 100, -1, np.int32)
    seg_info[0] = S
    for k in range(S):
        seg_info[1 + 4 * k:5 + 4 * k] = (k, ids[k] % 1000, ids[k], int((pan == k).sum()))
    pos, slot, n = ops.rle_events(torch.from_numpy(pan_ids)[None].cuda(), torch.from_numpy(seg_info)[None].cuda())
    rles = tubes.rle_from_events(pos[0].cpu().numpy(), slot[0].cpu().numpy(), int(n[0]), tubes.slot_ids(seg_info), H, W)
    assert sorted(rles) == sorted(int(i) for i in ids)
    for k in (0, 7, 19, 39):
        assert np.array_equal(tubes.rle_decode(rles[int(ids[k])], H, W).astype(bool), pan_ids == ids[k])
    total = sum(int(tubes.rle_decode(r, H, W).sum()) for r in rles.values())
    assert total == H * W                                         # the segments tile the frame


def test_training_gradient_matches_finite_difference_at_reference_batch():
    """Analytic gradient (forward_train_outputs + loss_single + backward at 4 clips x 2 frames @384x480, 12544 points) against
    the central finite difference of the same loss along a random direction.  The direction spans the parameters of the
    LAST decoder layer, which feed only the final prediction, and the two point sets are fixed, so the discrete parts of
    the objective (attention masks of the earlier layers, sampled points) do not move; the assignment is stable for the
    step sizes used.  Two step sizes must agree with each other and with the analytic value."""
    import openpvsg_b200 as pv
    from openpvsg_b200 import configs, synthetic as syn
    det = pv.build_detector(configs.mask2former_r50(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=3))
    det.cuda()
    head = det.panoptic_head
    data = syn.training_batch(4, device='cuda')
    with torch.no_grad():
        feats = det.extract_feat(data['ref_img'].flatten(0, 1))
    labels, masks = head.preprocess_gt(data['ref_gt_labels'], data['ref_gt_masks'], None, data['ref_gt_instance_ids'], data['ref_img_metas'])
    K = 12544
    g = torch.Generator(device='cuda').manual_seed(5)
    apts = torch.rand(1, K, 2, device='cuda', generator=g)
    lpts = torch.rand(sum(len(x) for x in labels), K, 2, device='cuda', generator=g)
    head.train_pixel_decoder = False                       # constants here: the direction lives in the last decoder layer
    params = [p for n, p in head.named_parameters() if n.startswith('transformer_decoder.layers.8.')]
    assert len(params) == 18

    def loss():
        cls_list, mask_list = head.forward_train_outputs(feats, 2)
        return sum(head.loss_single(cls_list[-1], mask_list[-1], labels, [m.float() for m in masks], None, assign_points=apts,
                                    loss_points=lpts, num_points=K))

    base = loss()
    base.backward()
    direction = [torch.randn(p.shape, device='cuda', generator=g) * p.detach().abs().mean().clamp_min(1e-3) for p in params]
    analytic = sum(float((p.grad.double() * v.double()).sum()) for p, v in zip(params, direction))
    fd = {}
    for eps in (4e-3, 2e-3):
        vals = []
        for sgn in (1.0, -1.0):
            with torch.no_grad():
                for p, v in zip(params, direction):
                    p.add_(v, alpha=sgn * eps)
                vals.append(float(loss().detach().double()))
                for p, v in zip(params, direction):
                    p.add_(v, alpha=-sgn * eps)
        fd[eps] = (vals[0] - vals[1]) / (2 * eps)
    assert abs(analytic) > 1e-2, analytic
    rel = {e: abs(v - analytic) / abs(analytic) for e, v in fd.items()}
    assert min(rel.values()) < 3e-2, (analytic, fd, float(base))
