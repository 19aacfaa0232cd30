"""Training slice: MSDeformAttn backward and the losses of loss_single, forward and backward,
against torch autograd through the CPU oracle (oracle/m2f.py::msda_core, oracle/losses.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    a, b = torch.as_tensor(a).detach().cpu().float(), torch.as_tensor(b).detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f'{what}: {err:.3e}'


@pytest.mark.parametrize('shapes,B,nq', [([(3, 5), (7, 10), (13, 21)], 2, 37), ([(15, 20), (30, 40), (60, 80)], 1, 6300)])
def test_msda_backward_vs_autograd(shapes, B, nq):
    from openpvsg_b200.losses import MultiScaleDeformableAttnFunction
    from oracle import m2f as om
    n = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(n + nq)
    value = torch.randn(B, n, 8, 32, generator=g)
    loc = torch.rand(B, nq, 8, 3, 4, 2, generator=g) * 1.3 - 0.15        # some samples cross the border / fall outside
    aw = torch.softmax(torch.randn(B, nq, 8, 12, generator=g), -1).view(B, nq, 8, 3, 4)
    gout = torch.randn(B, nq, 256, generator=g)
    v, l, a = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    om.msda_core(v, shapes, l, a).backward(gout)
    vd, ld, ad = (t.cuda().requires_grad_(True) for t in (value, loc, aw))
    out = MultiScaleDeformableAttnFunction.apply(vd, shapes, ld, ad)
    out.backward(gout.cuda())
    _close(out, om.msda_core(value, shapes, loc, aw), 1e-4, 'forward')
    _close(vd.grad, v.grad, 2e-4, 'grad value')
    _close(ad.grad, a.grad, 2e-4, 'grad attention weights')
    _close(ld.grad, l.grad, 5e-4, 'grad sampling locations')


def test_point_sample_and_losses_vs_autograd():
    from openpvsg_b200 import losses, ops
    from oracle import losses as ol
    g = torch.Generator().manual_seed(3)
    n, H, W, K = 7, 40, 56, 300
    maps = torch.randn(n, H, W, generator=g) * 3
    pts = torch.rand(n, K, 2, generator=g) * 1.1 - 0.05                  # a few points outside [0,1]: zero padding
    tgt = (torch.rand(n, H, W, generator=g) > 0.6).float()
    md = maps.cuda().requires_grad_(True)
    mc = maps.clone().requires_grad_(True)
    got = losses.point_sample(md[:, None], pts.cuda())
    want = ol.point_sample(mc[:, None], pts)
    _close(got, want, 1e-5, 'point_sample')
    shared = ops.point_sample(maps.cuda(), pts[0].cuda())
    _close(shared, ol.point_sample(maps[:, None], pts[:1].repeat(n, 1, 1)).squeeze(1), 1e-5, 'point_sample, shared points')
    pt = ol.point_sample(tgt[:, None], pts).squeeze(1)
    lm, ld = losses._MaskPointLosses.apply(got[:, 0], pt.cuda(), 5.0, 5.0, 5.0, 1.0)
    (lm * 0.7 + ld * 1.3).backward()
    wm = ol.mask_bce_loss(want.squeeze(1).reshape(-1), pt.reshape(-1), 5.0 * K)
    wd = ol.dice_loss(want.squeeze(1), pt, 5.0)
    (wm * 0.7 + wd * 1.3).backward()
    _close(lm, wm, 1e-5, 'loss_mask')
    _close(ld, wd, 1e-5, 'loss_dice')
    _close(md.grad, mc.grad, 1e-5, 'd(loss_mask + loss_dice) / d mask logits')
    # class loss
    x = torch.randn(200, 127, generator=g) * 2
    y = torch.randint(0, 127, (200,), generator=g)
    cw = torch.ones(127)
    cw[-1] = 0.1
    xd, xc = x.cuda().requires_grad_(True), x.clone().requires_grad_(True)
    lc = losses._WeightedCE.apply(xd, y.cuda(), cw.cuda(), None, 2.0)
    lc.backward()
    wc = ol.cross_entropy_loss(xc, y, cw, 2.0)
    wc.backward()
    _close(lc, wc, 1e-5, 'loss_cls')
    _close(xd.grad, xc.grad, 1e-6, 'd loss_cls / d logits')


def test_loss_single_vs_oracle():
    """Mask2FormerVideoHead.loss_single (mask2former_video_head.py:196-293) end to end on a synthetic batch: Hungarian
    targets (device cost matrix + scipy), the three losses and their gradients w.r.t. cls_scores / mask_preds, with the
    two random point sets fixed so both sides see the same points."""
    from openpvsg_b200 import losses
    from oracle import losses as ol
    g = torch.Generator().manual_seed(11)
    B, T, Q, h, w, K = 2, 2, 20, 24, 40, 500
    cls = torch.randn(B, Q, 127, generator=g)
    masks = torch.randn(B, T, Q, h, w, generator=g) * 2
    gt_labels = [torch.tensor([3, 40, 120]), torch.tensor([7, 7])]
    gt_masks = []
    for b, G in enumerate((3, 2)):
        m = torch.zeros(G, T, h, w)
        for k in range(G):
            m[k, :, 2 + 5 * k:12 + 5 * k, 4 + 8 * k:20 + 8 * k] = 1
            masks[b, :, 4 * k + 1] += 4 * (m[k] - 0.5)                   # make query 4k+1 resemble gt k
        gt_masks.append(m)
    apts = torch.rand(1, K, 2, generator=g)
    lpts = torch.rand(5, K, 2, generator=g)
    cc, mc = cls.clone().requires_grad_(True), masks.clone().requires_grad_(True)
    wc, wm, wd, wlabels, wpos = ol.loss_single(cc, mc, gt_labels, gt_masks, apts, lambda n: lpts[:n])
    (wc + wm + wd).backward()
    cd, md = cls.cuda().requires_grad_(True), masks.cuda().requires_grad_(True)
    lc, lm, ld = losses.loss_single(cd, md, [t.cuda() for t in gt_labels], [t.cuda() for t in gt_masks],
                                    assign_points=apts.cuda(), loss_points=lpts.cuda(), num_points=K)
    (lc + lm + ld).backward()
    for a, b, n in ((lc, wc, 'loss_cls'), (lm, wm, 'loss_mask'), (ld, wd, 'loss_dice')):
        _close(a, b, 2e-5, n)
    assert [r.tolist() for r, _ in wpos] == [[1, 5, 9], [1, 5]]           # the planted matches
    _close(cd.grad, cc.grad, 1e-6, 'grad cls_scores')
    _close(md.grad, mc.grad, 1e-6, 'grad mask_preds')
    # no ground truth at all: class loss only
    lc0, lm0, ld0 = losses.loss_single(cd.detach(), md.detach(), [torch.zeros(0, dtype=torch.int64).cuda()] * B,
                                       [torch.zeros(0, T, h, w).cuda()] * B, assign_points=apts.cuda(), num_points=K)
    assert float(lm0) == 0.0 and float(ld0) == 0.0 and float(lc0) > 0


def test_head_loss_on_forward_outputs():
    """head.forward (all decoder layers) -> head.loss: the reference's training objective evaluated on the B200 forward's
    outputs, 10 x 3 finite loss terms under the reference's key names; the last layer equals a direct loss_single call."""
    import openpvsg_b200 as pv
    from openpvsg_b200 import configs, synthetic as syn
    torch.manual_seed(0)
    det = pv.build_detector(configs.mask2former_r50(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=3))
    det.cuda()
    H, W = 96, 160
    img = syn.synthetic_frame(5, H, W)[None].cuda()
    head = det.panoptic_head
    cls_list, mask_list = head.forward(det.extract_feat(img), [[syn.frame_meta(H, W)]])
    assert len(cls_list) == 10 and mask_list[0].shape[:3] == (1, 1, 100)
    h, w = mask_list[0].shape[-2:]
    gt_masks = torch.zeros(2, 1, h, w, device='cuda')
    gt_masks[0, :, :h // 2, :w // 2] = 1
    gt_masks[1, :, h // 2:, w // 3:] = 1
    gt_labels = torch.tensor([5, 120], device='cuda')
    head.train_cfg = dict(num_points=256, oversample_ratio=3.0, importance_sample_ratio=0.75)
    d = head.loss(cls_list, mask_list, [gt_labels], [gt_masks], None)
    assert set(d) == {'loss_cls', 'loss_mask', 'loss_dice'} | {f'd{i}.{k}' for i in range(9) for k in ('loss_cls', 'loss_mask', 'loss_dice')}
    assert all(torch.isfinite(v).all() and float(v) >= 0 for v in d.values())
    with pytest.raises(NotImplementedError):
        head.forward_train(None)
