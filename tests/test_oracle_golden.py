"""CPU: the oracle restatement vs golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py), plus cross-checks of the restated mmcv/mmdet parts
against independent implementations available offline."""
import os

import numpy as np
import pytest
import torch

from openpvsg_b200 import synthetic as syn
from oracle import m2f as om
from oracle import relation as orel

TOL = 2e-5  # oracle and reference run the same fp32 torch CPU primitives


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _checksum(*tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors))


def _close(a, b, tol=TOL):
    a = torch.as_tensor(np.asarray(a)).float()
    b = torch.as_tensor(np.asarray(b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert err <= tol * max(1.0, b.abs().max().item()), err


def test_relation_golden(golden_dir):
    g = _load(golden_dir, 'rel_small.npz')
    sds = syn.relation_state_dicts(seed=int(g['weights_seed']))
    N, T, P = int(g['N']), int(g['T']), int(g['P'])
    feats = _randn(int(g['feats_seed']), N, T, 256)
    feats[3, 5:] = 0.0
    assert abs(_checksum(feats, *[v for sd in sds.values() for v in sd.values()
                                  if v.dtype.is_floating_point]) - float(g['in_checksum'])) < 1e-3
    out = orel.relation_forward(sds, feats, P)
    _close(out['sub'], g['sub'])
    _close(out['obj'], g['obj'])
    _close(out['pred_matrix'], g['pred_matrix'])
    assert out['pairs'] == g['pairs'].tolist()
    _close(orel.concatenate_sub_obj(out['sub'], out['obj'], out['pairs']), g['cat'])
    _close(out['span_pred'], g['span'], 5e-5)
    _close(out['prob'], g['prob'], 5e-5)
    vs, vp = orel.vanilla_model(sds['relation_model'], torch.as_tensor(g['cat']))
    _close(vs, g['vspan'])
    _close(vp, g['vprob'])
    # result lists are generated from the golden span/prob so ordering is compared exactly
    span, prob = torch.as_tensor(g['span']), torch.as_tensor(g['prob'])
    pw = orel.generate_pairwise_results(span, prob, g['pairs'].tolist())
    assert [[r['subject_index'], r['object_index'], r['relation']] for r in pw] == g['pw_triplets'].tolist()
    assert np.array_equal(np.array([r['relation_span'] for r in pw]).astype(np.uint8), g['pw_spans'])
    allr = orel.generate_results(span, prob, g['pairs'].tolist())[:200]
    assert [[r['subject_index'], r['object_index'], r['relation']] for r in allr] == g['all_triplets'].tolist()


def test_fusion_postprocess_golden(golden_dir):
    g = _load(golden_dir, 'fusion_post.npz')
    mask_cls, mask_pred = torch.as_tensor(g['mask_cls']), torch.as_tensor(g['mask_pred'])
    qf = torch.as_tensor(g['query_feats'])
    pan, qfd = om.panoptic_postprocess_with_query(mask_cls, mask_pred, qf)
    assert np.array_equal(pan.numpy(), g['pan'])
    assert sorted(qfd.keys()) == g['qf_keys'].tolist()
    assert [len(qfd[k]) for k in sorted(qfd.keys())] == g['qf_counts'].tolist()
    _close(torch.stack([qfd[k][0] for k in sorted(qfd.keys())]), g['qf_vals'], 0)
    labels, bboxes, binm = om.instance_postprocess(mask_cls, mask_pred)
    # topk(sorted=False) order is implementation-defined: compare as sets keyed by (label, score)
    def canon(lb, bb):
        rows = np.concatenate([np.asarray(lb, np.float64)[:, None], np.asarray(bb, np.float64)], 1)
        return rows[np.lexsort(rows.T[::-1])]
    np.testing.assert_allclose(canon(labels, bboxes), canon(g['ins_labels'], g['ins_bboxes']), atol=1e-6)
    meta = dict(img_shape=(36, 50, 3), ori_shape=(36, 50, 3))
    res = om.fusion_simple_test_with_query(mask_cls[None], mask_pred[None], qf[None], [meta], rescale=True)[0]
    assert np.array_equal(res['pan_results'].numpy(), g['crop_pan'])
    np.testing.assert_allclose(canon(res['ins_results'][0], res['ins_results'][1]),
                               canon(g['crop_ins_labels'], g['crop_ins_bboxes']), atol=1e-6)


def test_forward_head_golden(golden_dir):
    g = _load(golden_dir, 'forward_head.npz')
    sd = syn.mask2former_state_dict(seed=int(g['weights_seed']))
    dec_out, mf_img, mf_vid = _randn(31, 100, 2, 256), _randn(32, 2, 256, 12, 20), _randn(33, 2, 3, 256, 12, 20)
    assert abs(_checksum(dec_out, mf_img, mf_vid) - float(g['in_checksum'])) < 1e-3
    c1, m1, a1 = om.forward_head(sd, 'panoptic_head.', dec_out, mf_img, (3, 5))
    c2, m2, a2 = om.forward_head_video(sd, 'panoptic_head.', dec_out, mf_vid, (6, 10))
    _close(c1, g['cls_img'])
    _close(m1, g['mask_img'])
    _close(c2, g['cls_vid'])
    _close(m2, g['mask_vid'])
    assert list(a1.shape) == g['attn_img_shape'].tolist() and list(a2.shape) == g['attn_vid_shape'].tolist()
    assert np.array_equal(np.packbits(a1.numpy(), axis=-1), g['attn_img'])
    assert np.array_equal(np.packbits(a2.numpy(), axis=-1), g['attn_vid'])


def test_pe3d_and_match_golden(golden_dir):
    g = _load(golden_dir, 'pe3d.npz')
    _close(om.sine_pe_3d(1, 2, 5, 7), g['pos'], 1e-6)
    # with T = 1 the z term is constant and the 2-D encoding differs only by that constant
    g = _load(golden_dir, 'match_embds.npz')
    tgt = _randn(41, 100, 256)
    cur = tgt[torch.randperm(100, generator=torch.Generator().manual_seed(42))] + 0.3 * _randn(43, 100, 256)
    assert abs(_checksum(tgt, cur) - float(g['in_checksum'])) < 1e-3
    assert np.array_equal(np.asarray(om.match_from_embds(tgt, cur)), g['indices'])


def test_head_and_detector_golden(golden_dir):
    g = _load(golden_dir, 'head_forward.npz')
    sd = syn.mask2former_state_dict(seed=int(g['weights_seed']))
    H, W = int(g['H']), int(g['W'])
    img = syn.synthetic_frame(int(g['frame_seed']), H, W)[None]
    meta = syn.frame_meta(H, W)
    assert abs(_checksum(img) - float(g['in_checksum'])) < 1e-2
    with torch.no_grad():
        feats = om.resnet50(sd, img)
        vc, vm, vq = om.head_forward(sd, feats, video=True, num_frames=1)
        ic, im, iq = om.head_forward(sd, feats)
        vcls, vmask, vqf = om.head_simple_test_with_query(sd, feats, meta['batch_input_shape'], True, 1)
        icls, imask, iqf = om.head_simple_test_with_query(sd, feats, meta['batch_input_shape'])
    tol = 2e-4  # 15 stacked transformer layers, same primitives
    _close(vc[-1], g['v_cls_last'], tol)
    _close(vm[-1], g['v_mask_last'], tol)
    _close(vq, g['v_query'], tol)
    _close(vc[4], g['v_cls_mid'], tol)
    _close(vm[0], g['v_mask_first'], tol)
    _close(ic[-1], g['i_cls_last'], tol)
    _close(im[-1], g['i_mask_last'], tol)
    _close(iq, g['i_query'], tol)
    assert list(vmask.shape) == g['v_up_shape'].tolist() and list(imask.shape) == g['i_up_shape'].tolist()
    assert list(vqf.shape) == g['v_qf_shape'].tolist() and list(iqf.shape) == g['i_qf_shape'].tolist()
    _close(vmask[0, 0, ::7, ::5, ::5], g['v_up_sample'], tol)
    _close(imask[0, ::7, ::5, ::5], g['i_up_sample'], tol)

    d = _load(golden_dir, 'detector.npz')
    with torch.no_grad():
        vres = om.vps_simple_test(sd, img[None], [[meta]])[0][0]
        ires = om.ips_simple_test(sd, img, [meta])[0]
    assert np.array_equal(vres['pan_results'], d['v_pan'])
    assert sorted(vres['query_feats'].keys()) == d['v_keys'].tolist()
    _close(torch.stack([vres['query_feats'][k][0] for k in sorted(vres['query_feats'])]), d['v_feats'], tol)
    assert np.array_equal(ires['pan_results'], d['i_pan'])
    assert sorted(ires['query_feats'].keys()) == d['i_keys'].tolist()
    _close(np.stack([ires['query_feats'][k][0] for k in sorted(ires['query_feats'])]), d['i_feats'], tol)


def test_msda_core_vs_hf_transformers():
    """Cross-check of the restated mmcv op against HF transformers' independent
    multi_scale_deformable_attention (same grid_sample formulation)."""
    mod = pytest.importorskip('transformers.models.mask2former.modeling_mask2former')
    fn = getattr(mod, 'multi_scale_deformable_attention', None)
    if fn is None:
        pytest.skip('HF function not exposed in this transformers version')
    shapes = [(3, 5), (6, 10), (12, 20)]
    n = sum(h * w for h, w in shapes)
    value = _randn(1, 2, n, 8, 32)
    loc = torch.rand(2, n, 8, 3, 4, 2, generator=torch.Generator().manual_seed(2)) * 1.2 - 0.1
    aw = torch.softmax(_randn(3, 2, n, 8, 12), -1).view(2, n, 8, 3, 4)
    ours = om.msda_core(value, shapes, loc, aw)
    try:
        theirs = fn(value, torch.tensor(shapes), loc, aw)
    except Exception:
        theirs = fn(value, shapes, loc, aw)
    _close(ours, theirs, 1e-6)


def test_resnet_vs_torchvision():
    tv = pytest.importorskip('torchvision')
    sd = syn.mask2former_state_dict(seed=2)
    net = tv.models.resnet50(weights=None).eval()
    bsd = {k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}
    missing = net.load_state_dict(bsd, strict=False)
    assert set(missing.missing_keys) == {'fc.weight', 'fc.bias'} and not missing.unexpected_keys
    x = _randn(5, 1, 3, 64, 96)
    with torch.no_grad():
        y = net.maxpool(net.relu(net.bn1(net.conv1(x))))
        ref = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            y = layer(y)
            ref.append(y)
        ours = om.resnet50(sd, x)
    for a, b in zip(ours, ref):
        _close(a, b, 1e-5)


def test_rle_known_answer():
    # 3x4 mask, column-major runs: col0 = 0,1,1; col1 = 1,0,0; ...
    m = np.array([[0, 1, 0, 0], [1, 0, 0, 1], [1, 0, 0, 1]], np.uint8)
    assert om.rle_encode(m) == [1, 3, 6, 2]
    assert om.rle_encode(np.zeros((2, 2), np.uint8)) == [4]
    assert om.rle_encode(np.ones((2, 2), np.uint8)) == [0, 4]
    # pycocotools string coding: small counts map to chr(48 + x)
    assert om.rle_to_string([1, 3, 6, 2]) == '1365'[:0] + ''.join(chr(48 + c) for c in (1, 3, 6)) + chr(48 + ((2 - 3) & 0x1f))
