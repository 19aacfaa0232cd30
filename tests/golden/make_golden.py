#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):  ``python tests/golden/make_golden.py``.  The committed ``*.npz`` files are
what the tests read.

Two groups:

* relation head -- the reference classes in models/relation_head/*.py depend on torch
  only and are imported by file path, unmodified.
* in-tree Mask2Former logic -- models/mask2former*/.py import mmcv / mmdet, which are
  not installable here.  They are imported under *import stubs* (empty registries,
  ``nn.Module`` base classes, identity decorators); the reference functions that are
  executed (forward_head, forward_head_video, Mask2FormerVideoHead.forward,
  panoptic_postprocess_with_query, instance_postprocess, simple_test_with_query,
  SinePositionalEncoding3D.forward, match_from_embds,
  Mask2FormerVideoCustom.simple_test) are the reference's own, line for line.  Where
  such a function calls into an mmcv/mmdet component (pixel decoder, decoder layer,
  backbone, mask2bbox, bbox2result) the stub delegates to the oracle's restatement
  of that component, so these vectors pin the reference's orchestration / post-processing,
  not mmcv/mmdet internals.

Inputs and weights are regenerated from seeds by the tests (openpvsg_b200.synthetic +
torch.Generator); each file stores a checksum of them so RNG drift is detected.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from openpvsg_b200 import synthetic as syn  # noqa: E402
from oracle import m2f as om  # noqa: E402  (only for the L0 components behind the stubs)


def checksum(*tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors))


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ----------------------------------------------------------------------------------
# relation head: reference classes imported by path
# ----------------------------------------------------------------------------------
def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_relation():
    base = load_by_path('ref_rel_base', f'{REF}/models/relation_head/base.py')
    tr = load_by_path('ref_rel_transformer', f'{REF}/models/relation_head/transformer.py')
    tu = load_by_path('ref_rel_test_utils', f'{REF}/models/relation_head/test_utils.py')
    tru = load_by_path('ref_rel_train_utils', f'{REF}/models/relation_head/train_utils.py')
    sds = syn.relation_state_dicts(seed=1)
    sub_enc = base.ObjectEncoder(feature_dim=256).eval()
    obj_enc = base.ObjectEncoder(feature_dim=256).eval()
    ppn = base.PairProposalNetwork(256, 1024).eval()
    rel = tr.TemporalTransformer(512, 57).eval()
    van = base.VanillaModel(512, 57).eval()
    sub_enc.load_state_dict(sds['subject_encoder'])
    obj_enc.load_state_dict(sds['object_encoder'])
    ppn.load_state_dict(sds['pair_proposal_model'])
    rel.load_state_dict(sds['relation_model'])
    van.load_state_dict({k: v for k, v in sds['relation_model'].items()
                         if k.split('.')[0] in ('fc1', 'fc2', 'span_head', 'pred_head')})
    N, T, P = 14, 12, 30
    feats = randn(11, N, T, 256)
    feats[3, 5:] = 0.0  # absent frames are zero rows (utils/relation_matching.py:431-442)
    with torch.no_grad():
        sub = sub_enc(feats)
        obj = obj_enc(feats)
        pm = ppn(sub, obj)
        pairs = tu.pick_top_pairs_eval(pm, P)
        cat = tru.concatenate_sub_obj(sub, obj, pairs)
        span, prob = rel(cat)
        vspan, vprob = van(cat)
        res_pw = tu.generate_pairwise_results(span, prob, pairs)
        res_all = tu.generate_results(span, prob, pairs)
    np.savez_compressed(
        f'{HERE}/rel_small.npz', N=N, T=T, P=P, feats_seed=11, weights_seed=1,
        in_checksum=checksum(feats, *[v for sd in sds.values() for v in sd.values()
                                      if v.dtype.is_floating_point]),
        sub=sub.numpy(), obj=obj.numpy(), pred_matrix=pm.numpy(), pairs=np.array(pairs),
        cat=cat.numpy(), span=span.numpy(), prob=prob.numpy(), vspan=vspan.numpy(), vprob=vprob.numpy(),
        pw_triplets=np.array([[r['subject_index'], r['object_index'], r['relation']] for r in res_pw]),
        pw_spans=np.array([r['relation_span'] for r in res_pw]).astype(np.uint8),
        all_triplets=np.array([[r['subject_index'], r['object_index'], r['relation']] for r in res_all[:200]]),
    )
    print('rel_small.npz', len(pairs), 'pairs')
    # the two baseline relation models (tools/rel_test.py:167-175 `--model-name filter | conv`), reference classes
    conv = load_by_path('ref_rel_convolution', f'{REF}/models/relation_head/convolution.py')
    bsd = syn.relation_baseline_state_dicts(seed=2)
    filt = conv.HandcraftedFilter(512, 57).eval()
    lconv = conv.Learnable1DConv(512, 57).eval()
    filt.load_state_dict(bsd['filter'])
    lconv.load_state_dict(bsd['conv'])
    x = randn(12, 9, 21, 512)
    x[2, 7:] = 0.0
    with torch.no_grad():
        fs, fp = filt(x)
        cs, cp = lconv(x)
    np.savez_compressed(f'{HERE}/rel_baselines.npz', P=9, T=21, x_seed=12, weights_seed=2,
                        in_checksum=checksum(x, *[v for sd in bsd.values() for v in sd.values()]),
                        fspan=fs.numpy(), fprob=fp.numpy(), cspan=cs.numpy(), cprob=cp.numpy())
    print('rel_baselines.npz')


# ----------------------------------------------------------------------------------
# import stubs for mmcv / mmdet / pycocotools / unitrack
# ----------------------------------------------------------------------------------
class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


class _Any:
    """Callable / subscriptable placeholder for names that are imported but never run."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Any()


class _BaseFusion(nn.Module):
    # mmdet BasePanopticFusionHead.__init__ (L0): stores the class counts and test_cfg
    def __init__(self, num_things_classes=80, num_stuff_classes=53, test_cfg=None,
                 loss_panoptic=None, init_cfg=None, **kwargs):
        super().__init__()
        self.num_things_classes = num_things_classes
        self.num_stuff_classes = num_stuff_classes
        self.num_classes = num_things_classes + num_stuff_classes
        self.test_cfg = test_cfg


def _bbox2result(bboxes, labels, num_classes):
    # mmdet.core.bbox2result (L0)
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    bboxes = bboxes.detach().cpu().numpy()
    labels = labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


def install_stubs():
    def stub(name, **attrs):
        m = _StubModule(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    ident_deco = lambda *a, **k: (lambda f: f)  # noqa: E731
    stub('mmcv')
    stub('mmcv.cnn', Conv2d=nn.Conv2d)
    stub('mmcv.cnn.bricks')
    stub('mmcv.cnn.bricks.transformer', POSITIONAL_ENCODING=_Registry())
    stub('mmcv.ops')
    class _BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

    stub('mmcv.runner', ModuleList=nn.ModuleList, BaseModule=_BaseModule, force_fp32=ident_deco)
    stub('mmdet')
    stub('mmdet.utils')
    stub('mmdet.core', INSTANCE_OFFSET=1000, bbox2result=_bbox2result)
    stub('mmdet.core.visualization')
    stub('mmdet.core.evaluation')
    stub('mmdet.core.evaluation.panoptic_utils', INSTANCE_OFFSET=1000)
    stub('mmdet.core.mask', mask2bbox=om.mask2bbox)
    stub('mmdet.models')
    stub('mmdet.models.utils')
    stub('mmdet.models.builder', HEADS=_Registry(), DETECTORS=_Registry())
    stub('mmdet.models.dense_heads')
    stub('mmdet.models.dense_heads.anchor_free_head', AnchorFreeHead=type('AnchorFreeHead', (nn.Module,), {}))
    stub('mmdet.models.dense_heads.maskformer_head', MaskFormerHead=type('MaskFormerHead', (nn.Module,), {}))
    stub('mmdet.models.detectors')
    stub('mmdet.models.detectors.single_stage',
         SingleStageDetector=type('SingleStageDetector', (nn.Module,), {}))
    stub('mmdet.models.seg_heads')
    stub('mmdet.models.seg_heads.panoptic_fusion_heads')
    stub('mmdet.models.seg_heads.panoptic_fusion_heads.base_panoptic_fusion_head',
         BasePanopticFusionHead=_BaseFusion)
    stub('pycocotools')
    stub('pycocotools.mask')
    # package shells so relative imports inside the reference resolve without running models/__init__.py
    for name, path in (('models', f'{REF}/models'),
                       ('models.mask2former', f'{REF}/models/mask2former'),
                       ('models.mask2former_vps', f'{REF}/models/mask2former_vps')):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    stub('models.unitrack')
    stub('models.unitrack.utils')
    stub('models.unitrack.utils.log')
    stub('models.unitrack.utils.meter')
    stub('models.unitrack.utils.visualize')
    stub('models.unitrack.utils.io')


class _Layer(nn.Module):
    """Stands in for one mmdet DetrTransformerDecoderLayer: delegates to the oracle (L0)."""
    def __init__(self, sd, prefix):
        super().__init__()
        self.sd, self.prefix = sd, prefix

    def forward(self, query, key, value, query_pos, key_pos, attn_masks, **kw):
        return om.decoder_layer(self.sd, self.prefix, query, key, value, query_pos, key_pos, attn_masks[0])


def build_ref_head(cls, sd, pe):
    """Instantiate a reference head class without its mmcv-dependent __init__ and attach
    the parameters from ``sd`` (mmdet key names)."""
    ph = 'panoptic_head.'
    head = cls.__new__(cls)
    nn.Module.__init__(head)
    head.num_heads = 8
    head.num_queries = 100
    head.num_transformer_feat_level = 3
    head.num_transformer_decoder_layers = 9
    head.loss_sem_seg = None
    td = nn.Module()
    td.post_norm = nn.LayerNorm(256)
    td.layers = nn.ModuleList([_Layer(sd, f'{ph}transformer_decoder.layers.{i}.') for i in range(9)])
    head.transformer_decoder = td
    head.cls_embed = nn.Linear(256, 127)
    head.mask_embed = nn.Sequential(nn.Linear(256, 256), nn.ReLU(inplace=True), nn.Linear(256, 256),
                                    nn.ReLU(inplace=True), nn.Linear(256, 256))
    head.query_embed = nn.Embedding(100, 256)
    head.query_feat = nn.Embedding(100, 256)
    head.level_embed = nn.Embedding(3, 256)
    head.decoder_input_projs = nn.ModuleList([nn.Identity() for _ in range(3)])
    head.decoder_positional_encoding = pe
    own = {k[len(ph):]: v for k, v in sd.items() if k.startswith(ph) and 'pixel_decoder' not in k
           and '.layers.' not in k}
    missing = head.load_state_dict(own, strict=False)
    assert not missing.unexpected_keys, missing
    head.pixel_decoder = lambda feats: om.pixel_decoder(sd, feats, ph + 'pixel_decoder.')
    head.eval()
    return head


def golden_m2f():
    install_stubs()
    fusion_mod = importlib.import_module('models.mask2former.mask2former_fusion_head')
    head_mod = importlib.import_module('models.mask2former.mask2former_head')
    vhead_mod = importlib.import_module('models.mask2former_vps.mask2former_video_head')
    pe_mod = importlib.import_module('models.mask2former_vps.position_encoding')
    minvis_mod = importlib.import_module('models.mask2former_vps.mask2former_min_vis')
    vdet_mod = importlib.import_module('models.mask2former_vps.mask2former')
    idet_mod = importlib.import_module('models.mask2former.mask2former')

    test_cfg = dict(panoptic_on=True, semantic_on=False, instance_on=True, max_per_image=100,
                    iou_thr=0.8, filter_low_score=True, object_mask_thr=0.8, return_query=True)

    # ---- (a) fusion-head post-processing on crafted logits (things + stuff + skipped) ----
    Q, H, W = 24, 40, 56
    g = torch.Generator().manual_seed(21)
    mask_cls = torch.randn(Q, 127, generator=g)
    conf = torch.tensor([3, 117, 7, 7, 120, 126, 50, 3, 118, 117, 60, 61])  # confident classes
    for q, c in enumerate(conf.tolist()):
        mask_cls[q, c] += 12.0
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                            indexing='ij')
    mask_pred = torch.randn(Q, H, W, generator=g) * 0.5 - 3.0
    for q in range(Q):
        cy, cx = ((q % 12) // 4) * 13 + 6, (q % 4) * 14 + 7
        r = 3 + (q % 3)
        mask_pred[q] += 8.0 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))
    mask_pred[2] = mask_pred[3] + 0.01  # two queries competing for the same blob
    query_feats = torch.randn(Q, 256, generator=g)
    fh = fusion_mod.MaskFormerFusionHeadCustom(115, 11, test_cfg=test_cfg)
    with torch.no_grad():
        pan, qfd = fh.panoptic_postprocess_with_query(mask_cls, mask_pred, query_feats)
        labels, bboxes, binm = fh.instance_postprocess(mask_cls, mask_pred)
        meta = dict(img_shape=(36, 50, 3), ori_shape=(36, 50, 3))
        res = fh.simple_test_with_query(mask_cls[None], mask_pred[None], query_feats[None], [meta],
                                        rescale=True)[0]
    order = np.argsort(labels.numpy(), kind='stable')
    np.savez_compressed(
        f'{HERE}/fusion_post.npz', mask_cls=mask_cls.numpy(), mask_pred=mask_pred.numpy(),
        query_feats=query_feats.numpy(), pan=pan.numpy(),
        qf_keys=np.array(sorted(qfd.keys())),
        qf_vals=np.stack([qfd[k][0].numpy() for k in sorted(qfd.keys())]),
        qf_counts=np.array([len(qfd[k]) for k in sorted(qfd.keys())]),
        ins_labels=labels.numpy(), ins_bboxes=bboxes.numpy(),
        ins_masks=np.packbits(binm.numpy(), axis=-1), ins_order=order,
        crop_pan=res['pan_results'].numpy(), crop_ins_labels=res['ins_results'][0].numpy(),
        crop_ins_bboxes=res['ins_results'][1].numpy())
    print('fusion_post.npz segments:', sorted(qfd.keys()), 'ins', tuple(labels.shape))

    # ---- (b) forward_head / forward_head_video ----
    sd = syn.mask2former_state_dict(seed=3)
    pe3d = pe_mod.SinePositionalEncoding3D(num_feats=128, normalize=True)
    vhead = build_ref_head(vhead_mod.Mask2FormerVideoHead, sd, pe3d)
    # mmdet SinePositionalEncoding is L0 -> oracle restatement behind the stub
    ihead = build_ref_head(head_mod.Mask2FormerHeadCustom, sd, lambda mask: om.sine_pe_2d(*mask.shape))
    dec_out = randn(31, 100, 2, 256)
    mf_img = randn(32, 2, 256, 12, 20)
    mf_vid = randn(33, 2, 3, 256, 12, 20)
    with torch.no_grad():
        c1, m1, a1 = ihead.forward_head(dec_out, mf_img, (3, 5))
        c2, m2, a2 = vhead.forward_head_video(dec_out, mf_vid, (6, 10))
    np.savez_compressed(f'{HERE}/forward_head.npz', weights_seed=3,
                        in_checksum=checksum(dec_out, mf_img, mf_vid),
                        cls_img=c1.numpy(), mask_img=m1.numpy(), attn_img=np.packbits(a1.numpy(), axis=-1),
                        cls_vid=c2.numpy(), mask_vid=m2.numpy(), attn_vid=np.packbits(a2.numpy(), axis=-1),
                        attn_img_shape=np.array(a1.shape), attn_vid_shape=np.array(a2.shape))

    # ---- (c) SinePositionalEncoding3D ----
    with torch.no_grad():
        pos = pe3d(torch.zeros(1, 2, 5, 7, dtype=torch.bool))
    np.savez_compressed(f'{HERE}/pe3d.npz', pos=pos.numpy())

    # ---- (d) match_from_embds ----
    tgt = randn(41, 100, 256)
    cur = tgt[torch.randperm(100, generator=torch.Generator().manual_seed(42))] + 0.3 * randn(43, 100, 256)
    idx = minvis_mod.Mask2FormerVideoCustomMinVIS.match_from_embds(None, tgt, cur)
    np.savez_compressed(f'{HERE}/match_embds.npz', indices=np.asarray(idx), in_checksum=checksum(tgt, cur))

    # ---- (e) full head forward + detector simple_test (VPS, T=1) and IPS head ----
    Himg, Wimg = 96, 160
    img = syn.synthetic_frame(5, Himg, Wimg)[None]
    meta = syn.frame_meta(Himg, Wimg)
    feats = om.resnet50(sd, img)
    with torch.no_grad():
        vc, vm, vq = vhead.forward(list(feats), [[meta]], return_query=True)
        ic, im, iq = ihead.forward(list(feats), [meta], return_query=True)
        vcls, vmask, vqf = vhead.simple_test_with_query(list(feats), [[meta]])
        icls, imask, iqf = ihead.simple_test_with_query(list(feats), [meta])
    np.savez_compressed(
        f'{HERE}/head_forward.npz', weights_seed=3, frame_seed=5, H=Himg, W=Wimg,
        in_checksum=checksum(img),
        v_cls_last=vc[-1].numpy(), v_mask_last=vm[-1].numpy(), v_query=vq.numpy(),
        v_cls_mid=vc[4].numpy(), v_mask_first=vm[0].numpy(),
        i_cls_last=ic[-1].numpy(), i_mask_last=im[-1].numpy(), i_query=iq.numpy(),
        v_up_shape=np.array(vmask.shape), i_up_shape=np.array(imask.shape),
        v_up_sample=vmask[0, 0, ::7, ::5, ::5].numpy(), i_up_sample=imask[0, ::7, ::5, ::5].numpy(),
        v_qf_shape=np.array(vqf.shape), i_qf_shape=np.array(iqf.shape))

    # detectors: reference simple_test with backbone = oracle resnet, heads = reference heads
    vdet = vdet_mod.Mask2FormerVideoCustom.__new__(vdet_mod.Mask2FormerVideoCustom)
    nn.Module.__init__(vdet)
    vdet.extract_feat = lambda x: om.resnet50(sd, x)
    vdet.panoptic_head = vhead
    vdet.panoptic_fusion_head = fusion_mod.MaskFormerFusionHeadCustom(115, 11, test_cfg=test_cfg)
    vdet.num_things_classes, vdet.num_stuff_classes = 115, 11
    with torch.no_grad():
        vres = vdet.simple_test(img, [meta], ref_img=img[None], ref_img_metas=[[meta]], rescale=True)
    r = vres[0][0]
    idet = idet_mod.Mask2FormerCustom.__new__(idet_mod.Mask2FormerCustom)
    nn.Module.__init__(idet)
    idet.extract_feat = lambda x: om.resnet50(sd, x)
    idet.panoptic_head = ihead
    idet.panoptic_fusion_head = vdet.panoptic_fusion_head
    idet.num_things_classes, idet.num_stuff_classes = 115, 11
    with torch.no_grad():
        ires = idet.simple_test(img, [meta], rescale=True)[0]
    vkeys = sorted(r['query_feats'].keys())
    ikeys = sorted(ires['query_feats'].keys())
    np.savez_compressed(
        f'{HERE}/detector.npz', weights_seed=3, frame_seed=5, H=Himg, W=Wimg,
        v_pan=r['pan_results'], v_keys=np.array(vkeys),
        v_feats=np.stack([np.asarray(r['query_feats'][k][0]) for k in vkeys]) if vkeys else np.zeros((0, 256)),
        v_ins_boxes=np.concatenate(r['ins_results'][0], 0),
        v_ins_counts=np.array([len(x) for x in r['ins_results'][0]]),
        i_pan=ires['pan_results'], i_keys=np.array(ikeys),
        i_feats=np.stack([np.asarray(ires['query_feats'][k][0]) for k in ikeys]) if ikeys else np.zeros((0, 1, 256)),
        i_ins_boxes=np.concatenate(ires['ins_results'][0], 0),
        i_ins_counts=np.array([len(x) for x in ires['ins_results'][0]]))
    print('detector.npz VPS segments', vkeys, 'IPS segments', ikeys)


def golden_full_frames():
    """ORACLE-derived (not reference-derived) fixtures at the BASELINE full sizes: the CPU oracle
    takes ~10-40 s per frame at these sizes, so its outputs are stored once here instead of
    being recomputed on the GPU box.  The oracle itself is pinned by the vectors above."""
    sd = syn.mask2former_state_dict(seed=3)
    for name, (H, W) in (('frame_480x640', (480, 640)), ('frame_720x1280', (720, 1280))):
        img = syn.synthetic_frame(17, H, W)[None]
        meta = syn.frame_meta(H, W)
        with torch.no_grad():
            feats = om.resnet50(sd, img)
            cls, masks, query, ex = om.head_forward(sd, feats, video=True, num_frames=1, return_all=True)
            res = om.vps_simple_test(sd, img[None], [[meta]], instance_on=False)[0][0]
        keys = sorted(res['query_feats'].keys())
        near = [int((torch.nn.functional.interpolate(masks[i].flatten(0, 1), tuple(ex['memories'][i % 3].shape[-2:]),
                                                     mode='bilinear', align_corners=False).abs() < 1e-4).sum())
                for i in range(9)]
        np.savez_compressed(
            f'{HERE}/{name}.npz', weights_seed=3, frame_seed=17, H=H, W=W, in_checksum=checksum(img),
            c5_sample=feats[3][0, ::16, ::3, ::3].numpy(), c2_sample=feats[0][0, ::16, ::9, ::9].numpy(),
            mask_feature_sample=ex['mask_features'][0, 0, ::8, ::6, ::6].numpy(),
            cls_all=torch.stack([c[0] for c in cls]).numpy(), query=query.numpy(),
            mask_last_sample=masks[-1][0, 0, :, ::4, ::4].numpy(),
            mask_mid_sample=masks[4][0, 0, :, ::8, ::8].numpy(),
            attn_masks=np.concatenate([np.packbits(a[0].numpy().reshape(-1)) for a in ex['attn_masks']]),
            attn_shapes=np.array([list(a[0].shape) for a in ex['attn_masks']]),
            near_zero_mask_logits=np.array(near),
            pan=res['pan_results'], keys=np.array(keys),
            qfeats=np.stack([np.asarray(res['query_feats'][k][0]) for k in keys]) if keys else np.zeros((0, 256)))
        print(name, 'segments', keys, 'near-zero mask logits per layer', near)

# ----------------------------------------------------------------------------------
# Swin-B stage outputs at the BASELINE full size (ORACLE-derived, like frame_720x1280.npz: the reference has no
# Swin code; oracle/swin.py is pinned to torchvision in tests/test_swin.py).  Run: python tests/golden/make_golden.py swin
# ----------------------------------------------------------------------------------
def golden_swin_b_720p():
    from openpvsg_b200 import configs
    from oracle import swin as osw
    g = torch.Generator().manual_seed(5)
    sd = syn.swin_state_dict(g, prefix='', **configs.SWIN_B)
    img = syn.synthetic_frame(21, 720, 1280)[None]
    with torch.no_grad():
        outs = osw.swin_forward(sd, img, **configs.SWIN_B)
    data = dict(input_checksum=np.float64(checksum(img)), weight_checksum=np.float64(checksum(*[v.float() for v in sd.values()])))
    for i, o in enumerate(outs):
        data[f'stage{i}_shape'] = np.array(o.shape)
        data[f'stage{i}_sub'] = o[0, ::8, ::4, ::4].numpy().copy()          # every 8th channel, every 4th pixel
        data[f'stage{i}_abs_sum'] = np.float64(o.double().abs().sum().item())
    np.savez_compressed(os.path.join(HERE, 'swin_b_720x1280.npz'), **data)
    print('swin_b_720x1280.npz', {k: (v.shape if hasattr(v, 'shape') else v) for k, v in data.items()})



if __name__ == '__main__':
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == 'frames':
        golden_full_frames()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'swin':
        golden_swin_b_720p()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'relation':
        golden_relation()
        sys.exit(0)
    golden_relation()
    golden_m2f()
    golden_full_frames()
    golden_swin_b_720p()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')

