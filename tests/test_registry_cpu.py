"""CPU: the registry / config boundary -- reference config files build unmodified and
reference-layout checkpoints load strictly."""
import os

import pytest
import torch

import openpvsg_b200 as pv
from openpvsg_b200 import configs, synthetic as syn

REF_CFG = '/root/reference/configs/mask2former_vps/mask2former_video_r50_single_video_test.py'


def test_registered_names():
    # reference models/__init__.py:1-12 (inference-relevant names) + L0 type strings of its configs
    for n in ('Mask2FormerCustom', 'Mask2FormerVideoCustom'):
        assert n in pv.DETECTORS
    for n in ('Mask2FormerHeadCustom', 'Mask2FormerVideoHead', 'MaskFormerFusionHeadCustom', 'MaskFormerFusionHead'):
        assert n in pv.HEADS
    assert 'ResNet' in pv.BACKBONES
    for n in ('SinePositionalEncoding', 'SinePositionalEncoding3D'):
        assert n in pv.POSITIONAL_ENCODING


@pytest.mark.parametrize('video', [False, True])
def test_build_and_strict_load(video):
    det = pv.build_detector(configs.mask2former_r50(video))
    sd = syn.mask2former_state_dict(seed=1)
    res = det.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert not det.training
    assert det.train().training and not det.eval().training     # mode flag only: no batch statistics, no dropout
    with pytest.raises(NotImplementedError):
        det.aug_test(None, None)
    with pytest.raises(KeyError):
        pv.build_detector(dict(type='NoSuchDetector'))


def test_relation_state_dicts_load():
    sds = syn.relation_state_dicts(seed=0)
    pv.ObjectEncoder(feature_dim=256).load_state_dict(sds['subject_encoder'])
    pv.PairProposalNetwork(256, 1024).load_state_dict(sds['pair_proposal_model'])
    pv.TemporalTransformer(512, 57).load_state_dict(sds['relation_model'])
    for cls in (pv.VanillaModel, pv.HandcraftedFilter, pv.Learnable1DConv):
        cls(512, 57)


def test_no_fallback_without_cuda():
    """The product path must fail loudly, not fall back, when there is no CUDA tensor."""
    from openpvsg_b200 import lib, ops
    with pytest.raises(lib.PvsgError):
        ops.layernorm(torch.zeros(2, 256), torch.ones(256), torch.zeros(256))


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason='reference tree only exists in the build container')
def test_reference_config_file_builds_unmodified():
    cfg = pv.load_config(REF_CFG)
    assert cfg['model']['type'] == 'Mask2FormerVideoCustom'
    assert cfg['data']['test']['ref_seq_len_test'] == 1
    det = pv.build_detector(cfg['model'])
    assert type(det).__name__ == 'Mask2FormerVideoCustom'
    assert det.load_state_dict(syn.mask2former_state_dict(seed=1), strict=True)
    ours = configs.mask2former_r50(True)
    ref = cfg['model']
    assert ref['test_cfg'] == ours['test_cfg']
    for k in ('num_queries', 'num_things_classes', 'num_stuff_classes', 'feat_channels'):
        assert ref['panoptic_head'][k] == ours['panoptic_head'][k]
    ips = pv.load_config('/root/reference/configs/mask2former/'
                         'mask2former_r50_lsj_8x2_50e_coco-panoptic_custom_single_video_test.py')
    assert type(pv.build_detector(ips['model'])).__name__ == 'Mask2FormerCustom'


def test_preprocess_gt_matches_reference_golden():
    """Mask2FormerVideoHead.preprocess_gt against the reference's own preprocess_video_panoptic_gt (golden vectors written
    by tests/golden/make_golden_train_gt.py from models/mask2former_vps/utils.py:94-140): instance order, labels, per-frame
    masks, empty masks for frames without the instance, padding to pad_shape.  Pure index work: runs on the CPU."""
    import json
    import os
    import torch
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_gt.json')))
    head = pv.build_detector(configs.mask2former_r50(True)).panoptic_head
    for c in golden:
        metas = [dict(pad_shape=(c['pad'][0], c['pad'][1], 3)) for _ in c['masks']]
        labels, masks = head.preprocess_gt([torch.tensor(c['labels'])], [[torch.tensor(m, dtype=torch.uint8) for m in c['masks']]], None,
                                           [torch.tensor(c['ids'])], [metas])
        assert labels[0].tolist() == c['out_labels'] and labels[0].dtype == torch.int64
        assert masks[0].tolist() == c['out_masks'] and masks[0].dtype == torch.int64
    # a clip batch of two: one result per clip, in order
    a, b = golden[0], golden[1]
    labels, masks = head.preprocess_gt([torch.tensor(a['labels']), torch.tensor(b['labels'])],
                                       [[torch.tensor(m, dtype=torch.uint8) for m in a['masks']], [torch.tensor(m, dtype=torch.uint8) for m in b['masks']]],
                                       None, [torch.tensor(a['ids']), torch.tensor(b['ids'])],
                                       [[dict(pad_shape=(a['pad'][0], a['pad'][1], 3))] * len(a['masks']),
                                        [dict(pad_shape=(b['pad'][0], b['pad'][1], 3))] * len(b['masks'])])
    assert labels[1].tolist() == b['out_labels'] and masks[0].tolist() == a['out_masks']
