"""CPU: the reference's own tools run UNMODIFIED on top of openpvsg_b200/compat.

`tools/rel_test.py` and `tools/test.py` are imported by path from /root/reference (build container only; skipped
elsewhere) with the compat packages `mmcv`, `mmdet`, `models`, `datasets`, `utils` first on sys.path.  There is no GPU in
this container and the product has no CPU fallback, so for the numeric check the relation modules' device forward is
replaced by a TEST DOUBLE that calls the CPU oracle; everything else -- imports, constructors, state_dict loading, the
DataLoader over the product's PVSGRelationDataset, pair selection, the metric and CSV code -- is the real thing, and the
R@K the reference's evaluate() writes must equal the golden numbers of the reference's own run (releval.json)."""
import csv
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

import relset_fixture as fx
from openpvsg_b200 import relation_set as rs, synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
COMPAT = os.path.join(ROOT, 'openpvsg_b200', 'compat')
REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'tools', 'rel_test.py')),
                               reason='reference tree only exists in the build container')


@pytest.fixture
def compat_path():
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in [m for m in sys.modules if m.split('.')[0] in ('mmcv', 'mmdet', 'models', 'datasets', 'utils')]:
        del sys.modules[name]
    sys.path.insert(0, COMPAT)
    yield
    sys.path[:] = saved_path
    for name in [m for m in sys.modules if m.split('.')[0] in ('mmcv', 'mmdet', 'models', 'datasets', 'utils')]:
        del sys.modules[name]
    sys.modules.update({k: v for k, v in saved_mods.items() if k.split('.')[0] in ('mmcv', 'mmdet', 'models', 'datasets', 'utils')})


def _load_tool(name):
    spec = importlib.util.spec_from_file_location(f'ref_tool_{name}', os.path.join(REF, 'tools', f'{name}.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_ref
def test_rel_test_tool_runs_unmodified(compat_path, tmp_path, monkeypatch):
    from oracle import relation as orel
    tool = _load_tool('rel_test')
    from openpvsg_b200 import relation_head as rh
    assert tool.ObjectEncoder is rh.ObjectEncoder and tool.TemporalTransformer is rh.TemporalTransformer
    assert tool.HandcraftedFilter is rh.HandcraftedFilter and tool.PVSGRelationDataset is rs.PVSGRelationDataset
    golden = json.load(open(os.path.join(HERE, 'golden', 'releval.json')))
    sds = syn.relation_state_dicts(seed=1)
    mods = dict(subject_encoder=tool.ObjectEncoder(feature_dim=256), object_encoder=tool.ObjectEncoder(feature_dim=256),
                pair_proposal_model=tool.PairProposalNetwork(256, 1024), relation_model=tool.TemporalTransformer(512, 57))
    for k, m in mods.items():
        m.load_state_dict(sds[k])                       # what the tool's __main__ does with torch.load(...)[k]
    # TEST DOUBLE (no GPU here): device forwards -> CPU oracle on the same state_dicts
    monkeypatch.setattr(rh.ObjectEncoder, 'forward', lambda self, x: orel.object_encoder(
        {k: v for k, v in self.state_dict().items()}, x))
    monkeypatch.setattr(rh.PairProposalNetwork, 'forward', lambda self, s, o: orel.pair_proposal(self.state_dict(), s, o))
    monkeypatch.setattr(rh.TemporalTransformer, 'forward', lambda self, x: orel.temporal_transformer(self.state_dict(), x))
    monkeypatch.setattr(rh, 'pick_top_pairs_eval', orel.pick_top_pairs_eval)
    monkeypatch.setattr(tool, 'pick_top_pairs_eval', orel.pick_top_pairs_eval)
    monkeypatch.setattr(tool, 'concatenate_sub_obj', orel.concatenate_sub_obj)
    monkeypatch.setattr(tool, 'generate_pairwise_results', orel.generate_pairwise_results)
    monkeypatch.setattr(tool, 'generate_results', orel.generate_results)
    clip = fx.make_clip()
    linker = fx.link(clip)
    feats = rs.process_feats({t.track_id: t.qf_tube for t in rs.query_feat_tubes(linker)})
    rels = [dict(subject_index=g['subject_index'], object_index=g['object_index'], relation=g['relation'],
                 relation_span=np.array(g['relation_span'])) for g in golden['gt_relations']]
    ds = tool.PVSGRelationDataset(fx.make_anno(), 'train', memory={fx.VID: dict(feats=feats, relations=rels)})
    loader = tool.DataLoader(ds, batch_size=1, shuffle=False)
    csv_path = str(tmp_path / 'full_result.csv')
    tool.evaluate(mods['subject_encoder'], mods['object_encoder'], mods['pair_proposal_model'], mods['relation_model'],
                  loader, 100, [f'relation_{i}' for i in range(57)], torch.device('cpu'), csv_path, 'transformer_test')
    rows = list(csv.reader(open(csv_path)))
    assert rows[0][:2] == ['Model', 'Pair Recall'] and rows[1][0] == 'transformer_test'
    fm = golden['final_metrics']
    want = [f"{100 * fm[str(K)]['recall']:.2f}/{100 * fm[str(K)]['mean_recall']:.2f}" for K in (20, 50, 100)]
    assert rows[1][2:5] == want, (rows[1], want)
    assert rows[1][1] == f"{100 * np.mean(golden['pair_recall_list']):.2f}"


@needs_ref
def test_test_tool_imports_and_plumbing(compat_path, tmp_path, monkeypatch):
    """tools/test.py: imports resolve, the reference's config builds the B200 detector, the synthetic dataset /
    loader produce the forward_test call signature, and single_gpu_test collects what the model returns (the model's
    device forward is replaced by a recorder: no GPU in this container)."""
    tool = _load_tool('test')
    cfg = tool.Config.fromfile(os.path.join(REF, 'configs/mask2former_vps/mask2former_video_r50_single_video_test.py'))
    cfg.merge_from_dict({'data.test.type': 'SyntheticVPSDataset', 'data.test.num_frames': 3})
    cfg = tool.compat_cfg(cfg)
    cfg.data.test.test_mode = True
    dataset = tool.build_dataset(cfg.data.test)
    loader = tool.build_dataloader(dataset, samples_per_gpu=1, workers_per_gpu=0, dist=False, shuffle=False)
    cfg.model.train_cfg = None
    model = tool.build_detector(cfg.model, test_cfg=cfg.get('test_cfg'))
    assert type(model).__name__ == 'Mask2FormerVideoCustom'
    ckpt = str(tmp_path / 'ckpt.pth')
    torch.save(dict(state_dict=syn.mask2former_state_dict(seed=1), meta=dict(CLASSES=dataset.CLASSES)), ckpt)
    meta = tool.load_checkpoint(model, ckpt, map_location='cpu')
    assert 'CLASSES' in meta['meta']
    assert tool.fuse_conv_bn(model) is model
    seen = []

    def fake_simple_test(self, img, img_metas, ref_img, ref_img_metas, **kw):
        assert torch.is_tensor(ref_img) and ref_img.dim() == 5 and ref_img.shape[1] == 1
        assert ref_img_metas[0][0]['batch_input_shape'] == tuple(ref_img.shape[-2:]) and kw.get('rescale') is True
        seen.append(tuple(ref_img.shape))
        return [[dict(pan_results=np.zeros(ref_img.shape[-2:], np.int32))] for _ in range(ref_img.shape[0])]

    monkeypatch.setattr(type(model), 'simple_test', fake_simple_test)
    wrapped = tool.build_dp(model, 'cpu', device_ids=[0])
    outputs = tool.single_gpu_test(wrapped, loader, False, None, 0.3)
    assert len(outputs) == 3 and seen == [(1, 1, 3, 96, 160)] * 3
    tool.mmcv.dump(outputs, str(tmp_path / 'out.pkl'))
    assert len(tool.mmcv.load(str(tmp_path / 'out.pkl'))) == 3
    args = tool.parse_args.__globals__['argparse'].Namespace()      # DictAction parses --cfg-options like mmcv's
    tool.DictAction(['--cfg-options'], 'cfg_options')(None, args, ['a.b=1', 'c=[1,2]', 'd=true', 'e=x'])
    assert args.cfg_options == {'a.b': 1, 'c': [1, 2], 'd': True, 'e': 'x'}


@needs_ref
def test_train_tool_main_runs_unmodified(compat_path, tmp_path, monkeypatch):
    """tools/train.py main(), unmodified, on the reference's own VPS training config: every import resolves in compat, the
    config builds the B200 detector and the synthetic training set, and the tool reaches train_detector with them.  The
    model's train_step is replaced by a recorder (no GPU in this container); train_detector itself -- loader, optimizer
    groups, schedule, clipping, checkpoint -- is the real compat code (its GPU run: tests/test_training_slice.py)."""
    tool = _load_tool('train')
    calls = []

    def fake_train_step(self, data, optimizer=None):
        assert data['ref_img'].dim() == 5 and data['ref_img'].shape[:2] == (2, 2) and len(data['ref_gt_masks']) == 2
        assert data['ref_gt_labels'][0].shape[1] == 2 and data['ref_gt_instance_ids'][0].shape == data['ref_gt_labels'][0].shape
        loss = sum((p ** 2).sum() for n, p in self.named_parameters() if n.endswith('query_feat.weight'))
        calls.append(float(loss.detach()))
        return dict(loss=loss, log_vars=dict(loss=float(loss.detach())), num_samples=len(data['img_metas']))

    from openpvsg_b200.mask2former import Mask2FormerVideoCustom
    monkeypatch.setattr(Mask2FormerVideoCustom, 'train_step', fake_train_step)
    work = str(tmp_path / 'work')
    argv = ['train.py', os.path.join(REF, 'configs/mask2former_vps/mask2former_video_r50.py'), '--work-dir', work, '--no-validate',
            '--seed', '5', '--cfg-options', 'data.train.type=SyntheticVPSDataset', 'data.train.num_frames=4', 'data.train.test_mode=False',
            'data.samples_per_gpu=2', 'data.workers_per_gpu=0', 'runner.max_epochs=1', 'log_config.interval=1', 'device=cpu',
            'load_from=None']
    monkeypatch.setattr(sys, 'argv', argv)
    monkeypatch.setattr(tool, 'get_device', lambda: 'cpu')
    tool.main()
    assert len(calls) == 2                                     # 4 clips / 2 per batch, 1 epoch
    assert os.path.exists(os.path.join(work, 'epoch_1.pth')) and os.path.exists(os.path.join(work, 'mask2former_video_r50.py'))
    ck = torch.load(os.path.join(work, 'epoch_1.pth'), map_location='cpu', weights_only=False)
    assert ck['meta']['seed'] == 5 and len(ck['meta']['CLASSES']) == 126 and ck['meta']['iter'] == 2
