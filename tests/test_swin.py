"""Swin backbone (BASELINE configs[2]).

The reference has no Swin code, so parity is anchored on mmdet 2.25.0's published algorithm (oracle/swin.py,
"parity unpinned" by the reference) and the oracle itself is pinned to an independent implementation available
offline (torchvision's SwinTransformer).  GPU tests compare the CUDA path with the oracle: kernels <= 2e-5,
backbone / detector features <= 1e-3 (the north-star fp32 tolerance)."""
import numpy as np
import pytest
import torch

import openpvsg_b200 as pv
from openpvsg_b200 import configs, synthetic as syn
from oracle import swin as osw

SMALL = dict(embed_dims=128, depths=(2, 2, 2, 2), num_heads=(4, 8, 16, 32), window_size=12)


# ----------------------------------------------------------------------------- CPU ----------
def test_oracle_vs_torchvision():
    """oracle/swin.py (mmdet layout) == torchvision SwinTransformer on every stage, incl. window padding, shifted
    windows with masks, odd-sized patch merging."""
    from torchvision.models.swin_transformer import SwinTransformer
    torch.manual_seed(0)
    depths, heads = [2, 2, 2], [1, 2, 4]
    tv = SwinTransformer(patch_size=[4, 4], embed_dim=32, depths=depths, num_heads=heads, window_size=[4, 4],
                         stochastic_depth_prob=0.0).eval()
    with torch.no_grad():
        for n, p in tv.named_parameters():
            if 'relative_position_bias_table' in n:
                p.normal_(0, 0.5)
            elif p.dim() > 1:
                p.normal_(0, (1.0 / p.shape[1]) ** 0.5)
            elif 'bias' in n:
                p.normal_(0, 0.1)
            else:
                p.uniform_(0.8, 1.2)
    sd = osw.from_torchvision(tv.state_dict(), depths)
    C = 32
    for i in range(3):
        sd[f'norm{i}.weight'], sd[f'norm{i}.bias'] = torch.ones(C), torch.zeros(C)
        C *= 2
    img = torch.randn(2, 3, 72, 84)            # 18 x 21 patches -> padded windows, 9 x 11 -> odd merge, 5 x 6
    taps = []
    with torch.no_grad():
        osw.swin_forward(sd, img, 32, depths, heads, 4, 4, (0, 1, 2), stage_taps=taps)
        x = tv.features[0](img)
        for i in range(3):
            x = tv.features[2 * i + 1](x)
            t, hw = taps[i]
            assert tuple(x.shape[1:3]) == tuple(hw)
            assert (x.reshape(2, -1, x.shape[-1]) - t).abs().max().item() < 2e-5, i
            if i < 2:
                x = tv.features[2 * i + 2](x)


def test_relative_position_index_matches_standard_swin():
    for ws in (4, 7, 12):
        idx = osw.relative_position_index(ws)
        ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing='ij')
        c = torch.stack([ys.flatten(), xs.flatten()])
        rel = c[:, :, None] - c[:, None, :]
        want = (rel[0] + ws - 1) * (2 * ws - 1) + rel[1] + ws - 1
        assert torch.equal(idx, want)


def test_swin_b_builds_and_loads_strictly():
    assert 'SwinTransformer' in pv.BACKBONES
    det = pv.build_detector(configs.mask2former_swin(True))
    sd = syn.mask2former_state_dict(seed=1, in_channels=(128, 256, 512, 1024), backbone=dict(configs.SWIN_B))
    res = det.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(det.state_dict()) == set(sd)
    n = sum(p.numel() for p in det.backbone.parameters())
    assert 86e6 < n < 88e6                      # Swin-B backbone, 86.9 M parameters
    k = 'backbone.stages.2.blocks.17.attn.w_msa.relative_position_index'
    assert torch.equal(det.state_dict()[k], osw.relative_position_index(12))
    with pytest.raises(NotImplementedError):
        pv.build_backbone(dict(type='SwinTransformer', embed_dims=96, num_heads=(3, 6, 12, 25)))
    with pytest.raises(NotImplementedError):      # Swin-T widths: rejected at build time, not at the first forward
        pv.build_backbone(dict(type='SwinTransformer', embed_dims=96, num_heads=(3, 6, 12, 24)))
    with pytest.raises(NotImplementedError):
        pv.build_backbone(dict(type='SwinTransformer', embed_dims=128, num_heads=(4, 8, 16, 32), qkv_bias=False))
    with pytest.raises(pv.lib.PvsgError if hasattr(pv, 'lib') else Exception):
        det.backbone(torch.zeros(1, 3, 96, 96))  # CPU tensor: no fallback


def _golden_720p():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'swin_b_720x1280.npz'))


def _swin_b_720p_inputs():
    g = torch.Generator().manual_seed(5)
    sd = syn.swin_state_dict(g, prefix='', **configs.SWIN_B)
    img = syn.synthetic_frame(21, 720, 1280)[None]
    gold = _golden_720p()
    assert float(img.double().abs().sum()) == pytest.approx(float(gold['input_checksum']), rel=1e-9), 'frame RNG drifted'
    assert float(sum(v.double().abs().sum() for v in sd.values())) == pytest.approx(float(gold['weight_checksum']), rel=1e-9)
    return sd, img, gold


def test_swin_b_720p_golden_matches_oracle():
    """The committed full-size fixture (oracle-derived, tests/golden/make_golden.py swin) is what the oracle computes."""
    sd, img, gold = _swin_b_720p_inputs()
    with torch.no_grad():
        outs = osw.swin_forward(sd, img, **configs.SWIN_B)
    for i, o in enumerate(outs):
        assert list(o.shape) == gold[f'stage{i}_shape'].tolist()
        assert np.abs(o[0, ::8, ::4, ::4].numpy() - gold[f'stage{i}_sub']).max() < 1e-5
        assert float(o.double().abs().sum()) == pytest.approx(float(gold[f'stage{i}_abs_sum']), rel=1e-6)


# ----------------------------------------------------------------------------- GPU ----------
def _close(a, b, tol, name):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f'{name}: max abs err {err:.3e}'


@pytest.mark.gpu
@pytest.mark.parametrize('B,H,W,heads,ws,shift', [(2, 23, 40, 4, 12, 0), (2, 23, 40, 4, 12, 6), (1, 24, 36, 8, 12, 6),
                                                  (3, 5, 7, 2, 4, 2), (1, 9, 9, 2, 3, 1), (1, 7, 10, 1, 7, 3),
                                                  (1, 3, 5, 4, 12, 6)])
def test_window_attention_kernel(B, H, W, heads, ws, shift):
    from openpvsg_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + H + shift)
    C = heads * 32
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    bias = torch.randn(3 * C, generator=g) * 0.3
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    want = osw.window_attention_core(qkv, bias, table, heads, ws, shift)
    got = ops.window_attention(qkv.cuda(), bias.cuda(), table.cuda(), heads, ws, shift)
    # default = tensor-core kernel (split-bf16 products, fp32 accumulate): fp32-grade, gate 1e-4 like the other
    # MMA kernels; PVSG_WINATT_IMPL=1 selects the exact-fp32 SIMT kernel (measured 2e-5 on the same cases)
    _close(got, want, 1e-4, 'window_attention')
    both = ops.window_attention(qkv.cuda(), bias.cuda(), table.cuda(), heads, ws, shift, out_mode='both')
    assert torch.equal(both[0], got)
    if both[1] is not None:       # operand planes: hi = rn_bf16(v), lo = rn_bf16(v - hi), bit-exact from the fp32 result
        hi = got.to(torch.bfloat16)
        assert torch.equal(both[1].hi, hi) and torch.equal(both[1].lo, (got - hi.float()).to(torch.bfloat16))


@pytest.mark.gpu
def test_window_attention_rejects_unsupported():
    from openpvsg_b200 import ops
    from openpvsg_b200.lib import PvsgError
    with pytest.raises(PvsgError):           # head dim 16
        ops.window_attention(torch.zeros(1, 4, 4, 96).cuda(), torch.zeros(96).cuda(), torch.zeros(49, 2).cuda(), 2, 4, 0)
    with pytest.raises(PvsgError):           # CPU tensors
        ops.window_attention(torch.zeros(1, 4, 4, 96), torch.zeros(96), torch.zeros(49, 1), 1, 4, 0)


@pytest.mark.gpu
@pytest.mark.parametrize('B,H,W,C', [(2, 6, 8, 128), (1, 5, 7, 128), (2, 23, 40, 512), (1, 1, 1, 256)])
def test_patch_merge_ln_kernel(B, H, W, C):
    from openpvsg_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, H, W, C, generator=g)
    gamma, beta = torch.rand(4 * C, generator=g) + 0.5, torch.randn(4 * C, generator=g) * 0.1
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (0, W % 2, 0, H % 2))
    want = torch.nn.functional.layer_norm(torch.nn.functional.unfold(xp, 2, stride=2).transpose(1, 2), (4 * C,), gamma, beta)
    got = ops.patch_merge_ln(x.cuda(), gamma.cuda(), beta.cuda())
    _close(got.reshape(B, -1, 4 * C), want, 2e-5, 'patch_merge_ln')


@pytest.mark.gpu
@pytest.mark.parametrize('M', [64, 1000])
def test_linear_gelu_epilogue(M):
    """exact-erf GELU in the GEMM epilogues (skinny SIMT kernel for M <= 128, tcgen05 above), fp32 and plane outputs."""
    from openpvsg_b200 import ops
    g = torch.Generator().manual_seed(M)
    x, w, b = torch.randn(M, 128, generator=g), torch.randn(512, 128, generator=g) * 0.1, torch.randn(512, generator=g)
    want = torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b))
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), act=ops.ACT_GELU)
    _close(got, want, 1e-4, 'linear+gelu')
    sp = ops.linear(x.cuda(), w.cuda(), b.cuda(), act=ops.ACT_GELU, out_mode='split')
    if isinstance(sp, ops.Split):
        _close(sp.hi.float() + sp.lo.float(), want, 1e-4, 'linear+gelu planes')


def _swin_pair(cfg, seed, cuda):
    g = torch.Generator().manual_seed(seed)
    sd = syn.swin_state_dict(g, prefix='', **cfg)
    net = pv.build_backbone(dict(type='SwinTransformer', **cfg))
    assert net.load_state_dict(sd, strict=True)
    return net.to(cuda), sd


@pytest.mark.gpu
@pytest.mark.parametrize('hw', [(96, 160), (100, 130)])
def test_swin_backbone_vs_oracle(hw):
    """2-2-2-2 Swin at Swin-B widths: every stage output vs the oracle; 100 x 130 exercises patch-embed padding,
    window padding on every level and odd patch merging."""
    cuda = torch.device('cuda')
    net, sd = _swin_pair(SMALL, 11, cuda)
    img = syn.synthetic_frame(7, 96, 160)[None][..., :hw[0], :hw[1]].contiguous()
    img = torch.cat([img, img.flip(-1)], 0)
    with torch.no_grad():
        want = osw.swin_forward(sd, img, **SMALL)
        got = net(img.to(cuda))
    assert len(got) == 4
    for i, (a, b) in enumerate(zip(got, want)):
        _close(a, b, 1e-3, f'stage {i}')
        assert (a.cpu() - b).abs().mean().item() < 5e-5


@pytest.mark.gpu
def test_swin_b_detector_features_and_forward():
    """Full Swin-B Mask2Former-VPS: backbone features vs the oracle, then the detector end to end (pixel decoder
    fed by the Swin stage widths) against the oracle detector driven by the same features."""
    from oracle import m2f as om
    cuda = torch.device('cuda')
    sd = syn.mask2former_state_dict(seed=5, in_channels=(128, 256, 512, 1024), backbone=dict(configs.SWIN_B))
    det = pv.build_detector(configs.mask2former_swin(True))
    assert det.load_state_dict(sd, strict=True)
    det = det.to(cuda)
    H, W = 192, 256
    img = syn.synthetic_frame(21, H, W)[None]
    with torch.no_grad():
        want = osw.swin_forward(sd, img, prefix='backbone.', **configs.SWIN_B)
        feats = det.extract_feat(img.to(cuda))
    for i, (a, b) in enumerate(zip(feats, want)):
        _close(a, b, 1e-3, f'swin-b stage {i}')
    # pixel decoder + head on the ORACLE's backbone features (stage-level parity, as for R50)
    with torch.no_grad():
        ref_mf, ref_mem = om.pixel_decoder(sd, list(want))
        mf, mem = det.panoptic_head.pixel_decoder([f.to(cuda) for f in want])
    _close(mf, ref_mf, 1e-3, 'mask features')
    for a, b in zip(mem, ref_mem):
        _close(a, b, 1e-3, 'memory level')
    # the whole detector, free-running, against the oracle detector (Swin oracle backbone + the Mask2Former oracle),
    # tie-aware: panoptic ids, segment set, query features (oracle/parity.py)
    from oracle import parity
    meta = syn.frame_meta(H, W)
    head = det.panoptic_head
    head._capture_masks = []
    try:
        res = det.simple_test(None, None, ref_img=img.to(cuda)[None], ref_img_metas=[[meta]], rescale=True)[0][0]
        gpu_masks = [m[0].cpu().numpy() for m in head._capture_masks]
    finally:
        head._capture_masks = None
    pan = res['pan_results']
    assert pan.shape == (H, W) and pan.dtype == np.int32
    assert set(res['query_feats']) <= set(np.unique(pan).tolist())
    backbone = lambda sd_, x: osw.swin_forward(sd_, x, prefix='backbone.', **configs.SWIN_B)   # noqa: E731
    st = parity.check_frame(res, sd, img[0], meta, gpu_masks, what='swin-b detector', backbone=backbone)
    assert st['pan_mismatch_pixels'] <= 1e-4 * st['pixels'], st
    assert len(res['query_feats']) > 0, 'degenerate synthetic checkpoint: nothing kept'


@pytest.mark.gpu
def test_swin_b_full_size_720p_vs_golden():
    """BASELINE full size (720 x 1280 -> 184 x 320 patches, window padding on every level): all four stage outputs of
    the CUDA Swin-B against the committed oracle-derived fixture (sub-sampled values + whole-tensor sums)."""
    sd, img, gold = _swin_b_720p_inputs()
    net = pv.build_backbone(dict(type='SwinTransformer', **configs.SWIN_B))
    assert net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        outs = net.to('cuda')(img.cuda())
    for i, o in enumerate(outs):
        assert list(o.shape) == gold[f'stage{i}_shape'].tolist()
        sub = o[0, ::8, ::4, ::4].float().cpu().numpy()
        want = gold[f'stage{i}_sub']
        err = np.abs(sub - want).max()
        assert err <= 1e-3 * max(1.0, np.abs(want).max()), f'stage {i}: max abs err {err:.3e}'
        assert np.abs(sub - want).mean() < 5e-5
        assert float(o.double().abs().sum()) == pytest.approx(float(gold[f'stage{i}_abs_sum']), rel=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('shift', [0, 6])
def test_window_attention_full_size_properties(shift):
    """Stage-0 size of a 720p frame (184 x 320 tokens, 4 heads, window 12), no oracle: the output is linear in V,
    rows of a window that see identical keys are convex combinations of V (bounded by its extremes), and a constant V
    comes back unchanged (softmax rows sum to one, also for padded / masked windows)."""
    from openpvsg_b200 import ops
    g = torch.Generator().manual_seed(shift)
    B, H, W, heads = 1, 184, 320, 4
    C = 32 * heads
    qkv = torch.randn(B, H, W, 3 * C, generator=g).cuda()
    bias = torch.zeros(3 * C).cuda()          # zero bias: padded positions contribute V = 0 rows
    table = torch.randn(23 * 23, heads, generator=g).cuda()
    v2 = torch.randn(B, H, W, C, generator=g).cuda()
    a = ops.window_attention(qkv, bias, table, heads, 12, shift)
    q2 = qkv.clone()
    q2[..., 2 * C:] = v2
    b = ops.window_attention(q2, bias, table, heads, 12, shift)
    q3 = qkv.clone()
    q3[..., 2 * C:] = 0.5 * qkv[..., 2 * C:] - 2.0 * v2
    c = ops.window_attention(q3, bias, table, heads, 12, shift)
    assert (c - (0.5 * a - 2.0 * b)).abs().max().item() < 2e-4
    assert a.abs().max().item() <= qkv[..., 2 * C:].abs().max().item() + 1e-4
    # constant V with a matching bias row (so padded positions carry the same value): output == that constant
    q4 = qkv.clone()
    q4[..., 2 * C:] = 1.25
    bias4 = bias.clone()
    bias4[2 * C:] = 1.25
    d = ops.window_attention(q4, bias4, table, heads, 12, shift)
    assert (d - 1.25).abs().max().item() < 1e-4
