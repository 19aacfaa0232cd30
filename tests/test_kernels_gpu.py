"""GPU: every C-ABI kernel vs the CPU oracle / plain fp32 torch CPU ops on seeded inputs.

Tolerances are written next to each check; integer / mask / id outputs must be identical
except where a comment says which measure-zero ties are excluded.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from openpvsg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from openpvsg_b200 import ops as _ops
    return _ops


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(a, b, tol, what=''):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    d = (a - b).abs()
    err = d.max().item()
    if err > tol:   # say where: an intermittent mismatch is only diagnosable with the offending element
        i = int(d.flatten().argmax())
        idx = tuple(int(v) for v in np.unravel_index(i, tuple(a.shape)))
        raise AssertionError(f'{what}: max abs err {err:.3e} > {tol:.1e} at {idx}: got {a.flatten()[i].item():.9g}, '
                             f'expected {b.flatten()[i].item():.9g}; {int((d > tol).sum())} of {d.numel()} beyond tol')
    return err


# ------------------------------------------------------------------ dense ------------
@pytest.mark.parametrize('M,N,K', [(100, 256, 256), (19320, 288, 256), (100, 127, 256), (333, 1024, 256),
                                   (100, 256, 2048), (7, 57, 128), (130, 1, 1024), (50, 64, 147)])
def test_linear(ops, M, N, K):
    x, w, b = randn(1, M, K), randn(2, N, K) / K ** 0.5, randn(3, N)
    y = ops.linear(x.cuda(), w.cuda(), b.cuda())
    close(y, F.linear(x, w, b), 2e-4, 'linear')


def test_linear_epilogues_and_slices(ops):
    M, K = 100, 256
    x, pos, res = randn(1, M, K), randn(2, M, K), randn(3, M, 256)
    w, b = randn(4, 768, K) / 16, randn(5, 768)
    xc, pc, rc, wc, bc = x.cuda(), pos.cuda(), res.cuda(), w.cuda(), b.cuda()
    y = ops.linear(xc, wc[256:512], bc[256:512], add_input=pc, residual=rc, act=ops.ACT_RELU)
    ref = F.relu(F.linear(x + pos, w[256:512], b[256:512]) + res)
    close(y, ref, 2e-4, 'linear fused')
    # strided input / output views (fused qkv buffer)
    buf = torch.zeros(M, 768, device='cuda')
    ops.linear(xc, wc[:512], bc[:512], out=buf[:, :512])
    ops.linear(buf[:, :256], wc[512:, :], bc[512:], out=buf[:, 512:])
    ref1 = F.linear(x, w[:512], b[:512])
    close(buf[:, :512], ref1, 2e-4, 'linear strided out')
    close(buf[:, 512:], F.linear(ref1[:, :256], w[512:], b[512:]), 5e-4, 'linear strided in')


@pytest.mark.parametrize('cin,cout,k,stride,pad,hw', [(64, 64, 1, 1, 0, (46, 80)), (64, 64, 3, 1, 1, (46, 80)),
                                                      (128, 128, 3, 2, 1, (47, 81)), (3, 64, 7, 2, 3, (96, 160)),
                                                      (256, 512, 1, 2, 0, (46, 80)), (256, 256, 3, 1, 1, (23, 40))])
def test_conv2d(ops, cin, cout, k, stride, pad, hw):
    x = randn(1, 2, cin, *hw)
    w = randn(2, cout, cin, k, k) / (cin * k * k) ** 0.5
    b = randn(3, cout)
    ref = F.conv2d(x, w, b, stride, pad)
    res = randn(4, *ref.shape)
    y = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().cuda(), w.permute(0, 2, 3, 1).contiguous().cuda(),
                        b.cuda(), residual=res.permute(0, 2, 3, 1).contiguous().cuda(), stride=stride, pad=pad,
                        act=ops.ACT_RELU)
    close(y.permute(0, 3, 1, 2), F.relu(ref + res), 3e-4, 'conv')


def test_pool_layout_norms(ops):
    x = randn(1, 2, 64, 45, 77)
    xh = ops.nchw_to_nhwc(x.cuda())
    close(xh, x.permute(0, 2, 3, 1), 0, 'nchw_to_nhwc')
    close(ops.nhwc_to_nchw(xh), x, 0, 'nhwc_to_nchw')
    close(ops.maxpool3x3s2_nhwc(xh).permute(0, 3, 1, 2), F.max_pool2d(x, 3, 2, 1), 0, 'maxpool')
    for C in (256, 512):
        t, g, b = randn(2, 333, C) * 3 + 1, randn(3, C), randn(4, C)
        close(ops.layernorm(t.cuda(), g.cuda(), b.cuda()), F.layer_norm(t, (C,), g, b, 1e-5), 2e-5, 'layernorm')
    t = randn(5, 2, 256, 23, 40) * 2 + 0.5
    g, b = randn(6, 256), randn(7, 256)
    y = ops.groupnorm_nhwc(t.permute(0, 2, 3, 1).contiguous().cuda(), g.cuda(), b.cuda(), 32, act=ops.ACT_RELU)
    close(y.permute(0, 3, 1, 2), F.relu(F.group_norm(t, 32, g, b, 1e-5)), 3e-5, 'groupnorm')
    v = randn(8, 256)
    close(ops.add_rowvec(t.permute(0, 2, 3, 1).contiguous().cuda(), v.cuda()),
          t.permute(0, 2, 3, 1) + v, 0, 'add_rowvec')


@pytest.mark.parametrize('ihw,ohw', [((23, 40), (46, 80)), ((184, 320), (23, 40)), ((25, 33), (7, 9)),
                                     ((12, 20), (48, 80))])
def test_bilinear_resize(ops, ihw, ohw):
    x = randn(1, 2, 64, *ihw)
    ref = F.interpolate(x, size=ohw, mode='bilinear', align_corners=False)
    xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    close(ops.bilinear_resize_nhwc(xh, ohw).permute(0, 3, 1, 2), ref, 2e-6, 'bilinear')
    acc = randn(2, 2, *ohw, 64).cuda()
    base = acc.clone()
    ops.bilinear_resize_nhwc(xh, ohw, out=acc, accumulate=True)
    close(acc.permute(0, 3, 1, 2), ref + base.cpu().permute(0, 3, 1, 2), 3e-6, 'bilinear accumulate')


def test_sine_pe(ops, golden_dir):
    """Gate = the north_star tolerance (1e-3); the expected agreement is ~1e-7.  On two of ~25 boxes the
    first comparison came out at 1.5e-4 (not reproducible on request, every other run 6e-8): such a run
    is reported with the offending element as a warning so that it can be diagnosed, not hidden."""
    import warnings
    from oracle import m2f as om

    def check(got, ref, what):
        try:
            close(got, ref, 2e-5, what)
        except AssertionError as e:
            warnings.warn(f'sine_pe beyond the expected 2e-5: {e}')
            close(got, ref, 1e-3, what)

    pe = ops.sine_pe(23, 40, 'cuda')
    check(pe, om.sine_pe_2d(1, 23, 40)[0].flatten(1).t(), 'pe2d')
    lvl = randn(1, 256)
    pe = ops.sine_pe(6, 10, 'cuda', add_vec=lvl.cuda())
    check(pe, om.sine_pe_2d(1, 6, 10)[0].flatten(1).t() + lvl, 'pe2d+lvl')
    g = np.load(os.path.join(golden_dir, 'pe3d.npz'))
    pe3 = ops.sine_pe(5, 7, 'cuda', t=2)  # [(t h w), 256]
    ref = torch.as_tensor(g['pos'])[0].permute(0, 2, 3, 1).reshape(-1, 256)  # [t, c, h, w] -> tokens
    check(pe3, ref, 'pe3d vs reference golden')


# ------------------------------------------------------------------ MSDA -------------
SHAPES_A = [(15, 20), (30, 40), (60, 80)]     # 480x640 -> N = 6300
SHAPES_B = [(23, 40), (46, 80), (92, 160)]    # 736x1280 -> N = 19320


@pytest.mark.parametrize('shapes', [SHAPES_A, SHAPES_B, [(3, 5), (6, 10), (12, 20)]])
def test_msda_forward(ops, shapes):
    from oracle import m2f as om
    n = sum(h * w for h, w in shapes)
    B = 1 if n > 10000 else 2
    value = randn(1, B, n, 8, 32)
    g = torch.Generator().manual_seed(2)
    loc = torch.rand(B, n, 8, 3, 4, 2, generator=g) * 1.2 - 0.1   # exercises zero padding
    aw = torch.softmax(randn(3, B, n, 8, 12), -1).view(B, n, 8, 3, 4)
    ref = om.msda_core(value, shapes, loc, aw)
    out = ops.msda_forward(value.cuda(), shapes, loc.cuda(), aw.cuda())
    close(out, ref, 1e-4, 'msda')   # north_star bar is 1e-3; fp32 path expected ~1e-6


@pytest.mark.parametrize('shapes', [SHAPES_A, SHAPES_B])
def test_msda_fused(ops, shapes):
    """Fused softmax + location arithmetic + sampling vs the oracle's module path."""
    from oracle import m2f as om
    n = sum(h * w for h, w in shapes)
    value = randn(1, 1, n, 256)
    proj = torch.cat([randn(2, 1, n, 192) * 2.0, randn(3, 1, n, 96)], -1)
    refs = []
    for h, w in shapes:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                                indexing='ij')
        refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
    ref_pts = torch.cat(refs, 0)
    off = proj[..., :192].view(1, n, 8, 3, 4, 2)
    aw = proj[..., 192:].view(1, n, 8, 12).softmax(-1).view(1, n, 8, 3, 4)
    normalizer = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref_pts[None, :, None, None, None, :] + off / normalizer[None, None, None, :, None, :]
    ref = om.msda_core(value.view(1, n, 8, 32), shapes, loc, aw)
    out = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda())
    close(out, ref, 1e-4, 'msda fused')


def _fused_ref(value, shapes, proj, ref_pts):
    from oracle import m2f as om
    B, nq = proj.shape[:2]
    n = value.shape[1]
    off = proj[..., :192].view(B, nq, 8, 3, 4, 2)
    aw = proj[..., 192:].view(B, nq, 8, 12).softmax(-1).view(B, nq, 8, 3, 4)
    normalizer = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref_pts[None, :, None, None, None, :] + off / normalizer[None, None, None, :, None, :]
    return om.msda_core(value.view(B, n, 8, 32), shapes, loc, aw)


def test_msda_fused_ragged_batch_and_free_queries(ops):
    """The 8-lane-group kernel: level sizes that are not multiples of the 8 x 8 / 4-query tiles,
    batch 2, offsets far outside the maps, and queries that are NOT the pyramid tokens (Nq != N,
    not a multiple of 4: linear query order)."""
    shapes = [(3, 5), (7, 10), (13, 21)]
    n = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(11)
    value = torch.randn(2, n, 256, generator=g)
    for nq in (n, 37):
        proj = torch.cat([torch.randn(2, nq, 192, generator=g) * 6.0, torch.randn(2, nq, 96, generator=g) * 3.0], -1)
        ref_pts = torch.rand(nq, 2, generator=g)
        ref = _fused_ref(value, shapes, proj, ref_pts)
        out = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda())
        close(out, ref.reshape(2, nq, 256), 1e-4, f'msda fused nq={nq}')


@pytest.mark.parametrize('shapes,B,scale', [([(5, 6), (10, 12), (20, 24)], 2, 6.0), ([(2, 2), (4, 4), (8, 8)], 3, 3.0),
                                            ([(23, 40), (46, 80), (92, 160)], 2, 1.7), ([(15, 20), (30, 40), (60, 80)], 1, 8.0)])
def test_msda_region_tiled_kernel(ops, shapes, B, scale):
    """csrc/msda_tile.cu (TMA-staged value windows, the encoder's kernel): regions cut by the map border (sizes that are
    not multiples of the 16 x 8 region), maps smaller than one window, batches, and offsets far beyond the halo (the
    compacted global-memory pass) and beyond the map (zero padding by TMA fill) -- against the oracle, and against
    the lane-group kernel (PVSG_MSDA_IMPL=group), which must agree to re-association level."""
    import os
    n = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(int(scale * 10) + n)
    value = torch.randn(B, n, 256, generator=g)
    proj = torch.cat([torch.randn(B, n, 192, generator=g) * scale, torch.randn(B, n, 96, generator=g) * 2.0], -1)
    refs = []
    for h, w in shapes:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
        refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
    ref_pts = torch.cat(refs, 0)
    want = _fused_ref(value, shapes, proj, ref_pts)
    out = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda())
    close(out, want.reshape(B, n, 256), 1e-4, 'msda tile')
    sp = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda(), out_mode='split')
    close(sp.hi.float() + sp.lo.float(), want.reshape(B, n, 256), 1e-4, 'msda tile planes')
    os.environ['PVSG_MSDA_IMPL'] = 'group'
    try:
        old = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda())
    finally:
        del os.environ['PVSG_MSDA_IMPL']
    close(out, old, 2e-5, 'tile vs lane-group kernel')


# ------------------------------------------------------------------ attention --------
def _mha_ref(q, k, v, H, mask=None):
    B, Lq, E = q.shape
    D = E // H
    qh = q.view(B, Lq, H, D).transpose(1, 2)
    kh = k.view(B, -1, H, D).transpose(1, 2)
    vh = v.view(B, -1, H, D).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / D ** 0.5
    if mask is not None:
        s = s.masked_fill(mask[:, None].bool(), float('-inf'))
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E)


@pytest.mark.parametrize('B,H,Lq,Lk,E', [(1, 8, 100, 920, 256), (1, 8, 100, 14720, 256), (2, 8, 100, 100, 256),
                                         (12, 8, 14, 14, 256), (5, 4, 128, 128, 512), (3, 4, 37, 37, 512)])
def test_attention(ops, B, H, Lq, Lk, E):
    q, k, v = randn(1, B, Lq, E), randn(2, B, Lk, E), randn(3, B, Lk, E)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), H)
    close(out, _mha_ref(q, k, v, H), 2e-5, 'attention')


@pytest.mark.parametrize('B,Lq,Lk', [(2, 100, 920), (1, 100, 14720), (2, 37, 129), (1, 200, 3680), (3, 16, 64)])
def test_attention_tensor_core_planes(ops, B, Lq, Lk):
    """csrc/attention_mma.cu: K / V as split-bf16 planes, masks with fully blocked rows (open-row
    rule), ragged key counts (odd Lk exercises the byte mask loads), several query tiles."""
    H, E = 8, 256
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    q = torch.randn(B, Lq, E, generator=g)
    k = torch.randn(B, Lk, E, generator=g)
    v = torch.randn(B, Lk, E, generator=g)
    mask = (torch.rand(B, Lq, Lk, generator=g) < 0.6)
    mask[:, 3] = True                      # fully blocked row -> attends to everything
    mask[:, 5, : Lk // 2] = False
    row_open = (~mask).sum(-1).to(torch.int32)
    eff = mask.clone()
    eff[row_open == 0] = False
    ref = _mha_ref(q, k, v, H, eff)
    ks = ops.Split(*ops.split_bf16(k.cuda()))
    vs = ops.Split(*ops.split_bf16(v.cuda()))
    out = ops.attention(q.cuda(), ks, vs, H, mask=mask.to(torch.uint8).cuda(), row_open=row_open.cuda())
    close(out, ref, 2e-4, 'tensor-core attention (masked)')
    out = ops.attention(q.cuda(), ks, vs, H)
    close(out, _mha_ref(q, k, v, H), 2e-4, 'tensor-core attention (no mask)')
    # agrees with the fp32 SIMT kernel
    out2 = ops.attention(q.cuda(), k.cuda(), v.cuda(), H)
    close(out, out2.cpu(), 2e-4, 'tensor-core vs SIMT attention')


@pytest.mark.parametrize('B,H,Lq,Lk,E', [(5, 4, 128, 128, 512), (2, 4, 50, 77, 512), (3, 8, 200, 200, 256)])
def test_attention_tensor_core_relation_shapes(ops, B, H, Lq, Lk, E):
    """Head dim 128 (TemporalTransformer, transformer.py:20-25) and the ObjectEncoder shape, K / V as
    strided slices of a fused qkv plane buffer (the relation head's layout)."""
    g = torch.Generator().manual_seed(B * 100 + Lk)
    qkv = torch.randn(B, max(Lq, Lk), 3 * E, generator=g)
    q, k, v = qkv[:, :Lq, :E], qkv[:, :Lk, E:2 * E], qkv[:, :Lk, 2 * E:]
    ref = _mha_ref(q.contiguous(), k.contiguous(), v.contiguous(), H)
    dev = qkv.cuda()
    hi, lo = ops.split_bf16(dev)
    out = ops.attention(dev[:, :Lq, :E], ops.Split(hi[:, :Lk, E:2 * E], lo[:, :Lk, E:2 * E]),
                        ops.Split(hi[:, :Lk, 2 * E:], lo[:, :Lk, 2 * E:]), H)
    close(out, ref, 2e-4, 'tensor-core attention, strided planes')


def test_top_pairs_matches_torch_topk(ops):
    """pick_top_pairs_eval (test_utils.py:4-22): radix-select kernel vs torch.topk(sorted=True), with
    exact ties, negative scores and k larger than the number of off-diagonal entries."""
    g = torch.Generator().manual_seed(3)
    for N, k in ((200, 100), (7, 100), (33, 1)):
        m = torch.randn(N, N, generator=g)
        m[1, 2] = m[3, 4] = m[5, 6] = 9.0            # ties at the top: lower flat index first
        pairs, n = ops.top_pairs(m.cuda(), k)
        n = int(n.item())
        mm = m.clone()
        mm.fill_diagonal_(float('-inf'))
        kk = min(k, N * N)
        vals, idx = mm.flatten().topk(kk, sorted=True)
        keep = ~torch.isinf(vals)
        assert n == int(keep.sum())
        got = pairs[:n].cpu()
        got_vals = mm[got[:, 0].long(), got[:, 1].long()]
        assert torch.equal(got_vals, vals[keep])      # same scores in the same (descending) order
        ties = got[(got_vals == 9.0)]
        assert ties.tolist() == [[1, 2], [3, 4], [5, 6]][:min(3, n)]


def test_attention_masked_and_strided(ops):
    B, H, Lq, Lk, E = 1, 8, 100, 3680, 256
    q, k, v = randn(1, B, Lq, E), randn(2, B, Lk, E), randn(3, B, Lk, E)
    mask = (torch.rand(B, Lq, Lk, generator=torch.Generator().manual_seed(4)) < 0.7)
    mask[0, 5] = True   # fully blocked rows -> reset to all-open (mask2former_head.py:453-454)
    mask[0, 77] = True
    row_open = (~mask).sum(-1).to(torch.int32)
    fixed = mask.clone()
    fixed[torch.where(fixed.sum(-1) == fixed.shape[-1])] = False
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), H, mask=mask.to(torch.uint8).cuda(),
                        row_open=row_open.cuda())
    close(out, _mha_ref(q, k, v, H, fixed), 2e-5, 'masked attention')
    # seq-first storage [S, Bt, E] viewed as [Bt, S, E] (relation ObjectEncoder layout) + packed qkv
    S, Bt = 14, 12
    qkv = randn(5, S, Bt, 3 * E)
    qv = qkv.cuda().permute(1, 0, 2)
    out = ops.attention(qv[..., :E], qv[..., E:2 * E], qv[..., 2 * E:], H)
    ref = _mha_ref(*[t.permute(1, 0, 2).contiguous() for t in qkv.split(E, -1)], H)
    close(out, ref, 2e-5, 'strided attention')


# ------------------------------------------------------------------ mask logits ------
def test_mask_logits(ops):
    B, Q, C, h, w = 2, 100, 256, 24, 40
    embed, feat = randn(1, B, Q, C), randn(2, B, C, h, w)
    ref = torch.einsum('bqc,bchw->bqhw', embed, feat)
    ft = feat.permute(0, 2, 3, 1).reshape(B, h * w, C).contiguous().cuda()
    logits, mask, row_open = ops.mask_logits(embed.cuda(), ft, True, True)
    # N(0,1) operands, K = 256: |logit| reaches ~80; 1e-5 relative = split-bf16 / fp32 re-association level
    close(logits.view(B, Q, h, w), ref, 1e-5 * ref.abs().max().item(), 'mask logits')
    # sign mask: identical wherever the oracle logit is not within rounding distance of 0
    refm = ref.flatten(2) < 0
    safe = ref.flatten(2).abs() > 1e-3
    assert torch.equal(mask.cpu().bool()[safe], refm[safe])
    assert torch.equal(row_open.cpu(), (mask.cpu() == 0).sum(-1).to(torch.int32))
    # attention-mask path: downsample commutes with the contraction
    tgt = (6, 10)
    pooled = ops.bilinear_resize_nhwc(ft.view(B, h, w, C), tgt).view(B, -1, C)
    _, m2, ro2 = ops.mask_logits(embed.cuda(), pooled, False, True)
    down = F.interpolate(ref, tgt, mode='bilinear', align_corners=False).flatten(2)
    safe = down.abs() > 1e-3
    assert torch.equal(m2.cpu().bool()[safe], (down < 0)[safe])
    assert safe.float().mean() > 0.99


# ------------------------------------------------------------------ panoptic ---------
def test_panoptic_fuse_vs_reference_golden(ops, golden_dir):
    """Fused device post-processing vs the reference's own panoptic_postprocess_with_query
    outputs (tests/golden/fusion_post.npz): ids must be identical."""
    g = np.load(os.path.join(golden_dir, 'fusion_post.npz'))
    cls, mp = torch.as_tensor(g['mask_cls']), torch.as_tensor(g['mask_pred'])
    Q, H, W = mp.shape
    pan, info = ops.panoptic_fuse(cls.cuda(), mp.cuda(), (H, W), (H, W), (H, W), 115, 126)
    assert np.array_equal(pan.cpu().numpy(), g['pan'])
    info = info.cpu().numpy()
    n = info[0]
    rows = info[1:1 + 4 * n].reshape(n, 4)
    segs = sorted(set(int(s) for s in rows[:, 2] if s >= 0))
    assert segs == g['qf_keys'].tolist()
    # crop + rescale path (img_shape 36x50 of a 40x56 map, ori_shape = img_shape)
    pan2, _ = ops.panoptic_fuse(cls.cuda(), mp.cuda(), (H, W), (36, 50), (36, 50), 115, 126)
    assert np.array_equal(pan2.cpu().numpy(), g['crop_pan'])


def test_panoptic_fuse_upsampled(ops):
    """x4 upsample fused in: compare with the oracle on low-res logits (ids identical up to
    pixels whose top-2 probabilities / 0.5 threshold are within 1e-6 in the oracle)."""
    from oracle import m2f as om
    Q, h, w = 40, 24, 40
    H, W = 4 * h, 4 * w
    g = torch.Generator().manual_seed(7)
    cls = torch.randn(Q, 127, generator=g)
    for q in range(0, Q, 2):
        cls[q, (q * 5) % 126] += 14.0
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    mp = torch.randn(Q, h, w, generator=g) * 0.3 - 2.5
    for q in range(Q):
        cy, cx = (q * 5) % h, (q * 9) % w
        mp[q] += 7.0 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 8.0)
    up = F.interpolate(mp[None], size=(H, W), mode='bilinear', align_corners=False)[0]
    qf = torch.zeros(Q, 256)
    meta = dict(img_shape=(H - 8, W - 4, 3), ori_shape=(H - 8, W - 4, 3))
    ref = om.fusion_simple_test_with_query(cls[None], up[None], qf[None], [meta], rescale=True,
                                           instance_on=False)[0]['pan_results'].numpy()
    pan, _ = ops.panoptic_fuse(cls.cuda(), mp.cuda(), (H, W), (H - 8, W - 4), (H - 8, W - 4), 115, 126)
    pan = pan.cpu().numpy()
    mism = (pan != ref).mean()
    assert len(np.unique(ref)) > 4
    assert mism == 0.0, f'{mism:.2e} of pixels differ'
    # real rescale (ori_shape != img_shape)
    meta = dict(img_shape=(H - 8, W - 4, 3), ori_shape=(61, 97, 3))
    ref = om.fusion_simple_test_with_query(cls[None], up[None], qf[None], [meta], rescale=True,
                                           instance_on=False)[0]['pan_results'].numpy()
    pan, _ = ops.panoptic_fuse(cls.cuda(), mp.cuda(), (H, W), (H - 8, W - 4), (61, 97), 115, 126)
    assert (pan.cpu().numpy() != ref).mean() < 2e-3   # composite bilinear: only near-tie pixels may differ


@pytest.mark.parametrize('Q,NC,k', [(100, 126, 100), (7, 5, 35), (100, 126, 1), (33, 126, 64)])
def test_instance_select(ops, Q, NC, k):
    """softmax + flattened top-k (mask2former_fusion_head.py:214-222): same candidate SET and scores
    as torch.topk; the order of sorted=False is unspecified."""
    g = torch.Generator().manual_seed(Q + k)
    cls = torch.randn(Q, NC + 1, generator=g) * 3.0
    cls[1] = cls[0]                      # exact ties inside the scores
    scores = torch.softmax(cls, -1)[:, :-1].flatten()
    ref_s, ref_i = scores.topk(k, sorted=True)
    s, lab, qi = ops.instance_select(cls.cuda(), k)
    flat = (qi.cpu().long() * NC + lab.cpu().long())
    assert flat.unique().numel() == k
    close(s.cpu().sort(descending=True).values, ref_s, 1e-6, 'top-k scores')
    close(s.cpu(), scores[flat], 1e-6, 'score / index consistency')
    strictly_in = scores > ref_s[-1] + 1e-7          # everything clearly above the k-th value is selected
    assert set(strictly_in.nonzero().flatten().tolist()) <= set(flat.tolist())


def test_instance_finalize(ops):
    """Static-shape top-10 selection (models/mask2former_vps/mask2former.py:192-201) vs the torch
    formulation: things only, det_score = score * mask score, 1-based ids in candidate order."""
    g = torch.Generator().manual_seed(5)
    n, num_things, topk = 100, 115, 10
    scores = torch.rand(n, generator=g)
    labels = torch.randint(0, 126, (n,), generator=g).int()
    labels[::3] = 120                                    # plenty of stuff candidates
    query = torch.randint(0, 100, (n,), generator=g).int()
    stats = torch.stack([torch.rand(n, generator=g) * 500, torch.randint(0, 900, (n,), generator=g).float()], 1)
    stats[5] = 0.0                                        # empty mask -> score 0
    boxes = torch.randint(0, 700, (n, 4), generator=g).int()
    b6, lab, sel, cnt = ops.instance_finalize(scores.cuda(), labels.cuda(), query.cuda(), stats.cuda(), boxes.cuda(),
                                              num_things, topk)
    is_thing = labels < num_things
    det = scores * stats[:, 0] / (stats[:, 1] + 1e-6)
    ids = torch.cumsum(is_thing.float(), 0)
    order = torch.argsort(torch.where(is_thing, det, torch.full_like(det, -1.0)), descending=True, stable=True)[:topk]
    assert int(cnt.item()) == int(is_thing.sum())
    assert torch.equal(lab.cpu(), labels[order]) and torch.equal(sel.cpu(), query[order])
    ref = torch.cat([ids[order, None], boxes[order].float(), det[order, None]], 1)
    close(b6, ref, 1e-4 * float(ref.abs().max()), 'instance boxes')


def test_instance_masks(ops):
    from oracle import m2f as om
    Q, h, w = 12, 24, 40
    H, W = 4 * h, 4 * w
    mp = randn(3, Q, h, w) - 0.5
    mp[4] = -5.0  # empty mask -> zero box
    up = F.interpolate(mp[None], size=(H, W), mode='bilinear', align_corners=False)[0]
    idx = torch.tensor([3, 4, 0, 11, 3], dtype=torch.int32)
    stats, boxes, masks = ops.instance_masks(mp.cuda(), idx.cuda(), (H, W), (H, W), (H, W))
    sel = up[idx.long()]
    binm = sel > 0
    safe = sel.abs() > 1e-5
    assert torch.equal(masks.cpu().bool()[safe], binm[safe])
    ref_boxes = om.mask2bbox(masks.cpu().bool())
    assert torch.equal(boxes.cpu().float(), ref_boxes)
    ref_sum = (sel.sigmoid() * masks.cpu().float()).flatten(1).sum(1)
    close(stats[:, 0], ref_sum, 2e-2, 'mask score sum')
    assert torch.equal(stats[:, 1].cpu(), masks.cpu().float().flatten(1).sum(1))


# ------------------------------------------------------------------ relation ---------
def test_relation_kernels(ops, golden_dir):
    from oracle import relation as orel
    g = np.load(os.path.join(golden_dir, 'rel_small.npz'))
    sub, obj = torch.as_tensor(g['sub']), torch.as_tensor(g['obj'])
    close(ops.max_over_time(sub.cuda()), sub.max(1).values, 0, 'max_over_time')
    sds = syn.relation_state_dicts(seed=int(g['weights_seed']))
    sd = sds['pair_proposal_model']
    W1, b1, W2, b2 = (sd[k].cuda() for k in ('pair_ffn.0.weight', 'pair_ffn.0.bias', 'pair_ffn.2.weight',
                                             'pair_ffn.2.bias'))
    st, ot = ops.max_over_time(sub.cuda()), ops.max_over_time(obj.cuda())
    U = ops.linear(st, W1[:, :256], b1)
    V = ops.linear(ot, W1[:, 256:])
    pm = ops.pair_proposal(U, V, W2.view(-1), b2)
    close(pm, torch.as_tensor(g['pred_matrix']), 2e-4, 'pair proposal vs reference golden')
    # top pairs on the golden matrix itself: identical list
    pairs, n = ops.top_pairs(torch.as_tensor(g['pred_matrix']).cuda(), int(g['P']))
    n = int(n.item())
    assert pairs[:n].cpu().tolist() == g['pairs'].tolist()
    # k > N*N - N: diagonal (-inf) entries are dropped
    pairs2, n2 = ops.top_pairs(torch.as_tensor(g['pred_matrix']).cuda(), 14 * 14)
    assert int(n2.item()) == 14 * 13
    assert pairs2[:int(n2.item())].cpu().tolist() == orel.pick_top_pairs_eval(torch.as_tensor(g['pred_matrix']), 196)
    cat = ops.gather_pairs(sub.cuda(), obj.cuda(), torch.as_tensor(g['pairs'], dtype=torch.int32).cuda())
    close(cat, torch.as_tensor(g['cat']), 0, 'gather pairs vs reference golden')
    pe = sds['relation_model']['positional_encoding.pe'].view(5000, 512).cuda()
    cat_pe = ops.gather_pairs(sub.cuda(), obj.cuda(), torch.as_tensor(g['pairs'], dtype=torch.int32).cuda(), pe)
    close(cat_pe, torch.as_tensor(g['cat']) + pe[:int(g['T'])].cpu()[None], 1e-6, 'gather pairs + pe')


def test_error_codes(ops):
    from openpvsg_b200 import lib
    with pytest.raises(lib.PvsgError):
        ops.linear(torch.zeros(4, 8), torch.zeros(4, 8))          # CPU tensors are rejected
    with pytest.raises(lib.PvsgError):
        ops.layernorm(torch.zeros(4, 100, device='cuda'), torch.zeros(100, device='cuda'),
                      torch.zeros(100, device='cuda'))            # unsupported width -> error code
    with pytest.raises(lib.PvsgError):
        ops.msda_forward(torch.zeros(1, 10, 8, 32, device='cuda'), [(2, 3)],
                         torch.zeros(1, 10, 8, 1, 4, 2, device='cuda'), torch.zeros(1, 10, 8, 1, 4, device='cuda'))


@pytest.mark.gpu
def test_handle_create_workspace_destroy():
    """include/pvsg.h handle API: create on a device, grow-only workspace, info, destroy; bad device is refused."""
    from openpvsg_b200 import lib as l
    h = l.Handle(0)
    info = h.info()
    assert info['device'] == 0 and info['sm_count'] >= 100 and info['smem_optin_bytes'] >= 200 * 1024 and info['workspace_bytes'] == 0
    p1 = h.workspace(1000)
    assert p1 and h.info()['workspace_bytes'] == 1024
    assert h.workspace(512) == p1 and h.workspace(0) == p1            # never shrinks, same block
    p2 = h.workspace(1 << 20)
    assert p2 and h.info()['workspace_bytes'] == 1 << 20
    t = torch.zeros(4, device='cuda')                                  # the device still works after a grow
    assert float((t + 1).sum()) == 4.0
    h.close()
    h.close()                                                          # idempotent
    with pytest.raises(l.PvsgError):
        l.Handle(torch.cuda.device_count() + 3)
    assert l.handle(0) is l.handle(0)
