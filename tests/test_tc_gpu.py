"""GPU: the tcgen05 split-bf16 GEMM / conv engine vs float64 references.

The engine must be fp32-grade: max error relative to the row/column norms stays at the level of
fp32 re-association (a single-pass bf16 GEMM would sit near 4e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from openpvsg_b200 import ops as _ops
    _ops.set_engine('tc')
    yield _ops
    _ops.set_engine('tc')


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_split_planes(ops):
    x = randn(1, 1000, 64) * 3
    hi, lo = ops.split_bf16(x.cuda())
    rec = hi.float().cpu() + lo.float().cpu()
    assert ((rec - x).abs() <= x.abs() * 2.0 ** -16 + 1e-30).all()
    hi2, lo2 = ops.split_bf16(x.cuda(), x.cuda())
    assert torch.equal(hi2.cpu(), (2 * x).to(torch.bfloat16))


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (100, 256, 256), (19320, 288, 256), (333, 1024, 256),
                                   (100, 127, 256), (58880, 100, 256), (100, 256, 2048), (257, 57, 128),
                                   (920, 2048, 512)])
def test_linear_tc(ops, M, N, K):
    x, w, b = randn(1, M, K), randn(2, N, K) / K ** 0.5, randn(3, N)
    res = randn(4, M, N)
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), residual=res.cuda(), act=ops.ACT_RELU)
    ref = F.relu(F.linear(x.double(), w.double(), b.double()) + res.double())
    err = (y.cpu().double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err
    # agrees with the SIMT fp32 engine to re-association level
    ops.set_engine('simt')
    y2 = ops.linear(x.cuda(), w.cuda(), b.cuda(), residual=res.cuda(), act=ops.ACT_RELU)
    ops.set_engine('tc')
    assert (y - y2).abs().max().item() < 1e-4


def test_linear_tc_add_input_and_slices(ops):
    M, K = 100, 256
    x, pos = randn(1, M, K), randn(2, M, K)
    w, b = randn(4, 768, K) / 16, randn(5, 768)
    y = ops.linear(x.cuda(), w.cuda()[256:512], b.cuda()[256:512], add_input=pos.cuda())
    ref = F.linear((x + pos).double(), w[256:512].double(), b[256:512].double())
    assert (y.cpu().double() - ref).abs().max().item() < 3e-5 * ref.abs().max().item()
    buf = torch.zeros(M, 768, device='cuda')
    ops.linear(x.cuda(), w.cuda()[:512], b.cuda()[:512], out=buf[:, :512])
    ref1 = F.linear(x.double(), w[:512].double(), b[:512].double())
    assert (buf[:, :512].cpu().double() - ref1).abs().max().item() < 3e-5 * ref1.abs().max().item()


@pytest.mark.parametrize('cin,cout,k,pad,hw,B,stride', [
    (64, 64, 3, 1, (46, 80), 2, 1), (256, 256, 3, 1, (23, 40), 1, 1), (128, 128, 3, 1, (92, 160), 1, 1),
    (64, 256, 1, 0, (45, 77), 2, 1), (512, 512, 3, 1, (23, 40), 1, 1), (64, 64, 3, 1, (184, 320), 1, 1),
    # strided convs of the ResNet stage transitions (TMA elementStrides = 2), odd sizes included
    (128, 128, 3, 1, (92, 160), 1, 2), (256, 512, 1, 0, (92, 160), 1, 2), (64, 64, 3, 1, (47, 81), 2, 2),
    (256, 256, 3, 1, (46, 80), 1, 2), (512, 1024, 1, 0, (45, 79), 1, 2),
    # RGB stem: patches gathered straight into operand planes (pvsg_im2col_split) + GEMM
    (3, 64, 7, 3, (96, 160), 2, 2), (3, 64, 7, 3, (75, 131), 1, 2)])
def test_conv_tc(ops, cin, cout, k, pad, hw, B, stride):
    x = randn(1, B, cin, *hw)
    w = randn(2, cout, cin, k, k) / (cin * k * k) ** 0.5
    b = randn(3, cout)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride, pad)
    res = randn(4, *ref.shape)
    y = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().cuda(), w.permute(0, 2, 3, 1).contiguous().cuda(),
                        b.cuda(), residual=res.permute(0, 2, 3, 1).contiguous().cuda(), stride=stride, pad=pad,
                        act=ops.ACT_RELU)
    ref = F.relu(ref + res.double())
    err = (y.permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err


def _planes_value(sp):
    return sp.hi.float().cpu().double() + sp.lo.float().cpu().double()


@pytest.mark.parametrize('M,N,K', [(300, 256, 64), (1000, 288, 256), (19320, 256, 128), (333, 1024, 256),
                                   (257, 64, 576)])
@pytest.mark.parametrize('res_kind', ['none', 'f32', 'planes'])
@pytest.mark.parametrize('out_mode', ['f32', 'split', 'both'])
def test_linear_tc_output_and_residual_modes(ops, M, N, K, res_kind, out_mode):
    """Every epilogue staging scheme of gemm_tc_kernel: one / both kinds of TMA output x
    (no residual | fp32 residual box | residual carried as split planes), ragged M and N."""
    x, w, b = randn(1, M, K), randn(2, N, K) / K ** 0.5, randn(3, N)
    res = randn(4, M, N)
    planes = ops.Split(*ops.split_bf16(x.cuda()))
    if res_kind == 'planes':
        r = ops.Split(*ops.split_bf16(res.cuda()))
        res_val = _planes_value(r)
    elif res_kind == 'f32':
        r, res_val = res.cuda(), res.double()
    else:
        r, res_val = None, torch.zeros(M, N, dtype=torch.float64)
    y = ops.linear(planes, w.cuda(), b.cuda(), residual=r, act=ops.ACT_RELU, out_mode=out_mode)
    ref = F.relu(F.linear(_planes_value(planes), w.double(), b.double()) + res_val)
    tol = 3e-5 * max(1.0, ref.abs().max().item())
    f32, sp = (y if out_mode == 'both' else (y, None) if out_mode == 'f32' else (None, y))
    if f32 is not None:
        assert (f32.cpu().double() - ref).abs().max().item() < tol
    if sp is not None:
        assert isinstance(sp, ops.Split)
        assert (_planes_value(sp) - ref).abs().max().item() < tol


@pytest.mark.parametrize('cin,cout,k,pad,hw,B,stride', [(64, 256, 1, 0, (45, 77), 2, 1), (128, 128, 3, 1, (47, 81), 2, 1),
                                                        (256, 512, 1, 0, (46, 80), 1, 2)])
@pytest.mark.parametrize('out_mode', ['split', 'both'])
def test_conv_tc_plane_residual(ops, cin, cout, k, pad, hw, B, stride, out_mode):
    """The bottleneck's identity branch added from its operand planes (ResNet.forward)."""
    x = randn(1, B, cin, *hw)
    w = randn(2, cout, cin, k, k) / (cin * k * k) ** 0.5
    b = randn(3, cout)
    xs = ops.Split(*ops.split_bf16(x.permute(0, 2, 3, 1).contiguous().cuda()))
    xv = _planes_value(xs).permute(0, 3, 1, 2)
    ref = F.conv2d(xv, w.double(), b.double(), stride, pad)
    res = randn(4, *ref.shape)
    rs = ops.Split(*ops.split_bf16(res.permute(0, 2, 3, 1).contiguous().cuda()))
    y = ops.conv2d_nhwc(xs, w.permute(0, 2, 3, 1).contiguous().cuda(), b.cuda(), residual=rs, stride=stride, pad=pad,
                        act=ops.ACT_RELU, out_mode=out_mode)
    ref = F.relu(ref + _planes_value(rs).permute(0, 3, 1, 2))
    tol = 3e-5 * max(1.0, ref.abs().max().item())
    f32, sp = y if out_mode == 'both' else (None, y)
    assert (_planes_value(sp).permute(0, 3, 1, 2) - ref).abs().max().item() < tol
    if f32 is not None:
        assert (f32.permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item() < tol


@pytest.mark.parametrize('B,C,H,W', [(2, 3, 96, 160), (1, 3, 75, 131), (1, 3, 736, 1280)])
def test_stem7x7s2_rowpair_conv(ops, B, C, H, W):
    """ResNet conv1 through the row-pair packing (pvsg_stem7x7s2_pack) + 4x1 tcgen05 conv."""
    x = randn(1, B, C, H, W)
    w = randn(2, 64, C, 7, 7) / (C * 49) ** 0.5
    b = randn(3, 64)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), 2, 3))
    w2 = ops.stem_weight(w.permute(0, 2, 3, 1).contiguous().cuda())
    y = ops.stem7x7s2(x.cuda(), w2, b.cuda())
    assert tuple(y.shape) == (B, ref.shape[2], ref.shape[3], 64)
    err = (y.permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err


def test_plane_emitting_norms_and_msda(ops):
    """LayerNorm (y and y + pos planes), GroupNorm planes and MSDA planes equal the split of the fp32
    results of the same kernels."""
    x, pos = randn(1, 300, 256), randn(2, 300, 256)
    g, b = randn(3, 256), randn(4, 256)
    y, ys, qs = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), out_split=True, add=pos.cuda())
    ref = F.layer_norm(x, (256,), g, b)
    assert (y.cpu() - ref).abs().max().item() < 1e-5
    rel = 2.0 ** -16     # planes carry hi + lo: relative error <= 2^-17 per element
    assert (_planes_value(ys) - y.cpu().double()).abs().max().item() < rel * y.abs().max().item()
    assert (_planes_value(qs) - (y.cpu() + pos).double()).abs().max().item() < rel * (y.cpu() + pos).abs().max().item()
    xg = randn(5, 2, 20, 24, 256)
    yg = ops.groupnorm_nhwc(xg.cuda(), g.cuda(), b.cuda(), 32, act=ops.ACT_RELU)
    yf, sp = ops.groupnorm_nhwc(xg.cuda(), g.cuda(), b.cuda(), 32, act=ops.ACT_RELU, out_mode='both')
    sp2 = ops.groupnorm_nhwc(xg.cuda(), g.cuda(), b.cuda(), 32, act=ops.ACT_RELU, out_mode='split')
    assert torch.equal(yf, yg)
    assert (_planes_value(sp) - yg.cpu().double()).abs().max().item() < rel * yg.abs().max().item()
    assert torch.equal(sp2.hi, sp.hi) and torch.equal(sp2.lo, sp.lo)
    shapes = [(5, 7), (10, 14), (20, 28)]
    n = sum(h * w for h, w in shapes)
    value = randn(6, 2, n, 256)
    proj = torch.cat([randn(7, 2, n, 192) * 3, randn(8, 2, n, 96)], -1)
    ref_pts = torch.rand(n, 2, generator=torch.Generator().manual_seed(9))
    o = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda())
    osp = ops.msda_fused_forward(value.cuda(), shapes, proj.cuda(), ref_pts.cuda(), out_mode='split')
    assert isinstance(osp, ops.Split)
    assert (_planes_value(osp) - o.cpu().double()).abs().max().item() < rel * o.abs().max().item()
    hi, lo = ops.split_bf16(o)
    assert torch.equal(osp.hi, hi) and torch.equal(osp.lo, lo)     # same rounding as the split kernel
