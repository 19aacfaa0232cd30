"""torch-tensor front end of the C ABI (include/pvsg.h).

PyTorch is used here for device memory and streams only: every function checks its
tensors (CUDA, fp32, layout), passes raw pointers + sizes + the current stream to
libpvsg_sm100.so and returns the output tensor.  No torch arithmetic happens here and
there is no fallback path -- a missing library or a failed call raises ``PvsgError``.
All calls are stream-ordered and allocation-free on the native side, so a whole forward
can be captured with ``torch.cuda.graph``.
"""
import ctypes
import os

import torch

from . import lib as _l

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_raw_device = getattr(torch._C, '_cuda_getDevice', None)


def _stream():
    """cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() costs ~15 us of Python
    per call (device-index resolution, Stream object): 4400 library calls of a training step paid 65 ms for it; the raw
    accessor is ~0.3 us."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _f32(t, name='tensor'):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32):
        raise _l.PvsgError(f'{name}: expected a CUDA float32 tensor, got {t.device} {t.dtype}')
    if t.device.index != torch.cuda.current_device():
        # kernels are launched on the CURRENT device's stream with raw pointers: a tensor of another GPU would be
        # dereferenced on the wrong device (one process per GPU is the supported layout; use torch.cuda.device(...))
        raise _l.PvsgError(f'{name}: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}')
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rows(t, name):
    """View `t` as [rows, C] with unit inner stride; returns (tensor, rows, C, row_stride)."""
    _f32(t, name)
    if t.dim() != 2:
        if not t.is_contiguous():
            raise _l.PvsgError(f'{name}: non-2D tensors must be contiguous')
        t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1:
        raise _l.PvsgError(f'{name}: inner stride must be 1')
    return t, t.shape[0], t.shape[1], t.stride(0)


# ----------------------------------------------------------------------------------------
# engine selection: 'tc' = tcgen05 split-bf16 GEMM/conv (csrc/gemm_tc.cu) where the shape is
# covered, SIMT fp32 (csrc/gemm.cu) otherwise; 'simt' = SIMT everywhere.  Both are this
# library's CUDA kernels -- this is not a fallback to another backend.
# ----------------------------------------------------------------------------------------
ENGINE = [os.environ.get('PVSG_ENGINE', 'tc')]
SKINNY_M = 128   # rows at or below which linear() uses the skinny SIMT kernel instead of tcgen05
_wplanes = {}


def set_engine(name):
    if name not in ('tc', 'simt'):
        raise ValueError(name)
    ENGINE[0] = name


def split_bf16(x, add=None):
    """fp32 tensor -> (hi, lo) bf16 planes of the same shape (of x + add when given)."""
    lib = _l.load()
    _f32(x, 'x')
    if not x.is_contiguous() or (add is not None and not (add.is_contiguous() and add.shape == x.shape)):
        raise _l.PvsgError('split_bf16: contiguous tensors of equal shape required')
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _l.check(lib.pvsg_split_bf16(_ptr(x), _ptr(_f32(add)), _ptr(hi), _ptr(lo), x.numel(), _stream()),
             'pvsg_split_bf16')
    return hi, lo


class Split:
    """An fp32 activation carried as its two bf16 planes (the tcgen05 engine's operand format).
    Producers can emit it directly (``out_mode='split'`` / ``'both'``) so that GEMM -> GEMM chains
    never materialise the fp32 tensor nor run a separate split pass."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def dim(self):
        return self.hi.dim()

    def view(self, *shape):
        return Split(self.hi.view(*shape), self.lo.view(*shape))

    def __iter__(self):  # (hi, lo) unpacking
        return iter((self.hi, self.lo))


_split_cache = {}


def clear_split_cache():
    """Forget the planes remembered during the previous forward (called at the start of a frame)."""
    _split_cache.clear()


def remember_split(t, planes):
    """Associate operand planes with an fp32 tensor that crosses a module interface as a plain
    tensor (e.g. backbone stage outputs).  The tensor is held by the cache, so its address cannot
    be recycled for another tensor while the entry is alive."""
    if planes is not None:
        _split_cache[(t.data_ptr(), tuple(t.shape))] = (t, planes)


def recall_split(t):
    hit = _split_cache.get((t.data_ptr(), tuple(t.shape)))
    return hit[1] if hit is not None else None


def maybe_split(x, add=None):
    """Split planes of x (+ add) when the tcgen05 engine is active and the shape qualifies, else None."""
    if isinstance(x, Split):
        return x
    if ENGINE[0] == 'tc' and x.shape[-1] % 64 == 0 and x.is_contiguous() and x.numel() // x.shape[-1] > SKINNY_M:
        return Split(*split_bf16(x, add))
    return None


def _weight_planes(w):
    """Cached split planes of a (possibly sliced) weight matrix.  The cache lives ON the base
    tensor object (the nn.Parameter / prepared weight), so it dies with it and can never be hit
    by a different tensor that happens to reuse the same address; in-place updates
    (load_state_dict) bump the version counter and invalidate it."""
    base = w._base if w._base is not None else w
    cache = base.__dict__.get('_pvsg_planes')
    if cache is None:
        cache = {}
        base._pvsg_planes = cache
    # data_ptr / device: nn.Module._apply (.to, .float) swaps param.data without bumping the version
    key = (w.storage_offset(), tuple(w.shape), tuple(w.stride()), base._version, base.data_ptr(), str(base.device))
    hit = cache.get(key)
    if hit is None:
        for k in [k for k in cache if k[3:] != key[3:]]:
            del cache[k]
        hit = split_bf16(w.contiguous())
        cache[key] = hit
    return hit


_ident = {}


def _identity(dev):
    """bf16 [256,256] identity per device: the B operand of the residual-by-tensor-core k-blocks."""
    key = str(dev)
    if key not in _ident:
        _ident[key] = torch.eye(256, device=dev, dtype=torch.bfloat16).contiguous()
    return _ident[key]


def _tc_ok(K, lda):
    return ENGINE[0] == 'tc' and K % 64 == 0 and lda % 8 == 0


def linear(x, weight, bias=None, add_input=None, residual=None, act=ACT_NONE, out=None, out_mode='f32'):
    """act((x + add_input) @ weight.T + bias + residual); x [..., K], weight [N, K] (row slices ok).

    x may be a ``Split`` (operand planes produced upstream).  out_mode: 'f32' -> fp32 tensor;
    'split' -> ``Split`` planes only; 'both' -> (fp32, Split).  On the SIMT engine planes do not
    exist: 'split' returns the fp32 tensor and 'both' returns (fp32, None); consumers accept either.
    """
    lib = _l.load()
    w2, N, Kw, ldw = _rows(weight, 'weight')
    if isinstance(x, Split):
        lead, K = x.shape[:-1], x.shape[-1]
        M = x.hi.numel() // K
        if add_input is not None or Kw != K:
            raise _l.PvsgError('linear: Split input cannot take add_input / K mismatch')
        return _linear_tc(lib, x.view(M, K), w2, bias, residual, act, out, lead, M, N, K, out_mode)
    lead = x.shape[:-1]
    x2, M, K, lda = _rows(x, 'x')
    if Kw != K:
        raise _l.PvsgError(f'linear: K mismatch {K} vs {Kw}')
    # M <= 128 (the decoder's 100-query chains) goes to the latency-optimised skinny SIMT kernel
    # inside pvsg_linear: exact fp32, (x + pos) / bias / residual / ReLU fused, no split pass
    skinny = M <= SKINNY_M and K <= 512 and N <= 512   # measured: beyond that the tcgen05 kernel wins again
    if not skinny and _tc_ok(K, K) and x2.is_contiguous() and (add_input is None or add_input.is_contiguous()):
        planes = Split(*split_bf16(x2, None if add_input is None else add_input.reshape(M, K)))
        return _linear_tc(lib, planes, w2, bias, residual, act, out, lead, M, N, K, out_mode)
    a2 = None
    if add_input is not None:
        a2, M2, K2, lda2 = _rows(add_input, 'add_input')
        if (M2, K2, lda2) != (M, K, lda):
            raise _l.PvsgError('linear: add_input must match x layout')
    created = out is None
    if created:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    o2, Mo, No, ldc = _rows(out, 'out')
    if (Mo, No) != (M, N):
        raise _l.PvsgError('linear: bad out shape')
    r2, ldr = None, 0
    if residual is not None:
        r2, Mr, Nr, ldr = _rows(residual, 'residual')
        if (Mr, Nr) != (M, N):
            raise _l.PvsgError('linear: bad residual shape')
    if bias is not None and (_f32(bias, 'bias').numel() != N or not bias.is_contiguous()):
        raise _l.PvsgError('linear: bad bias')
    _l.check(lib.pvsg_linear(_ptr(x2), _ptr(a2), _ptr(w2), _ptr(bias), _ptr(r2), _ptr(o2), M, N, K, lda, ldw,
                             ldc, ldr, act, 1, 0, 0, 0, _stream()), 'pvsg_linear')
    out = out.reshape(*lead, N) if created else out
    return (out, None) if out_mode == 'both' else out


def _linear_tc(lib, planes, w2, bias, residual, act, out, lead, M, N, K, out_mode='f32'):
    """planes: Split [M,K].  Returns per out_mode (see linear)."""
    w_hi, w_lo = _weight_planes(w2)
    dev = planes.device
    want_f32 = out_mode in ('f32', 'both') or out is not None
    want_split = out_mode in ('split', 'both') and N % 8 == 0
    if not want_split and not want_f32:
        want_f32 = True  # odd N: planes cannot be emitted; caller gets fp32
    created = out is None
    o2, ldc = None, N
    if want_f32:
        if created:
            out = torch.empty(M, N, device=dev, dtype=torch.float32)
        o2, Mo, No, ldc = _rows(out, 'out')
        if (Mo, No) != (M, N):
            raise _l.PvsgError('linear: bad out shape')
    c_hi = c_lo = None
    if want_split:
        if ldc != N:
            raise _l.PvsgError('linear: split output needs a dense fp32 output')
        c_hi = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        c_lo = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    r2, ldr, r_hi, r_lo = None, 0, None, None
    if isinstance(residual, Split):      # residual carried as planes: r = hi + lo
        if residual.hi.numel() != M * N or not (residual.hi.is_contiguous() and residual.lo.is_contiguous()):
            raise _l.PvsgError('linear: bad residual planes')
        r_hi, r_lo, ldr = residual.hi, residual.lo, N
    elif residual is not None:
        r2, Mr, Nr, ldr = _rows(residual, 'residual')
        if (Mr, Nr) != (M, N):
            raise _l.PvsgError('linear: bad residual shape')
    if bias is not None and (_f32(bias, 'bias').numel() != N or not bias.is_contiguous()):
        raise _l.PvsgError('linear: bad bias')
    _l.check(lib.pvsg_linear_tc(_ptr(planes.hi), _ptr(planes.lo), K, _ptr(w_hi), _ptr(w_lo), K, _ptr(bias), _ptr(r2),
                                ldr, _ptr(o2), _ptr(c_hi), _ptr(c_lo), None, None, ldc, M, N, K, act, _ptr(r_hi),
                                _ptr(r_lo), _ptr(_identity(dev)) if r_hi is not None else None, _stream()),
             'pvsg_linear_tc')
    f32 = (out.reshape(*lead, N) if created else out) if want_f32 else None
    sp = Split(c_hi.view(*lead, N), c_lo.view(*lead, N)) if want_split else None
    if out_mode == 'both':
        return f32, sp
    if out_mode == 'split' and sp is not None:
        return sp
    return f32


def conv2d_nhwc(x, weight, bias=None, residual=None, stride=1, pad=0, act=ACT_NONE, out=None, out_mode='f32'):
    """x [B,H,W,Cin] token-major (tensor or Split), weight [Cout,R,S,Cin] -> [B,OH,OW,Cout].
    out_mode as in ``linear``."""
    lib = _l.load()
    if isinstance(x, Split) or (ENGINE[0] == 'tc' and stride in (1, 2) and x.dim() == 4 and x.shape[-1] % 64 == 0
                                and x.is_contiguous() and weight.is_contiguous() and out is None):
        B, H, W, Cin = x.shape
        Cout, R, S, _ = weight.shape
        if R == 1 and S == 1 and pad == 0 and stride == 1:
            r = residual.view(-1, Cout) if residual is not None else None
            if r is not None and isinstance(r, Split) and not isinstance(x, Split):
                raise _l.PvsgError('conv2d_nhwc: plane residual needs plane input')
            res = linear(x.view(B * H * W, Cin) if isinstance(x, Split) else x.view(B * H * W, Cin),
                         weight.view(Cout, Cin), bias, residual=r, act=act, out_mode=out_mode)
            shp = (B, H, W, Cout)
            if out_mode == 'both':
                return res[0].view(*shp), (res[1].view(*shp) if res[1] is not None else None)
            return res.view(*shp)
        xs = x if isinstance(x, Split) else Split(*split_bf16(x))
        w_hi, w_lo = _weight_planes(weight)
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        dev = xs.device
        want_split = out_mode in ('split', 'both') and Cout % 8 == 0
        want_f32 = out_mode in ('f32', 'both') or not want_split
        y = torch.empty(B, OH, OW, Cout, device=dev, dtype=torch.float32) if want_f32 else None
        y_hi = torch.empty(B, OH, OW, Cout, device=dev, dtype=torch.bfloat16) if want_split else None
        y_lo = torch.empty(B, OH, OW, Cout, device=dev, dtype=torch.bfloat16) if want_split else None
        r_hi = r_lo = None
        if isinstance(residual, Split):
            if tuple(residual.shape) != (B, OH, OW, Cout) or not residual.hi.is_contiguous():
                raise _l.PvsgError('conv2d_nhwc: bad residual planes')
            r_hi, r_lo, residual = residual.hi, residual.lo, None
        elif residual is not None and (tuple(residual.shape) != (B, OH, OW, Cout) or not residual.is_contiguous()):
            raise _l.PvsgError('conv2d_nhwc: bad residual')
        _l.check(lib.pvsg_conv2d_tc(_ptr(xs.hi), _ptr(xs.lo), _ptr(w_hi), _ptr(w_lo), _ptr(_f32(bias)),
                                    _ptr(_f32(residual)), _ptr(y), _ptr(y_hi), _ptr(y_lo), B, H, W, Cin, Cout, R, S,
                                    stride, pad, act, _ptr(r_hi), _ptr(r_lo), _stream()), 'pvsg_conv2d_tc')
        sp = Split(y_hi, y_lo) if want_split else None
        if out_mode == 'both':
            return y, sp
        return sp if (out_mode == 'split' and sp is not None) else y
    if (ENGINE[0] == 'tc' and x.dim() == 4 and x.shape[-1] % 64 != 0 and x.is_contiguous() and weight.is_contiguous()
            and out is None and weight.shape[1] * weight.shape[2] * weight.shape[3] <= 2048):
        # small-Cin conv (RGB stem): patches gathered straight into operand planes, then the GEMM path
        B, H, W, Cin = _f32(x, 'x').shape
        Cout, R, S, _ = _f32(weight, 'weight').shape
        kreal = R * S * Cin
        kpad = (kreal + 63) // 64 * 64
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        M = B * OH * OW
        hi = torch.empty(M, kpad, device=x.device, dtype=torch.bfloat16)
        lo = torch.empty(M, kpad, device=x.device, dtype=torch.bfloat16)
        _l.check(lib.pvsg_im2col_split(_ptr(x), _ptr(hi), _ptr(lo), B, H, W, Cin, R, S, stride, pad, kpad, _stream()),
                 'pvsg_im2col_split')
        cache = weight.__dict__.get('_pvsg_padded')
        if cache is None or cache[0] != (kpad, weight._version):
            wp = torch.zeros(Cout, kpad, device=x.device, dtype=torch.float32)
            wp[:, :kreal] = weight.view(Cout, kreal)
            cache = ((kpad, weight._version), wp)
            weight._pvsg_padded = cache
        r = residual.view(M, Cout) if residual is not None else None
        return _linear_tc(lib, Split(hi, lo), cache[1], bias, r, act, None, (B, OH, OW), M, Cout, kpad, out_mode)
    _f32(x, 'x'), _f32(weight, 'weight')
    if not (x.is_contiguous() and weight.is_contiguous() and x.dim() == 4 and weight.dim() == 4):
        raise _l.PvsgError('conv2d_nhwc: contiguous 4-D tensors required')
    B, H, W, Cin = x.shape
    Cout, R, S, Cw = weight.shape
    if Cw != Cin:
        raise _l.PvsgError('conv2d_nhwc: channel mismatch')
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty(B, OH, OW, Cout, device=x.device, dtype=torch.float32)
    if residual is not None and (tuple(residual.shape) != (B, OH, OW, Cout) or not residual.is_contiguous()):
        raise _l.PvsgError('conv2d_nhwc: bad residual')
    _l.check(lib.pvsg_conv2d_nhwc(_ptr(x), _ptr(weight), _ptr(_f32(bias)), _ptr(_f32(residual)), _ptr(out), B, H,
                                  W, Cin, Cout, R, S, stride, pad, act, _stream()), 'pvsg_conv2d_nhwc')
    return (out, None) if out_mode == 'both' else out


def stem_weight(weight):
    """[Cout,7,7,C<=4] stem weight (token-major taps) -> [Cout,4,1,64] for ``stem7x7s2``:
    W2[co, j, 0, par*32 + s*4 + c] = w[co, 2j+par, s, c]."""
    Cout, R, S, C = weight.shape
    if (R, S) != (7, 7) or C > 4:
        raise _l.PvsgError('stem_weight: 7x7 kernel with <= 4 input channels expected')
    w2 = torch.zeros(Cout, 4, 2, 8, 4, device=weight.device, dtype=torch.float32)
    wp = torch.zeros(Cout, 8, 8, 4, device=weight.device, dtype=torch.float32)
    wp[:, :7, :7, :C] = weight
    w2[:] = wp.view(Cout, 4, 2, 8, 4)
    return w2.view(Cout, 4, 1, 64).contiguous()


def stem7x7s2(x_nchw, w2, bias, act=ACT_RELU, out_mode='f32'):
    """ResNet conv1 (7x7 / stride 2 / pad 3) on the NCHW fp32 frame: one packing pass into row-pair
    operand planes, then the tcgen05 conv engine as a 4 x 1 convolution (no im2col buffer).
    Returns token-major [B,OH,OW,Cout]."""
    lib = _l.load()
    B, C, H, W = _f32(x_nchw, 'x').shape
    if not x_nchw.is_contiguous():
        raise _l.PvsgError('stem7x7s2: contiguous NCHW input required')
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    hi = torch.empty(B, OH + 3, OW, 64, device=x_nchw.device, dtype=torch.bfloat16)
    lo = torch.empty(B, OH + 3, OW, 64, device=x_nchw.device, dtype=torch.bfloat16)
    _l.check(lib.pvsg_stem7x7s2_pack(_ptr(x_nchw), _ptr(hi), _ptr(lo), B, C, H, W, _stream()), 'pvsg_stem7x7s2_pack')
    return conv2d_nhwc(Split(hi, lo), w2, bias, stride=1, pad=0, act=act, out_mode=out_mode)


def maxpool3x3s2_nhwc(x):
    lib = _l.load()
    B, H, W, C = _f32(x).shape
    out = torch.empty(B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_maxpool3x3s2_nhwc(_ptr(x.contiguous()), _ptr(out), B, H, W, C, _stream()), 'pvsg_maxpool')
    return out


def maxpool3x3s2_nhwc_backward(x, dy):
    lib = _l.load()
    B, H, W, C = _f32(x).shape
    dx = torch.empty_like(x)
    _l.check(lib.pvsg_maxpool3x3s2_nhwc_backward(_ptr(x.contiguous()), _ptr(_f32(dy).contiguous()), _ptr(dx), B, H, W, C, _stream()),
             'pvsg_maxpool_backward')
    return dx


def nchw_to_nhwc(x):
    lib = _l.load()
    B, C, H, W = _f32(x).shape
    out = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_nchw_to_nhwc(_ptr(x.contiguous()), _ptr(out), B, C, H, W, _stream()), 'pvsg_nchw_to_nhwc')
    return out


def nhwc_to_nchw(x):
    lib = _l.load()
    B, H, W, C = _f32(x).shape
    out = torch.empty(B, C, H, W, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_nhwc_to_nchw(_ptr(x.contiguous()), _ptr(out), B, C, H, W, _stream()), 'pvsg_nhwc_to_nchw')
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None, out_split=False, add=None):
    """LayerNorm over the last axis.  out_split=True additionally returns the Split planes of the
    result (None on the SIMT engine): (y, planes).  With ``add`` (same shape as x) also the planes of
    y + add: (y, planes, planes_of_sum)."""
    lib = _l.load()
    _f32(x)
    if not x.is_contiguous():
        raise _l.PvsgError('layernorm: contiguous input required')
    C = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    if add is not None and not out_split:
        raise _l.PvsgError('layernorm: add needs out_split')
    if out_split and ENGINE[0] == 'tc' and C % 64 == 0 and x.numel() // C > SKINNY_M:
        hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        if add is not None:
            if tuple(add.shape) != tuple(x.shape) or not add.is_contiguous():
                raise _l.PvsgError('layernorm: add must match x')
            shi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
            slo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
            _l.check(lib.pvsg_layernorm_split2(_ptr(x), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(out), _ptr(hi),
                                               _ptr(lo), _ptr(_f32(add)), _ptr(shi), _ptr(slo), x.numel() // C, C, eps,
                                               _stream()), 'pvsg_layernorm_split2')
            return out, Split(hi, lo), Split(shi, slo)
        _l.check(lib.pvsg_layernorm_split(_ptr(x), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(out), _ptr(hi), _ptr(lo),
                                          x.numel() // C, C, eps, _stream()), 'pvsg_layernorm_split')
        return out, Split(hi, lo)
    _l.check(lib.pvsg_layernorm(_ptr(x), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(out), x.numel() // C, C, eps,
                                _stream()), 'pvsg_layernorm')
    if add is not None:
        return out, None, None
    return (out, None) if out_split else out


def groupnorm_nhwc(x, gamma, beta, groups=32, eps=1e-5, act=ACT_NONE, out=None, out_mode='f32'):
    """x [B, H, W, C] (or [B, HW, C]) token-major.  out_mode as in ``linear`` ('split': planes only)."""
    lib = _l.load()
    _f32(x)
    if not x.is_contiguous():
        raise _l.PvsgError('groupnorm: contiguous input required')
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    stats = torch.empty(B * groups * 2, device=x.device, dtype=torch.float64)
    if out_mode != 'f32' and ENGINE[0] == 'tc' and C % 64 == 0 and out is None:
        y = torch.empty_like(x) if out_mode == 'both' else None
        hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        _l.check(lib.pvsg_groupnorm_nhwc_split(_ptr(x), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(y), _ptr(hi),
                                               _ptr(lo), _ptr(stats), B, HW, C, groups, eps, act, _stream()),
                 'pvsg_groupnorm_nhwc_split')
        return (y, Split(hi, lo)) if out_mode == 'both' else Split(hi, lo)
    if out is None:
        out = torch.empty_like(x)
    _l.check(lib.pvsg_groupnorm_nhwc(_ptr(x), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(out), _ptr(stats), B, HW,
                                     C, groups, eps, act, _stream()), 'pvsg_groupnorm_nhwc')
    return (out, None) if out_mode == 'both' else out


def add_rowvec(x, v, out=None):
    lib = _l.load()
    _f32(x), _f32(v)
    C = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    _l.check(lib.pvsg_add_rowvec(_ptr(x.contiguous()), _ptr(v.contiguous()), _ptr(out), x.numel() // C, C,
                                 _stream()), 'pvsg_add_rowvec')
    return out


def bilinear_resize_nhwc(src, out_hw, out=None, accumulate=False, out_split=False):
    """F.interpolate(mode='bilinear', align_corners=False) on token-major [B,H,W,C].  src may be a
    batch-strided view (frames that are slices of a longer token buffer).  out_split=True also returns
    the operand planes of the final values: (out, Split)."""
    lib = _l.load()
    B, IH, IW, C = _f32(src).shape
    OH, OW = out_hw
    if src.stride(3) != 1 or src.stride(2) != C or src.stride(1) != IW * C:
        src = src.contiguous()
    if out is None:
        if accumulate:
            raise _l.PvsgError('bilinear_resize: accumulate needs out')
        out = torch.empty(B, OH, OW, C, device=src.device, dtype=torch.float32)
    hi = lo = None
    if out_split and ENGINE[0] == 'tc' and C % 64 == 0:
        hi = torch.empty(B, OH, OW, C, device=src.device, dtype=torch.bfloat16)
        lo = torch.empty(B, OH, OW, C, device=src.device, dtype=torch.bfloat16)
    _l.check(lib.pvsg_bilinear_resize_nhwc_ex(_ptr(src), src.stride(0) if B > 1 else IH * IW * C, _ptr(out), _ptr(hi),
                                              _ptr(lo), B, IH, IW, OH, OW, C, 1 if accumulate else 0, _stream()),
             'pvsg_bilinear_resize_nhwc')
    if out_split:
        return out, (Split(hi, lo) if hi is not None else None)
    return out


def bilinear_resize_scaled(src, out_hw, scale):
    """F.interpolate(scale_factor=1/scale, mode='bilinear') on token-major [B,H,W,C]: explicit source-coordinate scale."""
    lib = _l.load()
    B, IH, IW, C = _f32(src).shape
    out = torch.empty(B, out_hw[0], out_hw[1], C, device=src.device, dtype=torch.float32)
    _l.check(lib.pvsg_bilinear_resize_scaled(_ptr(src.contiguous()), _ptr(out), B, IH, IW, out_hw[0], out_hw[1], C, float(scale),
                                             float(scale), _stream()), 'pvsg_bilinear_resize_scaled')
    return out


def sine_pe(h, w, device, t=0, num_feats=128, temperature=10000, scale=6.283185307179586, eps=1e-6,
            add_vec=None):
    """Token-major sine positional encoding [max(t,1)*h*w, 2*num_feats]."""
    lib = _l.load()
    # temperature tables exactly as the reference builds them (position_encoding.py:84-88)
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = (temperature ** (2 * (dim_t // 2) / num_feats)).to(device)
    dim_t_z = torch.arange(num_feats * 2, dtype=torch.float32)
    dim_t_z = (temperature ** (2 * (dim_t_z // 2) / (num_feats * 2))).to(device)
    out = torch.empty(max(t, 1) * h * w, 2 * num_feats, device=device, dtype=torch.float32)
    _l.check(lib.pvsg_sine_pe(_ptr(out), _ptr(dim_t), _ptr(dim_t_z), _ptr(_f32(add_vec)), t, h, w, num_feats,
                              scale, eps, _stream()), 'pvsg_sine_pe')
    return out


def _levels(spatial_shapes):
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    L = len(shapes)
    ss = (ctypes.c_int64 * (2 * L))(*[v for hw in shapes for v in hw])
    starts, acc = [], 0
    for h, w in shapes:
        starts.append(acc)
        acc += h * w
    ls = (ctypes.c_int64 * L)(*starts)
    return ss, ls, L, acc


def msda_forward(value, spatial_shapes, sampling_locations, attention_weights):
    """mmcv MultiScaleDeformableAttnFunction.forward: value [B,N,H,D], sampling_locations
    [B,Nq,H,L,P,2], attention_weights [B,Nq,H,L,P] -> [B,Nq,H*D]."""
    lib = _l.load()
    B, N, H, D = _f32(value).shape
    _, Nq, _, L, P, _ = _f32(sampling_locations).shape
    ss, ls, L2, tot = _levels(spatial_shapes)
    if L2 != L or tot != N:
        raise _l.PvsgError('msda_forward: spatial_shapes do not match value')
    out = torch.empty(B, Nq, H * D, device=value.device, dtype=torch.float32)
    _l.check(lib.pvsg_msda_forward(_ptr(value.contiguous()), ss, ls, _ptr(sampling_locations.contiguous()),
                                   _ptr(_f32(attention_weights).contiguous()), _ptr(out), B, N, Nq, H, D, L, P,
                                   _stream()), 'pvsg_msda_forward')
    return out


def msda_fused_forward(value, spatial_shapes, proj, ref, num_heads=8, num_points=4, out_mode='f32'):
    """value [B,N,H*D]; proj [B,Nq,H*L*P*3] raw (offsets | logits); ref [Nq,2].
    out_mode='split': the result leaves as operand planes only (tcgen05 engine, L*P <= 16)."""
    lib = _l.load()
    B, N, C = _f32(value).shape
    ss, ls, L, tot = _levels(spatial_shapes)
    Nq = proj.shape[1]
    D = C // num_heads
    if tot != N or _f32(proj).shape[2] != num_heads * L * num_points * 3 or tuple(_f32(ref).shape) != (Nq, 2):
        raise _l.PvsgError('msda_fused_forward: shape mismatch')
    if out_mode == 'split' and ENGINE[0] == 'tc' and L * num_points <= 16 and C % 64 == 0:
        hi = torch.empty(B, Nq, C, device=value.device, dtype=torch.bfloat16)
        lo = torch.empty(B, Nq, C, device=value.device, dtype=torch.bfloat16)
        _l.check(lib.pvsg_msda_fused_forward_split(_ptr(value.contiguous()), ss, ls, _ptr(proj.contiguous()),
                                                   _ptr(ref.contiguous()), None, _ptr(hi), _ptr(lo), B, N, Nq,
                                                   num_heads, D, L, num_points, _stream()),
                 'pvsg_msda_fused_forward_split')
        return Split(hi, lo)
    out = torch.empty(B, Nq, C, device=value.device, dtype=torch.float32)
    _l.check(lib.pvsg_msda_fused_forward(_ptr(value.contiguous()), ss, ls, _ptr(proj.contiguous()),
                                         _ptr(ref.contiguous()), _ptr(out), B, N, Nq, num_heads, D, L,
                                         num_points, _stream()), 'pvsg_msda_fused_forward')
    return out


def msda_backward(value, spatial_shapes, sampling_locations, attention_weights, grad_out):
    """mmcv MultiScaleDeformableAttnFunction.backward: value [B,N,H,D], grad_out [B,Nq,H*D] ->
    (grad_value, grad_sampling_locations, grad_attention_weights)."""
    lib = _l.load()
    B, N, H, D = _f32(value).shape
    _, Nq, _, L, P, _ = _f32(sampling_locations).shape
    ss, ls, L2, tot = _levels(spatial_shapes)
    if L2 != L or tot != N or tuple(_f32(grad_out).shape) != (B, Nq, H * D):
        raise _l.PvsgError('msda_backward: shape mismatch')
    gv = torch.empty_like(value, memory_format=torch.contiguous_format)
    gl = torch.empty(sampling_locations.shape, device=value.device, dtype=torch.float32)
    ga = torch.empty(attention_weights.shape, device=value.device, dtype=torch.float32)
    _l.check(lib.pvsg_msda_backward(_ptr(value.contiguous()), ss, ls, _ptr(sampling_locations.contiguous()),
                                    _ptr(_f32(attention_weights).contiguous()), _ptr(grad_out.contiguous()), _ptr(gv), _ptr(gl),
                                    _ptr(ga), B, N, Nq, H, D, L, P, _stream()), 'pvsg_msda_backward')
    return gv, gl, ga


def point_sample(maps, points):
    """maps [n,H,W], points [n,K,2] or [K,2] (x, y in [0,1]) -> [n,K] (mmcv.ops.point_sample, align_corners=False)."""
    lib = _l.load()
    n, H, W = _f32(maps).shape
    shared = points.dim() == 2
    K = points.shape[-2]
    if _f32(points).shape[-1] != 2 or (not shared and points.shape[0] != n):
        raise _l.PvsgError('point_sample: points must be [n,K,2] or [K,2]')
    out = torch.empty(n, K, device=maps.device, dtype=torch.float32)
    if n:
        _l.check(lib.pvsg_point_sample(_ptr(maps.contiguous()), _ptr(points.contiguous()), _ptr(out), n, H, W, K, 0 if shared else 1,
                                       _stream()), 'pvsg_point_sample')
    return out


def point_sample_backward(grad_out, points, hw):
    lib = _l.load()
    n, K = _f32(grad_out).shape
    gm = torch.empty(n, hw[0], hw[1], device=grad_out.device, dtype=torch.float32)
    if n:
        _l.check(lib.pvsg_point_sample_backward(_ptr(grad_out.contiguous()), _ptr(_f32(points).contiguous()), _ptr(gm), n, hw[0], hw[1],
                                                K, 0 if points.dim() == 2 else 1, _stream()), 'pvsg_point_sample_backward')
    return gm


def mask_point_losses(logits, targets, dice_eps=1.0, bce_grad_scale=0.0, dice_grad_scale=0.0, want_grad=False):
    """logits / targets [n,K] -> (sums fp32 [2] = (sum of BCE terms, sum of dice losses), grad or None)."""
    lib = _l.load()
    n, K = _f32(logits).shape
    sums = torch.empty(2, device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits, memory_format=torch.contiguous_format) if want_grad else None
    _l.check(lib.pvsg_mask_point_losses(_ptr(logits.contiguous()), _ptr(_f32(targets).contiguous()), n, K, float(dice_eps),
                                        float(bce_grad_scale), float(dice_grad_scale), _ptr(sums), _ptr(grad), _stream()),
             'pvsg_mask_point_losses')
    return sums, grad


def weighted_ce(logits, labels, class_weight, label_weight=None, grad_scale=0.0, want_grad=False):
    """logits [R,C], labels int64 [R] -> (sums fp32 [2] = (weighted loss sum, sum of class_weight[labels]), grad or None)."""
    lib = _l.load()
    R, C = _f32(logits).shape
    if labels.dtype != torch.int64 or labels.numel() != R or _f32(class_weight).numel() != C:
        raise _l.PvsgError('weighted_ce: int64 labels [R] and class_weight [C] expected')
    sums = torch.empty(2, device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits, memory_format=torch.contiguous_format) if want_grad else None
    _l.check(lib.pvsg_weighted_ce(_ptr(logits.contiguous()), _ptr(labels.contiguous()), _ptr(class_weight.contiguous()),
                                  _ptr(_f32(label_weight)), R, C, float(grad_scale), _ptr(sums), _ptr(grad), _stream()),
             'pvsg_weighted_ce')
    return sums, grad


def mask_match_cost(cls_logits, gt_labels, pred_points, gt_points, w_cls=2.0, w_mask=5.0, w_dice=5.0, dice_eps=1.0):
    """cls_logits [Q,C], gt_labels int64 [G], pred_points [Q,K], gt_points [G,K] -> cost [Q,G] (MaskHungarianAssigner)."""
    lib = _l.load()
    Q, C = _f32(cls_logits).shape
    G, K = _f32(gt_points).shape
    cost = torch.empty(Q, G, device=cls_logits.device, dtype=torch.float32)
    if G:
        _l.check(lib.pvsg_mask_match_cost(_ptr(cls_logits.contiguous()), _ptr(gt_labels.contiguous()), _ptr(_f32(pred_points).contiguous()),
                                          _ptr(gt_points.contiguous()), Q, G, C, K, float(w_cls), float(w_mask), float(w_dice),
                                          float(dice_eps), _ptr(cost), _stream()), 'pvsg_mask_match_cost')
    return cost


def groupnorm_nhwc_backward(x, gamma, beta, dy, groups=32, eps=1e-5, relu=False):
    """nn.GroupNorm (+ ReLU) backward on token-major x [B,...,C] -> (dx, dgamma, dbeta)."""
    lib = _l.load()
    x, dy = _f32(x).contiguous(), _f32(dy).contiguous()
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    dx = torch.empty_like(x)
    dg = torch.empty(C, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    stats = torch.empty(B * groups * 4, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_groupnorm_nhwc_backward(_ptr(x), _ptr(_f32(gamma).contiguous()), _ptr(_f32(beta).contiguous()), _ptr(dy), _ptr(dx),
                                              _ptr(dg), _ptr(db), _ptr(stats), B, HW, C, groups, eps, 1 if relu else 0, _stream()),
             'pvsg_groupnorm_nhwc_backward')
    return dx, dg, db


def bilinear_resize_nhwc_backward(dout, in_hw):
    """Adjoint of bilinear_resize_nhwc: dout [B,OH,OW,C] -> dsrc [B,IH,IW,C]."""
    lib = _l.load()
    dout = _f32(dout).contiguous()
    B, OH, OW, C = dout.shape
    dsrc = torch.empty(B, in_hw[0], in_hw[1], C, device=dout.device, dtype=torch.float32)
    _l.check(lib.pvsg_bilinear_resize_nhwc_backward(_ptr(dout), _ptr(dsrc), B, in_hw[0], in_hw[1], OH, OW, C, _stream()),
             'pvsg_bilinear_resize_nhwc_backward')
    return dsrc


def msda_proj_expand(proj, ref, spatial_shapes, num_heads=8, num_points=4):
    """raw projections [B,Nq,H*L*P*3] -> (sampling_locations [B,Nq,H,L,P,2], attention_weights [B,Nq,H,L,P])."""
    lib = _l.load()
    ss, ls, L, tot = _levels(spatial_shapes)
    B, Nq = _f32(proj).shape[:2]
    loc = torch.empty(B, Nq, num_heads, L, num_points, 2, device=proj.device, dtype=torch.float32)
    aw = torch.empty(B, Nq, num_heads, L, num_points, device=proj.device, dtype=torch.float32)
    _l.check(lib.pvsg_msda_proj_expand(_ptr(proj.contiguous()), _ptr(_f32(ref).contiguous()), ss, _ptr(loc), _ptr(aw), B, Nq,
                                       num_heads, L, num_points, _stream()), 'pvsg_msda_proj_expand')
    return loc, aw


def msda_proj_backward(aw, dloc, daw, spatial_shapes):
    lib = _l.load()
    ss, ls, L, tot = _levels(spatial_shapes)
    B, Nq, H, _, P = aw.shape
    dproj = torch.empty(B, Nq, H * L * P * 3, device=aw.device, dtype=torch.float32)
    _l.check(lib.pvsg_msda_proj_backward(_ptr(aw.contiguous()), _ptr(_f32(dloc).contiguous()), _ptr(_f32(daw).contiguous()), ss,
                                         _ptr(dproj), B, Nq, H, L, P, _stream()), 'pvsg_msda_proj_backward')
    return dproj


def splitk_plan(T, M, N):
    """Split-K plan of a token-reduction GEMM [M,T] x [N,T]^T: (chunks S, chunk length Kc, padded T = S * Kc).  Enough
    chunks that S * tiles covers the SMs twice, chunk length a multiple of the 64-column k-block."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    S = max(1, min((296 + tiles - 1) // tiles, (T + 255) // 256))
    Kc = ((T + S - 1) // S + 63) // 64 * 64
    S = (T + Kc - 1) // Kc
    return S, Kc, S * Kc


def alloc_planes(rows, T, ldt, device):
    """(hi, lo) bf16 [rows, ldt] for transpose_split: the kernel writes (zeros included) up to the 64-column tile that holds
    column T - 1, so only the columns beyond it are cleared here."""
    hi = torch.empty(rows, ldt, device=device, dtype=torch.bfloat16)
    lo = torch.empty(rows, ldt, device=device, dtype=torch.bfloat16)
    t64 = (T + 63) // 64 * 64
    if t64 < ldt:
        hi[:, t64:].zero_()
        lo[:, t64:].zero_()
    return hi, lo


def transpose_split(x, ldt, add=None, hi=None, lo=None, row0=0):
    """Token-major view x [.., C] (2-D [T,C] with any row stride, or 4-D [B,OH,OW,C] with any strides; unit channel
    stride) -> planes [C, ldt] (zero beyond T).  hi / lo + row0: write into rows row0.. of existing [R, ldt] planes."""
    lib = _l.load()
    _f32(x)
    if x.stride(-1) != 1 or x.dim() not in (2, 4):
        raise _l.PvsgError('transpose_split: 2-D or 4-D token-major view with unit channel stride required')
    C = x.shape[-1]
    if x.dim() == 2:
        T, OH, OW, sb, sh, sw = x.shape[0], 1, x.shape[0], 0, 0, x.stride(0)
    else:
        B, OH, OW, _ = x.shape
        T, sb, sh, sw = B * OH * OW, x.stride(0), x.stride(1), x.stride(2)
    if add is not None and (add.shape != x.shape or add.stride() != x.stride()):
        raise _l.PvsgError('transpose_split: add must share the layout of x')
    if hi is None:
        hi, lo = alloc_planes(C, T, ldt, x.device)
    if hi.shape[1] != ldt or row0 + C > hi.shape[0] or not hi.is_contiguous() or not lo.is_contiguous():
        raise _l.PvsgError('transpose_split: bad destination planes')
    _l.check(lib.pvsg_transpose_split(_ptr(x), _ptr(add), hi.data_ptr() + 2 * row0 * ldt, lo.data_ptr() + 2 * row0 * ldt, T, C,
                                      OH, OW, sb, sh, sw, ldt, _stream()), 'pvsg_transpose_split')
    return hi, lo


def splitk_gemm(a, w, S, Kc):
    """a = (hi, lo) [M, S*Kc], w = (hi, lo) [N, S*Kc] bf16 planes -> fp32 [M, N] = a w^T, reduced over S K-chunks: one
    batched tcgen05 launch over the chunk views, then a column sum over the partial results."""
    lib = _l.load()
    (a_hi, a_lo), (w_hi, w_lo) = a, w
    M, Tp = a_hi.shape
    N = w_hi.shape[0]
    if Tp != S * Kc or w_hi.shape[1] != Tp:
        raise _l.PvsgError('splitk_gemm: plane widths do not match the plan')
    part = torch.empty(S, M, N, device=a_hi.device, dtype=torch.float32)
    _l.check(lib.pvsg_linear_tc_batched(_ptr(a_hi), _ptr(a_lo), Tp, Kc if S > 1 else M * Tp, _ptr(w_hi), _ptr(w_lo), Tp,
                                        Kc if S > 1 else N * Tp, _ptr(part), None, None, N, S, M, N, Kc, _stream()),
             'pvsg_linear_tc_batched')
    return part[0] if S == 1 else colsum(part.view(S, M * N)).view(M, N)


def layernorm_backward(x, gamma, dy, eps=1e-5):
    """nn.LayerNorm backward over the last axis -> (dx, dgamma, dbeta)."""
    lib = _l.load()
    C = _f32(x).shape[-1]
    x, dy = x.contiguous(), _f32(dy).contiguous()
    dx = torch.empty_like(x)
    dg = torch.empty(C, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_layernorm_backward(_ptr(x), _ptr(_f32(gamma).contiguous()), _ptr(dy), _ptr(dx), _ptr(dg), _ptr(db),
                                         x.numel() // C, C, eps, _stream()), 'pvsg_layernorm_backward')
    return dx, dg, db


def relu_backward(dy, y):
    lib = _l.load()
    dy, y = _f32(dy).contiguous(), _f32(y).contiguous()
    dx = torch.empty_like(dy)
    _l.check(lib.pvsg_relu_backward(_ptr(dy), _ptr(y), _ptr(dx), dy.numel(), _stream()), 'pvsg_relu_backward')
    return dx


def colsum(x):
    """x [M,N] (row stride >= N) -> [N] column sums."""
    lib = _l.load()
    x2, M, N, ld = _rows(x, 'x')
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_colsum(_ptr(x2), _ptr(out), M, N, ld, _stream()), 'pvsg_colsum')
    return out


def _attn_train_args(q, k, v, num_heads, mask, row_open):
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _f32(t, n)
        if t.dim() != 3 or t.stride(2) != 1:
            raise _l.PvsgError(f'attention_train: {n} must be [B,L,E] with unit inner stride')
    B, Lq, E = q.shape
    Lk = k.shape[1]
    if E // num_heads != 32 or E % num_heads:
        raise _l.PvsgError('attention_train: head dim 32 only')
    if mask is not None and not (mask.dtype == torch.uint8 and mask.is_contiguous() and tuple(mask.shape) == (B, Lq, Lk)):
        raise _l.PvsgError('attention_train: mask must be contiguous uint8 [B,Lq,Lk]')
    if row_open is not None and not (row_open.dtype == torch.int32 and row_open.is_contiguous()):
        raise _l.PvsgError('attention_train: row_open must be contiguous int32')
    return B, Lq, Lk, E


def attention_train_forward(q, k, v, num_heads, mask=None, row_open=None):
    """Training-time attention: -> (out [B,Lq,E], lse [B,H,Lq]).  On the tcgen05 engine the forward is the inference kernel
    (csrc/attention_t5.cu) on freshly split K / V planes, with the row log-sum-exp as a second output; the exact-fp32 SIMT
    kernel (PVSG_ENGINE=simt, or layouts the tensor-map path does not take) computes the same pair."""
    lib = _l.load()
    B, Lq, Lk, E = _attn_train_args(q, k, v, num_heads, mask, row_open)
    out = torch.empty(B, Lq, E, device=q.device, dtype=torch.float32)
    lse = torch.empty(B, num_heads, Lq, device=q.device, dtype=torch.float32)
    if (ENGINE[0] == 'tc' and _l.ATTN_IMPL[0] == 't5' and B * num_heads <= 65535 and q.data_ptr() % 16 == 0
            and q.stride(0) % 4 == 0 and q.stride(1) % 4 == 0):
        k_hi, k_lo = split_bf16(k.contiguous())
        v_hi, v_lo = split_bf16(v.contiguous())
        ws = torch.empty(lib.pvsg_attention_t5_workspace_bytes(B, num_heads, Lq, Lk, 32), device=q.device, dtype=torch.uint8)
        _l.check(lib.pvsg_attention_t5_lse(_ptr(q), _ptr(k_hi), _ptr(k_lo), _ptr(v_hi), _ptr(v_lo), _ptr(mask), _ptr(row_open),
                                           _ptr(out), _ptr(lse), _ptr(ws), B, num_heads, Lq, Lk, 32, q.stride(0), q.stride(1),
                                           k_hi.stride(0), k_hi.stride(1), v_hi.stride(0), v_hi.stride(1), out.stride(0),
                                           out.stride(1), 32 ** -0.5, _stream()), 'pvsg_attention_t5_lse')
        return out, lse
    _l.check(lib.pvsg_attention_train_forward(_ptr(q), _ptr(k), _ptr(v), _ptr(mask), _ptr(row_open), _ptr(out), _ptr(lse), B,
                                              num_heads, Lq, Lk, 32, q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                                              v.stride(0), v.stride(1), out.stride(0), out.stride(1), 32 ** -0.5, _stream()),
             'pvsg_attention_train_forward')
    return out, lse


def attention_train_backward(q, k, v, num_heads, mask, row_open, out, dout, lse):
    """-> (dq [B,Lq,E], dk [B,Lk,E], dv [B,Lk,E])."""
    lib = _l.load()
    B, Lq, Lk, E = _attn_train_args(q, k, v, num_heads, mask, row_open)
    dout = _f32(dout).contiguous()
    if not out.is_contiguous():
        raise _l.PvsgError('attention_train_backward: contiguous forward output required')
    dq = torch.empty(B, Lq, E, device=q.device, dtype=torch.float32)
    dk = torch.empty(B, Lk, E, device=q.device, dtype=torch.float32)
    dv = torch.empty(B, Lk, E, device=q.device, dtype=torch.float32)
    delta = torch.empty(B, num_heads, Lq, device=q.device, dtype=torch.float32)
    _l.check(lib.pvsg_attention_train_backward(_ptr(q), _ptr(k), _ptr(v), _ptr(mask), _ptr(row_open), _ptr(out), _ptr(dout),
                                               _ptr(lse), _ptr(delta), _ptr(dq), _ptr(dk), _ptr(dv), B, num_heads, Lq, Lk, 32,
                                               q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1),
                                               out.stride(0), out.stride(1), 32 ** -0.5, _stream()),
             'pvsg_attention_train_backward')
    return dq, dk, dv


def attention(q, k, v, num_heads, mask=None, row_open=None, scale=None, out=None):
    """q [B,Lq,E], k/v [B,Lk,E] (any batch / token strides, unit inner stride) -> [B,Lq,E].
    mask uint8 [B,Lq,Lk] (non-zero = blocked), row_open int32 [B,Lq].
    k / v may be ``Split`` operand planes (head dim 32): tensor-core path (csrc/attention_mma.cu)."""
    lib = _l.load()
    if isinstance(k, Split) or isinstance(v, Split):
        if not (isinstance(k, Split) and isinstance(v, Split)):
            raise _l.PvsgError('attention: k and v must both be planes')
        _f32(q, 'q')
        B, Lq, E = q.shape
        Lk = k.shape[1]
        D = E // num_heads
        if D not in (32, 128) or q.stride(2) != 1 or tuple(k.shape) != (B, Lk, E) or tuple(v.shape) != (B, Lk, E):
            raise _l.PvsgError('attention: plane inputs need head dim 32 / 128 and [B,Lk,E] planes')
        for t in (k.hi, k.lo, v.hi, v.lo):
            if t.stride(2) != 1:
                raise _l.PvsgError('attention: planes need unit inner stride')
        if scale is None:
            scale = float(D) ** -0.5
        if out is None:
            out = torch.empty(B, Lq, E, device=q.device, dtype=torch.float32)
        if mask is not None and not (mask.dtype == torch.uint8 and mask.is_contiguous() and
                                     tuple(mask.shape) == (B, Lq, Lk)):
            raise _l.PvsgError('attention: mask must be contiguous uint8 [B,Lq,Lk]')
        if row_open is not None and not (row_open.dtype == torch.int32 and row_open.is_contiguous()):
            raise _l.PvsgError('attention: row_open must be contiguous int32')
        if (k.hi.stride() != k.lo.stride()) or (v.hi.stride() != v.lo.stride()):
            raise _l.PvsgError('attention: hi / lo planes must share their layout')
        t5 = (_l.ATTN_IMPL[0] == 't5' and k.hi.stride() == v.hi.stride() and q.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0
              and q.stride(0) % 4 == 0 and q.stride(1) % 4 == 0 and out.stride(0) % 4 == 0 and out.stride(1) % 4 == 0
              and all(t.data_ptr() % 16 == 0 for t in (k.hi, k.lo, v.hi, v.lo)) and B * num_heads <= 65535)
        if t5:      # tcgen05 / TMEM kernel (csrc/attention_t5.cu)
            ws = torch.empty(lib.pvsg_attention_t5_workspace_bytes(B, num_heads, Lq, Lk, D), device=q.device, dtype=torch.uint8)
            _l.check(lib.pvsg_attention_t5(_ptr(q), _ptr(k.hi), _ptr(k.lo), _ptr(v.hi), _ptr(v.lo), _ptr(mask), _ptr(row_open),
                                           _ptr(out), _ptr(ws), B, num_heads, Lq, Lk, D, q.stride(0), q.stride(1),
                                           k.hi.stride(0), k.hi.stride(1), v.hi.stride(0), v.hi.stride(1), out.stride(0),
                                           out.stride(1), scale, _stream()), 'pvsg_attention_t5')
            return out
        ws = torch.empty(lib.pvsg_attention_tc_workspace_bytes(B, num_heads, Lq, Lk, D), device=q.device,
                         dtype=torch.uint8)
        _l.check(lib.pvsg_attention_tc(_ptr(q), _ptr(k.hi), _ptr(k.lo), _ptr(v.hi), _ptr(v.lo), _ptr(mask),
                                       _ptr(row_open), _ptr(out), _ptr(ws), B, num_heads, Lq, Lk, D, q.stride(0),
                                       q.stride(1), k.hi.stride(0), k.hi.stride(1), v.hi.stride(0), v.hi.stride(1),
                                       out.stride(0), out.stride(1), scale, _stream()), 'pvsg_attention_tc')
        return out
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _f32(t, n)
        if t.dim() != 3 or t.stride(2) != 1:
            raise _l.PvsgError(f'attention: {n} must be [B,L,E] with unit inner stride')
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // num_heads
    if scale is None:
        scale = float(D) ** -0.5
    if out is None:
        out = torch.empty(B, Lq, E, device=q.device, dtype=torch.float32)
    if mask is not None and not (mask.dtype == torch.uint8 and mask.is_contiguous() and
                                 tuple(mask.shape) == (B, Lq, Lk)):
        raise _l.PvsgError('attention: mask must be contiguous uint8 [B,Lq,Lk]')
    if row_open is not None and not (row_open.dtype == torch.int32 and row_open.is_contiguous()):
        raise _l.PvsgError('attention: row_open must be contiguous int32')
    nbytes = lib.pvsg_attention_workspace_bytes(B, num_heads, Lq, Lk, D)
    ws = torch.empty(nbytes, device=q.device, dtype=torch.uint8)
    _l.check(lib.pvsg_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(mask), _ptr(row_open), _ptr(out), _ptr(ws), B,
                                num_heads, Lq, Lk, D, q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                                v.stride(0), v.stride(1), out.stride(0), out.stride(1), scale, _stream()),
             'pvsg_attention')
    return out


def mask_logits(embed, feat, want_logits=True, want_mask=False, feat_planes=None):
    """embed [B,Q,C], feat [B,P,C] token-major -> logits [B,Q,P] and/or (mask uint8, row_open int32).
    feat_planes: optional precomputed split_bf16(feat) (the mask features are reused by all ten
    prediction heads of a frame)."""
    lib = _l.load()
    B, Q, C = embed.shape
    P = _f32(feat).shape[1]
    if isinstance(embed, Split) or (ENGINE[0] == 'tc' and C % 64 == 0 and embed.is_contiguous()
                                    and feat.is_contiguous()):
        dev = embed.device
        logits = torch.empty(B, Q, P, device=dev, dtype=torch.float32) if want_logits else None
        mask = torch.empty(B, Q, P, device=dev, dtype=torch.uint8) if want_mask else None
        row_open = torch.zeros(B, Q, device=dev, dtype=torch.int32) if want_mask else None
        e_hi, e_lo = embed if isinstance(embed, Split) else split_bf16(embed)
        f_hi, f_lo = feat_planes if feat_planes is not None else split_bf16(feat)
        if not (e_hi.is_contiguous() and e_lo.is_contiguous() and f_hi.is_contiguous() and f_lo.is_contiguous()):
            raise _l.PvsgError('mask_logits: contiguous operand planes required')
        # all frames of the batch in one launch (3-D tensor maps)
        _l.check(lib.pvsg_linear_tc_batched(_ptr(e_hi), _ptr(e_lo), C, Q * C, _ptr(f_hi), _ptr(f_lo), C, P * C,
                                            _ptr(logits), _ptr(mask), _ptr(row_open), P, B, Q, P, C, _stream()),
                 'pvsg_linear_tc_batched')
        return logits, mask, row_open
    logits = torch.empty(B, Q, P, device=embed.device, dtype=torch.float32) if want_logits else None
    mask = torch.empty(B, Q, P, device=embed.device, dtype=torch.uint8) if want_mask else None
    row_open = torch.empty(B, Q, device=embed.device, dtype=torch.int32) if want_mask else None
    _l.check(lib.pvsg_mask_logits(_ptr(embed.contiguous()), _ptr(feat.contiguous()), _ptr(logits), _ptr(mask),
                                  _ptr(row_open), B, Q, P, C, _stream()), 'pvsg_mask_logits')
    return logits, mask, row_open


def panoptic_fuse(cls_logits, mask_logits_lr, in_hw, img_hw, out_hw, num_things, num_classes,
                  object_mask_thr=0.8, iou_thr=0.8, filter_low_score=True, instance_offset=1000):
    """cls_logits [Q,NC+1], mask_logits_lr [Q,h,w] -> (pan int32 [out_h,out_w], seg_info int32 [1+4Q])."""
    lib = _l.load()
    Q = _f32(cls_logits).shape[0]
    _, h, w = _f32(mask_logits_lr).shape
    dev = cls_logits.device
    pan = torch.empty(out_hw[0], out_hw[1], device=dev, dtype=torch.int32)
    seg_info = torch.empty(1 + 4 * Q, device=dev, dtype=torch.int32)
    work = torch.empty(4 * Q, device=dev, dtype=torch.int32)
    scores = torch.empty(Q, device=dev, dtype=torch.float32)
    pix = torch.empty(out_hw[0] * out_hw[1], device=dev, dtype=torch.int16)
    _l.check(lib.pvsg_panoptic_fuse(_ptr(cls_logits.contiguous()), _ptr(mask_logits_lr.contiguous()), Q,
                                    num_classes, num_things, h, w, in_hw[0], in_hw[1], img_hw[0], img_hw[1],
                                    out_hw[0], out_hw[1], object_mask_thr, iou_thr, 1 if filter_low_score else 0,
                                    instance_offset, _ptr(pan), _ptr(seg_info), _ptr(work), _ptr(scores),
                                    _ptr(pix), _stream()), 'pvsg_panoptic_fuse')
    return pan, seg_info


def instance_masks(mask_logits_lr, query_idx, in_hw, img_hw, out_hw, want_masks=True):
    lib = _l.load()
    _, h, w = _f32(mask_logits_lr).shape
    n = query_idx.numel()
    dev = mask_logits_lr.device
    stats = torch.empty(n, 2, device=dev, dtype=torch.float32)
    boxes = torch.empty(n, 4, device=dev, dtype=torch.int32)
    masks = torch.empty(n, out_hw[0], out_hw[1], device=dev, dtype=torch.uint8) if want_masks else None
    if n == 0:
        return stats, boxes, masks
    _l.check(lib.pvsg_instance_masks(_ptr(mask_logits_lr.contiguous()), _ptr(query_idx.to(torch.int32).contiguous()),
                                     n, h, w, in_hw[0], in_hw[1], img_hw[0], img_hw[1], out_hw[0], out_hw[1],
                                     _ptr(stats), _ptr(boxes), _ptr(masks), _stream()), 'pvsg_instance_masks')
    return stats, boxes, masks


def instance_select(cls_logits, k):
    """softmax(cls_logits)[:, :-1].flatten().topk(k, sorted=False) -> (scores [k], labels int32 [k],
    query index int32 [k]), in ascending flat-index order."""
    lib = _l.load()
    Q, C1 = _f32(cls_logits).shape
    dev = cls_logits.device
    scores = torch.empty(k, device=dev, dtype=torch.float32)
    labels = torch.empty(k, device=dev, dtype=torch.int32)
    query = torch.empty(k, device=dev, dtype=torch.int32)
    _l.check(lib.pvsg_instance_select(_ptr(cls_logits.contiguous()), Q, C1 - 1, k, _ptr(scores), _ptr(labels),
                                      _ptr(query), _stream()), 'pvsg_instance_select')
    return scores, labels, query


def instance_finalize(scores, labels, query, stats, boxes, num_things, topk):
    """Top-`topk` instances by class score x mask score -> (boxes6 [topk,6], labels int32 [topk],
    query int32 [topk], count int32 [1]); see pvsg_instance_finalize."""
    lib = _l.load()
    n = scores.numel()
    dev = scores.device
    topk = min(topk, n)
    boxes6 = torch.empty(topk, 6, device=dev, dtype=torch.float32)
    out_labels = torch.empty(topk, device=dev, dtype=torch.int32)
    sel_query = torch.empty(topk, device=dev, dtype=torch.int32)
    count = torch.empty(1, device=dev, dtype=torch.int32)
    if labels.dtype != torch.int32 or query.dtype != torch.int32 or boxes.dtype != torch.int32:
        raise _l.PvsgError('instance_finalize: int32 labels / query / boxes expected')
    _l.check(lib.pvsg_instance_finalize(_ptr(_f32(scores).contiguous()), _ptr(labels.contiguous()),
                                        _ptr(query.contiguous()), _ptr(_f32(stats).contiguous()),
                                        _ptr(boxes.contiguous()), n, num_things, topk, _ptr(boxes6), _ptr(out_labels),
                                        _ptr(sel_query), _ptr(count), _stream()), 'pvsg_instance_finalize')
    return boxes6, out_labels, sel_query, count


def postprocess_batched(cls_logits, mask_logits_lr, in_hw, img_hw, out_hw, num_things, num_classes, object_mask_thr,
                        iou_thr, filter_low_score, instance_offset, instance_on, max_per_image, topk):
    """Fused panoptic + instance post-processing of a BATCH of frames (static shapes, graph
    capturable): cls_logits [B,Q,NC+1], mask_logits_lr [B,Q,h,w] -> dict of [B, ...] tensors
    (pan int32 [B,H,W], seg_info int32 [B,1+4Q], and with instance_on: ins_boxes [B,topk,6],
    ins_labels [B,topk], ins_count [B,1], ins_masks uint8 [B,topk,H,W])."""
    lib = _l.load()
    B, Q, C1 = _f32(cls_logits).shape
    _, _, h, w = _f32(mask_logits_lr).shape
    dev = cls_logits.device
    cls_logits, mask_logits_lr = cls_logits.contiguous(), mask_logits_lr.contiguous()
    H, W = out_hw
    out = {}
    pan = torch.empty(B, H, W, device=dev, dtype=torch.int32)
    seg_info = torch.empty(B, 1 + 4 * Q, device=dev, dtype=torch.int32)
    work = torch.empty(B, 4 * Q, device=dev, dtype=torch.int32)
    scores = torch.empty(B, Q, device=dev, dtype=torch.float32)
    pix = torch.empty(B, H * W, device=dev, dtype=torch.int16)
    _l.check(lib.pvsg_panoptic_fuse_batched(_ptr(cls_logits), _ptr(mask_logits_lr), B, Q, num_classes, num_things, h, w,
                                            in_hw[0], in_hw[1], img_hw[0], img_hw[1], H, W, object_mask_thr, iou_thr,
                                            1 if filter_low_score else 0, instance_offset, _ptr(pan), _ptr(seg_info),
                                            _ptr(work), _ptr(scores), _ptr(pix), _stream()),
             'pvsg_panoptic_fuse_batched')
    out['pan'], out['seg_info'] = pan, seg_info
    if instance_on:
        n = min(max_per_image, Q * num_classes)
        topk = min(topk, n)
        ts = torch.empty(B, n, device=dev, dtype=torch.float32)
        tl = torch.empty(B, n, device=dev, dtype=torch.int32)
        tq = torch.empty(B, n, device=dev, dtype=torch.int32)
        _l.check(lib.pvsg_instance_select_batched(_ptr(cls_logits), B, Q, C1 - 1, n, _ptr(ts), _ptr(tl), _ptr(tq),
                                                  _stream()), 'pvsg_instance_select_batched')
        stats = torch.empty(B, n, 2, device=dev, dtype=torch.float32)
        boxes = torch.empty(B, n, 4, device=dev, dtype=torch.int32)
        geom = (h, w, in_hw[0], in_hw[1], img_hw[0], img_hw[1], H, W)
        _l.check(lib.pvsg_instance_masks_batched(_ptr(mask_logits_lr), _ptr(tq), B, Q, n, *geom, _ptr(stats), _ptr(boxes),
                                                 None, _stream()), 'pvsg_instance_masks_batched')
        boxes6 = torch.empty(B, topk, 6, device=dev, dtype=torch.float32)
        labels = torch.empty(B, topk, device=dev, dtype=torch.int32)
        sel = torch.empty(B, topk, device=dev, dtype=torch.int32)
        count = torch.empty(B, 1, device=dev, dtype=torch.int32)
        _l.check(lib.pvsg_instance_finalize_batched(_ptr(ts), _ptr(tl), _ptr(tq), _ptr(stats), _ptr(boxes), B, n,
                                                    num_things, topk, _ptr(boxes6), _ptr(labels), _ptr(sel),
                                                    _ptr(count), _stream()), 'pvsg_instance_finalize_batched')
        stats2 = torch.empty(B, topk, 2, device=dev, dtype=torch.float32)
        boxes2 = torch.empty(B, topk, 4, device=dev, dtype=torch.int32)
        masks = torch.empty(B, topk, H, W, device=dev, dtype=torch.uint8)
        _l.check(lib.pvsg_instance_masks_batched(_ptr(mask_logits_lr), _ptr(sel), B, Q, topk, *geom, _ptr(stats2),
                                                 _ptr(boxes2), _ptr(masks), _stream()), 'pvsg_instance_masks_batched')
        out.update(ins_boxes=boxes6, ins_labels=labels, ins_count=count, ins_masks=masks)
    return out


def rle_events(pan, seg_info, cap=1 << 17):
    """Run boundaries of every kept segment of pan int32 [B,H,W] (seg_info [B,1+4Q] from the panoptic
    post-process), column-major like pycocotools: (ev_pos uint32-as-int32 [B,cap], ev_slot int16 [B,cap],
    n_events int32 [B]).  See ``tubes.rle_from_events`` for the host side."""
    lib = _l.load()
    if pan.dtype != torch.int32 or seg_info.dtype != torch.int32 or not pan.is_cuda:
        raise _l.PvsgError('rle_events: CUDA int32 pan / seg_info expected')
    B, H, W = pan.shape
    Q = (seg_info.shape[1] - 1) // 4
    dev = pan.device
    ws = torch.empty(2, B, W, device=dev, dtype=torch.int32)
    ev_pos = torch.empty(B, cap, device=dev, dtype=torch.int32)
    ev_slot = torch.empty(B, cap, device=dev, dtype=torch.int16)
    n_events = torch.empty(B, device=dev, dtype=torch.int32)
    _l.check(lib.pvsg_rle_events(_ptr(pan.contiguous()), _ptr(seg_info.contiguous()), B, Q, H, W, _ptr(ws), _ptr(ev_pos),
                                 _ptr(ev_slot), _ptr(n_events), cap, _stream()), 'pvsg_rle_events')
    return ev_pos, ev_slot, n_events


def window_attention(qkv, qkv_bias, bias_table, num_heads, window, shift, out_mode='f32'):
    """Swin (shifted-)window attention between the qkv and proj linears (mmdet swin.py ShiftWindowMSA / WindowMSA):
    qkv [B,H,W,3C] of the unpadded map -> [B,H,W,C]; padding, roll, partition, bias, mask, reverse, crop inside.
    out_mode 'split': the result leaves as ``Split`` operand planes only (tcgen05 engine), 'both': (fp32, Split)."""
    lib = _l.load()
    _f32(qkv, 'qkv')
    if qkv.dim() != 4 or not qkv.is_contiguous():
        raise _l.PvsgError('window_attention: contiguous [B,H,W,3C] expected')
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if _f32(qkv_bias).numel() != C3 or tuple(_f32(bias_table).shape) != ((2 * window - 1) ** 2, num_heads):
        raise _l.PvsgError('window_attention: bad qkv_bias / relative_position_bias_table shape')
    want_split = out_mode in ('split', 'both') and ENGINE[0] == 'tc' and B * H * W > SKINNY_M
    want_f32 = out_mode in ('f32', 'both') or not want_split
    out = torch.empty(B, H, W, C, device=qkv.device, dtype=torch.float32) if want_f32 else None
    hi = torch.empty(B, H, W, C, device=qkv.device, dtype=torch.bfloat16) if want_split else None
    lo = torch.empty(B, H, W, C, device=qkv.device, dtype=torch.bfloat16) if want_split else None
    _l.check(lib.pvsg_window_attention(_ptr(qkv), _ptr(qkv_bias.contiguous()), _ptr(bias_table.contiguous()), _ptr(out),
                                       _ptr(hi), _ptr(lo), B, H, W, C, num_heads, window, shift, _stream()),
             'pvsg_window_attention')
    sp = Split(hi, lo) if want_split else None
    if out_mode == 'both':
        return out, sp
    return sp if (out_mode == 'split' and sp is not None) else out


def patch_merge_ln(x, gamma, beta, eps=1e-5):
    """mmdet PatchMerging up to the reduction linear: x [B,H,W,C] -> LayerNorm(unfold 2x2) [B,ceil(H/2),ceil(W/2),4C]."""
    lib = _l.load()
    _f32(x, 'x')
    if x.dim() != 4 or not x.is_contiguous():
        raise _l.PvsgError('patch_merge_ln: contiguous [B,H,W,C] expected')
    B, H, W, C = x.shape
    if _f32(gamma).numel() != 4 * C or _f32(beta).numel() != 4 * C:
        raise _l.PvsgError('patch_merge_ln: norm over 4C expected')
    y = torch.empty(B, (H + 1) // 2, (W + 1) // 2, 4 * C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_patch_merge_ln(_ptr(x), _ptr(gamma.contiguous()), _ptr(beta.contiguous()), _ptr(y), B, H, W, C, eps,
                                     _stream()), 'pvsg_patch_merge_ln')
    return y


def tube_overlap(gt, pan, seg_info, num_gt):
    """counts int32 [B, num_gt + 1, Q + 1]: joint histogram of GT object ids (gt int32 [B,H,W]) and the kept
    segments of pan int32 [B,H,W] (slot order of ``tubes.slot_ids``); row num_gt = other GT ids, column Q = pixels
    of no kept segment.  One pass over both maps (``relation_set.match_clip`` derives every IoU from it)."""
    lib = _l.load()
    if gt.dtype != torch.int32 or pan.dtype != torch.int32 or seg_info.dtype != torch.int32 or not pan.is_cuda \
            or not gt.is_cuda or gt.shape != pan.shape:
        raise _l.PvsgError('tube_overlap: CUDA int32 gt / pan of one shape and int32 seg_info expected')
    B, H, W = pan.shape
    Q = (seg_info.shape[1] - 1) // 4
    counts = torch.empty(B, num_gt + 1, Q + 1, device=pan.device, dtype=torch.int32)
    _l.check(lib.pvsg_tube_overlap(_ptr(gt.contiguous()), _ptr(pan.contiguous()), _ptr(seg_info.contiguous()), B, Q, H, W,
                                   int(num_gt), _ptr(counts), _stream()), 'pvsg_tube_overlap')
    return counts


def reconsdot(trk, det, tmp=100.0):
    """trk [ntrk,nst,d], det [ndet,nsd,d] zero-padded position-major embeddings -> cost [ntrk,ndet] (pvsg_reconsdot)."""
    lib = _l.load()
    ntrk, nst, d = _f32(trk, 'trk').shape
    ndet, nsd, d2 = _f32(det, 'det').shape
    if d2 != d:
        raise _l.PvsgError('reconsdot: feature widths differ')
    cost = torch.empty(ntrk, ndet, device=trk.device, dtype=torch.float32)
    ws = torch.empty(lib.pvsg_reconsdot_workspace_bytes(ntrk, nst, ndet, nsd, d), device=trk.device, dtype=torch.uint8)
    _l.check(lib.pvsg_reconsdot(_ptr(trk.contiguous()), _ptr(det.contiguous()), _ptr(cost), _ptr(ws), ntrk, nst, ndet, nsd, d,
                                float(tmp), _stream()), 'pvsg_reconsdot')
    return cost


def lap_assign(cost, cost_limit):
    """cost [n,m] fp32 (+inf = forbidden) -> (x int32 [n], y int32 [m]) of lap.lapjv(cost, extend_cost=True, cost_limit)."""
    lib = _l.load()
    n, m = _f32(cost, 'cost').shape
    x = torch.empty(n, device=cost.device, dtype=torch.int32)
    y = torch.empty(m, device=cost.device, dtype=torch.int32)
    _l.check(lib.pvsg_lap_assign(_ptr(cost.contiguous()), n, m, float(cost_limit), _ptr(x), _ptr(y), _stream()), 'pvsg_lap_assign')
    return x, y


def minvis_chain(embeds):
    """embeds [T,Q,C] raw per-frame query embeddings -> sigma int32 [T-1,Q]: sigma[t][i] = query of frame t+1 matched to
    query i of frame t (cosine cost, exact assignment; all T-1 problems solved concurrently)."""
    lib = _l.load()
    T, Q, C = _f32(embeds, 'embeds').shape
    cost = torch.empty(T - 1, Q, Q, device=embeds.device, dtype=torch.float32)
    _l.check(lib.pvsg_cosine_chain_cost(_ptr(embeds.contiguous()), _ptr(cost), T, Q, C, _stream()), 'pvsg_cosine_chain_cost')
    x = torch.empty(T - 1, Q, device=embeds.device, dtype=torch.int32)
    y = torch.empty(T - 1, Q, device=embeds.device, dtype=torch.int32)
    _l.check(lib.pvsg_lap_square_batched(_ptr(cost), T - 1, Q, _ptr(x), _ptr(y), _stream()), 'pvsg_lap_square_batched')
    return x


def perm_chain(sigma, Q):
    """sigma int32 [T-1,Q] (may be empty) -> perms int32 [T,Q]: perms[0] = identity, perms[t] = sigma[t-1][perms[t-1]]."""
    lib = _l.load()
    T = sigma.shape[0] + 1
    perms = torch.empty(T, Q, device=sigma.device, dtype=torch.int32)
    _l.check(lib.pvsg_perm_chain(_ptr(sigma.contiguous()) if T > 1 else None, _ptr(perms), T, Q, _stream()), 'pvsg_perm_chain')
    return perms


def max_over_time(x):
    lib = _l.load()
    N, T, C = _f32(x).shape
    out = torch.empty(N, C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_max_over_time(_ptr(x.contiguous()), _ptr(out), N, T, C, _stream()), 'pvsg_max_over_time')
    return out


def temporal_fir(x, w):
    """x [P,T,C], w [K] (odd K) -> zero-padded cross-correlation along T (depthwise F.conv1d with one shared filter)."""
    lib = _l.load()
    Pn, T, C = _f32(x).shape
    y = torch.empty_like(x, memory_format=torch.contiguous_format)
    _l.check(lib.pvsg_temporal_fir(_ptr(x.contiguous()), _ptr(_f32(w).contiguous()), _ptr(y), Pn, T, C, w.numel(), _stream()),
             'pvsg_temporal_fir')
    return y


def temporal_unfold(x, k):
    """x [P,T,C] -> [P,T,k*C]: the k zero-padded temporal taps of every frame side by side (tap-major)."""
    lib = _l.load()
    Pn, T, C = _f32(x).shape
    y = torch.empty(Pn, T, k * C, device=x.device, dtype=torch.float32)
    _l.check(lib.pvsg_temporal_unfold(_ptr(x.contiguous()), _ptr(y), Pn, T, C, k, _stream()), 'pvsg_temporal_unfold')
    return y


def pair_proposal(U, V, w2, b2):
    lib = _l.load()
    N, Hd = _f32(U).shape
    out = torch.empty(N, N, device=U.device, dtype=torch.float32)
    _l.check(lib.pvsg_pair_proposal(_ptr(U.contiguous()), _ptr(_f32(V).contiguous()), _ptr(_f32(w2).contiguous()),
                                    _ptr(_f32(b2)), _ptr(out), N, Hd, _stream()), 'pvsg_pair_proposal')
    return out


def top_pairs(pair, k):
    lib = _l.load()
    N = _f32(pair).shape[0]
    pairs = torch.zeros(max(1, min(k, N * N)), 2, device=pair.device, dtype=torch.int32)
    n_out = torch.zeros(1, device=pair.device, dtype=torch.int32)
    _l.check(lib.pvsg_top_pairs(_ptr(pair.contiguous()), N, k, _ptr(pairs), _ptr(n_out), _stream()),
             'pvsg_top_pairs')
    return pairs, n_out


def gather_pairs(sub, obj, pairs, pe=None):
    """sub/obj [N,T,F], pairs int32 [P,2] -> [P,T,2F] (+ pe[:T])."""
    lib = _l.load()
    N, T, Fd = _f32(sub).shape
    Pn = pairs.shape[0]
    if tuple(_f32(obj).shape) != (N, T, Fd) or pairs.dtype != torch.int32 or pairs.dim() != 2 or pairs.shape[1] != 2:
        raise _l.PvsgError('gather_pairs: sub / obj [N,T,F] of one shape and int32 pairs [P,2] expected')
    if pe is not None and (_f32(pe).dim() != 2 or pe.shape[0] < T or pe.shape[1] != 2 * Fd or not pe.is_contiguous()):
        # the reference's PositionalEncoding fails the same way for clips longer than max_len (transformer.py:59-81)
        raise _l.PvsgError(f'gather_pairs: positional table {tuple(pe.shape)} does not cover T={T}, 2F={2 * Fd}')
    out = torch.empty(Pn, T, 2 * Fd, device=sub.device, dtype=torch.float32)
    if Pn == 0:
        return out
    _l.check(lib.pvsg_gather_pairs(_ptr(sub.contiguous()), _ptr(_f32(obj).contiguous()), _ptr(pairs.contiguous()),
                                   _ptr(_f32(pe)), _ptr(out), Pn, T, Fd, _stream()), 'pvsg_gather_pairs')
    return out
