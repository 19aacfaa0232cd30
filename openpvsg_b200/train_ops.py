"""Differentiable building blocks of the decoder head's training pass (second slice).

``torch.autograd.Function``s whose forward AND backward are libpvsg_sm100.so kernels: dense layers (``pvsg_linear`` three
times: y = act(x W^T + b), dx = dy W, dW^T = x^T dy), LayerNorm, the decoder's masked multi-head attention, the
query x pixel mask contraction, and the two broadcast adds of the head (``level_embed``, the query embeddings).  autograd
only does the bookkeeping: the tape, slicing ``in_proj_weight`` and accumulating gradients that meet at a tensor.

Reference: the modules ``Mask2FormerVideoHead.forward`` calls in training mode
(``models/mask2former_vps/mask2former_video_head.py:361-462``): mmcv ``MultiheadAttention`` / ``FFN`` / ``nn.LayerNorm``
inside ``DetrTransformerDecoderLayer`` (config ``mask2former_video_r50_base.py:63-88``), ``forward_head_video`` (:337-359).
"""
import torch

from . import ops


def _t(x):
    """Transposed contiguous copy of a 2-D tensor (data movement for the backward GEMMs)."""
    return x.t().contiguous()


def _grad_weight(dz2, sources):
    """dW [N, sum C_i] = dz2^T [N,T] . concat_i(src_i) [T, C_i]: both operands are transposed into token-minor bf16 planes
    in one pass each (ops.transpose_split), the GEMM runs split-K over token chunks on the tcgen05 engine
    (ops.splitk_gemm).  sources: list of (view [T,C] or [B,OH,OW,C], add or None)."""
    N = dz2.shape[1]
    T = dz2.shape[0]
    widths = [src.shape[-1] for src, _ in sources]
    S, Kc, Tp = ops.splitk_plan(T, N, sum(widths))
    dzp = ops.transpose_split(dz2, Tp)
    if len(sources) == 1:
        wp = ops.transpose_split(sources[0][0], Tp, add=sources[0][1])
    else:
        hi, lo = ops.alloc_planes(sum(widths), T, Tp, dz2.device)
        row = 0
        for (src, add), c in zip(sources, widths):
            ops.transpose_split(src, Tp, add=add, hi=hi, lo=lo, row0=row)
            row += c
        wp = (hi, lo)
    return ops.splitk_gemm(dzp, wp, S, Kc)


class _Linear(torch.autograd.Function):
    """y = act((x + add_input) W^T + bias + residual); act in {none, relu}."""

    @staticmethod
    def forward(ctx, x, weight, bias, add_input, residual, act):
        y = ops.linear(x, weight, bias, add_input=add_input, residual=residual, act=act)
        ctx.act = act
        ctx.has = (bias is not None, add_input is not None, residual is not None)
        ctx.save_for_backward(x, weight, add_input, y if act == ops.ACT_RELU else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, add_input, y = ctx.saved_tensors
        has_bias, has_add, has_res = ctx.has
        N, K = weight.shape
        dz = dy.contiguous()
        if ctx.act == ops.ACT_RELU:                                      # y = relu(z + residual): the mask applies to both
            dz = ops.relu_backward(dz, y)
        d_res = dz if has_res and ctx.needs_input_grad[4] else None
        dz2 = dz.reshape(-1, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0] or (has_add and ctx.needs_input_grad[3]):
            dx = ops.linear(dz2, _t(weight)).reshape(x.shape)           # dy W
        if ctx.needs_input_grad[1]:                                      # dW = dz^T (x + add): reduction over the tokens
            x2 = x.reshape(-1, K).contiguous()
            a2 = add_input.reshape(-1, K).contiguous() if has_add else None
            dw = _grad_weight(dz2, [(x2, a2)])
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dz2)
        return (dx if ctx.needs_input_grad[0] else None, dw, db, dx if has_add and ctx.needs_input_grad[3] else None, d_res, None)


def linear(x, weight, bias=None, add_input=None, residual=None, act=ops.ACT_NONE):
    return _Linear.apply(x, weight, bias, add_input, residual, act)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        ctx.eps = eps
        ctx.save_for_backward(x, gamma)
        return ops.layernorm(x, gamma, beta, eps)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_backward(x, gamma, dy, ctx.eps)
        return dx, dg, db, None


def layernorm(x, norm):
    return _LayerNorm.apply(x, norm.weight, norm.bias, norm.eps)


class _Attention(torch.autograd.Function):
    """softmax(q k^T / sqrt(32) + mask) v per head; mask / row_open as produced by ``ops.mask_logits`` (no gradient)."""

    @staticmethod
    def forward(ctx, q, k, v, num_heads, mask, row_open):
        out, lse = ops.attention_train_forward(q, k, v, num_heads, mask, row_open)
        ctx.num_heads = num_heads
        ctx.save_for_backward(q, k, v, mask, row_open, out, lse)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, mask, row_open, out, lse = ctx.saved_tensors
        dq, dk, dv = ops.attention_train_backward(q, k, v, ctx.num_heads, mask, row_open, out, dout, lse)
        return dq, dk, dv, None, None, None


def attention(q, k, v, num_heads, mask=None, row_open=None):
    return _Attention.apply(q, k, v, num_heads, mask, row_open)


class _MaskLogits(torch.autograd.Function):
    """einsum('bqc,bpc->bqp') (mask2former_video_head.py:345, pixels token-major)."""

    @staticmethod
    def forward(ctx, embed, feat):
        ctx.save_for_backward(embed, feat)
        return ops.mask_logits(embed.contiguous(), feat.contiguous(), True, False)[0]

    @staticmethod
    def backward(ctx, dl):
        embed, feat = ctx.saved_tensors
        dl = dl.contiguous()
        d_embed = d_feat = None
        B, Q, C = embed.shape
        P = feat.shape[1]
        if ctx.needs_input_grad[0]:                      # dL F: reduction over the pixels, split-K
            S, Kc, Tp = ops.splitk_plan(P, Q, C)
            rows = []
            for b in range(B):
                a = ops.split_bf16(torch.nn.functional.pad(dl[b], (0, Tp - P)).contiguous())
                rows.append(ops.splitk_gemm(a, ops.transpose_split(feat[b], Tp), S, Kc))
            d_embed = torch.stack(rows)
        if ctx.needs_input_grad[1]:                      # dL^T E: reduction over the queries (padded to one k-block pair)
            Qp = (Q + 63) // 64 * 64
            d_feat = torch.stack([ops.splitk_gemm(ops.transpose_split(dl[b], Qp), ops.transpose_split(embed[b], Qp), 1, Qp)
                                  for b in range(B)])
        return d_embed, d_feat


def mask_logits(embed, feat):
    return _MaskLogits.apply(embed, feat)


class _AddRowvec(torch.autograd.Function):
    """x [.., C] + v [C] (``level_embed``, mask2former_video_head.py:397)."""

    @staticmethod
    def forward(ctx, x, v):
        return ops.add_rowvec(x.contiguous(), v.contiguous())

    @staticmethod
    def backward(ctx, dy):
        dv = ops.colsum(dy.reshape(-1, dy.shape[-1]).contiguous()) if ctx.needs_input_grad[1] else None
        return (dy if ctx.needs_input_grad[0] else None), dv


def add_rowvec(x, v):
    return _AddRowvec.apply(x, v)


class _ExpandBatch(torch.autograd.Function):
    """w [Q,C] -> [B,Q,C] (``query_feat.weight.unsqueeze(1).repeat``, :409-410); the gradient is the sum over the batch."""

    @staticmethod
    def forward(ctx, w, B):
        return w[None].expand(B, -1, -1).contiguous()

    @staticmethod
    def backward(ctx, dy):
        B = dy.shape[0]
        return ops.colsum(dy.contiguous().view(B, -1)).view(dy.shape[1:]), None


def expand_batch(w, B):
    return _ExpandBatch.apply(w, B)


# ---------------------------------------------------------------------------------------------- pixel decoder pieces
class _GroupNorm(torch.autograd.Function):
    """mmcv ConvModule norm (+ ReLU) on token-major maps [B,H,W,C]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, relu):
        x = x.contiguous()
        ctx.cfg = (groups, eps, relu)
        ctx.save_for_backward(x, gamma, beta)
        return ops.groupnorm_nhwc(x, gamma, beta, groups, eps, ops.ACT_RELU if relu else ops.ACT_NONE)

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta = ctx.saved_tensors
        groups, eps, relu = ctx.cfg
        dx, dg, db = ops.groupnorm_nhwc_backward(x, gamma, beta, dy, groups, eps, relu)
        return dx, dg, db, None, None, None


def groupnorm(x, gn, relu=False):
    return _GroupNorm.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, relu)


class _Conv(torch.autograd.Function):
    """act(conv(x, w) + bias + residual) on token-major maps; x [B,H,W,Cin], w [Cout,R,S,Cin] (the kernel layout), square
    filters, stride 1 or 2.  Backward through the forward engine:
      dX = stride-1 convolution of dZ (zero-inserted for stride 2) with the mirrored, transposed filter, pad R-1-pad;
      dW[:, r, s, :] = dZ^T X_shift(r, s): one GEMM per filter tap on a (strided) shifted copy of the padded input."""

    @staticmethod
    def forward(ctx, x, w, bias, residual, stride, pad, act):
        x, w = x.contiguous(), w.contiguous()
        y = ops.conv2d_nhwc(x, w, bias, residual=residual, stride=stride, pad=pad, act=act)
        ctx.cfg = (stride, pad, act, bias is not None, residual is not None)
        ctx.save_for_backward(x, w, y if act == ops.ACT_RELU else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        stride, pad, act, has_bias, has_res = ctx.cfg
        B, H, W, Cin = x.shape
        Cout, R, S, _ = w.shape
        dz = dy.contiguous()
        if act == ops.ACT_RELU:
            dz = ops.relu_backward(dz, y)
        OH, OW = dz.shape[1:3]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if stride == 1:
                z = dz
            else:                                        # zero insertion (data movement)
                z = dz.new_zeros(B, H - R + 1 + 2 * pad, W - S + 1 + 2 * pad, Cout)
                z[:, 0:stride * OH:stride, 0:stride * OW:stride] = dz
            if R == 1 and S == 1:
                dx = ops.linear(z.view(-1, Cout), _t(w.view(Cout, Cin))).view(B, H, W, Cin)
            else:
                dx = ops.conv2d_nhwc(z, w.flip(1, 2).permute(3, 1, 2, 0).contiguous(), None, pad=R - 1 - pad)
        if ctx.needs_input_grad[1]:                      # all filter taps in ONE token-reduction GEMM: N = R * S * Cin
            xp = torch.nn.functional.pad(x, (0, 0, pad, pad, pad, pad)) if pad else x
            taps = [(xp[:, r:r + stride * OH:stride, q:q + stride * OW:stride], None) for r in range(R) for q in range(S)]
            dw = _grad_weight(dz.view(-1, Cout), taps).view(Cout, R, S, Cin)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dz.view(-1, Cout))
        return dx, dw, db, (dz if has_res and ctx.needs_input_grad[3] else None), None, None, None


def conv(x, w, bias=None, residual=None, stride=1, pad=0, act=ops.ACT_NONE):
    return _Conv.apply(x, w, bias, residual, stride, pad, act)


def conv3x3(x, weight):
    """weight in the nn.Conv2d layout [Cout,Cin,3,3] (the permute is on the tape)."""
    return _Conv.apply(x, weight.permute(0, 2, 3, 1), None, None, 1, 1, ops.ACT_NONE)


class _MaxPool(torch.autograd.Function):
    """3x3 / stride 2 / pad 1 max pooling (ResNet stem); the gradient goes to the first maximum of each window in scan
    order, as ATen's max_pool2d backward."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool3x3s2_nhwc(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool3x3s2_nhwc_backward(x, dy)


def maxpool3x3s2(x):
    return _MaxPool.apply(x)


class _ResizeAdd(torch.autograd.Function):
    """base + bilinear_resize(src -> base's size) (the FPN top-down step; F.interpolate bilinear, align_corners=False)."""

    @staticmethod
    def forward(ctx, base, src):
        ctx.in_hw = tuple(src.shape[1:3])
        out = base.contiguous().clone()
        ops.bilinear_resize_nhwc(src.contiguous(), out.shape[1:3], out=out, accumulate=True)
        return out

    @staticmethod
    def backward(ctx, dy):
        dsrc = ops.bilinear_resize_nhwc_backward(dy, ctx.in_hw) if ctx.needs_input_grad[1] else None
        return (dy if ctx.needs_input_grad[0] else None), dsrc


def resize_add(base, src):
    return _ResizeAdd.apply(base, src)


class _MSDAFused(torch.autograd.Function):
    """MultiScaleDeformableAttention core from the raw projections (ops.msda_fused_forward); the backward expands the
    projections into explicit sampling locations / attention weights, runs pvsg_msda_backward and maps their gradients
    back (softmax over L*P, division by the level sizes)."""

    @staticmethod
    def forward(ctx, value, proj, ref, shapes, num_heads, num_points):
        value, proj = value.contiguous(), proj.contiguous()
        ctx.cfg = (shapes, num_heads, num_points)
        ctx.save_for_backward(value, proj, ref)
        return ops.msda_fused_forward(value, shapes, proj, ref, num_heads, num_points)

    @staticmethod
    def backward(ctx, dout):
        value, proj, ref = ctx.saved_tensors
        shapes, H, P = ctx.cfg
        B, N, C = value.shape
        loc, aw = ops.msda_proj_expand(proj, ref, shapes, H, P)
        gv, gl, ga = ops.msda_backward(value.view(B, N, H, C // H), shapes, loc, aw, dout.contiguous())
        dproj = ops.msda_proj_backward(aw, gl, ga, shapes)
        return gv.view(B, N, C), dproj, None, None, None, None


def msda_fused(value, proj, ref, shapes, num_heads, num_points):
    return _MSDAFused.apply(value, proj, ref, tuple(shapes), num_heads, num_points)
