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
        d_res = dz if has_res and ctx.needs_input_grad[4] else None      # the residual bypasses the activation only
        if ctx.act == ops.ACT_RELU:                                      # when act is none (the head never mixes them)
            if has_res:
                raise ops._l.PvsgError('linear backward: residual with a fused ReLU is not used by the head')
            dz = ops.relu_backward(dz, y)
        dz2 = dz.reshape(-1, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0] or (has_add and ctx.needs_input_grad[3]):
            dx = ops.linear(dz2, _t(weight)).reshape(x.shape)           # dy W
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, K)
            a2 = _t(add_input.reshape(-1, K)) if has_add else None
            dw = _t(ops.linear(_t(x2), _t(dz2), add_input=a2))           # (x + add)^T dy, transposed back -> [N, K]
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
        if ctx.needs_input_grad[0]:
            d_embed = torch.stack([ops.linear(dl[b], _t(feat[b])) for b in range(embed.shape[0])])        # dL F
        if ctx.needs_input_grad[1]:
            d_feat = torch.stack([ops.linear(_t(dl[b]), _t(embed[b])) for b in range(embed.shape[0])])    # dL^T E
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
