"""B200 backend of the OpenPVSG Mask2Former / Mask2Former-VPS inference path.

Same registry names, constructor arguments, ``forward`` / ``simple_test`` signatures and
``state_dict`` key layout as the reference (and the mmdet 2.25 / mmcv 1.4 modules its
configs name), so a config ``type=`` string or an existing checkpoint resolves to these
classes unchanged -- but every tensor operation is a call into libpvsg_sm100.so
(openpvsg_b200/ops.py).  Inference only: training methods raise NotImplementedError.

Internal layout is token-major (NHWC).  Tensors crossing the reference's interfaces keep
their reference SHAPES ([B,C,H,W] feature maps, [B,Q,h,w] mask logits); feature maps are
handed around as channels_last views so no transposition is needed between modules.

Reference files: models/mask2former/{mask2former,mask2former_head,mask2former_fusion_head}.py,
models/mask2former_vps/{mask2former,mask2former_video_head,position_encoding}.py; L0 modules
per SURVEY.md appendix A.
"""
import copy
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .registry import (ATTENTION, BACKBONES, DETECTORS, FEEDFORWARD_NETWORK, HEADS, PLUGIN_LAYERS,
                       POSITIONAL_ENCODING, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, ConfigDict,
                       build_backbone, build_head, build_plugin_layer, build_positional_encoding,
                       build_transformer_layer_sequence, to_cfg)

INSTANCE_OFFSET = 1000


def _tokens(x):
    """[B,C,H,W] tensor (any memory format) -> token-major [B,H,W,C] contiguous."""
    if x.dim() != 4:
        raise ValueError('expected a [B,C,H,W] tensor')
    t = x.permute(0, 2, 3, 1)
    if t.is_contiguous():
        return t
    if t.stride(3) == 1:            # token-major data behind a batch-strided view: one plain copy
        return t.contiguous()
    return ops.nchw_to_nhwc(x.contiguous())


def _as_nchw(t):
    """token-major [B,H,W,C] -> logical [B,C,H,W] (channels_last view, no copy)."""
    return t.permute(0, 3, 1, 2)


class _Prepared(nn.Module):
    """Modules cache kernel-layout copies of their parameters; loading a checkpoint or
    moving the module invalidates the cache."""

    def __init__(self):
        super().__init__()
        self._prep = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._prep = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **k):
        self._prep = None
        return super()._apply(fn, *a, **k)

    def _stale(self):
        """True when a parameter changed in place since the kernel-layout copies were made (an optimizer step bumps the
        version counters; load_state_dict / .to() are caught by the hooks above)."""
        sig = tuple(p._version for p in self.parameters())
        if sig != self.__dict__.get('_prep_sig'):
            self.__dict__['_prep_sig'] = sig
            return True
        return False


# ======================================================================================
# backbone: mmdet ResNet (L0, A1)
# ======================================================================================
class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.stride = stride
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))
        else:
            self.downsample = None


def _fold(conv, bn):
    """conv (no bias) + eval-mode BN -> ([Cout,R,S,Cin] weight, bias), fp32."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * scale[:, None, None, None]).permute(0, 2, 3, 1).contiguous()
    b = (bn.bias - bn.running_mean * scale).contiguous()
    return w, b


@BACKBONES.register_module()
class ResNet(_Prepared):
    """mmdet ``ResNet`` (depth 50/101, style='pytorch', norm_eval) -- torchvision key names.
    cfg: configs/mask2former_vps/mask2former_video_r50_base.py:7-16."""
    arch = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}

    def __init__(self, depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=-1, norm_cfg=None,
                 norm_eval=True, style='pytorch', init_cfg=None, **kwargs):
        super().__init__()
        if depth not in self.arch or style != 'pytorch' or num_stages != 4:
            raise NotImplementedError('ResNet: only depth 50/101, style="pytorch", 4 stages')
        self.out_indices = tuple(out_indices)
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inpl = 64
        for i, (n, planes) in enumerate(zip(self.arch[depth], (64, 128, 256, 512))):
            blocks = []
            for b in range(n):
                stride = 2 if (b == 0 and i > 0) else 1
                blocks.append(_Bottleneck(inpl, planes, stride, b == 0))
                inpl = planes * 4
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
        if norm_cfg is not None and dict(norm_cfg).get('requires_grad', True) is False:     # base cfg :12 (BN frozen)
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad_(False)
        self.eval()

    def init_weights(self):
        pass

    @torch.no_grad()
    def _prepare(self):
        p = {'stem': _fold(self.conv1, self.bn1), 'blocks': []}
        for i in range(4):
            for blk in getattr(self, f'layer{i + 1}'):
                d = dict(c1=_fold(blk.conv1, blk.bn1), c2=_fold(blk.conv2, blk.bn2), c3=_fold(blk.conv3, blk.bn3),
                         stride=blk.stride, stage=i)
                d['ds'] = _fold(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                p['blocks'].append(d)
        self._prep = p

    def forward_train(self, x):
        """``forward`` on the autograd tape (train_ops): eval-mode BatchNorm folded into every convolution as in
        inference (norm_eval=True: running statistics, base cfg :7-16); the convolution weights and -- when
        norm_cfg.requires_grad, as in the VPS config -- the BN affine parameters receive gradients through the fold; fp32
        maps instead of operand planes."""
        from . import train_ops as T

        def fold(conv, bn):       # weight prep on the tape: the BN affine parameters train when norm_cfg.requires_grad (VPS cfg :13)
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            return (conv.weight * scale[:, None, None, None]).permute(0, 2, 3, 1), shift

        x = T.conv(_tokens(x), *fold(self.conv1, self.bn1), stride=2, pad=3, act=ops.ACT_RELU)
        x = T.maxpool3x3s2(x)
        outs = []
        for i in range(4):
            for blk in getattr(self, f'layer{i + 1}'):
                o = T.conv(x, *fold(blk.conv1, blk.bn1), act=ops.ACT_RELU)
                o = T.conv(o, *fold(blk.conv2, blk.bn2), stride=blk.stride, pad=1, act=ops.ACT_RELU)
                if blk.downsample is None:
                    idt = x
                else:
                    sub = x[:, ::blk.stride, ::blk.stride].contiguous() if blk.stride > 1 else x
                    idt = T.conv(sub, *fold(blk.downsample[0], blk.downsample[1]))
                x = T.conv(o, *fold(blk.conv3, blk.bn3), residual=idt, act=ops.ACT_RELU)
            if i in self.out_indices:
                outs.append(_as_nchw(x))
        return tuple(outs)

    @torch.no_grad()
    def forward(self, x):
        if self._prep is None or self._stale():
            self._prepare()
        p = self._prep
        ops.clear_split_cache()
        if ops.ENGINE[0] == 'tc' and x.is_contiguous():
            if 'stem2' not in p:
                p['stem2'] = ops.stem_weight(p['stem'][0])
            x = ops.stem7x7s2(x, p['stem2'], p['stem'][1])
        else:
            x = ops.conv2d_nhwc(_tokens(x), *p['stem'], stride=2, pad=3, act=ops.ACT_RELU)
        x = ops.maxpool3x3s2_nhwc(x)
        # On the tcgen05 engine activations live as operand planes only: the 1x1 -> 3x3 -> 1x1 chain
        # hands planes from epilogue to TMA loader (out_mode='split'), and the identity branch is
        # added from its planes too (r = hi + lo), so no fp32 copy of a block output is written or
        # read.  Only the stage outputs that leave the backbone are also materialised as fp32.
        xs = ops.maybe_split(x)
        outs = []
        nblk = len(p['blocks'])
        for j, d in enumerate(p['blocks']):
            inp = xs if xs is not None else x
            o = ops.conv2d_nhwc(inp, *d['c1'], act=ops.ACT_RELU, out_mode='split')
            o = ops.conv2d_nhwc(o, *d['c2'], stride=d['stride'], pad=1, act=ops.ACT_RELU, out_mode='split')
            idt = inp if d['ds'] is None else ops.conv2d_nhwc(inp, *d['ds'], stride=d['stride'], out_mode='split')
            leaves = (j + 1 == nblk or p['blocks'][j + 1]['stage'] != d['stage']) and d['stage'] in self.out_indices
            if leaves or xs is None:
                x, xs = ops.conv2d_nhwc(o, *d['c3'], residual=idt, act=ops.ACT_RELU, out_mode='both')
            else:
                x, xs = None, ops.conv2d_nhwc(o, *d['c3'], residual=idt, act=ops.ACT_RELU, out_mode='split')
            if leaves:
                ops.remember_split(x, xs)   # the pixel decoder's 1x1 convs reuse these planes
                outs.append(_as_nchw(x))
        return tuple(outs)


# ======================================================================================
# positional encodings
# ======================================================================================
@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(nn.Module):
    """mmdet SinePositionalEncoding (L0, A4).  ``forward(mask)`` takes the all-valid bool mask
    [B,h,w] the reference passes (mask2former_head.py:428-432) and returns [B,2F,h,w]."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * 3.141592653589793, eps=1e-6,
                 offset=0., init_cfg=None):
        super().__init__()
        if not normalize or offset != 0.:
            raise NotImplementedError('SinePositionalEncoding: normalize=True, offset=0 only')
        self.num_feats, self.temperature, self.scale, self.eps = num_feats, temperature, scale, eps

    def tokens(self, h, w, device, t=0, add_vec=None):
        return ops.sine_pe(h, w, device, t=t, num_feats=self.num_feats, temperature=self.temperature,
                           scale=self.scale, eps=self.eps, add_vec=add_vec)

    def forward(self, mask):
        b, h, w = mask.shape
        pe = self.tokens(h, w, mask.device).view(1, h, w, -1)
        return _as_nchw(pe).expand(b, -1, -1, -1)


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding3D(SinePositionalEncoding):
    """models/mask2former_vps/position_encoding.py:9-109; mask [B,T,h,w] -> [B,T,2F,h,w]."""

    def forward(self, mask):
        assert mask.dim() == 4
        b, t, h, w = mask.shape
        pe = self.tokens(h, w, mask.device, t=t).view(1, t, h, w, -1)
        return pe.permute(0, 1, 4, 2, 3).expand(b, -1, -1, -1, -1)


# ======================================================================================
# pixel decoder (L0, A2/A3)
# ======================================================================================
@ATTENTION.register_module()
class MultiScaleDeformableAttention(_Prepared):
    """mmcv MultiScaleDeformableAttention (cfg mask2former_video_r50_base.py:38-47).

    ``forward`` keeps mmcv's signature on seq-first tensors; the pixel decoder calls
    ``forward_tokens`` (batch-first token-major, fused projections + fused sampling kernel).
    """

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__()
        if embed_dims // num_heads != 32:
            raise NotImplementedError('MultiScaleDeformableAttention kernel: 32 channels per head')
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.batch_first = batch_first
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)

    @torch.no_grad()
    def _prepare(self):
        self._prep = dict(
            w=torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0).contiguous(),
            b=torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0).contiguous())

    @torch.no_grad()
    def forward_tokens(self, x, pos, ref, spatial_shapes, x_planes=None, q_planes=None):
        """x, pos [B,N,C]; ref [N,2] -> output_proj(msda) + x.  x_planes / q_planes: operand planes of
        x and of x + pos when the producer (the previous LayerNorm) already emitted them."""
        if self._prep is None or self._stale():
            self._prepare()
        proj = ops.linear(q_planes, self._prep['w'], self._prep['b']) if q_planes is not None else \
            ops.linear(x, self._prep['w'], self._prep['b'], add_input=pos)
        value = ops.linear(x_planes if x_planes is not None else x, self.value_proj.weight, self.value_proj.bias)
        samp = ops.msda_fused_forward(value, spatial_shapes, proj, ref, self.num_heads, self.num_points,
                                      out_mode='split')   # planes for the output projection
        return ops.linear(samp, self.output_proj.weight, self.output_proj.bias, residual=x)

    def forward_tokens_train(self, x, pos, ref, spatial_shapes):
        """``forward_tokens`` on the autograd tape (train_ops)."""
        from . import train_ops as T
        w = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
        b = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0)
        proj = T.linear(x, w, b, add_input=pos)
        value = T.linear(x, self.value_proj.weight, self.value_proj.bias)
        samp = T.msda_fused(value, proj, ref, spatial_shapes, self.num_heads, self.num_points)
        return T.linear(samp, self.output_proj.weight, self.output_proj.bias, residual=x)

    @torch.no_grad()
    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        """mmcv signature (seq-first unless batch_first).  reference_points [B,Nq,L,2]."""
        if value is None:
            value = query
        if identity is None:
            identity = query
        if key_padding_mask is not None and bool(key_padding_mask.any()):
            raise NotImplementedError('key_padding_mask with masked positions')
        q = query if query_pos is None else None
        if not self.batch_first:
            query, value, identity = (t.permute(1, 0, 2).contiguous() for t in (query, value, identity))
            if query_pos is not None:
                query_pos = query_pos.permute(1, 0, 2).contiguous()
        B, Nq, C = query.shape
        shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()] if torch.is_tensor(spatial_shapes) \
            else list(spatial_shapes)
        H, L, P = self.num_heads, self.num_levels, self.num_points
        v = ops.linear(value, self.value_proj.weight, self.value_proj.bias)
        off = ops.linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias, add_input=query_pos)
        aw = ops.linear(query, self.attention_weights.weight, self.attention_weights.bias, add_input=query_pos)
        # general reference points (per level): use the unfused op with explicit locations
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=query.device)
        loc = reference_points[:, :, None, :, None, :] + off.view(B, Nq, H, L, P, 2) / norm[None, None, None, :, None, :]
        aw = aw.view(B, Nq, H, L * P).softmax(-1).view(B, Nq, H, L, P)
        out = ops.msda_forward(v.view(B, -1, H, C // H), shapes, loc.contiguous(), aw.contiguous())
        out = ops.linear(out, self.output_proj.weight, self.output_proj.bias, residual=identity)
        del q
        return out if self.batch_first else out.permute(1, 0, 2)


@FEEDFORWARD_NETWORK.register_module()
class FFN(nn.Module):
    """mmcv FFN: keys ``layers.0.0`` (Linear+act) and ``layers.1`` (Linear); add_identity."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0.,
                 dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        if num_fcs != 2:
            raise NotImplementedError('FFN: num_fcs=2 only')
        self.add_identity = add_identity
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                                                  nn.Dropout(ffn_drop)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    @torch.no_grad()
    def forward(self, x, identity=None):
        h = ops.linear(x, self.layers[0][0].weight, self.layers[0][0].bias, act=ops.ACT_RELU)
        res = (x if identity is None else identity) if self.add_identity else None
        return ops.linear(h, self.layers[1].weight, self.layers[1].bias, residual=res)


class _Norm(nn.LayerNorm):
    @torch.no_grad()
    def forward(self, x):
        return ops.layernorm(x if x.is_contiguous() else x.contiguous(), self.weight, self.bias, self.eps)


@TRANSFORMER_LAYER.register_module()
class BaseTransformerLayer(nn.Module):
    """mmcv BaseTransformerLayer restricted to the encoder form used by the pixel decoder:
    operation_order = ('self_attn', 'norm', 'ffn', 'norm') with MultiScaleDeformableAttention."""

    def __init__(self, attn_cfgs=None, ffn_cfgs=None, operation_order=None, norm_cfg=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        if tuple(operation_order) != ('self_attn', 'norm', 'ffn', 'norm'):
            raise NotImplementedError(f'BaseTransformerLayer: unsupported operation_order {operation_order}')
        attn_cfgs = to_cfg(attn_cfgs)
        self.attentions = nn.ModuleList([ATTENTION.build(attn_cfgs)])
        self.embed_dims = self.attentions[0].embed_dims
        ffn = dict(to_cfg(ffn_cfgs))
        ffn.pop('type', None)
        ffn.setdefault('embed_dims', self.embed_dims)
        self.ffns = nn.ModuleList([FFN(**ffn)])
        self.norms = nn.ModuleList([_Norm(self.embed_dims), _Norm(self.embed_dims)])

    @torch.no_grad()
    def forward_tokens(self, x, pos, ref, spatial_shapes, x_planes=None, q_planes=None, next_q=True):
        """Returns (x, planes of x, planes of x + pos): both LayerNorms emit the operand planes of
        their output for the GEMMs that consume it (the last one also those of x + pos, the next
        layer's query), and the FFN hidden layer only ever exists as planes."""
        n0, n1, ffn = self.norms[0], self.norms[1], self.ffns[0]
        x = self.attentions[0].forward_tokens(x, pos, ref, spatial_shapes, x_planes, q_planes)
        x, xs = ops.layernorm(x, n0.weight, n0.bias, n0.eps, out_split=True)
        h = ops.linear(xs if xs is not None else x, ffn.layers[0][0].weight, ffn.layers[0][0].bias,
                       act=ops.ACT_RELU, out_mode='split')
        x = ops.linear(h, ffn.layers[1].weight, ffn.layers[1].bias, residual=x if ffn.add_identity else None)
        if next_q:
            return ops.layernorm(x, n1.weight, n1.bias, n1.eps, out_split=True, add=pos)
        return ops.layernorm(x, n1.weight, n1.bias, n1.eps, out_split=True) + (None,)


    def _forward_tokens_train(self, x, pos, ref, spatial_shapes):
        from . import train_ops as T
        n0, n1, ffn = self.norms[0], self.norms[1], self.ffns[0]
        x = T.layernorm(self.attentions[0].forward_tokens_train(x, pos, ref, spatial_shapes), n0)
        h = T.linear(x, ffn.layers[0][0].weight, ffn.layers[0][0].bias, act=ops.ACT_RELU)
        x = T.linear(h, ffn.layers[1].weight, ffn.layers[1].bias, residual=x if ffn.add_identity else None)
        return T.layernorm(x, n1)


BaseTransformerLayer.forward_tokens_train = BaseTransformerLayer._forward_tokens_train


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerEncoder(nn.Module):
    def __init__(self, transformerlayers=None, num_layers=6, post_norm_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        cfg = dict(to_cfg(transformerlayers))
        cfg.pop('type', None)
        self.layers = nn.ModuleList([BaseTransformerLayer(**copy.deepcopy(cfg)) for _ in range(num_layers)])
        self.embed_dims = self.layers[0].embed_dims


class _ConvModule(nn.Module):
    """mmcv ConvModule: conv -> GN -> act, keys ``conv`` / ``gn``."""

    def __init__(self, cin, cout, k, padding=0, bias=False, act=False, groups=32):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=bias)
        self.gn = nn.GroupNorm(groups, cout)
        self.act = act
        self.pad = padding
        self._w = None

    @torch.no_grad()
    def forward_tokens(self, x, planes=None):
        if self._w is None or self._w.device != self.conv.weight.device or self._w_ver != self.conv.weight._version:
            self._w = self.conv.weight.permute(0, 2, 3, 1).contiguous()
            self._w_ver = self.conv.weight._version
        if planes is None:
            planes = ops.recall_split(x)   # e.g. backbone stage outputs already carry operand planes
        y = ops.conv2d_nhwc(planes if planes is not None else x, self._w, self.conv.bias, pad=self.pad)
        return y

    def forward_tokens_train(self, x):
        """conv -> GN (-> ReLU) on the autograd tape; x [B,H,W,Cin] token-major."""
        from . import train_ops as T
        if self.conv.kernel_size == (1, 1):
            y = T.linear(x, self.conv.weight.view(self.conv.out_channels, -1), self.conv.bias)
        else:
            assert self.conv.kernel_size == (3, 3) and self.conv.bias is None
            y = T.conv3x3(x, self.conv.weight)
        return T.groupnorm(y, self.gn, relu=self.act)

    def norm_tokens(self, y, out_mode='f32'):
        return ops.groupnorm_nhwc(y, self.gn.weight, self.gn.bias, self.gn.num_groups, self.gn.eps,
                                  ops.ACT_RELU if self.act else ops.ACT_NONE, out_mode=out_mode)

    def _load_from_state_dict(self, *a, **k):
        self._w = None
        return super()._load_from_state_dict(*a, **k)


@PLUGIN_LAYERS.register_module()
class MSDeformAttnPixelDecoder(_Prepared):
    """mmdet MSDeformAttnPixelDecoder (cfg mask2former_video_r50_base.py:27-59)."""

    def __init__(self, in_channels=(256, 512, 1024, 2048), strides=(4, 8, 16, 32), feat_channels=256,
                 out_channels=256, num_outs=3, norm_cfg=None, act_cfg=None, encoder=None,
                 positional_encoding=None, init_cfg=None):
        super().__init__()
        self.strides = list(strides)
        self.num_input_levels = len(in_channels)
        enc = to_cfg(encoder)
        self.num_encoder_levels = enc.transformerlayers.attn_cfgs.num_levels
        self.num_outs = num_outs
        groups = (norm_cfg or {}).get('num_groups', 32)
        self.input_convs = nn.ModuleList([
            _ConvModule(in_channels[i], feat_channels, 1, bias=True, groups=groups)
            for i in range(self.num_input_levels - 1, self.num_input_levels - self.num_encoder_levels - 1, -1)])
        self.encoder = TRANSFORMER_LAYER_SEQUENCE.build(enc)
        self.postional_encoding = build_positional_encoding(positional_encoding)  # (sic) mmdet attribute name
        self.level_encoding = nn.Embedding(self.num_encoder_levels, feat_channels)
        self.lateral_convs = nn.ModuleList()
        self.output_convs = nn.ModuleList()
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            self.lateral_convs.append(_ConvModule(in_channels[i], feat_channels, 1, groups=groups))
            self.output_convs.append(_ConvModule(feat_channels, feat_channels, 3, padding=1, act=True, groups=groups))
        self.mask_feature = nn.Conv2d(feat_channels, out_channels, 1)
        self._shape_cache = {}

    def init_weights(self):
        pass

    @torch.no_grad()
    def _shape_consts(self, shapes, device):
        """level positional encodings + reference points; depend on the shapes only."""
        key = (tuple(shapes), str(device))
        if key not in self._shape_cache or self._prep is None or self._stale():
            pos, refs = [], []
            for i, (h, w) in enumerate(shapes):
                pos.append(self.postional_encoding.tokens(h, w, device, add_vec=self.level_encoding.weight[i].contiguous()))
                ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                                        indexing='ij')
                refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
            self._shape_cache = {key: (torch.cat(pos, 0).contiguous(), torch.cat(refs, 0).to(device).contiguous())}
            self._prep = True
        return self._shape_cache[key]

    def _batched_pos(self, pos, B):
        """pos broadcast over the batch, materialised once per (shape, batch)."""
        if B == 1:
            return pos[None]
        key = ('posb', pos.data_ptr(), tuple(pos.shape), B)
        hit = self._shape_cache.get(key)
        if hit is None:
            hit = pos[None].expand(B, -1, -1).contiguous()
            if not (pos.is_cuda and torch.cuda.is_current_stream_capturing()):
                self._shape_cache[key] = hit   # tensors created during graph capture belong to the graph
        return hit

    def forward_train(self, feats):
        """``forward`` on the autograd tape: same outputs (token-major views, logical NCHW), differentiable w.r.t. every
        parameter of the pixel decoder (the backbone maps ``feats`` are constants)."""
        from . import train_ops as T
        B = feats[0].shape[0]
        toks, shapes = [], []
        for i in range(self.num_encoder_levels):
            y = self.input_convs[i].forward_tokens_train(_tokens(feats[self.num_input_levels - i - 1]))
            shapes.append((y.shape[1], y.shape[2]))
            toks.append(y.view(B, -1, y.shape[-1]))
        x = torch.cat(toks, 1)
        ref = self._shape_consts(shapes, x.device)[1]
        with torch.no_grad():
            sine = [self.postional_encoding.tokens(h, w, x.device) for h, w in shapes]
        pos = torch.cat([T.add_rowvec(pe, self.level_encoding.weight[i]) for i, pe in enumerate(sine)], 0)
        posb = pos[None].expand(B, -1, -1).contiguous() if B > 1 else pos[None]
        for layer in self.encoder.layers:
            x = layer.forward_tokens_train(x, posb, ref, shapes)
        outs, start = [], 0
        for (h, w) in shapes:
            outs.append(x[:, start:start + h * w].reshape(B, h, w, -1))
            start += h * w
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            j = self.num_input_levels - self.num_encoder_levels - 1 - i
            cur = T.resize_add(self.lateral_convs[j].forward_tokens_train(_tokens(feats[i])), outs[-1])
            outs.append(self.output_convs[j].forward_tokens_train(cur))
        wmf = self.mask_feature.weight.view(self.mask_feature.out_channels, -1)
        mf = T.linear(outs[-1], wmf, self.mask_feature.bias)
        return _as_nchw(mf), [_as_nchw(o) for o in outs[:self.num_outs]]

    @torch.no_grad()
    def forward(self, feats):
        """feats: 4 maps [B,C,H,W] (C2..C5) -> (mask_feature [B,C,h4,w4], [m32, m16, m8])."""
        B = feats[0].shape[0]
        toks, shapes = [], []
        for i in range(self.num_encoder_levels):
            f = _tokens(feats[self.num_input_levels - i - 1])
            cm = self.input_convs[i]
            y = cm.norm_tokens(cm.forward_tokens(f))
            shapes.append((y.shape[1], y.shape[2]))
            toks.append(y.view(B, -1, y.shape[-1]))
        x = torch.cat(toks, 1)  # pure data movement (torch.cat = cudaMemcpy-class op)
        pos, ref = self._shape_consts(shapes, x.device)
        posb = self._batched_pos(pos, B)
        xs = qs = None
        nlay = len(self.encoder.layers)
        for li, layer in enumerate(self.encoder.layers):
            x, xs, qs = layer.forward_tokens(x, posb, ref, shapes, xs, qs, next_q=li + 1 < nlay)
        outs, start = [], 0
        for (h, w) in shapes:
            outs.append(x[:, start:start + h * w].reshape(B, h, w, -1))
            start += h * w
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            j = self.num_input_levels - self.num_encoder_levels - 1 - i
            lat = self.lateral_convs[j]
            cur = lat.norm_tokens(lat.forward_tokens(_tokens(feats[i])))
            # += upsampled coarser map (read through its batch-strided view), planes for the 3x3 conv
            cur, cur_planes = ops.bilinear_resize_nhwc(outs[-1], cur.shape[1:3], out=cur, accumulate=True, out_split=True)
            oc = self.output_convs[j]
            # maps beyond num_outs only feed the next FPN step / the mask-feature conv: planes suffice
            last = i == 0 and len(outs) >= self.num_outs
            outs.append(oc.norm_tokens(oc.forward_tokens(cur, cur_planes), out_mode='split' if last else 'f32'))
        wmf = self.mask_feature.weight.view(self.mask_feature.out_channels, -1)
        # fp32 for the API / pooling, planes for the ten mask-logit contractions
        mf, mf_planes = ops.linear(outs[-1], wmf, self.mask_feature.bias, out_mode='both')
        ops.remember_split(mf, mf_planes)
        return _as_nchw(mf), [_as_nchw(o) for o in outs[:self.num_outs]]


# ======================================================================================
# transformer decoder (L0, A5)
# ======================================================================================
@ATTENTION.register_module()
class MultiheadAttention(nn.Module):
    """mmcv MultiheadAttention wrapper: parameters live in ``attn`` (torch nn.MultiheadAttention
    key names); forward = identity + MHA(q + query_pos, k + key_pos, v)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, batch_first=False,
                 init_cfg=None, **kwargs):
        super().__init__()
        if batch_first:
            raise NotImplementedError('MultiheadAttention: batch_first=False only (as in the reference cfg)')
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)

    @torch.no_grad()
    def forward_tokens(self, query, query_pos, kproj=None, vproj=None, mask=None, row_open=None):
        """batch-first tokens.  query/query_pos [B,Q,E].  Cross attention passes the
        precomputed key / value projections [B,Lk,E]; self attention passes none."""
        E = self.embed_dims
        w, b = self.attn.in_proj_weight, self.attn.in_proj_bias
        if kproj is None:  # self attention: q, k from query + pos; v from query
            qk = ops.linear(query, w[:2 * E], b[:2 * E], add_input=query_pos)
            v = ops.linear(query, w[2 * E:], b[2 * E:])
            q, k = qk[..., :E], qk[..., E:]
        else:
            q = ops.linear(query, w[:E], b[:E], add_input=query_pos)
            k, v = kproj, vproj
        o = ops.attention(q, k, v, self.num_heads, mask=mask, row_open=row_open)
        return ops.linear(o, self.attn.out_proj.weight, self.attn.out_proj.bias, residual=query)

    @torch.no_grad()
    def project_kv(self, key, key_pos, value, key_planes=None, value_planes=None):
        """key_planes / value_planes: operand planes of (key + key_pos) and value when the caller
        has them (each decoder level is projected by three different layers)."""
        E = self.embed_dims
        w, b = self.attn.in_proj_weight, self.attn.in_proj_bias
        # on the tcgen05 engine the projections leave as operand planes: the tensor-core attention
        # kernel consumes them directly (head dim 32)
        mode = 'split' if E // self.num_heads == 32 else 'f32'
        k = ops.linear(key_planes, w[E:2 * E], b[E:2 * E], out_mode=mode) if key_planes is not None else \
            ops.linear(key, w[E:2 * E], b[E:2 * E], add_input=key_pos, out_mode=mode)
        v = ops.linear(value_planes if value_planes is not None else value, w[2 * E:], b[2 * E:], out_mode=mode)
        if isinstance(k, ops.Split) != isinstance(v, ops.Split):   # mixed engines cannot happen for equal shapes
            raise ops._l.PvsgError('project_kv: inconsistent operand formats')
        return k, v

    @torch.no_grad()
    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        """mmcv signature, seq-first [L,B,E]; attn_mask bool [B*H,Lq,Lk] (True = blocked) or None."""
        if key is None:
            key = query
        if value is None:
            value = key
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if key_padding_mask is not None and bool(key_padding_mask.any()):
            raise NotImplementedError('key_padding_mask with masked positions')
        bf = lambda t: None if t is None else t.permute(1, 0, 2).contiguous()  # noqa: E731
        q, k, v, qp, kp = bf(query), bf(key), bf(value), bf(query_pos), bf(key_pos)
        mask = row_open = None
        if attn_mask is not None:
            B = q.shape[0]
            m = attn_mask.view(B, self.num_heads, *attn_mask.shape[1:])
            if not bool((m == m[:, :1]).all()):
                raise NotImplementedError('per-head attention masks')
            mask = m[:, 0].to(torch.uint8).contiguous()
            row_open = None  # an explicit mask is used as given
        kproj, vproj = self.project_kv(k, kp, v)
        out = self.forward_tokens(q, qp, kproj, vproj, mask, row_open)
        if identity is not None:
            out = out - q + bf(identity)
        return out.permute(1, 0, 2)


@TRANSFORMER_LAYER.register_module()
class DetrTransformerDecoderLayer(nn.Module):
    """mmdet DetrTransformerDecoderLayer, operation_order
    ('cross_attn','norm','self_attn','norm','ffn','norm') -- mask2former_video_r50_base.py:63-88."""

    def __init__(self, attn_cfgs=None, ffn_cfgs=None, feedforward_channels=None, operation_order=None,
                 norm_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        if tuple(operation_order) != ('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'):
            raise NotImplementedError(f'DetrTransformerDecoderLayer: unsupported operation_order {operation_order}')
        a = dict(to_cfg(attn_cfgs))
        a.pop('type', None)
        self.attentions = nn.ModuleList([MultiheadAttention(**a), MultiheadAttention(**a)])
        self.embed_dims = self.attentions[0].embed_dims
        f = dict(to_cfg(ffn_cfgs or {}))
        f.pop('type', None)
        f.setdefault('embed_dims', self.embed_dims)
        if feedforward_channels is not None:
            f['feedforward_channels'] = feedforward_channels
        self.ffns = nn.ModuleList([FFN(**f)])
        self.norms = nn.ModuleList([_Norm(self.embed_dims) for _ in range(3)])

    @torch.no_grad()
    def forward_tokens(self, query, query_pos, kproj, vproj, mask, row_open, q_planes=None):
        """One decoder layer on batch-first tokens.  Returns (query, planes of query + query_pos): the
        LayerNorms emit the operand planes their consumers need (y, and y + pos for the projections that
        take the positional embedding), so the chain runs without separate split passes.
        q_planes: planes of query + query_pos from the previous layer (None on the first / SIMT path)."""
        ca, sa, ffn = self.attentions[0], self.attentions[1], self.ffns[0]
        n0, n1, n2 = self.norms
        E = self.embed_dims
        # cross attention
        w, b = ca.attn.in_proj_weight, ca.attn.in_proj_bias
        q = ops.linear(q_planes, w[:E], b[:E]) if q_planes is not None else \
            ops.linear(query, w[:E], b[:E], add_input=query_pos)
        o = ops.attention(q, kproj, vproj, ca.num_heads, mask=mask, row_open=row_open)
        x = ops.linear(o, ca.attn.out_proj.weight, ca.attn.out_proj.bias, residual=query)
        x, xs, xq = ops.layernorm(x, n0.weight, n0.bias, n0.eps, out_split=True, add=query_pos)
        # self attention: q, k from x + pos; v from x
        w, b = sa.attn.in_proj_weight, sa.attn.in_proj_bias
        qk = ops.linear(xq, w[:2 * E], b[:2 * E]) if xq is not None else \
            ops.linear(x, w[:2 * E], b[:2 * E], add_input=query_pos)
        v = ops.linear(xs if xs is not None else x, w[2 * E:], b[2 * E:])
        o = ops.attention(qk[..., :E], qk[..., E:], v, sa.num_heads)
        x = ops.linear(o, sa.attn.out_proj.weight, sa.attn.out_proj.bias, residual=x)
        x, xs = ops.layernorm(x, n1.weight, n1.bias, n1.eps, out_split=True)
        # FFN: the hidden layer only exists as planes
        h = ops.linear(xs if xs is not None else x, ffn.layers[0][0].weight, ffn.layers[0][0].bias,
                       act=ops.ACT_RELU, out_mode='split')
        x = ops.linear(h, ffn.layers[1].weight, ffn.layers[1].bias, residual=x if ffn.add_identity else None)
        x, _, xq = ops.layernorm(x, n2.weight, n2.bias, n2.eps, out_split=True, add=query_pos)
        return x, xq

    def forward_tokens_train(self, query, query_pos, kproj, vproj, mask, row_open):
        """``forward_tokens`` on the autograd tape (train_ops): same operation order, fp32 tensors, every forward and
        backward step a library kernel.  Dropout rates of the reference config are all 0 (base cfg :70-83)."""
        from . import train_ops as T
        ca, sa, ffn = self.attentions[0], self.attentions[1], self.ffns[0]
        n0, n1, n2 = self.norms
        E = self.embed_dims
        w, b = ca.attn.in_proj_weight, ca.attn.in_proj_bias
        q = T.linear(query, w[:E], b[:E], add_input=query_pos)
        o = T.attention(q, kproj, vproj, ca.num_heads, mask, row_open)
        x = T.layernorm(T.linear(o, ca.attn.out_proj.weight, ca.attn.out_proj.bias, residual=query), n0)
        w, b = sa.attn.in_proj_weight, sa.attn.in_proj_bias
        qk = T.linear(x, w[:2 * E], b[:2 * E], add_input=query_pos)
        v = T.linear(x, w[2 * E:], b[2 * E:])
        o = T.attention(qk[..., :E], qk[..., E:], v, sa.num_heads)
        x = T.layernorm(T.linear(o, sa.attn.out_proj.weight, sa.attn.out_proj.bias, residual=x), n1)
        h = T.linear(x, ffn.layers[0][0].weight, ffn.layers[0][0].bias, act=ops.ACT_RELU)
        x = T.linear(h, ffn.layers[1].weight, ffn.layers[1].bias, residual=x if ffn.add_identity else None)
        return T.layernorm(x, n2)

    @torch.no_grad()
    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        """mmcv BaseTransformerLayer signature (seq-first), as called at mask2former_head.py:457-468."""
        attn_masks = attn_masks or [None, None]
        x = self.attentions[0](query, key, value, None, query_pos=query_pos, key_pos=key_pos,
                               attn_mask=attn_masks[0], key_padding_mask=key_padding_mask)
        x = self.norms[0](x.contiguous())
        x = self.attentions[1](x, x, x, None, query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[1],
                               key_padding_mask=query_key_padding_mask)
        x = self.norms[1](x.contiguous())
        x = self.ffns[0](x)
        return self.norms[2](x)


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerDecoder(nn.Module):
    def __init__(self, transformerlayers=None, num_layers=9, return_intermediate=False, post_norm_cfg=None,
                 init_cfg=None, **kwargs):
        super().__init__()
        cfg = dict(to_cfg(transformerlayers))
        cfg.pop('type', None)
        self.layers = nn.ModuleList([DetrTransformerDecoderLayer(**copy.deepcopy(cfg)) for _ in range(num_layers)])
        self.embed_dims = self.layers[0].embed_dims
        self.post_norm = _Norm(self.embed_dims)
        self.return_intermediate = return_intermediate


# ======================================================================================
# heads
# ======================================================================================
class _Mask2FormerHeadBase(_Prepared):
    """Shared implementation of Mask2FormerHeadCustom (models/mask2former/mask2former_head.py)
    and Mask2FormerVideoHead (models/mask2former_vps/mask2former_video_head.py)."""
    video = False

    def __init__(self, in_channels, feat_channels, out_channels, num_things_classes=80, num_stuff_classes=53,
                 num_queries=100, num_transformer_feat_level=3, pixel_decoder=None,
                 enforce_decoder_input_project=False, transformer_decoder=None, positional_encoding=None,
                 loss_cls=None, loss_mask=None, loss_dice=None, train_cfg=None, test_cfg=None, init_cfg=None,
                 **kwargs):
        super().__init__()
        transformer_decoder = to_cfg(transformer_decoder)
        pixel_decoder = to_cfg(pixel_decoder)
        self.num_things_classes = num_things_classes
        self.num_stuff_classes = num_stuff_classes
        self.num_classes = num_things_classes + num_stuff_classes
        self.num_queries = num_queries
        self.num_transformer_feat_level = num_transformer_feat_level
        self.num_heads = transformer_decoder.transformerlayers.attn_cfgs.num_heads
        self.num_transformer_decoder_layers = transformer_decoder.num_layers
        assert pixel_decoder.encoder.transformerlayers.attn_cfgs.num_levels == num_transformer_feat_level
        pixel_decoder_ = copy.deepcopy(pixel_decoder)
        pixel_decoder_.update(in_channels=in_channels, feat_channels=feat_channels, out_channels=out_channels)
        self.pixel_decoder = build_plugin_layer(pixel_decoder_)[1]
        self.transformer_decoder = build_transformer_layer_sequence(transformer_decoder)
        self.decoder_embed_dims = self.transformer_decoder.embed_dims
        if self.decoder_embed_dims != feat_channels or enforce_decoder_input_project:
            raise NotImplementedError('decoder_input_projs other than Identity')
        self.decoder_input_projs = nn.ModuleList([nn.Identity() for _ in range(num_transformer_feat_level)])
        self.decoder_positional_encoding = build_positional_encoding(positional_encoding)
        self.query_embed = nn.Embedding(num_queries, feat_channels)
        self.query_feat = nn.Embedding(num_queries, feat_channels)
        self.level_embed = nn.Embedding(num_transformer_feat_level, feat_channels)
        self.cls_embed = nn.Linear(feat_channels, self.num_classes + 1)
        self.mask_embed = nn.Sequential(nn.Linear(feat_channels, feat_channels), nn.ReLU(inplace=True),
                                        nn.Linear(feat_channels, feat_channels), nn.ReLU(inplace=True),
                                        nn.Linear(feat_channels, out_channels))
        self.test_cfg, self.train_cfg = test_cfg, train_cfg
        self._pe_cache = {}
        # debug / parity hook: when set to a list, ``_run`` appends the raw sign masks (uint8 [B,Q,hw], non-zero =
        # blocked) it computes for decoder layers 0..L-1, so a test can resolve near-threshold ties the same way
        self._capture_masks = None
        self.train_pixel_decoder = True     # forward_train: False freezes the pixel decoder (it then runs the inference kernels)

    def init_weights(self):
        pass

    # ---- training: decoder head on the autograd tape, pixel decoder / backbone frozen ----
    def _embeds_train(self, query):
        from . import train_ops as T
        x = T.layernorm(query, self.transformer_decoder.post_norm)
        cls_pred = T.linear(x, self.cls_embed.weight, self.cls_embed.bias)
        me = T.linear(x, self.mask_embed[0].weight, self.mask_embed[0].bias, act=ops.ACT_RELU)
        me = T.linear(me, self.mask_embed[2].weight, self.mask_embed[2].bias, act=ops.ACT_RELU)
        return cls_pred, T.linear(me, self.mask_embed[4].weight, self.mask_embed[4].bias)

    def forward_train_outputs(self, feats, num_frames=1):
        """The training-mode ``forward`` (mask2former_video_head.py:361-462 / mask2former_head.py:397-479): ->
        (all_cls_scores [L+1] x [B,Q,NC+1], all_mask_preds [L+1] x [B,T,Q,h,w] (video) / [B,Q,h,w] (image)), differentiable
        w.r.t. every parameter of the head: pixel decoder (``train_pixel_decoder``, default on), transformer decoder,
        prediction heads, query / level embeddings.  The backbone maps ``feats`` are constants."""
        from . import train_ops as T
        if self.train_pixel_decoder:
            mask_features, memories = self.pixel_decoder.forward_train(feats)
        else:
            with torch.no_grad():
                mask_features, memories = self.pixel_decoder(feats)
        BT = mask_features.shape[0]
        Tn = num_frames
        B = BT // Tn
        assert B * Tn == BT
        mf = _tokens(mask_features)
        _, h4, w4, C = mf.shape
        mf_flat = mf.view(B, Tn * h4 * w4, C)
        lvl_shapes = [tuple(m.shape[-2:]) for m in memories]
        toks = [_tokens(m) for m in memories]
        with torch.no_grad():
            pooled = [ops.bilinear_resize_nhwc(mf.detach(), s).view(B, -1, C) for s in lvl_shapes]
            dec_pe = [self._batched(self._decoder_pe((Tn, h, w), mf.device), B) for h, w in lvl_shapes]
        dec_in = [T.add_rowvec(tok, self.level_embed.weight[i]).view(B, -1, C) for i, tok in enumerate(toks)]
        Q = self.num_queries
        query = T.expand_batch(self.query_feat.weight, B)
        qpos = T.expand_batch(self.query_embed.weight, B)
        layers = self.transformer_decoder.layers
        nl = self.num_transformer_decoder_layers
        E = self.decoder_embed_dims

        def predict(query, nxt):
            cls_pred, me = self._embeds_train(query)
            mask_pred = T.mask_logits(me, mf_flat).view(B, Q, Tn, h4, w4).transpose(1, 2)
            with torch.no_grad():          # attn_mask = (interpolate(mask_pred).sigmoid() < 0.5).detach(), :346-357
                _, mask, row_open = ops.mask_logits(me.detach(), pooled[nxt], False, True)
            return cls_pred, mask_pred, mask, row_open

        cls_list, mask_list = [], []
        cls_pred, mask_pred, mask, row_open = predict(query, 0)
        cls_list.append(cls_pred)
        mask_list.append(mask_pred)
        for i in range(nl):
            lvl = i % self.num_transformer_feat_level
            ca = layers[i].attentions[0]
            w, b = ca.attn.in_proj_weight, ca.attn.in_proj_bias
            kproj = T.linear(dec_in[lvl], w[E:2 * E], b[E:2 * E], add_input=dec_pe[lvl])
            vproj = T.linear(dec_in[lvl], w[2 * E:], b[2 * E:])
            if self._capture_masks is not None:
                self._capture_masks.append(mask)
            query = layers[i].forward_tokens_train(query, qpos, kproj, vproj, mask, row_open)
            cls_pred, mask_pred, mask, row_open = predict(query, (i + 1) % self.num_transformer_feat_level)
            cls_list.append(cls_pred)
            mask_list.append(mask_pred)
        if not self.video:
            mask_list = [m[:, 0] for m in mask_list]
        return cls_list, mask_list

    def preprocess_gt(self, gt_labels_list, gt_masks_list, gt_semantic_seg, gt_instance_ids, img_metas):
        """maskformer_video_head.py:138-180 + utils.py:94-140 (``preprocess_video_panoptic_gt``): per clip, one label and
        one [T,H,W] mask stack per instance id; frames without the instance get an empty mask.  ``gt_labels`` /
        ``gt_instance_ids`` [n,2] = (frame, value); ``gt_masks``: per frame a [n_f,H,W] tensor or a BitmapMasks-like object
        (``.pad(shape, pad_val).to_tensor(dtype, device)``).  Index bookkeeping only."""
        labels_out, masks_out = [], []
        for gt_labels, gt_masks, ids, metas in zip(gt_labels_list, gt_masks_list, gt_instance_ids, img_metas):
            dev = gt_labels.device
            frames = []
            for f, meta in enumerate(metas):
                m = gt_masks[f]
                if hasattr(m, 'pad'):
                    m = m.pad(meta['pad_shape'][:2], pad_val=0).to_tensor(dtype=torch.bool, device=dev)
                else:
                    ph, pw = meta['pad_shape'][:2]
                    m = torch.nn.functional.pad(m.to(dev).bool(), (0, pw - m.shape[-1], 0, ph - m.shape[-2]))
                frames.append(m)
            labels, things = [], []
            for inst in torch.unique(ids[:, 1]):
                pos = torch.nonzero(ids[:, 1] == inst, as_tuple=True)[0]
                lab = gt_labels[:, 1][pos]
                assert bool((lab == lab[0]).all())
                labels.append(lab[0])
                in_frames = ids[:, 0][pos].to(torch.int32).tolist()
                stack = []
                for f, meta in enumerate(metas):
                    if f not in in_frames:
                        stack.append(torch.zeros(tuple(meta['pad_shape'][:2]), dtype=torch.bool, device=dev))
                    else:
                        frame_ids = ids[ids[:, 0] == f, 1]
                        stack.append(frames[f][int(torch.nonzero(frame_ids == inst, as_tuple=True)[0].item())])
                things.append(torch.stack(stack))
            labels_out.append(torch.stack(labels).long())
            masks_out.append(torch.stack(things).long())
        return labels_out, masks_out

    def preprocess_gt_image(self, gt_labels_list, gt_masks_list, gt_semantic_segs, img_metas):
        """mmdet MaskFormerHead.preprocess_gt -> ``preprocess_panoptic_gt`` (mmdet 2.25 models/utils/panoptic_gt_processing.py,
        called from mask2former_head.py:519): per image, the instance masks padded to pad_shape, followed by one mask per
        stuff class present in the semantic map (labels in [num_things, num_classes); 255 = void).  Index bookkeeping."""
        if gt_semantic_segs is None:
            gt_semantic_segs = [None] * len(gt_labels_list)
        labels_out, masks_out = [], []
        for labels, gm, sem, meta in zip(gt_labels_list, gt_masks_list, gt_semantic_segs, img_metas):
            dev = labels.device
            ph, pw = meta['pad_shape'][:2]
            if hasattr(gm, 'pad'):
                things = gm.pad((ph, pw), pad_val=0).to_tensor(dtype=torch.bool, device=dev)
            else:
                things = torch.nn.functional.pad(gm.to(dev).bool(), (0, pw - gm.shape[-1], 0, ph - gm.shape[-2]))
            if sem is not None:
                sem = sem.to(dev).squeeze(0)
                stuff = [c for c in torch.unique(sem).tolist() if self.num_things_classes <= c < self.num_classes]
                if stuff:
                    things = torch.cat([things, torch.stack([sem == c for c in stuff])], 0)
                    labels = torch.cat([labels, torch.tensor(stuff, dtype=labels.dtype, device=dev)], 0)
            labels_out.append(labels.long())
            masks_out.append(things.long())
        return labels_out, masks_out

    def forward_train(self, feats, img_metas, gt_bboxes, gt_labels, gt_masks, gt_semantic_seg=None, gt_instance_ids=None,
                      gt_bboxes_ignore=None):
        """mask2former_video_head.py:464-522 (``loss_sem_seg=None``, the shipped configs): forward -> preprocess_gt ->
        loss.  Returns the reference's loss dict; ``.backward()`` on its sum fills the gradients of the decoder head."""
        assert gt_bboxes_ignore is None
        if not self.video:         # mask2former_head.py:481-523: (feats, img_metas, gt_bboxes, gt_labels, gt_masks, gt_semantic_seg)
            all_cls_scores, all_mask_preds = self.forward_train_outputs(feats, 1)
            labels, masks = self.preprocess_gt_image(gt_labels, gt_masks, gt_semantic_seg, img_metas)
            return self.loss(all_cls_scores, all_mask_preds, labels, masks, img_metas)
        num_frames = len(img_metas[0])
        all_cls_scores, all_mask_preds = self.forward_train_outputs(feats, num_frames)
        labels, masks = self.preprocess_gt(gt_labels, gt_masks, gt_semantic_seg, gt_instance_ids, img_metas)
        return self.loss(all_cls_scores, all_mask_preds, labels, masks, img_metas)

    def loss_single(self, cls_scores, mask_preds, gt_labels_list, gt_masks_list, img_metas=None, **point_sets):
        """mask2former_video_head.py:196-293 (mask_preds [B,T,Q,h,w]) / mask2former_head.py:233-318 ([B,Q,h,w]):
        -> (loss_cls, loss_mask, loss_dice), differentiable w.r.t. cls_scores / mask_preds."""
        from . import losses
        if mask_preds.dim() == 4:                       # image head: one frame
            mask_preds = mask_preds[:, None]
            gt_masks_list = [g[:, None] for g in gt_masks_list]
        tc = dict(self.train_cfg or {})
        kw = dict(num_points=tc.get('num_points', 12544), oversample_ratio=tc.get('oversample_ratio', 3.0),
                  importance_sample_ratio=tc.get('importance_sample_ratio', 0.75))
        kw.update(point_sets)
        return losses.loss_single(cls_scores, mask_preds, gt_labels_list, gt_masks_list, img_metas, num_classes=self.num_classes, **kw)

    def loss(self, all_cls_scores, all_mask_preds, gt_labels_list, gt_masks_list, img_metas=None):
        """mask2former_video_head.py:524-634, the default (loss_split_th_st=False) branch: loss_single per decoder layer,
        the last layer under the plain names, earlier ones as d{i}.loss_*."""
        gt_masks_list = [g if g.is_floating_point() else g.float() for g in gt_masks_list]     # once, not per decoder layer
        per_layer = [self.loss_single(c, m, gt_labels_list, gt_masks_list, img_metas)
                     for c, m in zip(all_cls_scores, all_mask_preds)]
        out = dict(loss_cls=per_layer[-1][0], loss_mask=per_layer[-1][1], loss_dice=per_layer[-1][2])
        for i, (lc, lm, ld) in enumerate(per_layer[:-1]):
            out[f'd{i}.loss_cls'], out[f'd{i}.loss_mask'], out[f'd{i}.loss_dice'] = lc, lm, ld
        return out

    # ---- per-layer prediction heads (mask2former_head.py:355-395 / video :337-359) ----
    @torch.no_grad()
    def _embeds(self, query):
        """query [B,Q,C] -> (cls_pred [B,Q,NC+1], mask_embed [B,Q,C])."""
        pn = self.transformer_decoder.post_norm
        x, xs = ops.layernorm(query, pn.weight, pn.bias, pn.eps, out_split=True)
        xin = xs if xs is not None else x
        cls_pred = ops.linear(xin, self.cls_embed.weight, self.cls_embed.bias)
        # the 3-layer mask-embed MLP and the mask-logit contraction chain on operand planes
        me = ops.linear(xin, self.mask_embed[0].weight, self.mask_embed[0].bias, act=ops.ACT_RELU, out_mode='split')
        me = ops.linear(me, self.mask_embed[2].weight, self.mask_embed[2].bias, act=ops.ACT_RELU, out_mode='split')
        me = ops.linear(me, self.mask_embed[4].weight, self.mask_embed[4].bias, out_mode='split')
        return cls_pred, me

    @torch.no_grad()
    def _decoder_pe(self, shape_thw, device):
        key = (shape_thw, str(device))
        if key not in self._pe_cache:
            t, h, w = shape_thw
            self._pe_cache[key] = self.decoder_positional_encoding.tokens(h, w, device, t=t if self.video else 0)
        return self._pe_cache[key]

    def _batched(self, t, B):
        """t broadcast over the batch, materialised once per (tensor, batch)."""
        if B == 1:
            return t[None]
        key = ('b', t.data_ptr(), t._version, tuple(t.shape), B)   # _version: parameters reloaded in place
        hit = self._pe_cache.get(key)
        if hit is None:
            hit = t[None].expand(B, -1, -1).contiguous()
            if not (t.is_cuda and torch.cuda.is_current_stream_capturing()):
                self._pe_cache[key] = hit
        return hit

    @torch.no_grad()
    def _run(self, feats, num_frames, want_all, force_masks=None):
        """Core forward on token-major tensors.

        Returns dict(cls [B,Q,NC+1] list, mask_lr [B,T,Q,h,w] list (all layers only when
        want_all), query [B,Q,C]).  Work elision (SURVEY.md 7 "exact-parity work elision"):
        intermediate layers only need the sign of the bilinearly down-sampled logits, which is
        the sign of embed . (down-sampled mask features), so the full-resolution contraction
        runs for the last layer only unless ``want_all``.
        """
        mask_features, memories = self.pixel_decoder(feats)
        BT = mask_features.shape[0]
        T = num_frames
        B = BT // T
        assert B * T == BT  # mask2former_video_head.py:384
        mf = _tokens(mask_features)                       # [BT,h4,w4,C]
        _, h4, w4, C = mf.shape
        mf_flat = mf.view(B, T * h4 * w4, C)
        lvl_shapes = [tuple(m.shape[-2:]) for m in memories]
        # attention-mask features: mask features resized to each decoder level (linear, so it
        # commutes with the contraction -- see include/pvsg.h pvsg_mask_logits)
        pooled = [ops.bilinear_resize_nhwc(mf, s).view(B, -1, C) for s in lvl_shapes]
        # operand planes of the (re-used) mask features for the tcgen05 engine, split once per frame
        mf_planes = ops.recall_split(mf)    # emitted by the mask-feature conv
        mf_planes = mf_planes.view(B, T * h4 * w4, C) if mf_planes is not None else ops.maybe_split(mf_flat)
        pooled_planes = [ops.maybe_split(pl) for pl in pooled]
        dec_in, dec_pe = [], []
        for i, m in enumerate(memories):
            tok = _tokens(m)                              # [BT,h,w,C]
            h, w = lvl_shapes[i]
            dec_in.append(ops.add_rowvec(tok, self.level_embed.weight[i].contiguous()).view(B, T * h * w, C))
            pe = self._decoder_pe((T, h, w), tok.device)
            dec_pe.append(self._batched(pe, B))
        kin_planes = [ops.maybe_split(d, pe) for d, pe in zip(dec_in, dec_pe)]
        vin_planes = [ops.maybe_split(d) for d in dec_in]
        Q = self.num_queries
        query = self.query_feat.weight[None].expand(B, -1, -1).contiguous()
        qpos = self._batched(self.query_embed.weight, B)
        layers = self.transformer_decoder.layers
        nl = self.num_transformer_decoder_layers
        # K / V projections of every (layer, level) pair: independent of the queries
        cls_list, mask_list = [], []
        cls_pred, me = self._embeds(query)
        cls_list.append(cls_pred)
        if want_all:
            mask_list.append(ops.mask_logits(me, mf_flat, True, False, mf_planes)[0].view(B, Q, T, h4, w4).transpose(1, 2))
        _, mask, row_open = ops.mask_logits(me, pooled[0], False, True, pooled_planes[0])
        if self._capture_masks is not None:
            self._capture_masks.append(mask)
        if force_masks is not None:  # tests: teacher-force the discrete masks (see tests/test_models_gpu.py)
            mask, row_open = self._forced(force_masks[0])
        q_planes = None
        for i in range(nl):
            lvl = i % self.num_transformer_feat_level
            kproj, vproj = layers[i].attentions[0].project_kv(dec_in[lvl], dec_pe[lvl], dec_in[lvl],
                                                              kin_planes[lvl], vin_planes[lvl])
            query, q_planes = layers[i].forward_tokens(query, qpos, kproj, vproj, mask, row_open, q_planes)
            cls_pred, me = self._embeds(query)
            cls_list.append(cls_pred)
            last = i == nl - 1
            if want_all or last:
                mask_list.append(ops.mask_logits(me, mf_flat, True, False, mf_planes)[0].view(B, Q, T, h4, w4).transpose(1, 2))
            if not last:
                nxt = (i + 1) % self.num_transformer_feat_level
                _, mask, row_open = ops.mask_logits(me, pooled[nxt], False, True, pooled_planes[nxt])
                if self._capture_masks is not None:
                    self._capture_masks.append(mask)
                if force_masks is not None:
                    mask, row_open = self._forced(force_masks[i + 1])
        return dict(cls=cls_list, masks=mask_list, query=query)

    @staticmethod
    def _forced(mask):
        mask = mask.to(torch.uint8).contiguous()
        return mask, (mask == 0).sum(-1).to(torch.int32).contiguous()

    @torch.no_grad()
    def _upsample(self, mask_lr, size):
        """[N,Q,h,w] low-res logits -> [N,Q,H,W] (F.interpolate bilinear, align_corners=False)."""
        N, Q, h, w = mask_lr.shape
        # a [N*Q, h, w] plane stack is token-major with C = 1: resize through the C=4 kernel by
        # moving Q into the channel axis
        t = ops.nchw_to_nhwc(mask_lr.contiguous())                    # [N,h,w,Q]
        up = ops.bilinear_resize_nhwc(t, size)                        # Q = 100 -> multiple of 4
        return ops.nhwc_to_nchw(up)


@HEADS.register_module()
class Mask2FormerHeadCustom(_Mask2FormerHeadBase):
    """models/mask2former/mask2former_head.py:20 (image panoptic head)."""
    video = False

    @torch.no_grad()
    def forward(self, feats, img_metas, return_query=False):
        """mask2former_head.py:397-479: returns (cls_pred_list, mask_pred_list[, query_feat])."""
        r = self._run(feats, 1, want_all=True)
        masks = [m[:, 0] for m in r['masks']]
        qf = r['query'].transpose(0, 1)  # [Q,B,C]
        return (r['cls'], masks, qf) if return_query else (r['cls'], masks)

    @torch.no_grad()
    def simple_test_with_query(self, feats, img_metas, upsample=True, **kwargs):
        """mask2former_head.py:650-681.  upsample=False (used by the detector's fused path)
        returns the low-resolution logits instead of the 4x bilinear upsample."""
        r = self._run(feats, 1, want_all=False)
        mask = r['masks'][-1][:, 0]
        if upsample:
            mask = self._upsample(mask, tuple(img_metas[0]['batch_input_shape']))
        return r['cls'][-1], mask, r['query'].transpose(0, 1).unsqueeze(0)

    def simple_test(self, feats, img_metas, **kwargs):
        c, m, _ = self.simple_test_with_query(feats, img_metas, **kwargs)
        return c, m


@HEADS.register_module()
class Mask2FormerVideoHead(_Mask2FormerHeadBase):
    """models/mask2former_vps/mask2former_video_head.py:20 (video head; (b, t) batch axis)."""
    video = True

    @torch.no_grad()
    def forward(self, feats, img_metas, return_query=False):
        """mask2former_video_head.py:361-462; img_metas = list over batch of lists over frames."""
        T = len(img_metas[0])
        r = self._run(feats, T, want_all=True)
        qf = r['query'].transpose(0, 1)
        return (r['cls'], r['masks'], qf) if return_query else (r['cls'], r['masks'])

    @torch.no_grad()
    def simple_test_with_query(self, feats, img_metas, upsample=True, **kwargs):
        """mask2former_video_head.py:637-669 -> (cls [B,Q,NC+1], masks [B,T,Q,H,W], query [Q,B,C])."""
        T = len(img_metas[0])
        r = self._run(feats, T, want_all=False)
        mask = r['masks'][-1]
        if upsample:
            B = mask.shape[0]
            mask = self._upsample(mask.flatten(0, 1), tuple(img_metas[0][0]['batch_input_shape']))
            mask = mask.unflatten(0, (B, T))
        return r['cls'][-1], mask, r['query'].transpose(0, 1)


# ======================================================================================
# fusion head (models/mask2former/mask2former_fusion_head.py)
# ======================================================================================
@HEADS.register_module()
class MaskFormerFusionHeadCustom(nn.Module):
    def __init__(self, num_things_classes=80, num_stuff_classes=53, test_cfg=None, loss_panoptic=None,
                 init_cfg=None, **kwargs):
        super().__init__()
        self.num_things_classes = num_things_classes
        self.num_stuff_classes = num_stuff_classes
        self.num_classes = num_things_classes + num_stuff_classes
        self.test_cfg = dict(test_cfg or {})

    def forward_train(self, **kwargs):
        return dict()

    @torch.no_grad()
    def _panoptic(self, mask_cls, mask_lr, in_hw, img_hw, out_hw):
        return ops.panoptic_fuse(mask_cls, mask_lr, in_hw, img_hw, out_hw, self.num_things_classes, self.num_classes,
                                 float(self.test_cfg.get('object_mask_thr', 0.8)),
                                 float(self.test_cfg.get('iou_thr', 0.8)),
                                 bool(self.test_cfg.get('filter_low_score', False)), INSTANCE_OFFSET)

    @staticmethod
    def _query_dict(seg_info, query_feat):
        """seg_info (host int32 array) -> {seg_id: [feat]} in the reference's insertion order."""
        n = int(seg_info[0])
        rows = seg_info[1:1 + 4 * n].reshape(n, 4)
        d = defaultdict(list)
        for q, _cls, seg, _area in rows:
            if seg >= 0:
                d[int(seg)].append(query_feat[int(q)])
        return d

    @torch.no_grad()
    def panoptic_postprocess_with_query(self, mask_cls, mask_pred, query_feats):
        """mask2former_fusion_head.py:96-171 on full-resolution logits [Q,H,W]."""
        H, W = mask_pred.shape[-2:]
        pan, info = self._panoptic(mask_cls, mask_pred, (H, W), (H, W), (H, W))
        return pan, self._query_dict(info.cpu().numpy(), query_feats)

    @torch.no_grad()
    def instance_postprocess(self, mask_cls, mask_lr, in_hw=None, img_hw=None, out_hw=None, want_masks=True):
        """mask2former_fusion_head.py:192-242.  Class-score top-k runs on the tiny [Q, NC]
        score matrix on the host side of the stream (torch.topk, 12.6k values); all per-pixel
        work (binary masks, mask scores, boxes) is pvsg_instance_masks."""
        H, W = mask_lr.shape[-2:]
        in_hw, img_hw, out_hw = in_hw or (H, W), img_hw or (H, W), out_hw or (H, W)
        return self._instance_finish(self._instance_device(mask_cls, mask_lr, in_hw, img_hw, out_hw, want_masks))

    @torch.no_grad()
    def _instance_device(self, mask_cls, mask_lr, in_hw, img_hw, out_hw, want_masks=True):
        """Static-shape device part (CUDA-graph capturable): all ``max_per_image`` candidates are
        evaluated, stuff candidates are dropped afterwards by ``_instance_finish``."""
        max_per_image = min(self.test_cfg.get('max_per_image', 100), mask_cls.shape[0] * self.num_classes)
        scores_per_image, labels_per_image, query_indices = ops.instance_select(mask_cls, max_per_image)
        stats, boxes, masks = ops.instance_masks(mask_lr, query_indices, in_hw, img_hw, out_hw, want_masks)
        return dict(scores=scores_per_image, labels=labels_per_image.long(), stats=stats, boxes=boxes, masks=masks,
                    query=query_indices, labels32=labels_per_image)

    @torch.no_grad()
    def _instance_finish(self, d):
        is_thing = d['labels'] < self.num_things_classes
        stats = d['stats'][is_thing]
        mask_scores = stats[:, 0] / (stats[:, 1] + 1e-6)
        det_scores = d['scores'][is_thing] * mask_scores
        bboxes = torch.cat([d['boxes'][is_thing].float(), det_scores[:, None]], dim=-1)
        masks = d['masks'][is_thing].bool() if d['masks'] is not None else None
        return d['labels'][is_thing], bboxes, masks

    @torch.no_grad()
    def simple_test_with_query(self, mask_cls_results, mask_pred_results, query_feats, img_metas, rescale=False,
                               lowres=False, **kwargs):
        """mask2former_fusion_head.py:325-404.  With ``lowres=True`` mask_pred_results are the
        un-upsampled logits and the x4 upsample is fused into the post-processing kernels."""
        panoptic_on = self.test_cfg.get('panoptic_on', True)
        semantic_on = self.test_cfg.get('semantic_on', False)
        instance_on = self.test_cfg.get('instance_on', False)
        assert not semantic_on, 'segmantic segmentation results are not supported yet.'
        results = []
        for mask_cls, mask_pred, qf, meta in zip(mask_cls_results, mask_pred_results, query_feats, img_metas):
            img_hw = tuple(meta['img_shape'][:2])
            in_hw = tuple(meta['batch_input_shape']) if lowres else tuple(mask_pred.shape[-2:])
            out_hw = tuple(meta['ori_shape'][:2]) if rescale else img_hw
            result = dict()
            if panoptic_on:
                pan, info = self._panoptic(mask_cls, mask_pred, in_hw, img_hw, out_hw)
                result['pan_results'] = pan
                result['query_feats'] = self._query_dict(info.cpu().numpy(), qf)
            if instance_on:
                result['ins_results'] = self.instance_postprocess(mask_cls, mask_pred, in_hw, img_hw, out_hw)
            results.append(result)
        return results


HEADS.register_module(name='MaskFormerFusionHead', module=MaskFormerFusionHeadCustom)


def bbox2result(bboxes, labels, num_classes):
    """mmdet.core.bbox2result (L0)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, bboxes.shape[1]), dtype=np.float32) for _ in range(num_classes)]
    bboxes = bboxes.detach().cpu().numpy()
    labels = labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


# ======================================================================================
# detectors
# ======================================================================================
class _DetectorBase(nn.Module):
    def __init__(self, backbone, neck=None, panoptic_head=None, panoptic_fusion_head=None, train_cfg=None,
                 test_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        if neck is not None:
            raise NotImplementedError('neck')
        self.backbone = build_backbone(backbone)
        ph = copy.deepcopy(to_cfg(panoptic_head))
        ph.update(train_cfg=train_cfg, test_cfg=test_cfg)
        self.panoptic_head = build_head(ph)
        pf = copy.deepcopy(to_cfg(panoptic_fusion_head))
        pf.update(test_cfg=test_cfg)
        self.panoptic_fusion_head = build_head(pf)
        self.num_things_classes = self.panoptic_head.num_things_classes
        self.num_stuff_classes = self.panoptic_head.num_stuff_classes
        self.num_classes = self.panoptic_head.num_classes
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.eval()

    def train(self, mode=True):
        """Training mode only changes the flag: every normalisation layer of the model is batch-independent (LayerNorm,
        GroupNorm, eval-mode BatchNorm folded into the convolutions: norm_eval=True) and all dropout rates are 0."""
        return super().train(mode)

    def init_weights(self):
        """mmcv BaseModule.init_weights: pretrained weights come from ``load_state_dict`` / ``load_checkpoint`` here."""

    # Captured CUDA graphs (engine.FrameRunner) hold raw pointers to kernel-layout copies of the weights:
    # reloading, moving or casting the detector starts a new weights epoch and drops them (engine.get_runner).
    def _load_from_state_dict(self, *args, **kwargs):
        self._weights_gen = getattr(self, '_weights_gen', 0) + 1
        if getattr(self, '_runners', None):
            self._runners.clear()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **k):
        self._weights_gen = getattr(self, '_weights_gen', 0) + 1
        if getattr(self, '_runners', None):
            self._runners.clear()
        return super()._apply(fn, *a, **k)

    def extract_feat(self, img):
        return self.backbone(img)

    def forward_train(self, *a, **k):
        raise NotImplementedError('forward_train: Mask2FormerCustom (image) and Mask2FormerVideoCustom (video) implement it')

    def _parse_losses(self, losses):
        """mmdet BaseDetector._parse_losses (single process: no all_reduce): -> (total loss, log_vars)."""
        log_vars = {}
        for name, value in losses.items():
            log_vars[name] = value.mean() if torch.is_tensor(value) else sum(v.mean() for v in value)
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        return loss, {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in log_vars.items()}

    def train_step(self, data, optimizer=None):
        """mmdet BaseDetector.train_step: the runner's entry point (``tools/train.py`` -> ``EpochBasedRunner``)."""
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))

    def aug_test(self, imgs, img_metas, **kwargs):
        raise NotImplementedError  # as the reference: mask2former.py:193-194

    def forward(self, img=None, img_metas=None, return_loss=True, **kwargs):
        """mmdet BaseDetector.forward (return_loss defaults to True there: train_step calls self(**data))."""
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)


@DETECTORS.register_module()
class Mask2FormerCustom(_DetectorBase):
    """models/mask2former/mask2former.py:14 (image panoptic segmentation)."""
    train_backbone = True

    def forward_train(self, img, img_metas, gt_bboxes=None, gt_labels=None, gt_masks=None, gt_semantic_seg=None,
                      gt_bboxes_ignore=None, **kwargs):
        """models/mask2former/mask2former.py:75-116."""
        for m in img_metas:
            m['batch_input_shape'] = tuple(img.shape[-2:])
        if self.train_backbone and hasattr(self.backbone, 'forward_train'):
            x = self.backbone.forward_train(img)
        else:
            with torch.no_grad():
                x = self.extract_feat(img)
        return self.panoptic_head.forward_train(x, img_metas, gt_bboxes, gt_labels, gt_masks, gt_semantic_seg,
                                                gt_bboxes_ignore=gt_bboxes_ignore)

    def forward_test(self, imgs, img_metas, **kwargs):
        """mmdet BaseDetector.forward_test (single augmentation): adds batch_input_shape."""
        img, metas = (imgs[0], img_metas[0]) if isinstance(imgs, (list, tuple)) else (imgs, img_metas)
        for m in metas:
            m['batch_input_shape'] = tuple(img.shape[-2:])
        return self.simple_test(img, metas, **kwargs)

    @torch.no_grad()
    def simple_test(self, imgs, img_metas, **kwargs):
        """mask2former.py:121-191."""
        feats = self.extract_feat(imgs)
        mask_cls, mask_lr, query_feats = self.panoptic_head.simple_test_with_query(feats, img_metas, upsample=False)
        results = self.panoptic_fusion_head.simple_test_with_query(mask_cls, mask_lr, query_feats, img_metas,
                                                                   lowres=True, **kwargs)
        for r in results:
            if 'pan_results' in r:
                r['pan_results'] = r['pan_results'].cpu().numpy()
            if 'query_feats' in r:
                r['query_feats'] = {k: [x.cpu().numpy() for x in v] for k, v in r['query_feats'].items()}
            if 'ins_results' in r:
                labels, bboxes, masks = r['ins_results']
                bbox_results = bbox2result(bboxes, labels, self.num_things_classes)
                mask_results = [[] for _ in range(self.num_things_classes)]
                masks_np = masks.cpu().numpy()
                for j, label in enumerate(labels.tolist()):
                    mask_results[label].append(masks_np[j])
                r['ins_results'] = bbox_results, mask_results
        if self.num_stuff_classes == 0:
            results = [res['ins_results'] for res in results]
        return results


@DETECTORS.register_module()
class Mask2FormerVideoCustom(_DetectorBase):
    """models/mask2former_vps/mask2former.py:33 (video panoptic segmentation)."""

    def __init__(self, *args, dataset='kitti-step', **kwargs):
        super().__init__(*args, **kwargs)
        self.dataset = dataset
        self.train_backbone = True

    def forward_train(self, img, img_metas, gt_bboxes=None, gt_labels=None, gt_masks=None, gt_semantic_seg=None,
                      gt_bboxes_ignore=None, *, ref_img=None, ref_img_metas=None, ref_gt_bboxes=None, ref_gt_labels=None,
                      ref_gt_bboxes_ignore=None, ref_gt_masks=None, ref_gt_semantic_seg=None, ref_gt_instance_ids=None,
                      **kwargs):
        """models/mask2former_vps/mask2former.py:85-123.  ``self.train_backbone`` (default True) puts the ResNet on the
        autograd tape too (BatchNorm frozen as the reference config has it); a Swin backbone runs frozen."""
        bs, num_frame, three, h, w = ref_img.size()
        for metas in ref_img_metas:
            for m in metas:
                m['batch_input_shape'] = (h, w)
        frames = ref_img.reshape(bs * num_frame, three, h, w)
        if self.train_backbone and hasattr(self.backbone, 'forward_train'):
            video_x = self.backbone.forward_train(frames)
        else:
            with torch.no_grad():
                video_x = self.extract_feat(frames)
        return self.panoptic_head.forward_train(video_x, ref_img_metas, ref_gt_bboxes, ref_gt_labels, ref_gt_masks,
                                                ref_gt_semantic_seg, ref_gt_instance_ids, gt_bboxes_ignore=None)

    def forward_test(self, imgs, img_metas, **kwargs):
        """mask2former.py:225-240."""
        for img, img_meta in zip(imgs, img_metas):
            for m in img_meta:
                m['batch_input_shape'] = tuple(img.size()[-2:])
        ref = kwargs['ref_img'][0] if isinstance(kwargs['ref_img'], (list, tuple)) else kwargs['ref_img']
        for sample_metas in kwargs['ref_img_metas']:          # [batch][frame] dicts, as the reference indexes them
            for frame_meta in sample_metas:
                frame_meta['batch_input_shape'] = tuple(ref.size()[-2:])
        kwargs['ref_img'] = ref
        return self.simple_test(img=imgs, img_metas=img_metas, **kwargs)

    @torch.no_grad()
    def simple_test(self, img, img_metas, ref_img, ref_img_metas, **kwargs):
        """mask2former.py:125-223 for the shipped configuration (clip length 1 per sample)."""
        bs, num_frame, three, h, w = ref_img.size()
        if num_frame != 1:
            # frames >= 2 call self.match_from_embds, which the reference class does not define
            # (mask2former.py:155) -- same failure mode here.
            raise AttributeError("'Mask2FormerVideoCustom' object has no attribute 'match_from_embds'")
        if getattr(self, '_runners', None) is not None:
            # CUDA-graph replay of the same kernels (openpvsg_b200/engine.py); a batch of samples
            # (samples_per_gpu > 1, all of one shape) goes through one batched replay
            from .engine import get_runner
            metas = [m[0] for m in ref_img_metas]
            key = lambda d: (tuple(d['batch_input_shape']), tuple(d['img_shape']), tuple(d['ori_shape']))  # noqa: E731
            if bs == 1:
                runner = get_runner(self, metas[0], kwargs.get('rescale', False))
                return [[runner.run(ref_img)]]
            if all(key(d) == key(metas[0]) for d in metas):
                # one synchronous call, pipelined inside: the batch goes through the runner in SYNC_CHUNKS pieces, so the
                # host->device copy of piece i+1 and the device->host copy + result building of piece i-1 overlap the
                # graph replay of piece i (the call still returns only when every result is on the host)
                from . import engine
                chunks = engine.SYNC_CHUNKS if bs >= 4 * engine.SYNC_CHUNKS and bs % engine.SYNC_CHUNKS == 0 else 1
                per = bs // chunks
                runner = get_runner(self, metas[0], kwargs.get('rescale', False), batch=per)
                out, queue = [], []
                for c in range(chunks):
                    queue.append(runner.submit([ref_img[i, 0] for i in range(c * per, (c + 1) * per)]))
                    if len(queue) == 2:
                        out += runner.collect(queue.pop(0))
                while queue:
                    out += runner.collect(queue.pop(0))
                return [[r] for r in out]
        video_x = self.extract_feat(ref_img.reshape(bs * num_frame, three, h, w))
        results = [[] for _ in range(bs)]
        for i in range(bs):
            feats = [f[i:i + 1] for f in video_x]
            mask_cls, mask_lr, query = self.panoptic_head.simple_test_with_query(feats, [ref_img_metas[i]],
                                                                                 upsample=False)
            out_logits = mask_cls                      # [1,Q,NC+1]
            out_embds = query.permute(1, 0, 2)         # [1,Q,C]
            for frame_id in range(num_frame):
                res = self.panoptic_fusion_head.simple_test_with_query(
                    out_logits, mask_lr[:, frame_id], out_embds, [ref_img_metas[i][frame_id]], lowres=True,
                    **kwargs)[0]
                res['pan_results'] = res['pan_results'].cpu().numpy()
                if 'ins_results' in res:
                    labels, bboxes, masks = res['ins_results']
                    ids = torch.arange(len(bboxes), dtype=bboxes.dtype, device=bboxes.device)[:, None] + 1
                    bboxes = torch.cat([ids, bboxes], dim=1)
                    inds = torch.argsort(bboxes[:, -1], descending=True)[:10]
                    labels, bboxes, masks = labels[inds], bboxes[inds], masks[inds]
                    bbox_results = bbox2result(bboxes, labels, self.num_things_classes)
                    mask_results = [[] for _ in range(self.num_things_classes)]
                    masks_np = masks.cpu().numpy()
                    for j, label in enumerate(labels.tolist()):
                        mask_results[label].append(masks_np[j])
                    res['ins_results'] = bbox_results, mask_results
                results[i].append(res)
        return results


def _fusion_simple_test(self, mask_cls_results, mask_pred_results, img_metas, rescale=False, lowres=False, **kwargs):
    """MaskFormerFusionHeadCustom.simple_test (mask2former_fusion_head.py:244-321): the
    query-less form used by the MinVIS detector."""
    dummy = [None] * len(img_metas)
    out = []
    for mask_cls, mask_pred, _, meta in zip(mask_cls_results, mask_pred_results, dummy, img_metas):
        img_hw = tuple(meta['img_shape'][:2])
        in_hw = tuple(meta['batch_input_shape']) if lowres else tuple(mask_pred.shape[-2:])
        out_hw = tuple(meta['ori_shape'][:2]) if rescale else img_hw
        result = dict()
        if self.test_cfg.get('panoptic_on', True):
            result['pan_results'], _info = self._panoptic(mask_cls, mask_pred, in_hw, img_hw, out_hw)
        if self.test_cfg.get('instance_on', False):
            result['ins_results'] = self.instance_postprocess(mask_cls, mask_pred, in_hw, img_hw, out_hw)
        out.append(result)
    return out


MaskFormerFusionHeadCustom.simple_test = _fusion_simple_test


@DETECTORS.register_module()
class Mask2FormerVideoCustomMinVIS(Mask2FormerVideoCustom):
    """models/mask2former_vps/mask2former_min_vis.py:35 -- per-frame inference, MinVIS query
    matching between consecutive frames, clip-averaged class logits, per-frame fusion."""

    @torch.no_grad()
    def match_from_embds(self, tgt_embds, cur_embds):
        """mask2former_min_vis.py:244-258: cosine cost and exact linear assignment (rows: target queries, columns:
        current queries), both on the device; returns for every target row the matched current query."""
        sigma = ops.minvis_chain(torch.stack([tgt_embds, cur_embds]))
        return sigma[0].long()

    @torch.no_grad()
    def link_clip(self, query_list):
        """The whole clip's matching at once.  The reference matches frame t against the RE-ORDERED queries
        of frame t-1 (mask2former_min_vis.py:176-181); re-ordering the target only permutes the rows of the cost
        matrix, so the T-1 assignment problems are independent on the raw embeddings: they are solved concurrently
        (one warp each) and the clip's orderings are the running composition of the per-pair matchings."""
        embeds = torch.stack(list(query_list))                  # [T, Q, C]
        T, Q = embeds.shape[:2]
        sigma = ops.minvis_chain(embeds) if T > 1 else torch.empty(0, Q, device=embeds.device, dtype=torch.int32)
        return ops.perm_chain(sigma, Q).long()                  # [T, Q]

    @torch.no_grad()
    def simple_test(self, img, img_metas, ref_img, ref_img_metas, **kwargs):
        """mask2former_min_vis.py:132-231 (batch size 1, as the reference's squeeze() implies)."""
        bs, num_frame, three, h, w = ref_img.size()
        assert bs == 1, 'MinVIS inference handles one clip at a time'
        video_x = self.extract_feat(ref_img.reshape(bs * num_frame, three, h, w))
        pred_logits, mask_lr_list, query_list = [], [], []
        for i in range(num_frame):
            feats = [f[i:i + 1] for f in video_x]
            cls, mask_lr, query = self.panoptic_head.simple_test_with_query(feats, [[ref_img_metas[0][i]]],
                                                                            upsample=False)
            pred_logits.append(cls[0])
            mask_lr_list.append(mask_lr[0, 0])
            query_list.append(query[:, 0])
        perms = self.link_clip(query_list)
        out_logits = [pred_logits[i][perms[i], :] for i in range(num_frame)]
        out_masks = [mask_lr_list[i][perms[i], :, :] for i in range(num_frame)]
        logits = (sum(out_logits) / len(out_logits)).unsqueeze(0)
        results = [[]]
        for frame_id in range(num_frame):
            res = self.panoptic_fusion_head.simple_test(logits, out_masks[frame_id][None],
                                                        [ref_img_metas[0][frame_id]], lowres=True, **kwargs)[0]
            res['pan_results'] = res['pan_results'].cpu().numpy()
            if 'ins_results' in res:
                labels, bboxes, masks = res['ins_results']
                ids = torch.arange(len(bboxes), dtype=bboxes.dtype, device=bboxes.device)[:, None] + 1
                bboxes = torch.cat([ids, bboxes], dim=1)
                inds = torch.argsort(bboxes[:, -1], descending=True)[:10]
                labels, bboxes, masks = labels[inds], bboxes[inds], masks[inds]
                mask_results = [[] for _ in range(self.num_things_classes)]
                masks_np = masks.cpu().numpy()
                for j, label in enumerate(labels.tolist()):
                    mask_results[label].append(masks_np[j])
                res['ins_results'] = bbox2result(bboxes, labels, self.num_things_classes), mask_results
            results[0].append(res)
        return results
