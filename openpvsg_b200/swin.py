"""Swin Transformer backbone (BASELINE configs[2] "Mask2Former-VPS Swin-B").

The reference repo ships no Swin code: its configs build backbones through mmdet's ``BACKBONES`` registry, and
the Swin-B Mask2Former configuration is mmdet 2.25.0's (``configs/mask2former/mask2former_swin-b-p4-w12-384_lsj_
8x2_50e_coco-panoptic.py``: embed 128, depths 2-2-18-2, heads 4-8-16-32, window 12, head in_channels
[128,256,512,1024]).  This module mirrors ``mmdet/models/backbones/swin.py`` ``SwinTransformer``: same registry
name, constructor arguments, ``forward(img) -> tuple of [B,C_i,H_i,W_i]`` and ``state_dict`` keys
(``patch_embed.projection|norm``, ``stages.{i}.blocks.{j}.{norm1,attn.w_msa.{relative_position_bias_table,
relative_position_index,qkv,proj},norm2,ffn.layers.{0.0,1}}``, ``stages.{i}.downsample.{norm,reduction}``,
``norm{i}``) so an mmdet checkpoint loads with ``strict=True``.

Device path per block (7 launches): LayerNorm (+ operand planes) -> qkv GEMM (tcgen05) -> ``pvsg_window_attention``
(pad / roll / partition / relative-position bias / shift mask / reverse / crop folded into addressing, result emitted as operand planes) -> proj GEMM
(+ residual) -> LayerNorm (+ planes) -> fc1 GEMM with exact-GELU epilogue emitting planes -> fc2 GEMM (+ residual).
Patch merging = ``pvsg_patch_merge_ln`` + reduction GEMM.  Tokens stay token-major [B,H,W,C] throughout.
"""
import torch
import torch.nn as nn

from . import ops
from .mask2former import _Prepared, _as_nchw, _tokens
from .registry import BACKBONES


class _WindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, qkv_bias=True):
        super().__init__()
        ws = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        seq = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)[:, None] + torch.arange(ws)[None, :]
        coords = seq.reshape(1, -1)
        self.register_buffer('relative_position_index', (coords + coords.T).flip(1).contiguous())
        self.qkv = nn.Linear(embed_dims, embed_dims * 3, bias=qkv_bias)
        self.proj = nn.Linear(embed_dims, embed_dims)


class _ShiftWindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, shift_size, qkv_bias=True):
        super().__init__()
        self.window_size, self.shift_size, self.num_heads = window_size, shift_size, num_heads
        self.w_msa = _WindowMSA(embed_dims, num_heads, window_size, qkv_bias)


class _FFN(nn.Module):
    """mmcv FFN key layout: layers.0.0 = fc1, layers.1 = fc2."""

    def __init__(self, embed_dims, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, hidden), nn.GELU(), nn.Dropout(0.0)),
                                    nn.Linear(hidden, embed_dims), nn.Dropout(0.0))


class _SwinBlock(nn.Module):
    def __init__(self, embed_dims, num_heads, hidden, window_size, shift, qkv_bias=True):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dims)
        self.attn = _ShiftWindowMSA(embed_dims, num_heads, window_size, window_size // 2 if shift else 0, qkv_bias)
        self.norm2 = nn.LayerNorm(embed_dims)
        self.ffn = _FFN(embed_dims, hidden)

    @torch.no_grad()
    def forward_tokens(self, x):
        """x fp32 [B,H,W,C] -> same (mmdet SwinBlock.forward :329-356)."""
        a = self.attn
        w = a.w_msa
        y, yp = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, out_split=True)
        qkv = ops.linear(yp if yp is not None else y, w.qkv.weight, w.qkv.bias)
        att = ops.window_attention(qkv, w.qkv.bias, w.relative_position_bias_table, a.num_heads, a.window_size, a.shift_size,
                                   out_mode='split')
        x = ops.linear(att, w.proj.weight, w.proj.bias, residual=x)
        y, yp = ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, out_split=True)
        fc1, fc2 = self.ffn.layers[0][0], self.ffn.layers[1]
        h = ops.linear(yp if yp is not None else y, fc1.weight, fc1.bias, act=ops.ACT_GELU, out_mode='split')
        return ops.linear(h, fc2.weight, fc2.bias, residual=x)


class _PatchMerging(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.norm = nn.LayerNorm(4 * in_channels)
        self.reduction = nn.Linear(4 * in_channels, out_channels, bias=False)

    @torch.no_grad()
    def forward_tokens(self, x):
        return ops.linear(ops.patch_merge_ln(x, self.norm.weight, self.norm.bias, self.norm.eps), self.reduction.weight)


class _PatchEmbed(nn.Module):
    def __init__(self, in_channels, embed_dims, patch_size, norm):
        super().__init__()
        self.projection = nn.Conv2d(in_channels, embed_dims, patch_size, patch_size)
        self.norm = nn.LayerNorm(embed_dims) if norm else None


class _SwinBlockSequence(nn.Module):
    def __init__(self, embed_dims, num_heads, hidden, depth, window_size, qkv_bias, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([_SwinBlock(embed_dims, num_heads, hidden, window_size, j % 2 == 1, qkv_bias)
                                     for j in range(depth)])
        self.downsample = downsample


@BACKBONES.register_module()
class SwinTransformer(_Prepared):
    """mmdet 2.25 ``SwinTransformer`` (inference).  Swin-B for BASELINE configs[2]:
    ``embed_dims=128, depths=(2,2,18,2), num_heads=(4,8,16,32), window_size=12``."""

    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4, window_size=7, mlp_ratio=4,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3),
                 qkv_bias=True, qk_scale=None, patch_norm=True, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1,
                 use_abs_pos_embed=False, act_cfg=None, norm_cfg=None, with_cp=False, pretrained=None,
                 convert_weights=False, frozen_stages=-1, init_cfg=None):
        super().__init__()
        if use_abs_pos_embed or qk_scale is not None or tuple(strides) != (patch_size, 2, 2, 2)[:len(strides)]:
            raise NotImplementedError('SwinTransformer: absolute position embedding / qk_scale / custom strides')
        if (act_cfg or {}).get('type', 'GELU') != 'GELU' or (norm_cfg or {}).get('type', 'LN') != 'LN':
            raise NotImplementedError('SwinTransformer: GELU + LN only')
        if not qkv_bias:
            raise NotImplementedError('SwinTransformer: qkv_bias=False (the padded window positions read the qkv bias)')
        if any(embed_dims * 2 ** i != h * 32 for i, h in enumerate(num_heads)) or window_size > 12:
            raise NotImplementedError('SwinTransformer: head dim 32 and window <= 12 (every published Swin variant)')
        if any(embed_dims * 2 ** i not in (128, 256, 512, 1024) for i in range(len(depths))):
            raise NotImplementedError('SwinTransformer: stage widths must be in {128, 256, 512, 1024} (pvsg_layernorm); '
                                      'Swin-B (embed_dims=128) is the supported variant')
        self.out_indices = tuple(out_indices)
        self.patch_size = patch_size
        self.patch_embed = _PatchEmbed(in_channels, embed_dims, patch_size, patch_norm)
        self.stages = nn.ModuleList()
        self.num_features = [int(embed_dims * 2 ** i) for i in range(len(depths))]
        for i, depth in enumerate(depths):
            C = self.num_features[i]
            down = _PatchMerging(C, 2 * C) if i < len(depths) - 1 else None
            self.stages.append(_SwinBlockSequence(C, num_heads[i], int(mlp_ratio * C), depth, window_size, qkv_bias, down))
        for i in self.out_indices:
            self.add_module(f'norm{i}', nn.LayerNorm(self.num_features[i]))
        self.eval()

    def init_weights(self):
        pass

    @torch.no_grad()
    def forward(self, x):
        ops.clear_split_cache()
        P = self.patch_size
        if x.shape[-2] % P or x.shape[-1] % P:      # PatchEmbed adaptive 'corner' padding (layout glue)
            x = nn.functional.pad(x, (0, (P - x.shape[-1] % P) % P, 0, (P - x.shape[-2] % P) % P))
        if self._prep is None:
            self._prep = self.patch_embed.projection.weight.permute(0, 2, 3, 1).contiguous()   # [Cout,R,S,Cin]
        x = ops.conv2d_nhwc(_tokens(x.contiguous()), self._prep, self.patch_embed.projection.bias, stride=P)
        if self.patch_embed.norm is not None:
            n = self.patch_embed.norm
            x = ops.layernorm(x, n.weight, n.bias, n.eps)
        outs = []
        for i, stage in enumerate(self.stages):
            for blk in stage.blocks:
                x = blk.forward_tokens(x)
            if i in self.out_indices:
                n = getattr(self, f'norm{i}')
                o, op = ops.layernorm(x, n.weight, n.bias, n.eps, out_split=True)
                if op is not None:
                    ops.remember_split(o, op)       # the pixel decoder's 1x1 convs reuse these planes
                outs.append(_as_nchw(o))
            if stage.downsample is not None:
                x = stage.downsample.forward_tokens(x)
        return tuple(outs)
