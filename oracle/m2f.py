"""fp32 CPU oracle of the Mask2Former / Mask2Former-VPS inference forward.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Every function cites the
reference file:line it follows (paths relative to /root/reference).  Parts
that live in mmcv-full 1.4.0 / mmdet 2.25.0 (not vendored by the reference)
are restated from their published algorithm and say so ("L0").

All functions are functional: they take the mmdet-style ``state_dict`` (``sd``)
plus a key prefix, so the oracle shares no module code with the product.
"""
import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

INSTANCE_OFFSET = 1000  # mmdet/core/evaluation/panoptic_utils.py (L0)


# --------------------------------------------------------------------------
# backbone: mmdet ResNet(depth=50, style='pytorch', norm_eval=True)  (L0, A1)
# cfg: configs/mask2former_vps/mask2former_video_r50_base.py:7-16
# --------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], False, 0.0, 1e-5)


def resnet50(sd, x, prefix='backbone.'):
    """Returns (C2, C3, C4, C5); torchvision-resnet50 topology, stride on 3x3."""
    x = F.conv2d(x, sd[prefix + 'conv1.weight'], None, 2, 3)
    x = F.relu(_bn(sd, prefix + 'bn1', x))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nblk in enumerate((3, 4, 6, 3)):
        for b in range(nblk):
            p = f'{prefix}layer{li + 1}.{b}.'
            stride = 2 if (b == 0 and li > 0) else 1
            idt = x
            o = F.relu(_bn(sd, p + 'bn1', F.conv2d(x, sd[p + 'conv1.weight'])))
            o = F.relu(_bn(sd, p + 'bn2', F.conv2d(o, sd[p + 'conv2.weight'], None, stride, 1)))
            o = _bn(sd, p + 'bn3', F.conv2d(o, sd[p + 'conv3.weight']))
            if b == 0:
                idt = _bn(sd, p + 'downsample.1',
                          F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride))
            x = F.relu(o + idt)
        outs.append(x)
    return tuple(outs)


# --------------------------------------------------------------------------
# mmdet SinePositionalEncoding(num_feats=128, normalize=True)  (L0, A4)
# --------------------------------------------------------------------------
def sine_pe_2d(b, h, w, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    not_mask = torch.ones(b, h, w, dtype=torch.int)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).view(b, h, w, -1)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).view(b, h, w, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def sine_pe_3d(b, t, h, w, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """models/mask2former_vps/position_encoding.py:55-99 with an all-valid mask."""
    not_mask = torch.ones(b, t, h, w, dtype=torch.int)
    z_embed = not_mask.cumsum(1, dtype=torch.float32)
    y_embed = not_mask.cumsum(2, dtype=torch.float32)
    x_embed = not_mask.cumsum(3, dtype=torch.float32)
    z_embed = z_embed / (z_embed[:, -1:, :, :] + eps) * scale
    y_embed = y_embed / (y_embed[:, :, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    dim_t_z = torch.arange(num_feats * 2, dtype=torch.float32)
    dim_t_z = temperature ** (2 * (dim_t_z // 2) / (num_feats * 2))
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_z = z_embed[..., None] / dim_t_z
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=5).view(b, t, h, w, -1)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=5).view(b, t, h, w, -1)
    pos_z = torch.stack((pos_z[..., 0::2].sin(), pos_z[..., 1::2].cos()), dim=5).view(b, t, h, w, -1)
    return (torch.cat((pos_y, pos_x), dim=4) + pos_z).permute(0, 1, 4, 2, 3)


# --------------------------------------------------------------------------
# mmcv MultiScaleDeformableAttention  (L0, A3)
# cfg: configs/mask2former_vps/mask2former_video_r50_base.py:38-47
# --------------------------------------------------------------------------
def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """mmcv ``multi_scale_deformable_attn_pytorch`` (the CPU path of the op).

    value [B, N, H, D]; spatial_shapes list[(h, w)]; sampling_locations
    [B, Nq, H, L, P, 2] in [0, 1] (x, y); attention_weights [B, Nq, H, L, P].
    Returns [B, Nq, H*D].
    """
    bs, _, nh, d = value.shape
    _, nq, _, nl, npt, _ = sampling_locations.shape
    value_list = value.split([h * w for h, w in spatial_shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        v = value_list[lvl].flatten(2).transpose(1, 2).reshape(bs * nh, d, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)  # [B*H, Nq, P, 2]
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros',
                                     align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(bs * nh, 1, nq, nl * npt)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1)
    return out.view(bs, nh * d, nq).transpose(1, 2).contiguous()


def msda_module(sd, p, query, query_pos, reference_points, spatial_shapes,
                num_heads=8, num_levels=3, num_points=4):
    """mmcv MultiScaleDeformableAttention.forward, batch-first tensors [B, N, C].

    value = query (no pos); offsets / weights from query + query_pos;
    returns output_proj(core) + identity.
    """
    b, n, c = query.shape
    identity = query
    q = query + query_pos
    value = F.linear(query, sd[p + 'value_proj.weight'], sd[p + 'value_proj.bias'])
    value = value.view(b, n, num_heads, c // num_heads)
    off = F.linear(q, sd[p + 'sampling_offsets.weight'], sd[p + 'sampling_offsets.bias'])
    off = off.view(b, n, num_heads, num_levels, num_points, 2)
    aw = F.linear(q, sd[p + 'attention_weights.weight'], sd[p + 'attention_weights.bias'])
    aw = aw.view(b, n, num_heads, num_levels * num_points).softmax(-1)
    aw = aw.view(b, n, num_heads, num_levels, num_points)
    normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=torch.float32)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = msda_core(value, spatial_shapes, loc, aw)
    out = F.linear(out, sd[p + 'output_proj.weight'], sd[p + 'output_proj.bias'])
    return out + identity


# --------------------------------------------------------------------------
# mmdet MSDeformAttnPixelDecoder  (L0, A2)
# cfg: configs/mask2former_vps/mask2former_video_r50_base.py:27-59
# --------------------------------------------------------------------------
def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def pixel_decoder(sd, feats, prefix='panoptic_head.pixel_decoder.', num_layers=6,
                  return_intermediate=False):
    """feats = (C2, C3, C4, C5) -> (mask_feature [B,256,h4,w4], [m32, m16, m8])."""
    b = feats[0].shape[0]
    tokens, pos_list, shapes, refs = [], [], [], []
    for i in range(3):
        level_idx = 3 - i  # C5, C4, C3
        f = feats[level_idx]
        f = F.conv2d(f, sd[f'{prefix}input_convs.{i}.conv.weight'], sd[f'{prefix}input_convs.{i}.conv.bias'])
        f = _gn(sd, f'{prefix}input_convs.{i}.gn', f)
        h, w = f.shape[-2:]
        pos = sine_pe_2d(b, h, w)
        lvl = sd[prefix + 'level_encoding.weight'][i].view(1, -1, 1, 1)
        pos_list.append((lvl + pos).flatten(2).permute(0, 2, 1))
        tokens.append(f.flatten(2).permute(0, 2, 1))
        shapes.append((h, w))
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32),
                                torch.arange(w, dtype=torch.float32), indexing='ij')
        # MlvlPointGenerator offset 0.5, then / ([w, h] * stride)
        refs.append(torch.stack(((xs.flatten() + 0.5) / w, (ys.flatten() + 0.5) / h), -1))
    x = torch.cat(tokens, 1)
    pos = torch.cat(pos_list, 1)
    ref = torch.cat(refs, 0)[None, :, None, :].repeat(b, 1, 3, 1)
    inter = []
    for l in range(num_layers):
        p = f'{prefix}encoder.layers.{l}.'
        x = msda_module(sd, p + 'attentions.0.', x, pos, ref, shapes)
        x = _ln(sd, p + 'norms.0', x)
        y = F.relu(F.linear(x, sd[p + 'ffns.0.layers.0.0.weight'], sd[p + 'ffns.0.layers.0.0.bias']))
        x = x + F.linear(y, sd[p + 'ffns.0.layers.1.weight'], sd[p + 'ffns.0.layers.1.bias'])
        x = _ln(sd, p + 'norms.1', x)
        inter.append(x)
    outs = []
    start = 0
    for (h, w) in shapes:
        outs.append(x[:, start:start + h * w].permute(0, 2, 1).reshape(b, -1, h, w))
        start += h * w
    c2 = feats[0]
    cur = _gn(sd, prefix + 'lateral_convs.0.gn', F.conv2d(c2, sd[prefix + 'lateral_convs.0.conv.weight']))
    y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode='bilinear', align_corners=False)
    y = F.relu(_gn(sd, prefix + 'output_convs.0.gn',
                   F.conv2d(y, sd[prefix + 'output_convs.0.conv.weight'], None, 1, 1)))
    mask_feature = F.conv2d(y, sd[prefix + 'mask_feature.weight'], sd[prefix + 'mask_feature.bias'])
    if return_intermediate:
        return mask_feature, outs, inter
    return mask_feature, outs


# --------------------------------------------------------------------------
# mmdet DetrTransformerDecoderLayer + mmcv MultiheadAttention / FFN  (L0, A5)
# --------------------------------------------------------------------------
def mha(sd, p, query, key, value, query_pos, key_pos, attn_mask, num_heads=8):
    """mmcv MultiheadAttention.forward on seq-first tensors; returns identity + out."""
    identity = query
    q = query + query_pos
    k = key + key_pos
    out = F.multi_head_attention_forward(
        q, k, value, q.shape[-1], num_heads,
        sd[p + 'attn.in_proj_weight'], sd[p + 'attn.in_proj_bias'], None, None, False, 0.0,
        sd[p + 'attn.out_proj.weight'], sd[p + 'attn.out_proj.bias'],
        training=False, need_weights=False, attn_mask=attn_mask)[0]
    return identity + out


def decoder_layer(sd, p, query, key, value, query_pos, key_pos, attn_mask):
    """operation_order = cross_attn, norm, self_attn, norm, ffn, norm (post-norm)."""
    x = mha(sd, p + 'attentions.0.', query, key, value, query_pos, key_pos, attn_mask)
    x = _ln(sd, p + 'norms.0', x)
    x = mha(sd, p + 'attentions.1.', x, x, x, query_pos, query_pos, None)
    x = _ln(sd, p + 'norms.1', x)
    y = F.relu(F.linear(x, sd[p + 'ffns.0.layers.0.0.weight'], sd[p + 'ffns.0.layers.0.0.bias']))
    x = x + F.linear(y, sd[p + 'ffns.0.layers.1.weight'], sd[p + 'ffns.0.layers.1.bias'])
    return _ln(sd, p + 'norms.2', x)


# --------------------------------------------------------------------------
# head: models/mask2former/mask2former_head.py, models/mask2former_vps/mask2former_video_head.py
# --------------------------------------------------------------------------
def forward_head(sd, prefix, decoder_out, mask_feature, attn_mask_target_size, num_heads=8, return_attn_logits=False):
    """models/mask2former/mask2former_head.py:355-395.

    decoder_out [Q, B, C]; mask_feature [B, C, h, w].
    """
    decoder_out = _ln(sd, prefix + 'transformer_decoder.post_norm', decoder_out)
    decoder_out = decoder_out.transpose(0, 1)
    cls_pred = F.linear(decoder_out, sd[prefix + 'cls_embed.weight'], sd[prefix + 'cls_embed.bias'])
    me = F.relu(F.linear(decoder_out, sd[prefix + 'mask_embed.0.weight'], sd[prefix + 'mask_embed.0.bias']))
    me = F.relu(F.linear(me, sd[prefix + 'mask_embed.2.weight'], sd[prefix + 'mask_embed.2.bias']))
    me = F.linear(me, sd[prefix + 'mask_embed.4.weight'], sd[prefix + 'mask_embed.4.bias'])
    mask_pred = torch.einsum('bqc,bchw->bqhw', me, mask_feature)
    attn_mask = F.interpolate(mask_pred, attn_mask_target_size, mode='bilinear', align_corners=False)
    attn_mask = attn_mask.flatten(2).unsqueeze(1).repeat((1, num_heads, 1, 1)).flatten(0, 1)
    if return_attn_logits:
        return cls_pred, mask_pred, attn_mask.sigmoid() < 0.5, attn_mask
    attn_mask = attn_mask.sigmoid() < 0.5
    return cls_pred, mask_pred, attn_mask


def forward_head_video(sd, prefix, decoder_out, mask_feature, attn_mask_target_size, num_heads=8,
                       return_attn_logits=False):
    """models/mask2former_vps/mask2former_video_head.py:337-359.

    mask_feature [B, T, C, h, w] -> mask_pred [B, T, Q, h, w].
    """
    decoder_out = _ln(sd, prefix + 'transformer_decoder.post_norm', decoder_out)
    decoder_out = decoder_out.transpose(0, 1)
    cls_pred = F.linear(decoder_out, sd[prefix + 'cls_embed.weight'], sd[prefix + 'cls_embed.bias'])
    me = F.relu(F.linear(decoder_out, sd[prefix + 'mask_embed.0.weight'], sd[prefix + 'mask_embed.0.bias']))
    me = F.relu(F.linear(me, sd[prefix + 'mask_embed.2.weight'], sd[prefix + 'mask_embed.2.bias']))
    me = F.linear(me, sd[prefix + 'mask_embed.4.weight'], sd[prefix + 'mask_embed.4.bias'])
    mask_pred = torch.einsum('bqc,btchw->btqhw', me, mask_feature)
    bs, nf = mask_pred.shape[:2]
    attn_mask = F.interpolate(mask_pred.flatten(0, 1), attn_mask_target_size, mode='bilinear',
                              align_corners=False).unflatten(0, (bs, nf))
    attn_mask = attn_mask.flatten(3).unsqueeze(1).repeat((1, num_heads, 1, 1, 1)).flatten(0, 1)
    attn_mask = attn_mask.transpose(1, 2).flatten(2)
    if return_attn_logits:
        return cls_pred, mask_pred, attn_mask.sigmoid() < 0.5, attn_mask
    attn_mask = attn_mask.sigmoid() < 0.5
    return cls_pred, mask_pred, attn_mask


def _resolve_ties(attn_mask, attn_logits, other, tie_eps, stats, layer):
    """Tie-aware comparison of a discrete decision (tests only).  The attention mask is a SIGN TEST on fp32
    logits (mask2former_head.py:391), so two correct fp32 implementations may disagree on a bit whose logit is
    within re-association noise of zero.  ``other`` is another implementation's mask for the same layer
    ([B, Q, hw] or [B*heads, Q, hw], non-zero = blocked): where it differs from ours AND our logit is within
    ``tie_eps`` of the threshold, its decision is adopted (so the two runs stay comparable downstream); differing
    bits farther than ``tie_eps`` from the threshold are genuine errors and are only counted."""
    other = torch.as_tensor(other).bool()
    heads = attn_mask.shape[0] // other.shape[0]
    if heads > 1:      # [B, Q, hw] -> [B*heads, Q, hw] in the reference's order (index = b * heads + h)
        other = other.unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1)
    diff = other != attn_mask
    tie = attn_logits.abs() < tie_eps
    per_head = lambda m: int(m.sum()) // heads      # noqa: E731  (every head carries the same mask)
    stats.append(dict(layer=layer, flipped_ties=per_head(diff & tie), flipped_non_ties=per_head(diff & ~tie),
                      near_threshold=per_head(tie),
                      max_abs_logit_of_flips=float(attn_logits[diff].abs().max()) if diff.any() else 0.0))
    return torch.where(diff & tie, other, attn_mask)


def head_forward(sd, feats, video=False, num_frames=1, prefix='panoptic_head.', num_layers=9,
                 return_all=False, tie_masks=None, tie_eps=1e-3):
    """Mask2FormerHeadCustom.forward (models/mask2former/mask2former_head.py:397-479) /
    Mask2FormerVideoHead.forward (models/mask2former_vps/mask2former_video_head.py:361-462).

    Returns (cls_pred_list, mask_pred_list, query_feat[, extras]).
    tie_masks (tests only): the attention masks ANOTHER implementation computed for layers 0..num_layers-1
    (raw sign masks, before the all-blocked-row rule); near-threshold disagreements are resolved in its favour
    (``_resolve_ties``) and reported in extras['tie_stats'].
    """
    mask_features, memories = pixel_decoder(sd, feats, prefix + 'pixel_decoder.')
    if video:
        bs_nf = mask_features.shape[0]
        batch_size = bs_nf // num_frames
        assert batch_size * num_frames == bs_nf  # mask2former_video_head.py:384
        mask_features = mask_features.reshape((batch_size, num_frames) + mask_features.shape[1:])
        memories = [m.reshape((batch_size, num_frames) + m.shape[1:]) for m in memories]
    else:
        batch_size = mask_features.shape[0]
    decoder_inputs, decoder_pos = [], []
    for i in range(3):
        lvl = sd[prefix + 'level_embed.weight'][i].view(1, 1, -1)
        if video:
            di = memories[i].flatten(3).permute(1, 3, 0, 2).flatten(0, 1) + lvl
            _, t, _, h, w = memories[i].shape
            pe = sine_pe_3d(batch_size, t, h, w).flatten(3).permute(1, 3, 0, 2).flatten(0, 1)
        else:
            di = memories[i].flatten(2).permute(2, 0, 1) + lvl
            h, w = memories[i].shape[-2:]
            pe = sine_pe_2d(batch_size, h, w).flatten(2).permute(2, 0, 1)
        decoder_inputs.append(di)
        decoder_pos.append(pe)
    query_feat = sd[prefix + 'query_feat.weight'].unsqueeze(1).repeat((1, batch_size, 1))
    query_embed = sd[prefix + 'query_embed.weight'].unsqueeze(1).repeat((1, batch_size, 1))
    fh = forward_head_video if video else forward_head
    cls_list, mask_list, attn_list, raw_list, tie_stats = [], [], [], [], []
    cls_pred, mask_pred, attn_mask, attn_logits = fh(sd, prefix, query_feat, mask_features, memories[0].shape[-2:],
                                                     return_attn_logits=True)
    cls_list.append(cls_pred)
    mask_list.append(mask_pred)
    for i in range(num_layers):
        level_idx = i % 3
        attn_mask = attn_mask.clone()
        raw_list.append(attn_mask.clone())      # the sign mask before tie resolution / the all-blocked-row rule
        if tie_masks is not None:
            attn_mask = _resolve_ties(attn_mask, attn_logits, tie_masks[i], tie_eps, tie_stats, i)
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
        attn_list.append(attn_mask)
        query_feat = decoder_layer(
            sd, f'{prefix}transformer_decoder.layers.{i}.', query_feat,
            decoder_inputs[level_idx], decoder_inputs[level_idx], query_embed,
            decoder_pos[level_idx], attn_mask)
        cls_pred, mask_pred, attn_mask, attn_logits = fh(sd, prefix, query_feat, mask_features,
                                                         memories[(i + 1) % 3].shape[-2:], return_attn_logits=True)
        cls_list.append(cls_pred)
        mask_list.append(mask_pred)
    if return_all:
        return cls_list, mask_list, query_feat, dict(mask_features=mask_features, memories=memories,
                                                      attn_masks=attn_list, raw_attn_masks=raw_list, tie_stats=tie_stats)
    return cls_list, mask_list, query_feat


def head_simple_test_with_query(sd, feats, batch_input_shape, video=False, num_frames=1, tie_masks=None, tie_eps=1e-3,
                                tie_stats=None):
    """models/mask2former/mask2former_head.py:650-681,
    models/mask2former_vps/mask2former_video_head.py:637-669.
    tie_masks / tie_eps: see ``head_forward`` (tests only); the per-layer statistics are appended to ``tie_stats``."""
    cls_list, mask_list, query_feat, extras = head_forward(sd, feats, video, num_frames, return_all=True,
                                                           tie_masks=tie_masks, tie_eps=tie_eps)
    if tie_stats is not None:
        tie_stats.extend(extras['tie_stats'])
    mask_cls = cls_list[-1]
    mask_pred = mask_list[-1]
    if video:
        bs, nf = mask_pred.shape[:2]
        mask_pred = F.interpolate(mask_pred.flatten(0, 1), size=tuple(batch_input_shape),
                                  mode='bilinear', align_corners=False).unflatten(0, (bs, nf))
        return mask_cls, mask_pred, query_feat
    mask_pred = F.interpolate(mask_pred, size=tuple(batch_input_shape), mode='bilinear',
                              align_corners=False)
    return mask_cls, mask_pred, query_feat.unsqueeze(0)


# --------------------------------------------------------------------------
# fusion head: models/mask2former/mask2former_fusion_head.py
# --------------------------------------------------------------------------
def panoptic_postprocess_with_query(mask_cls, mask_pred, query_feats, num_things=115, num_stuff=11,
                                    object_mask_thr=0.8, iou_thr=0.8, filter_low_score=True):
    """models/mask2former/mask2former_fusion_head.py:96-171."""
    num_classes = num_things + num_stuff
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    mask_pred = mask_pred.sigmoid()
    keep = labels.ne(num_classes) & (scores > object_mask_thr)
    cur_scores = scores[keep]
    cur_classes = labels[keep]
    cur_masks = mask_pred[keep]
    cur_query_feats = query_feats[keep]
    cur_prob_masks = cur_scores.view(-1, 1, 1) * cur_masks
    h, w = cur_masks.shape[-2:]
    panoptic_seg = torch.full((h, w), num_classes, dtype=torch.int32)
    query_feat_dict = defaultdict(list)
    if cur_masks.shape[0] > 0:
        cur_mask_ids = cur_prob_masks.argmax(0)
        instance_id = 1
        for k in range(cur_classes.shape[0]):
            pred_class = int(cur_classes[k].item())
            isthing = pred_class < num_things
            mask = cur_mask_ids == k
            mask_area = mask.sum().item()
            original_area = (cur_masks[k] >= 0.5).sum().item()
            if filter_low_score:
                mask = mask & (cur_masks[k] >= 0.5)
            if mask_area > 0 and original_area > 0:
                if mask_area / original_area < iou_thr:
                    continue
                if not mask.any():  # :156-157, :161-162
                    continue
                if not isthing:
                    panoptic_seg[mask] = pred_class
                    query_feat_dict[pred_class].append(cur_query_feats[k])
                else:
                    seg_id = pred_class + instance_id * INSTANCE_OFFSET
                    panoptic_seg[mask] = seg_id
                    query_feat_dict[seg_id].append(cur_query_feats[k])
                    instance_id += 1
    return panoptic_seg, query_feat_dict


def mask2bbox(masks):
    """mmdet/core/mask/utils.py mask2bbox (L0, A6): (x0, y0, x1+1, y1+1), zeros if empty."""
    n = masks.shape[0]
    bboxes = masks.new_zeros((n, 4), dtype=torch.float32)
    x_any = torch.any(masks, dim=1)
    y_any = torch.any(masks, dim=2)
    for i in range(n):
        x = torch.where(x_any[i, :])[0]
        y = torch.where(y_any[i, :])[0]
        if len(x) > 0 and len(y) > 0:
            bboxes[i, :] = bboxes.new_tensor([x[0], y[0], x[-1] + 1, y[-1] + 1])
    return bboxes


def instance_postprocess(mask_cls, mask_pred, num_things=115, num_stuff=11, max_per_image=100):
    """models/mask2former/mask2former_fusion_head.py:192-242."""
    num_classes = num_things + num_stuff
    num_queries = mask_cls.shape[0]
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(num_queries, 1).flatten(0, 1)
    scores_per_image, top_indices = scores.flatten(0, 1).topk(max_per_image, sorted=False)
    labels_per_image = labels[top_indices]
    query_indices = top_indices // num_classes
    mask_pred = mask_pred[query_indices]
    is_thing = labels_per_image < num_things
    scores_per_image = scores_per_image[is_thing]
    labels_per_image = labels_per_image[is_thing]
    mask_pred = mask_pred[is_thing]
    mask_pred_binary = (mask_pred > 0).float()
    mask_scores = (mask_pred.sigmoid() * mask_pred_binary).flatten(1).sum(1) / (
        mask_pred_binary.flatten(1).sum(1) + 1e-6)
    det_scores = scores_per_image * mask_scores
    mask_pred_binary = mask_pred_binary.bool()
    bboxes = mask2bbox(mask_pred_binary)
    bboxes = torch.cat([bboxes, det_scores[:, None]], dim=-1)
    return labels_per_image, bboxes, mask_pred_binary


def fusion_simple_test_with_query(mask_cls_results, mask_pred_results, query_feats, img_metas,
                                  rescale=False, instance_on=True, **cfg):
    """models/mask2former/mask2former_fusion_head.py:325-404."""
    results = []
    for mask_cls, mask_pred, qf, meta in zip(mask_cls_results, mask_pred_results, query_feats, img_metas):
        ih, iw = meta['img_shape'][:2]
        mask_pred = mask_pred[:, :ih, :iw]
        if rescale:
            oh, ow = meta['ori_shape'][:2]
            mask_pred = F.interpolate(mask_pred[:, None], size=(oh, ow), mode='bilinear',
                                      align_corners=False)[:, 0]
        pan, qfd = panoptic_postprocess_with_query(mask_cls, mask_pred, qf, **cfg)
        result = dict(pan_results=pan, query_feats=qfd)
        if instance_on:
            result['ins_results'] = instance_postprocess(mask_cls, mask_pred)
        results.append(result)
    return results


# --------------------------------------------------------------------------
# detectors
# --------------------------------------------------------------------------
def ips_simple_test(sd, img, img_metas, rescale=True, instance_on=True, return_raw=False, backbone=None, **tie):
    """Mask2FormerCustom.simple_test, models/mask2former/mask2former.py:121-191
    (up to the numpy conversion of pan_results / query feats).  **tie: tie_masks / tie_eps / tie_stats (tests)."""
    feats = (backbone or resnet50)(sd, img)
    mask_cls, mask_pred, query_feats = head_simple_test_with_query(
        sd, feats, img_metas[0]['batch_input_shape'], **tie)
    results = fusion_simple_test_with_query(mask_cls, mask_pred, query_feats, img_metas,
                                            rescale=rescale, instance_on=instance_on)
    for r in results:
        r['pan_results'] = r['pan_results'].numpy()
        r['query_feats'] = {k: [x.numpy() for x in v] for k, v in r['query_feats'].items()}
    if return_raw:
        return results, dict(cls=mask_cls, masks=mask_pred, embds=query_feats)
    return results


def match_from_embds(tgt_embds, cur_embds):
    """models/mask2former_vps/mask2former_min_vis.py:244-258."""
    from scipy.optimize import linear_sum_assignment
    cur_embds = cur_embds / cur_embds.norm(dim=1)[:, None]
    tgt_embds = tgt_embds / tgt_embds.norm(dim=1)[:, None]
    cos_sim = torch.mm(cur_embds, tgt_embds.transpose(0, 1))
    C = 1.0 * (1 - cos_sim)
    indices = linear_sum_assignment(C.transpose(0, 1))
    return indices[1]


def vps_simple_test(sd, ref_img, ref_img_metas, rescale=True, instance_on=True, return_raw=False, backbone=None,
                    tie_masks=None, tie_eps=1e-3, tie_stats=None):
    """Mask2FormerVideoCustom.simple_test, models/mask2former_vps/mask2former.py:125-223.

    ref_img [B, T, 3, H, W]; the shipped test config uses T = 1 and B = 1.
    tie_masks (tests only): per frame of the flattened batch, the list of attention masks of another implementation
    (see ``head_forward``); statistics are appended to ``tie_stats``.
    """
    bs, num_frame, three, h, w = ref_img.shape
    video_x = (backbone or resnet50)(sd, ref_img.reshape(bs * num_frame, three, h, w))
    pred_logits, mask_pred_list, query_pred_list = [], [], []
    for i in range(video_x[0].shape[0]):
        cur = [f[i].unsqueeze(0) for f in video_x]
        mask_cls, mask_pred, query_fea = head_simple_test_with_query(
            sd, cur, ref_img_metas[0][0]['batch_input_shape'], video=True, num_frames=1,
            tie_masks=None if tie_masks is None else tie_masks[i], tie_eps=tie_eps, tie_stats=tie_stats)
        pred_logits.append(mask_cls.squeeze())
        mask_pred_list.append(mask_pred.squeeze())
        query_pred_list.append(query_fea.permute(0, 2, 1).squeeze())
    out_logits, out_masks, out_embds = [pred_logits[0]], [mask_pred_list[0]], [query_pred_list[0]]
    for i in range(1, len(pred_logits)):
        indices = match_from_embds(out_embds[-1], query_pred_list[i])
        out_logits.append(pred_logits[i][indices, :])
        out_masks.append(mask_pred_list[i][indices, :, :])
        out_embds.append(query_pred_list[i][indices, :])
    out_logits = (sum(out_logits) / len(out_logits)).unsqueeze(0)
    out_masks = torch.stack(out_masks, dim=0).unsqueeze(0)
    out_embds = (sum(out_embds) / len(out_embds)).unsqueeze(0)
    results = [[] for _ in range(bs)]
    for frame_id in range(num_frame):
        result = fusion_simple_test_with_query(
            out_logits, out_masks[:, frame_id], out_embds,
            [ref_img_metas[idx][frame_id] for idx in range(bs)], rescale=rescale,
            instance_on=instance_on)
        for i in range(len(result)):
            result[i]['pan_results'] = result[i]['pan_results'].numpy()
            results[i].append(result[i])
    if return_raw:
        return results, dict(cls=out_logits, masks=out_masks, embds=out_embds)
    return results


# --------------------------------------------------------------------------
# tube linking: models/mask2former_vps/utils.py:20-89 (in-memory part)
# --------------------------------------------------------------------------
def rle_encode(mask):
    """COCO RLE 'counts' (uncompressed list) of a [H, W] uint8 mask, column-major,
    as pycocotools.mask.encode produces before its string compression (L0)."""
    flat = np.asarray(mask, dtype=np.uint8).flatten(order='F')
    counts, prev, run = [], 0, 0
    for v in flat:
        if v != prev:
            counts.append(run)
            run, prev = 0, v
        run += 1
    counts.append(run)
    return counts


def rle_to_string(counts):
    """pycocotools rleToString (L0): LEB128-like, delta-coded from the 3rd count."""
    out = []
    for i, x in enumerate(counts):
        x = int(x)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return ''.join(out)


def concat_seq(outputs):
    """models/mask2former_vps/utils.py:20-89 without the file / image side effects.

    outputs: list over frames of [dict(pan_results, query_feats)].
    Returns (results rows [(frame, tid, cls, h, w, rle_string)], feat_tubes_dict).
    """
    object_list, feat_tubes, rows = [], {}, []
    for frame_id, output in enumerate(outputs):
        output = output[0]
        for ins_id, feat in output['query_feats'].items():
            if ins_id not in object_list:
                object_list.append(ins_id)
                feat_tubes[object_list.index(ins_id) + 1] = {}
            tid = object_list.index(ins_id) + 1
            feat_tubes[tid][frame_id] = dict(
                query_feat=np.asarray(feat[0], dtype=np.float32).reshape(-1),
                cls_id=int(ins_id % 1000))
            mask = (output['pan_results'] == ins_id).astype(np.uint8)
            rows.append((frame_id + 1, tid, int(ins_id % 1000), mask.shape[0], mask.shape[1],
                         rle_to_string(rle_encode(mask))))
    return rows, feat_tubes


def minvis_simple_test(sd, ref_img, ref_img_metas, rescale=True, backbone=None, tie_masks=None, tie_eps=1e-3,
                       tie_stats=None, return_masks=False):
    """Mask2FormerVideoCustomMinVIS.simple_test, models/mask2former_vps/mask2former_min_vis.py:132-231
    (panoptic branch): per-frame heads, MinVIS matching, clip-averaged logits, per-frame fusion."""
    bs, num_frame, three, h, w = ref_img.shape
    video_x = (backbone or resnet50)(sd, ref_img.reshape(bs * num_frame, three, h, w))
    pred_logits, mask_pred_list, query_pred_list = [], [], []
    for i in range(num_frame):
        cur = [f[i].unsqueeze(0) for f in video_x]
        mask_cls, mask_pred, query_fea = head_simple_test_with_query(
            sd, cur, ref_img_metas[0][0]['batch_input_shape'], video=True, num_frames=1,
            tie_masks=None if tie_masks is None else tie_masks[i], tie_eps=tie_eps, tie_stats=tie_stats)
        pred_logits.append(mask_cls.squeeze())
        mask_pred_list.append(mask_pred.squeeze())
        query_pred_list.append(query_fea.permute(0, 2, 1).squeeze())
    out_logits, out_masks, out_embds = [pred_logits[0]], [mask_pred_list[0]], [query_pred_list[0]]
    perms = []
    for i in range(1, num_frame):
        indices = match_from_embds(out_embds[-1], query_pred_list[i])
        perms.append(np.asarray(indices))
        out_logits.append(pred_logits[i][indices, :])
        out_masks.append(mask_pred_list[i][indices, :, :])
        out_embds.append(query_pred_list[i][indices, :])
    logits = (sum(out_logits) / len(out_logits)).unsqueeze(0)
    masks = torch.stack(out_masks, dim=0).unsqueeze(0)
    results = []
    for frame_id in range(num_frame):
        meta = ref_img_metas[0][frame_id]
        mp = masks[0, frame_id][:, :meta['img_shape'][0], :meta['img_shape'][1]]
        if rescale:
            mp = F.interpolate(mp[:, None], size=tuple(meta['ori_shape'][:2]), mode='bilinear',
                               align_corners=False)[:, 0]
        pan, _ = panoptic_postprocess_with_query(logits[0], mp, torch.zeros(mp.shape[0], 1))
        results.append(pan.numpy())
    if return_masks:
        return results, perms, logits, masks
    return results, perms, logits
