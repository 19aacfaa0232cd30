"""Seeded synthetic checkpoints and inputs.

No checkpoint of the reference is available offline (``load_from`` is a URL,
configs/mask2former_vps/mask2former_video_r50.py:9-11; the relation
``epoch_100.pth`` is absent), so tests / bench / smoke use one seeded random
state_dict with the reference's key layout (SURVEY.md section 8b "Ownership").
Weights are scaled so that activations stay O(1) through the 50-layer backbone
and so that the fusion head keeps a non-trivial set of segments (otherwise the
panoptic map would be all-void and the id-parity test vacuous).
"""
import math

import numpy as np
import torch

IMG_MEAN = (123.675, 116.28, 103.53)   # configs/_base_/datasets/pvsg_vps.py:4-6
IMG_STD = (58.395, 57.12, 57.375)


def _conv_w(g, cout, cin, k, gain=2.0):
    std = math.sqrt(gain / (cin * k * k))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _lin(g, sd, name, cout, cin, gain=1.0, bias_std=0.02):
    sd[name + '.weight'] = torch.randn(cout, cin, generator=g) * math.sqrt(gain / cin)
    sd[name + '.bias'] = torch.randn(cout, generator=g) * bias_std


def _norm(g, sd, name, c, lo=0.8, hi=1.2):
    sd[name + '.weight'] = torch.empty(c).uniform_(lo, hi, generator=g)
    sd[name + '.bias'] = torch.randn(c, generator=g) * 0.05


def _bn(g, sd, name, c, gamma=1.0):
    sd[name + '.weight'] = torch.empty(c).uniform_(0.8, 1.2, generator=g) * gamma
    sd[name + '.bias'] = torch.randn(c, generator=g) * 0.05
    sd[name + '.running_mean'] = torch.randn(c, generator=g) * 0.1
    sd[name + '.running_var'] = torch.empty(c).uniform_(0.5, 1.5, generator=g)
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def resnet50_state_dict(g, prefix='backbone.'):
    sd = {}
    sd[prefix + 'conv1.weight'] = _conv_w(g, 64, 3, 7)
    _bn(g, sd, prefix + 'bn1', 64)
    inpl = 64
    for li, (nblk, planes) in enumerate(((3, 64), (4, 128), (6, 256), (3, 512))):
        for b in range(nblk):
            p = f'{prefix}layer{li + 1}.{b}.'
            sd[p + 'conv1.weight'] = _conv_w(g, planes, inpl, 1)
            _bn(g, sd, p + 'bn1', planes)
            sd[p + 'conv2.weight'] = _conv_w(g, planes, planes, 3)
            _bn(g, sd, p + 'bn2', planes)
            sd[p + 'conv3.weight'] = _conv_w(g, planes * 4, planes, 1)
            _bn(g, sd, p + 'bn3', planes * 4, gamma=0.3)
            if b == 0:
                sd[p + 'downsample.0.weight'] = _conv_w(g, planes * 4, inpl, 1, gain=1.0)
                _bn(g, sd, p + 'downsample.1', planes * 4)
            inpl = planes * 4
    return sd


def swin_state_dict(g, prefix='backbone.', embed_dims=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=12,
                    patch_size=4, mlp_ratio=4, out_indices=(0, 1, 2, 3)):
    """mmdet 2.25 ``SwinTransformer`` key layout (mmdet/models/backbones/swin.py), incl. the persistent
    ``relative_position_index`` buffers.  Residual-branch outputs are damped so 24 blocks stay O(1)."""
    sd = {}
    p = prefix
    sd[p + 'patch_embed.projection.weight'] = _conv_w(g, embed_dims, 3, patch_size, gain=1.0)
    sd[p + 'patch_embed.projection.bias'] = torch.randn(embed_dims, generator=g) * 0.02
    _norm(g, sd, p + 'patch_embed.norm', embed_dims)
    ws = window_size
    seq = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)[:, None] + torch.arange(ws)[None, :]
    coords = seq.reshape(1, -1)
    rel_index = (coords + coords.T).flip(1).contiguous()
    C = embed_dims
    for i, depth in enumerate(depths):
        for j in range(depth):
            b = f'{p}stages.{i}.blocks.{j}.'
            _norm(g, sd, b + 'norm1', C)
            sd[b + 'attn.w_msa.relative_position_bias_table'] = torch.randn((2 * ws - 1) ** 2, num_heads[i], generator=g) * 0.5
            sd[b + 'attn.w_msa.relative_position_index'] = rel_index.clone()
            _lin(g, sd, b + 'attn.w_msa.qkv', 3 * C, C, gain=2.0, bias_std=0.1)
            _lin(g, sd, b + 'attn.w_msa.proj', C, C, gain=0.25)
            _norm(g, sd, b + 'norm2', C)
            _lin(g, sd, b + 'ffn.layers.0.0', mlp_ratio * C, C, gain=2.0, bias_std=0.1)
            _lin(g, sd, b + 'ffn.layers.1', C, mlp_ratio * C, gain=0.25)
        if i < len(depths) - 1:
            _norm(g, sd, f'{p}stages.{i}.downsample.norm', 4 * C)
            sd[f'{p}stages.{i}.downsample.reduction.weight'] = torch.randn(2 * C, 4 * C, generator=g) * math.sqrt(1.0 / (4 * C))
        if i in out_indices:
            _norm(g, sd, f'{p}norm{i}', C)
        C *= 2
    return sd


def mask2former_state_dict(seed=0, num_classes=126, num_queries=100, enc_layers=6, dec_layers=9,
                           in_channels=(256, 512, 1024, 2048), cls_gain=40.0, mask_shift=14.0,
                           dec_gain=0.02, backbone=None):
    """Full detector state_dict (backbone + panoptic_head) with mmdet 2.25 key names.  backbone: None = ResNet-50;
    a dict of ``swin_state_dict`` keyword arguments = Swin (then in_channels must be its stage widths)."""
    g = torch.Generator().manual_seed(seed)
    sd = resnet50_state_dict(g) if backbone is None else swin_state_dict(g, **backbone)
    C = 256
    ph = 'panoptic_head.'
    pd = ph + 'pixel_decoder.'
    for i, cin in enumerate(reversed(in_channels[1:])):  # C5, C4, C3
        sd[f'{pd}input_convs.{i}.conv.weight'] = _conv_w(g, C, cin, 1, gain=1.0)
        sd[f'{pd}input_convs.{i}.conv.bias'] = torch.randn(C, generator=g) * 0.02
        _norm(g, sd, f'{pd}input_convs.{i}.gn', C)
    for l in range(enc_layers):
        p = f'{pd}encoder.layers.{l}.'
        a = p + 'attentions.0.'
        # offsets: a few pixels of spread so samples leave the reference cell and hit borders
        sd[a + 'sampling_offsets.weight'] = torch.randn(192, C, generator=g) * (0.6 / math.sqrt(C))
        sd[a + 'sampling_offsets.bias'] = torch.randn(192, generator=g) * 1.5
        _lin(g, sd, a + 'attention_weights', 96, C, gain=1.0)
        _lin(g, sd, a + 'value_proj', C, C)
        _lin(g, sd, a + 'output_proj', C, C, gain=0.5)
        _lin(g, sd, p + 'ffns.0.layers.0.0', 1024, C, gain=2.0)
        _lin(g, sd, p + 'ffns.0.layers.1', C, 1024, gain=0.5)
        _norm(g, sd, p + 'norms.0', C)
        _norm(g, sd, p + 'norms.1', C)
    sd[pd + 'level_encoding.weight'] = torch.randn(3, C, generator=g)
    sd[pd + 'lateral_convs.0.conv.weight'] = _conv_w(g, C, in_channels[0], 1, gain=1.0)
    _norm(g, sd, pd + 'lateral_convs.0.gn', C)
    sd[pd + 'output_convs.0.conv.weight'] = _conv_w(g, C, C, 3)
    _norm(g, sd, pd + 'output_convs.0.gn', C)
    sd[pd + 'mask_feature.weight'] = _conv_w(g, C, C, 1, gain=1.0)
    sd[pd + 'mask_feature.bias'] = torch.randn(C, generator=g) * 0.02
    td = ph + 'transformer_decoder.'
    for i in range(dec_layers):
        p = f'{td}layers.{i}.'
        for a in (0, 1):
            q = f'{p}attentions.{a}.attn.'
            sd[q + 'in_proj_weight'] = torch.randn(3 * C, C, generator=g) * math.sqrt(2.0 / C)
            sd[q + 'in_proj_bias'] = torch.randn(3 * C, generator=g) * 0.02
            sd[q + 'out_proj.weight'] = torch.randn(C, C, generator=g) * math.sqrt(dec_gain / C)
            sd[q + 'out_proj.bias'] = torch.randn(C, generator=g) * 0.02
        _lin(g, sd, p + 'ffns.0.layers.0.0', 2048, C, gain=2.0)
        _lin(g, sd, p + 'ffns.0.layers.1', C, 2048, gain=dec_gain)
        for k in range(3):
            _norm(g, sd, f'{p}norms.{k}', C)
    _norm(g, sd, td + 'post_norm', C)
    sd[ph + 'query_embed.weight'] = torch.randn(num_queries, C, generator=g)
    sd[ph + 'query_feat.weight'] = torch.randn(num_queries, C, generator=g)
    sd[ph + 'level_embed.weight'] = torch.randn(3, C, generator=g)
    # confident classes so that `score > object_mask_thr` keeps a subset of queries
    _lin(g, sd, ph + 'cls_embed', num_classes + 1, C, gain=cls_gain)
    _lin(g, sd, ph + 'mask_embed.0', C, C, gain=2.0)
    _lin(g, sd, ph + 'mask_embed.2', C, C, gain=2.0)
    _lin(g, sd, ph + 'mask_embed.4', C, C, gain=1.0)
    # channel 0 of mask_feature is ~constant; a negative embed bias on it shifts every
    # mask logit down so masks are sparse blobs instead of half-planes.
    sd[pd + 'mask_feature.bias'][0] = 4.0
    sd[pd + 'mask_feature.weight'][0] *= 0.05
    sd[ph + 'mask_embed.4.weight'][0] *= 0.0
    sd[ph + 'mask_embed.4.bias'][0] = -mask_shift / 4.0
    return sd


def relation_state_dicts(seed=0, feature_dim=256, hidden_dim=1024, num_relations=57):
    """The four state_dicts saved by tools/rel_train.py:223-231, torch key names
    (models/relation_head/base.py:26-62, transformer.py:8-33)."""
    g = torch.Generator().manual_seed(seed)

    def enc_layer(sd, p, d, ffn):
        sd[p + 'self_attn.in_proj_weight'] = torch.randn(3 * d, d, generator=g) * math.sqrt(1.5 / d)
        sd[p + 'self_attn.in_proj_bias'] = torch.randn(3 * d, generator=g) * 0.02
        sd[p + 'self_attn.out_proj.weight'] = torch.randn(d, d, generator=g) * math.sqrt(1.0 / d)
        sd[p + 'self_attn.out_proj.bias'] = torch.randn(d, generator=g) * 0.02
        _lin(g, sd, p + 'linear1', ffn, d, gain=2.0)
        _lin(g, sd, p + 'linear2', d, ffn, gain=1.0)
        _norm(g, sd, p + 'norm1', d)
        _norm(g, sd, p + 'norm2', d)

    out = {}
    for name in ('subject_encoder', 'object_encoder'):
        sd = {}
        for l in range(2):
            enc_layer(sd, f'transformer_encoder.layers.{l}.', feature_dim, 512)
        out[name] = sd
    sd = {}
    _lin(g, sd, 'pair_ffn.0', hidden_dim, 2 * feature_dim, gain=2.0)
    _lin(g, sd, 'pair_ffn.2', 1, hidden_dim, gain=1.0)
    out['pair_proposal_model'] = sd
    d = 2 * feature_dim
    sd = {}
    position = torch.arange(5000).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
    pe = torch.zeros(5000, 1, d)
    pe[:, 0, 0::2] = torch.sin(position * div_term)
    pe[:, 0, 1::2] = torch.cos(position * div_term)
    sd['positional_encoding.pe'] = pe
    enc_layer(sd, 'transformer_encoder.layers.0.', d, 512)
    _norm(g, sd, 'layer_norm', d)
    _lin(g, sd, 'fc1', d // 2, d, gain=2.0)
    _lin(g, sd, 'fc2', d // 4, d // 2, gain=2.0)
    _lin(g, sd, 'span_head', num_relations, d // 4)
    _lin(g, sd, 'pred_head', num_relations, d // 4)
    out['relation_model'] = sd
    return out


def relation_baseline_state_dicts(seed=0, input_dim=512, num_relations=57, kernel_size=5, num_layers=1):
    """state_dicts of the two baseline relation models of models/relation_head/convolution.py (``--model-name filter`` /
    ``conv`` of tools/rel_test.py:167-175): HandcraftedFilter (heads only, its filter is a constant) and Learnable1DConv."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in ('filter', 'conv'):
        sd = {}
        if name == 'conv':
            for l in range(num_layers):
                sd[f'conv_layers.{2 * l}.weight'] = torch.randn(input_dim, input_dim, kernel_size, generator=g) * \
                    math.sqrt(2.0 / (input_dim * kernel_size))
                sd[f'conv_layers.{2 * l}.bias'] = torch.randn(input_dim, generator=g) * 0.02
        _lin(g, sd, 'fc1', input_dim // 2, input_dim, gain=2.0)
        _lin(g, sd, 'fc2', input_dim // 4, input_dim // 2, gain=2.0)
        _lin(g, sd, 'span_head', num_relations, input_dim // 4)
        _lin(g, sd, 'pred_head', num_relations, input_dim // 4)
        out[name] = sd
    return out


def synthetic_frame(seed, height=720, width=1280, size_divisor=32):
    """One normalised, padded frame [3, Hp, Wp] fp32:
    uint8 ~ U{0..255} from default_rng(seed), SeqNormalize (to_rgb=False), SeqPad(32)."""
    rng = np.random.default_rng(seed)
    # low-frequency content: an upsampled coarse random field plus pixel noise, so the
    # network sees image-like structure instead of white noise
    ch, cw = (height + 39) // 40, (width + 39) // 40
    coarse = rng.integers(0, 256, size=(ch, cw, 3)).astype(np.float32)
    img = np.repeat(np.repeat(coarse, 40, axis=0), 40, axis=1)[:height, :width]
    img = np.clip(img + rng.integers(-20, 21, size=(height, width, 3)), 0, 255).astype(np.uint8)
    x = (img.astype(np.float32) - np.array(IMG_MEAN, np.float32)) / np.array(IMG_STD, np.float32)
    hp = (height + size_divisor - 1) // size_divisor * size_divisor
    wp = (width + size_divisor - 1) // size_divisor * size_divisor
    out = np.zeros((3, hp, wp), np.float32)
    out[:, :height, :width] = x.transpose(2, 0, 1)
    return torch.from_numpy(out)


def frame_meta(height=720, width=1280, size_divisor=32):
    hp = (height + size_divisor - 1) // size_divisor * size_divisor
    wp = (width + size_divisor - 1) // size_divisor * size_divisor
    return dict(img_shape=(height, width, 3), ori_shape=(height, width, 3),
                pad_shape=(hp, wp, 3), batch_input_shape=(hp, wp))


def training_batch(clips=16, height=384, width=480, num_frames=2, num_gt=4, device='cpu', seed=0):
    """A synthetic training batch in the reference's format (configs/_base_/datasets/pvsg_vps.py: 2-frame clips resized to
    360 x 480 and padded to a multiple of 32; mask2former_video_r50.py: samples_per_gpu = 16): the keyword arguments of
    ``Mask2FormerVideoCustom.forward_train`` -- ref_img [B,T,3,H,W], per clip the (frame, label) / (frame, instance id) pairs
    and per frame the instance masks [n,H,W].  ``num_gt`` box-shaped instances per clip, present in every frame."""
    frames = torch.stack([torch.stack([synthetic_frame(seed + 7 * c + t, height, width) for t in range(num_frames)])
                          for c in range(clips)]).to(device)
    metas = [[dict(frame_meta(height, width), pad_shape=(height, width, 3)) for _ in range(num_frames)] for _ in range(clips)]
    gt_masks, gt_labels, gt_ids = [], [], []
    for c in range(clips):
        m = torch.zeros(num_gt, num_frames, height, width, dtype=torch.bool)
        for k in range(num_gt):
            y0, x0 = (37 * k + 11 * c) % (height // 2), (53 * k + 29 * c) % (width // 2)
            m[k, :, y0:y0 + height // 3, x0:x0 + width // 3] = True
        gt_masks.append([m[:, t].to(device) for t in range(num_frames)])
        gt_labels.append(torch.tensor([[t, (10 * k + 3 + c) % 126] for t in range(num_frames) for k in range(num_gt)], device=device))
        gt_ids.append(torch.tensor([[t, 100 + k] for t in range(num_frames) for k in range(num_gt)], device=device))
    return dict(img=frames[:, 0], img_metas=[m[0] for m in metas], return_loss=True, ref_img=frames, ref_img_metas=metas,
                ref_gt_bboxes=None, ref_gt_labels=gt_labels, ref_gt_masks=gt_masks, ref_gt_semantic_seg=None,
                ref_gt_instance_ids=gt_ids)
