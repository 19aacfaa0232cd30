"""TEST INFRASTRUCTURE (oracle): fp32 CPU restatement of mmdet 2.25.0 ``SwinTransformer`` (the backbone
BASELINE configs[2] names), functional over an mmdet-layout ``state_dict``.

**Parity unpinned by the reference**: LilyDaytoy/OpenPVSG ships no Swin code, config or test (SURVEY.md 8d
"Config 3"); the algorithm is restated from the pinned third-party version (mmdet==2.25.0, README.md:125):
``mmdet/models/backbones/swin.py`` -- WindowMSA.forward :84-117 (q scaled by head_dim**-0.5, relative position
bias table indexed by ``relative_position_index``, optional mask, softmax, proj), ShiftWindowMSA.forward :175-246
(pad bottom/right to a window multiple AFTER norm1, roll by -shift, img_mask regions from the slices (0,-ws),
(-ws,-shift), (-shift,None) with -100 fill, window partition / reverse, roll back, crop), SwinBlock :288-356
(x + attn(norm1 x); x + ffn(norm2 x), GELU), SwinBlockSequence :359-437 (shift on odd blocks, downsample after
the blocks, returns the pre-downsample map), SwinTransformer.forward :745-772 (per-stage ``norm{i}`` on the outputs)
-- and ``mmdet/models/utils/transformer.py`` PatchEmbed :115-232 (conv k=s=patch, flatten, LayerNorm),
PatchMerging :235-352 (nn.Unfold 2x2 channel order, LayerNorm(4C), Linear(4C, 2C, bias=False)).
Cross-checked against an independent implementation available offline: torchvision ``SwinTransformer``
(tests/test_swin.py::test_oracle_vs_torchvision, weights converted between the two patch-merging layouts).
Imported by tests/ and bench.py's CPU baseline only.
"""
import torch
import torch.nn.functional as F


def relative_position_index(ws):
    """WindowMSA.__init__ :60-66: double_step_seq(2*Ww-1, Wh, 1, Ww); coords + coords.T; flip(1)."""
    seq1 = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)
    seq2 = torch.arange(0, ws, 1)
    coords = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
    return (coords + coords.T).flip(1).contiguous()


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, H, W, ws):
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def window_msa(sd, pre, x, num_heads, ws, mask=None):
    B, N, C = x.shape
    qkv = F.linear(x, sd[pre + 'qkv.weight'], sd[pre + 'qkv.bias'])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (C // num_heads) ** -0.5
    attn = q @ k.transpose(-2, -1)
    index = sd.get(pre + 'relative_position_index', relative_position_index(ws))
    bias = sd[pre + 'relative_position_bias_table'][index.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def shift_window_msa(sd, pre, query, hw, num_heads, ws, shift):
    B, L, C = query.shape
    H, W = hw
    assert L == H * W
    query = query.view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    query = F.pad(query, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = query.shape[1], query.shape[2]
    mask = None
    if shift > 0:
        query = torch.roll(query, shifts=(-shift, -shift), dims=(1, 2))
        img_mask = torch.zeros((1, Hp, Wp, 1))
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                img_mask[:, hs, wsl, :] = cnt
                cnt += 1
        mw = window_partition(img_mask, ws).view(-1, ws * ws)
        mask = mw.unsqueeze(1) - mw.unsqueeze(2)
        mask = mask.masked_fill(mask != 0, -100.0).masked_fill(mask == 0, 0.0)
    windows = window_partition(query, ws).view(-1, ws * ws, C)
    out = window_msa(sd, pre + 'w_msa.', windows, num_heads, ws, mask).view(-1, ws, ws, C)
    x = window_reverse(out, Hp, Wp, ws)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    if pad_r > 0 or pad_b > 0:
        x = x[:, :H, :W, :].contiguous()
    return x.view(B, H * W, C)


def swin_block(sd, pre, x, hw, num_heads, ws, shift):
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'])
    x = x + shift_window_msa(sd, pre + 'attn.', y, hw, num_heads, ws, shift)
    y = F.layer_norm(x, (C,), sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias'])
    y = F.gelu(F.linear(y, sd[pre + 'ffn.layers.0.0.weight'], sd[pre + 'ffn.layers.0.0.bias']))
    return x + F.linear(y, sd[pre + 'ffn.layers.1.weight'], sd[pre + 'ffn.layers.1.bias'])


def patch_merging(sd, pre, x, hw):
    B, L, C = x.shape
    H, W = hw
    x = x.view(B, H, W, C).permute(0, 3, 1, 2)
    x = F.pad(x, (0, W % 2, 0, H % 2))                    # AdaptivePadding('corner') for kernel = stride = 2
    H, W = x.shape[-2:]
    x = F.unfold(x, kernel_size=2, stride=2).transpose(1, 2)
    x = F.layer_norm(x, (4 * C,), sd[pre + 'norm.weight'], sd[pre + 'norm.bias'])
    return F.linear(x, sd[pre + 'reduction.weight']), (H // 2, W // 2)


def swin_forward(sd, img, embed_dims=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=12, patch_size=4,
                 out_indices=(0, 1, 2, 3), prefix='', stage_taps=None):
    """img [B,3,H,W] -> tuple of [B,C_i,H_i,W_i] (SwinTransformer.forward).  stage_taps: optional list that
    receives every stage's un-normalised output tokens (for the torchvision cross-check)."""
    p = prefix
    H, W = img.shape[-2:]
    img = F.pad(img, (0, (patch_size - W % patch_size) % patch_size, 0, (patch_size - H % patch_size) % patch_size))
    x = F.conv2d(img, sd[p + 'patch_embed.projection.weight'], sd[p + 'patch_embed.projection.bias'], stride=patch_size)
    hw = tuple(x.shape[-2:])
    x = x.flatten(2).transpose(1, 2)
    if p + 'patch_embed.norm.weight' in sd:
        x = F.layer_norm(x, (embed_dims,), sd[p + 'patch_embed.norm.weight'], sd[p + 'patch_embed.norm.bias'])
    outs = []
    for i, depth in enumerate(depths):
        for j in range(depth):
            x = swin_block(sd, f'{p}stages.{i}.blocks.{j}.', x, hw, num_heads[i], window_size,
                           0 if j % 2 == 0 else window_size // 2)
        out, out_hw = x, hw
        if stage_taps is not None:
            stage_taps.append((out, out_hw))
        if i < len(depths) - 1:
            x, hw = patch_merging(sd, f'{p}stages.{i}.downsample.', x, hw)
        if i in out_indices:
            C = out.shape[-1]
            o = F.layer_norm(out, (C,), sd[f'{p}norm{i}.weight'], sd[f'{p}norm{i}.bias'])
            outs.append(o.view(-1, *out_hw, C).permute(0, 3, 1, 2).contiguous())
    return tuple(outs)


# ------------------------------------------------------------------ helpers for the tests --------
def window_attention_core(qkv, qkv_bias, bias_table, num_heads, ws, shift):
    """What ``pvsg_window_attention`` computes: qkv [B,H,W,3C] of the unpadded map -> [B,H,W,C], stated with the
    functions above (the proj linear replaced by the identity, zero-input padding = bias rows)."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    # padded rows must read qkv_bias: subtract it, zero-pad, add it back after the partition
    x = (qkv - qkv_bias).reshape(B, H * W, C3)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    q = F.pad(x.view(B, H, W, C3), (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = q.shape[1], q.shape[2]
    mask = None
    if shift > 0:
        q = torch.roll(q, shifts=(-shift, -shift), dims=(1, 2))
        img_mask = torch.zeros((1, Hp, Wp, 1))
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                img_mask[:, hs, wsl, :] = cnt
                cnt += 1
        mw = window_partition(img_mask, ws).view(-1, ws * ws)
        mask = mw.unsqueeze(1) - mw.unsqueeze(2)
        mask = mask.masked_fill(mask != 0, -100.0).masked_fill(mask == 0, 0.0)
    win = window_partition(q, ws).view(-1, ws * ws, C3) + qkv_bias
    Bw, N, _ = win.shape
    t = win.reshape(Bw, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    qq, kk, vv = t[0] * (C // num_heads) ** -0.5, t[1], t[2]
    attn = qq @ kk.transpose(-2, -1)
    bias = bias_table[relative_position_index(ws).view(-1)].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bw // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, num_heads, N, N)
    out = (attn.softmax(-1) @ vv).transpose(1, 2).reshape(Bw, N, C).view(-1, ws, ws, C)
    xo = window_reverse(out, Hp, Wp, ws)
    if shift > 0:
        xo = torch.roll(xo, shifts=(shift, shift), dims=(1, 2))
    return xo[:, :H, :W, :].contiguous()


def from_torchvision(tv_sd, depths):
    """torchvision SwinTransformer state_dict -> mmdet layout (the inverse of mmdet's swin_converter for the
    patch-merging channel order: torchvision concatenates [x(0,0), x(1,0), x(0,1), x(1,1)] tap-major, nn.Unfold is
    channel-major with taps (kh, kw) row-major)."""
    sd = {'patch_embed.projection.weight': tv_sd['features.0.0.weight'], 'patch_embed.projection.bias': tv_sd['features.0.0.bias'],
          'patch_embed.norm.weight': tv_sd['features.0.2.weight'], 'patch_embed.norm.bias': tv_sd['features.0.2.bias']}
    tap_of = [0, 2, 1, 3]       # torchvision block k -> unfold tap kh*2 + kw
    for i, depth in enumerate(depths):
        f = f'features.{2 * i + 1}.'
        for j in range(depth):
            a, b = f'{f}{j}.', f'stages.{i}.blocks.{j}.'
            for n in ('norm1', 'norm2'):
                sd[b + n + '.weight'], sd[b + n + '.bias'] = tv_sd[a + n + '.weight'], tv_sd[a + n + '.bias']
            for n in ('qkv', 'proj'):
                sd[b + f'attn.w_msa.{n}.weight'], sd[b + f'attn.w_msa.{n}.bias'] = tv_sd[a + f'attn.{n}.weight'], tv_sd[a + f'attn.{n}.bias']
            sd[b + 'attn.w_msa.relative_position_bias_table'] = tv_sd[a + 'attn.relative_position_bias_table']
            sd[b + 'ffn.layers.0.0.weight'], sd[b + 'ffn.layers.0.0.bias'] = tv_sd[a + 'mlp.0.weight'], tv_sd[a + 'mlp.0.bias']
            sd[b + 'ffn.layers.1.weight'], sd[b + 'ffn.layers.1.bias'] = tv_sd[a + 'mlp.3.weight'], tv_sd[a + 'mlp.3.bias']
        if i < len(depths) - 1:
            m = f'features.{2 * i + 2}.'
            C4 = tv_sd[m + 'norm.weight'].numel()
            C = C4 // 4
            perm = torch.empty(C4, dtype=torch.long)      # perm[mmdet index] = torchvision index
            for k in range(4):
                for c in range(C):
                    perm[c * 4 + tap_of[k]] = k * C + c
            sd[f'stages.{i}.downsample.norm.weight'] = tv_sd[m + 'norm.weight'][perm]
            sd[f'stages.{i}.downsample.norm.bias'] = tv_sd[m + 'norm.bias'][perm]
            sd[f'stages.{i}.downsample.reduction.weight'] = tv_sd[m + 'reduction.weight'][:, perm]
    return sd
