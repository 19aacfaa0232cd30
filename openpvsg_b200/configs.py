"""Model dicts of the reference's shipped configurations, restated as a function.

The reference's config FILES are the API and are consumed unmodified when present
(``registry.load_config('configs/mask2former_vps/mask2former_video_r50_single_video_test.py')``;
tests/test_registry.py checks that path against /root/reference).  They do not travel to
the GPU box, so tests / bench / smoke build the same dict from here.
Values: configs/mask2former_vps/mask2former_video_r50_base.py:1-142 and
mask2former_video_r50_single_video_test.py:42-59; IPS:
configs/mask2former/mask2former_r50_lsj_8x2_50e_coco-panoptic_custom_single_video_test.py.
"""


def mask2former_r50(video=True, num_things=115, num_stuff=11, num_queries=100, instance_on=True):
    enc_layer = dict(
        type='BaseTransformerLayer',
        attn_cfgs=dict(type='MultiScaleDeformableAttention', embed_dims=256, num_heads=8, num_levels=3,
                       num_points=4, im2col_step=64, dropout=0.0, batch_first=False, norm_cfg=None, init_cfg=None),
        ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.0,
                      act_cfg=dict(type='ReLU', inplace=True)),
        operation_order=('self_attn', 'norm', 'ffn', 'norm'))
    dec_layer = dict(
        type='DetrTransformerDecoderLayer',
        attn_cfgs=dict(type='MultiheadAttention', embed_dims=256, num_heads=8, attn_drop=0.0, proj_drop=0.0,
                       dropout_layer=None, batch_first=False),
        ffn_cfgs=dict(embed_dims=256, feedforward_channels=2048, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                      ffn_drop=0.0, dropout_layer=None, add_identity=True),
        feedforward_channels=2048,
        operation_order=('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'))
    return dict(
        type='Mask2FormerVideoCustom' if video else 'Mask2FormerCustom',
        backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=-1,
                      norm_cfg=dict(type='SyncBN' if video else 'BN', requires_grad=video), norm_eval=True,
                      style='pytorch'),
        panoptic_head=dict(
            type='Mask2FormerVideoHead' if video else 'Mask2FormerHeadCustom',
            in_channels=[256, 512, 1024, 2048], strides=[4, 8, 16, 32], feat_channels=256, out_channels=256,
            num_things_classes=num_things, num_stuff_classes=num_stuff, num_queries=num_queries,
            num_transformer_feat_level=3,
            pixel_decoder=dict(type='MSDeformAttnPixelDecoder', num_outs=3, norm_cfg=dict(type='GN', num_groups=32),
                               act_cfg=dict(type='ReLU'),
                               encoder=dict(type='DetrTransformerEncoder', num_layers=6, transformerlayers=enc_layer,
                                            init_cfg=None),
                               positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True),
                               init_cfg=None),
            enforce_decoder_input_project=False,
            positional_encoding=dict(type='SinePositionalEncoding3D' if video else 'SinePositionalEncoding',
                                     num_feats=128, normalize=True),
            transformer_decoder=dict(type='DetrTransformerDecoder', return_intermediate=True, num_layers=9,
                                     transformerlayers=dec_layer, init_cfg=None),
            loss_cls=None, loss_mask=None, loss_dice=None),
        panoptic_fusion_head=dict(type='MaskFormerFusionHeadCustom', num_things_classes=num_things,
                                  num_stuff_classes=num_stuff, loss_panoptic=None, init_cfg=None),
        train_cfg=None,
        test_cfg=dict(panoptic_on=True, semantic_on=False, instance_on=instance_on, max_per_image=100,
                      iou_thr=0.8, filter_low_score=True, object_mask_thr=0.8, return_query=True),
        init_cfg=None)


SWIN_B = dict(embed_dims=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=12)


def mask2former_swin(video=True, swin=None, **kwargs):
    """Mask2Former(-VPS) with a Swin backbone (BASELINE configs[2]: Swin-B).  Not a reference config file -- the
    reference ships R50 only; the backbone block and ``in_channels`` follow mmdet 2.25.0
    ``configs/mask2former/mask2former_swin-b-p4-w12-384_lsj_8x2_50e_coco-panoptic.py`` (via its swin-t base),
    everything else is the R50 video config above."""
    sw = dict(SWIN_B, **(swin or {}))
    cfg = mask2former_r50(video, **kwargs)
    cfg['backbone'] = dict(type='SwinTransformer', pretrain_img_size=384, embed_dims=sw['embed_dims'], depths=sw['depths'],
                           num_heads=sw['num_heads'], window_size=sw['window_size'], mlp_ratio=4, qkv_bias=True,
                           qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.3, patch_norm=True,
                           out_indices=(0, 1, 2, 3), with_cp=False, convert_weights=True, frozen_stages=-1, init_cfg=None)
    cfg['panoptic_head']['in_channels'] = [sw['embed_dims'] * 2 ** i for i in range(4)]
    return cfg
