"""Training slice of the head: losses of ``Mask2FormerVideoHead.loss_single`` and the backward of
the pixel decoder's deformable attention, as ``torch.autograd.Function``s over libpvsg_sm100.so kernels.

Reference: ``models/mask2former_vps/mask2former_video_head.py:162-293`` (``_get_target_single``, ``loss_single``);
third-party pieces it calls (absent from the reference tree, restated from the pinned versions' published algorithm):
mmcv 1.4 ``ops.point_sample`` and ``MultiScaleDeformableAttnFunction``; mmdet 2.25 ``CrossEntropyLoss``, ``DiceLoss``
(``naive_dice=True, eps=1.0``), ``MaskHungarianAssigner`` with ``ClassificationCost`` / ``CrossEntropyLossCost`` / ``DiceCost``,
``MaskPseudoSampler``, ``get_uncertain_point_coords_with_randomness``; loss weights and ``class_weight`` from
``configs/mask2former_vps/mask2former_video_r50_base.py:89-127``.

This is the FIRST slice of the training row: forward + backward of the losses w.r.t. the head's outputs
(``cls_scores``, ``mask_preds``) and of MSDeformAttn w.r.t. its inputs.  The backward of the GEMM / attention engine
(the rest of ``forward_train``) is not built; ``Mask2FormerVideoHead.forward_train`` still raises.
"""
import numpy as np
import torch

from . import ops


class MultiScaleDeformableAttnFunction(torch.autograd.Function):
    """mmcv.ops.multi_scale_deform_attn.MultiScaleDeformableAttnFunction (forward: pvsg_msda_forward; backward:
    pvsg_msda_backward).  value [B,N,H,D], spatial_shapes [(h,w)], sampling_locations [B,Nq,H,L,P,2],
    attention_weights [B,Nq,H,L,P] -> [B,Nq,H*D]."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, sampling_locations, attention_weights):
        shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
        ctx.shapes = shapes
        ctx.save_for_backward(value, sampling_locations, attention_weights)
        return ops.msda_forward(value, shapes, sampling_locations, attention_weights)

    @staticmethod
    def backward(ctx, grad_output):
        value, loc, aw = ctx.saved_tensors
        gv, gl, ga = ops.msda_backward(value, ctx.shapes, loc, aw, grad_output.contiguous())
        return gv, None, gl, ga


class _PointSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, points):
        ctx.hw = tuple(maps.shape[-2:])
        ctx.save_for_backward(points)
        return ops.point_sample(maps, points)

    @staticmethod
    def backward(ctx, grad_out):
        (points,) = ctx.saved_tensors
        return ops.point_sample_backward(grad_out.contiguous(), points, ctx.hw), None


def point_sample(input, points, align_corners=False):
    """mmcv.ops.point_sample for single-channel maps: input [n,1,H,W] (or [n,H,W]), points [n,K,2] in [0,1] -> [n,1,K]."""
    if align_corners:
        raise NotImplementedError('point_sample: align_corners=False (the reference call)')
    squeeze = input.dim() == 4
    if squeeze and input.shape[1] != 1:
        raise NotImplementedError('point_sample: single-channel maps')
    out = _PointSample.apply(input[:, 0] if squeeze else input, points)
    return out[:, None] if squeeze else out


class _MaskPointLosses(torch.autograd.Function):
    """(loss_mask, loss_dice) of loss_single :270-289 from sampled logits / targets [n,K]."""

    @staticmethod
    def forward(ctx, logits, targets, num_total_masks, w_mask, w_dice, eps):
        n, K = logits.shape
        sums, _ = ops.mask_point_losses(logits, targets, eps)
        ctx.save_for_backward(logits, targets)
        ctx.cfg = (float(num_total_masks), float(w_mask), float(w_dice), float(eps), K)
        s = sums.cpu()                     # two scalars; the reference's reductions end in python floats too (.item())
        ntm = float(num_total_masks)
        return (logits.new_tensor(float(s[0]) / (ntm * K) * w_mask), logits.new_tensor(float(s[1]) / ntm * w_dice))

    @staticmethod
    def backward(ctx, g_mask, g_dice):
        logits, targets = ctx.saved_tensors
        ntm, w_mask, w_dice, eps, K = ctx.cfg
        _, grad = ops.mask_point_losses(logits, targets, eps, float(g_mask) * w_mask / (ntm * K), float(g_dice) * w_dice / ntm,
                                        want_grad=True)
        return grad, None, None, None, None, None


class _WeightedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_weight, label_weight, loss_weight):
        sums, _ = ops.weighted_ce(logits, labels, class_weight, label_weight)
        s = sums.cpu()
        ctx.save_for_backward(logits, labels, class_weight, label_weight)
        ctx.cfg = (float(s[1]), float(loss_weight))
        return logits.new_tensor(float(s[0]) / float(s[1]) * loss_weight)       # avg_factor = class_weight[labels].sum()

    @staticmethod
    def backward(ctx, g):
        logits, labels, cw, lw = ctx.saved_tensors
        avg, w = ctx.cfg
        _, grad = ops.weighted_ce(logits, labels, cw, lw, float(g) * w / avg, want_grad=True)
        return grad, None, None, None, None


def get_uncertain_point_coords_with_randomness(mask_pred, labels, num_points, oversample_ratio, importance_sample_ratio,
                                               generator=None):
    """mmdet.models.utils.point_sample: oversample random points, keep the most uncertain ones (-|logit|) plus random
    ones.  mask_pred [n,1,H,W] -> [n,num_points,2].  (Random draws and top-k are torch calls: data generation /
    selection, no arithmetic on the path's values.)"""
    n = mask_pred.shape[0]
    dev = mask_pred.device
    num_sampled = int(num_points * oversample_ratio)
    coords = torch.rand(n, num_sampled, 2, device=dev, generator=generator)
    logits = ops.point_sample(mask_pred[:, 0].detach(), coords)
    num_uncertain = int(importance_sample_ratio * num_points)
    idx = torch.topk(-logits.abs(), k=num_uncertain, dim=1)[1]
    picked = torch.gather(coords, 1, idx[..., None].expand(-1, -1, 2))
    if num_points - num_uncertain > 0:
        picked = torch.cat([picked, torch.rand(n, num_points - num_uncertain, 2, device=dev, generator=generator)], 1)
    return picked


def reduce_mean(value, device):
    """mmdet.core.reduce_mean (mask2former_video_head.py:246): the positive count averaged over the ranks, so that every
    rank normalises its mask losses by the same factor; the identity in a single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float32, device=device)
    dist.all_reduce(t)
    return float(t.item()) / dist.get_world_size()


def hungarian_assign(cls_score, mask_points_pred, gt_labels, gt_points_masks, w_cls=2.0, w_mask=5.0, w_dice=5.0, eps=1.0):
    """mmdet MaskHungarianAssigner.assign + MaskPseudoSampler: -> (pos_inds, pos_assigned_gt_inds) int64, sorted by query."""
    from scipy.optimize import linear_sum_assignment
    dev = cls_score.device
    if gt_labels.numel() == 0:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z
    cost = ops.mask_match_cost(cls_score, gt_labels, mask_points_pred, gt_points_masks, w_cls, w_mask, w_dice, eps)
    rows, cols = linear_sum_assignment(cost.detach().cpu().numpy())
    order = np.argsort(rows)
    return torch.as_tensor(rows[order], device=dev), torch.as_tensor(cols[order], device=dev)


def loss_single(cls_scores, mask_preds, gt_labels_list, gt_masks_list, img_metas=None, num_classes=126, class_weight=None,
                num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75, loss_weights=(2.0, 5.0, 5.0),
                assign_points=None, loss_points=None, generator=None):
    """mask2former_video_head.py:196-293 (point_loss=True, single process: reduce_mean is the identity).
    cls_scores [B,Q,C+1]; mask_preds [B,T,Q,h,w]; gt_masks_list[b] [G_b,T,h,w]; gt_labels_list[b] int64 [G_b].
    assign_points [1,K,2] / loss_points [n_pos,K,2]: fix the two random point sets (tests)."""
    B, Q, C1 = cls_scores.shape
    dev = cls_scores.device
    if class_weight is None:
        class_weight = torch.ones(num_classes + 1, device=dev)
        class_weight[-1] = 0.1
    from scipy.optimize import linear_sum_assignment
    labels = torch.full((B, Q), num_classes, dtype=torch.int64, device=dev)
    # phase 1 (device, no host round trip): sampled points and the assignment cost matrix of every clip
    all_pred = mask_preds.transpose(2, 1).flatten(2, 3)                         # [B, Q, T*h, w] (one copy, on the tape)
    preds, gts, costs = [], [], []
    for b in range(B):
        g = gt_masks_list[b]
        gt_masks = (g if g.is_floating_point() else g.float()).flatten(1, 2)    # [G, T*h, w]: frames as one long image
        mask_pred = all_pred[b]                                                 # [Q, T*h, w]
        preds.append(mask_pred)
        gts.append(gt_masks)
        if gt_masks.shape[0] == 0:
            costs.append(None)
            continue
        pts = assign_points if assign_points is not None else torch.rand(1, num_points, 2, device=dev, generator=generator)
        pred_pts = ops.point_sample(mask_pred.detach().contiguous(), pts[0])
        gt_pts = ops.point_sample(gt_masks.contiguous(), pts[0])
        costs.append(ops.mask_match_cost(cls_scores[b].detach(), gt_labels_list[b], pred_pts, gt_pts, *loss_weights, 1.0))
    # phase 2: ONE device->host transfer of all cost matrices, the assignments on the host (scipy, as mmdet's
    # MaskHungarianAssigner + MaskPseudoSampler), ONE transfer of the index lists back
    pos_flat, gt_flat, n_gt = [], [], 0
    if any(c is not None for c in costs):
        host = torch.cat([c.reshape(-1) for c in costs if c is not None]).cpu().numpy()
        off = 0
        for b in range(B):
            G = gts[b].shape[0]
            if costs[b] is not None:
                rows, cols = linear_sum_assignment(host[off:off + Q * G].reshape(Q, G))
                off += Q * G
                order = np.argsort(rows)
                pos_flat.append(b * Q + rows[order])
                gt_flat.append(n_gt + cols[order])
            n_gt += G
    if pos_flat:
        idx = torch.as_tensor(np.stack([np.concatenate(pos_flat), np.concatenate(gt_flat)]), device=dev)
        pos_idx, gt_idx = idx[0], idx[1]
        labels.view(-1)[pos_idx] = torch.cat(list(gt_labels_list))[gt_idx]
        mask_pos = all_pred.flatten(0, 1)[pos_idx]                              # [n_pos, T*h, w]
        mask_targets = torch.cat(gts, 0)[gt_idx] if len(gts) > 1 else gts[0][gt_idx]
    else:
        mask_pos = preds[0][:0]
        mask_targets = gts[0][:0]
    loss_cls = _WeightedCE.apply(cls_scores.flatten(0, 1), labels.flatten(), class_weight, None, loss_weights[0])
    num_total_masks = max(reduce_mean(float(mask_pos.shape[0]), dev), 1.0)
    if mask_targets.shape[0] == 0:
        zero = mask_pos.sum()
        return loss_cls, zero, zero
    with torch.no_grad():
        pts = loss_points if loss_points is not None else get_uncertain_point_coords_with_randomness(
            mask_pos.unsqueeze(1), None, num_points, oversample_ratio, importance_sample_ratio, generator)
        point_targets = ops.point_sample(mask_targets.contiguous(), pts)
    point_preds = _PointSample.apply(mask_pos.contiguous(), pts)
    loss_mask, loss_dice = _MaskPointLosses.apply(point_preds, point_targets, num_total_masks, loss_weights[1], loss_weights[2], 1.0)
    return loss_cls, loss_mask, loss_dice
