"""CPU oracle of the training slice -- TEST INFRASTRUCTURE (oracle/__init__.py).

Restates, in plain torch with autograd, ``Mask2FormerVideoHead.loss_single`` / ``_get_target_single``
(models/mask2former_vps/mask2former_video_head.py:162-293) together with the mmcv 1.4 / mmdet 2.25 pieces it calls
(absent from /root/reference: **parity unpinned by the reference**, restated from the pinned versions' published
algorithm; ``point_sample`` = ``F.grid_sample(2p - 1, align_corners=False)``).  Gradients come from torch autograd.
Independent pin: dice / sigmoid-BCE losses, the three assignment costs and the point sampling equal the implementation of
the same published losses in HF transformers' Mask2Former, values and gradients
(tests/test_oracle_golden.py::test_training_losses_vs_hf_mask2former).
"""
import torch
import torch.nn.functional as F


def point_sample(input, points, align_corners=False):
    """mmcv.ops.point_sample: input [N,C,H,W], points [N,P,2] in [0,1] -> [N,C,P]."""
    out = F.grid_sample(input, (2.0 * points - 1.0).unsqueeze(2), align_corners=align_corners)
    return out.squeeze(3)


def cross_entropy_loss(cls_scores, labels, class_weight, loss_weight=2.0):
    """mmdet CrossEntropyLoss(use_sigmoid=False, reduction='mean', class_weight) called with
    avg_factor=class_weight[labels].sum() (loss_single :236-244)."""
    loss = F.cross_entropy(cls_scores, labels, weight=class_weight, reduction='none')
    return loss_weight * loss.sum() / class_weight[labels].sum()


def mask_bce_loss(pred, target, avg_factor, loss_weight=5.0):
    """mmdet CrossEntropyLoss(use_sigmoid=True, reduction='mean') on flattened point logits."""
    return loss_weight * F.binary_cross_entropy_with_logits(pred, target.float(), reduction='none').sum() / avg_factor


def dice_loss(pred, target, avg_factor, eps=1.0, loss_weight=5.0):
    """mmdet DiceLoss(use_sigmoid=True, activate=True, naive_dice=True, eps=1.0, reduction='mean')."""
    inp = pred.sigmoid().flatten(1)
    tgt = target.flatten(1).float()
    a = torch.sum(inp * tgt, 1)
    d = (2 * a + eps) / (torch.sum(inp, 1) + torch.sum(tgt, 1) + eps)
    return loss_weight * (1 - d).sum() / avg_factor


def match_cost(cls_score, gt_labels, pred_pts, gt_pts, w_cls=2.0, w_mask=5.0, w_dice=5.0, eps=1.0):
    """mmdet MaskHungarianAssigner: ClassificationCost + CrossEntropyLossCost(use_sigmoid) + DiceCost(pred_act, naive_dice)."""
    cls_cost = -cls_score.softmax(-1)[:, gt_labels] * w_cls
    n = pred_pts.shape[1]
    pos = F.binary_cross_entropy_with_logits(pred_pts, torch.ones_like(pred_pts), reduction='none')
    neg = F.binary_cross_entropy_with_logits(pred_pts, torch.zeros_like(pred_pts), reduction='none')
    mask_cost = (torch.einsum('nc,mc->nm', pos, gt_pts) + torch.einsum('nc,mc->nm', neg, 1 - gt_pts)) / n * w_mask
    p = pred_pts.sigmoid()
    num = 2 * torch.einsum('nc,mc->nm', p, gt_pts)
    den = p.sum(-1)[:, None] + gt_pts.sum(-1)[None, :]
    dcost = (1 - (num + eps) / (den + eps)) * w_dice
    return cls_cost + mask_cost + dcost


def loss_single(cls_scores, mask_preds, gt_labels_list, gt_masks_list, assign_points, loss_points_fn, num_classes=126,
                class_weight=None, loss_weights=(2.0, 5.0, 5.0)):
    """loss_single :196-293 with the random point sets supplied by the caller: assign_points [1,K,2];
    loss_points_fn(n_pos) -> [n_pos,K,2].  Returns (loss_cls, loss_mask, loss_dice, labels, pos indices per image)."""
    from scipy.optimize import linear_sum_assignment
    B, Q, _ = cls_scores.shape
    if class_weight is None:
        class_weight = torch.ones(num_classes + 1)
        class_weight[-1] = 0.1
    labels = torch.full((B, Q), num_classes, dtype=torch.long)
    pos_pred, pos_tgt, pos_all = [], [], []
    for b in range(B):
        gt_masks = gt_masks_list[b].flatten(1, 2)
        mask_pred = mask_preds[b].transpose(1, 0).flatten(1, 2)
        G = gt_labels_list[b].shape[0]
        pred_pts = point_sample(mask_pred.unsqueeze(1), assign_points.repeat(Q, 1, 1)).squeeze(1)
        gt_pts = point_sample(gt_masks.unsqueeze(1).float(), assign_points.repeat(G, 1, 1)).squeeze(1)
        cost = match_cost(cls_scores[b].detach(), gt_labels_list[b], pred_pts.detach(), gt_pts)
        rows, cols = linear_sum_assignment(cost.numpy())
        rows, cols = torch.as_tensor(rows), torch.as_tensor(cols)
        order = torch.argsort(rows)
        rows, cols = rows[order], cols[order]
        labels[b, rows] = gt_labels_list[b][cols]
        pos_pred.append(mask_pred[rows])
        pos_tgt.append(gt_masks[cols])
        pos_all.append((rows, cols))
    loss_cls = cross_entropy_loss(cls_scores.flatten(0, 1), labels.flatten(), class_weight, loss_weights[0])
    mask_pos, mask_targets = torch.cat(pos_pred, 0), torch.cat(pos_tgt, 0)
    ntm = max(float(mask_pos.shape[0]), 1.0)
    pts = loss_points_fn(mask_pos.shape[0])
    with torch.no_grad():
        point_targets = point_sample(mask_targets.unsqueeze(1).float(), pts).squeeze(1)
    point_preds = point_sample(mask_pos.unsqueeze(1), pts).squeeze(1)
    loss_d = dice_loss(point_preds, point_targets, ntm, 1.0, loss_weights[2])
    K = point_preds.shape[1]
    loss_m = mask_bce_loss(point_preds.reshape(-1), point_targets.reshape(-1), ntm * K, loss_weights[1])
    return loss_cls, loss_m, loss_d, labels, pos_all
