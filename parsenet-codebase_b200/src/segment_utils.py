"""Drop-in for the hot-path helpers of reference src/segment_utils.py: to_one_hot (:283), relaxed_iou_fast (:356),
SIOU_matched_segments (:139) with its host-side IoU bookkeeping.  Hungarian matching is
scipy.optimize.linear_sum_assignment (the reference uses lapsolver.solve_dense; same optimum, possibly a different
assignment on exact cost ties)."""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


def solve_dense(cost):
    return linear_sum_assignment(np.asarray(cost))


def to_one_hot(target, maxx=50, device_id=0):
    if isinstance(target, np.ndarray):
        target = torch.from_numpy(target.astype(np.int64)).cuda(device_id)
    out = torch.zeros((target.shape[0], maxx), device=target.device)
    return out.scatter_(1, target.long().unsqueeze(1), 1)


def relaxed_iou_fast(pred, gt, max_clusters=50):
    """pred, gt (B,N,K) soft/one-hot memberships -> (B,K,K) relaxed IoU matrix"""
    dots = pred.transpose(1, 2) @ gt
    np_ = pred.sum(1).unsqueeze(2)
    ng = gt.sum(1).unsqueeze(1)
    return dots / (np_ + ng - dots + 1e-7)


def primitive_type_segment_torch(pred, weights):
    """pred (N,L) one-hot types, weights (N,K) memberships -> (K,) majority type of every segment"""
    return torch.max(pred.t() @ weights, 0)[1]


def _merge_types(p):
    p = p.copy() if isinstance(p, np.ndarray) else p
    for src, dst in ((0, 9), (6, 9), (7, 9), (8, 2)):
        p[p == src] = dst
    return p


def mean_IOU_primitive_segment(matching, predicted_labels, labels, pred_prim, gt_prim):
    ious, prim_ious, pairs = [], [], []
    for b in range(labels.shape[0]):
        iou_b, prim_b, pairs = [], [], []
        for r, c in zip(*matching[b]):
            pi, gi = predicted_labels[b] == r, labels[b] == c
            if gi.sum() == 0 or pi.sum() == 0 or gi.sum() < 100:
                continue
            iou_b.append(np.logical_and(pi, gi).sum() / (np.logical_or(pi, gi).sum() + 1e-8))
            g_t, p_t = gt_prim[b][gi][0], pred_prim[b][r]
            prim_b.append(g_t == p_t)
            pairs.append([g_t, p_t])
        ious.append(np.mean(iou_b))
        prim_ious.append(np.mean(prim_b))
    return np.mean(ious), np.mean(prim_ious), pairs


def SIOU_matched_segments(target, pred_labels, primitives_pred, primitives, weights):
    """segment IoU + primitive-type IoU over Hungarian-matched (predicted, gt) segments.
    NOTE: like the reference, the primitive-id arrays are merged in place (0,6,7 -> 9; 8 -> 2)."""
    for arr in (primitives, primitives_pred):
        for src, dst in ((0, 9), (6, 9), (7, 9), (8, 2)):
            arr[arr == src] = dst
    dev = weights.device.index if weights.is_cuda else 0
    lab_hot, clu_hot = to_one_hot(target, device_id=dev), to_one_hot(pred_labels, device_id=dev)
    cost = 1.0 - relaxed_iou_fast(clu_hot.unsqueeze(0).float(), lab_hot.unsqueeze(0).float()).data.cpu().numpy()
    matching = [list(solve_dense(cost[0]))]
    prim_hot = to_one_hot(primitives_pred, 10, dev).float()
    prim_pred = primitive_type_segment_torch(prim_hot, weights).data.cpu().numpy()
    s_iou, p_iou, pairs = mean_IOU_primitive_segment(matching, pred_labels[None], target[None], prim_pred[None],
                                                     primitives[None])
    return s_iou, p_iou, matching, pairs
