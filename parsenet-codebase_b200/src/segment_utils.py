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
    """segment IoU and primitive-type agreement over the matched (predicted r, gt c) pairs (reference :66-112).
    Intersections / unions come from one confusion matrix per shape (exact integer counts) instead of 2 x 50 boolean
    passes over the points; the gt type of a segment is the type of its first point, as in the reference."""
    ious, prim_ious, pairs = [], [], []
    for b in range(labels.shape[0]):
        iou_b, prim_b, pairs = [], [], []
        pl, gl = np.asarray(predicted_labels[b]).astype(np.int64), np.asarray(labels[b]).astype(np.int64)
        K = int(max(pl.max(), gl.max(), max(matching[b][0]), max(matching[b][1]))) + 1
        conf = np.bincount(pl * K + gl, minlength=K * K).reshape(K, K)
        n_p, n_g = conf.sum(1), conf.sum(0)
        first = np.full(K, -1, dtype=np.int64)
        first[gl[::-1]] = np.arange(gl.shape[0] - 1, -1, -1)      # last write wins -> lowest index of every label
        for r, c in zip(*matching[b]):
            if n_g[c] == 0 or n_p[r] == 0 or n_g[c] < 100:
                continue
            inter = conf[r, c]
            iou_b.append(inter / (n_p[r] + n_g[c] - inter + 1e-8))
            g_t, p_t = gt_prim[b][first[c]], pred_prim[b][r]
            prim_b.append(g_t == p_t)
            pairs.append([g_t, p_t])
        ious.append(np.mean(iou_b))
        prim_ious.append(np.mean(prim_b))
    return np.mean(ious), np.mean(prim_ious), pairs


def siou_prepare(target, pred_labels, primitives, matching):
    """the part of SIOU_matched_segments that needs no predicted segment types: the segment IoU of one shape and, per counted
    pair, (gt type, predicted cluster id).  The training step computes this while the fit kernels still run and finishes with
    siou_finish once the predicted types have been read back.  `primitives` is merged in place like SIOU_matched_segments does."""
    lut = np.arange(max(int(primitives.max()) + 1, 10))
    for src, dst in _MERGE:
        lut[src] = dst
    primitives[...] = lut[primitives]
    pl, gl = np.asarray(pred_labels).astype(np.int64), np.asarray(target).astype(np.int64)
    K = int(max(pl.max(), gl.max(), max(matching[0]), max(matching[1]))) + 1
    conf = np.bincount(pl * K + gl, minlength=K * K).reshape(K, K)
    n_p, n_g = conf.sum(1), conf.sum(0)
    first = np.full(K, -1, dtype=np.int64)
    first[gl[::-1]] = np.arange(gl.shape[0] - 1, -1, -1)
    iou_b, todo = [], []
    for r, c in zip(*matching):
        if n_g[c] == 0 or n_p[r] == 0 or n_g[c] < 100:
            continue
        inter = conf[r, c]
        iou_b.append(inter / (n_p[r] + n_g[c] - inter + 1e-8))
        todo.append((primitives[first[c]], r))
    return np.mean(iou_b), todo


def siou_finish(prepared, prim_pred_seg):
    """(segment IoU, primitive-type IoU) of one shape from siou_prepare's result and the (K,) predicted segment types"""
    s_iou, todo = prepared
    return np.mean([s_iou]), np.mean(np.mean([g_t == prim_pred_seg[r] for g_t, r in todo]))


_MERGE = ((0, 9), (6, 9), (7, 9), (8, 2))          # closed splines -> 9, open splines -> 2 (reference :151-159)
_MERGE_LUT = {}


def iou_cost_host(pred_labels, target, K=50):
    """1 - relaxed IoU between the one-hot memberships of two label vectors, from their K x K confusion matrix.
    Bit-identical to `1 - relaxed_iou_fast(to_one_hot(pred), to_one_hot(target))`: every count is an integer
    < 2^24, so the fp32 matmul of the reference is exact and the remaining arithmetic is the same fp32 expression."""
    pred_labels = np.asarray(pred_labels).astype(np.int64)
    target = np.asarray(target).astype(np.int64)
    if pred_labels.min() < 0 or target.min() < 0 or pred_labels.max() >= K or target.max() >= K:
        raise ValueError(f"labels must lie in [0, {K})")
    conf = np.bincount(pred_labels * K + target, minlength=K * K).reshape(K, K).astype(np.float32)
    n_p, n_g = conf.sum(1, keepdims=True), conf.sum(0, keepdims=True)
    return np.float32(1.0) - conf / (n_p + n_g - conf + np.float32(1e-7))


def segment_types_device(prim_pred_point, weights_kn):
    """per-point predicted primitive ids (N,) int64 + membership weights (K,N) -> (K,) majority type of every cluster
    (reference: to_one_hot(primitives_pred, 10) after the type merge, primitive_type_segment_torch).  Device only."""
    dev = weights_kn.device
    lut = _MERGE_LUT.get(str(dev))
    if lut is None:
        table = np.arange(10)
        for src, dst in _MERGE:
            table[src] = dst
        lut = _MERGE_LUT[str(dev)] = torch.from_numpy(table).to(dev)
    merged = lut[prim_pred_point]
    hot = torch.zeros((merged.shape[0], 10), device=dev).scatter_(1, merged.unsqueeze(1), 1.0)
    return torch.max(hot.t() @ weights_kn.t(), 0)[1]


def segment_types_batched(prim_pred_bn, weights_bns):
    """segment_types_device for a batch: per-point predicted ids (B,N) int64 + similarities (B,N,S) -> (B,S) majority type
    of every weight column (one bmm for all shapes)"""
    dev = weights_bns.device
    lut = _MERGE_LUT.get(str(dev))
    if lut is None:
        table = np.arange(10)
        for src, dst in _MERGE:
            table[src] = dst
        lut = _MERGE_LUT[str(dev)] = torch.from_numpy(table).to(dev)
    merged = lut[prim_pred_bn]
    B, N = merged.shape
    hot = torch.zeros((B, N, 10), device=dev).scatter_(2, merged.unsqueeze(2), 1.0)
    return torch.max(torch.bmm(hot.transpose(1, 2), weights_bns), 1)[1]


def SIOU_matched_segments(target, pred_labels, primitives_pred, primitives, weights, prim_pred_seg=None,
                          matching=None):
    """segment IoU + primitive-type IoU over Hungarian-matched (predicted, gt) segments.
    NOTE: like the reference, the primitive-id arrays are merged in place (0,6,7 -> 9; 8 -> 2).
    prim_pred_seg: optional precomputed (K,) majority type per cluster (segment_types_device), in which case
    primitives_pred / weights are not needed and nothing touches the device.
    matching: optional (rows, cols) of the Hungarian solve on the same (pred_labels, target) pair (fitting_utils.match
    computes exactly this cost matrix), to avoid solving it twice."""
    for arr in (primitives, primitives_pred):
        if arr is None:
            continue
        lut = np.arange(max(int(arr.max()) + 1, 10))
        for src, dst in _MERGE:
            lut[src] = dst
        arr[...] = lut[arr]
    matching = [list(matching)] if matching is not None else [list(solve_dense(iou_cost_host(pred_labels, target)))]
    if prim_pred_seg is None:
        dev = weights.device
        pp = torch.from_numpy(np.asarray(primitives_pred).astype(np.int64)).to(dev)
        prim_pred_seg = segment_types_device(pp, weights.t()).data.cpu().numpy()
    s_iou, p_iou, pairs = mean_IOU_primitive_segment(matching, pred_labels[None], target[None], prim_pred_seg[None],
                                                     primitives[None])
    return s_iou, p_iou, matching, pairs


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
