"""Segment matching on the device (csrc/assign.cu): the K x K IoU cost of two label vectors and its optimal assignment, for a
batch of shapes in two launches.  Replaces to_one_hot + relaxed_iou_fast + lapsolver.solve_dense of the reference's `match` /
`SIOU_matched_segments` (src/fitting_utils.py:362-376, src/segment_utils.py:165-174)."""
import torch

from .cabi import call
from .ops import _need_cuda, _ptr, _stream


def iou_cost(pred_bn, gt_bn, K=50):
    """pred, gt (B,N) integer labels in [0,K) on the device -> (B,K,K) float32 cost = 1 - relaxed IoU (rows: predicted)"""
    _need_cuda(pred_bn, gt_bn)
    pred = pred_bn.detach().to(torch.int32).contiguous()
    gt = gt_bn.detach().to(torch.int32).contiguous()
    B, N = pred.shape
    cost = torch.empty((B, K, K), dtype=torch.float32, device=pred.device)
    bad = torch.empty((1,), dtype=torch.int32, device=pred.device)
    call("pn_iou_cost", _ptr(pred), _ptr(gt), B, N, K, _ptr(cost), _ptr(bad), _stream())
    return cost, bad


def hungarian(cost_bnn):
    """(B,n,n) float32 costs on the device -> (B,n) int32: the column assigned to every row by the optimal assignment"""
    _need_cuda(cost_bnn)
    cost = cost_bnn.detach().to(torch.float32).contiguous()
    B, n, _ = cost.shape
    out = torch.empty((B, n), dtype=torch.int32, device=cost.device)
    call("pn_hungarian", _ptr(cost), B, n, _ptr(out), _stream())
    return out


def match_batched(pred_bn, gt_bn, K=50):
    """the reference's `match` for a batch: (B,K) column (gt segment) matched to every predicted cluster id, and the cost
    matrices; nothing is read back"""
    cost, bad = iou_cost(pred_bn, gt_bn, K)
    return hungarian(cost), cost, bad
