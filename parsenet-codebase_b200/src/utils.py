"""Drop-in for the hot-path part of reference src/utils.py: Chamfer distances (:273-358), rescale_input_outputs
(:361-390), grad_norm (:393).  The nearest-neighbour searches run in csrc/chamfer.cu (argmin kept for backward)."""
import numpy as np
import torch

from pnb200.fitting import nearest_sqdist
from src.guard import guard_sqrt


def _as_cuda(t):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t.astype(np.float32)).cuda()
    return t


def chamfer_distance(pred, gt, sqrt=False):
    """pred (B,Np,3), gt (B,M,3) -> mean over the batch of (mean_i min_j + mean_j min_i) / 2"""
    pred, gt = _as_cuda(pred), _as_cuda(gt)
    d_pred = nearest_sqdist(pred, gt)          # per predicted point: nearest gt   (B,Np)
    d_gt = nearest_sqdist(gt, pred)            # per gt point: nearest prediction  (B,M)
    if sqrt:
        d_pred, d_gt = guard_sqrt(d_pred), guard_sqrt(d_gt)
    return torch.mean(d_pred.mean(1) + d_gt.mean(1)) / 2.0


def chamfer_argmin(pred, gt):
    """index of the nearest gt point of every predicted point, (B,Np) int32 (the matcher="nearest" option of the post-fit
    optimisers)"""
    from pnb200.fitting import nearest_index
    return nearest_index(_as_cuda(pred), _as_cuda(gt))


def chamfer_distance_one_side(pred, gt, side=1):
    """side 0: every predicted point to its nearest gt; side 1: every gt point to its nearest prediction"""
    pred, gt = _as_cuda(pred), _as_cuda(gt)
    d = nearest_sqdist(pred, gt) if side == 0 else nearest_sqdist(gt, pred)
    return torch.mean(d.mean(1))


def chamfer_distance_single_shape(pred, gt, one_side=False, sqrt=False, reduce=True):
    """pred (Np,3), gt (M,3).  one_side: gt -> nearest prediction only.  reduce=False returns per-point vectors."""
    pred, gt = _as_cuda(pred), _as_cuda(gt)
    cd_gt = nearest_sqdist(gt.unsqueeze(0), pred.unsqueeze(0))[0]          # (M,)  min over predictions
    if sqrt:
        cd_gt = guard_sqrt(cd_gt)
    if one_side:
        return cd_gt.mean(0) if reduce else cd_gt
    cd_pred = nearest_sqdist(pred.unsqueeze(0), gt.unsqueeze(0))[0]        # (Np,) min over gt
    if sqrt:
        cd_pred = guard_sqrt(cd_pred)
    if reduce:
        return (cd_pred.mean() + cd_gt.mean()) / 2.0
    return (cd_pred + cd_gt) / 2.0


def rescale_input_outputs(scales, output, points, control_points, batch_size):
    """undo the anisotropic normalisation (per-axis extents `scales`) relative to the largest extent"""
    s = torch.from_numpy(np.stack(scales, 0).astype(np.float32)).to(output.device).reshape(batch_size, 1, 3)
    smax = s.reshape(batch_size, 3).max(1)[0]
    output = output * s / smax.reshape(batch_size, 1, 1)
    points = points * s.reshape(batch_size, 3, 1) / smax.reshape(batch_size, 1, 1)
    control_points = control_points * s.reshape(batch_size, 1, 1, 3) / smax.reshape(batch_size, 1, 1, 1)
    return s, output, points, control_points


def grad_norm(model):
    total = sum(p.grad.data.norm(2) for p in model.parameters()).item()
    return bool(np.isnan(total) or np.isinf(total))


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
