"""Drop-in for reference src/residual_utils.py Evaluation (:49-378): embedding -> mean-shift clusters -> membership
weights -> Hungarian match to gt -> per-segment primitive / spline fit -> residual loss (train mode).
residual_eval_mode (:210, open3d outlier removal / up-sampling / meshes) is outside the hot path."""
import numpy as np
import torch
from scipy import stats

from src.fitting_optimization import FittingModule
from src.fitting_utils import match, to_one_hot, weights_normalize
from src.mean_shift import MeanShift
from src.primitive_forward import fit_one_shape_torch
from src.primitives import ResidualLoss
from src.segment_utils import SIOU_matched_segments


def convert_to_one_hot(data):
    return to_one_hot(torch.max(data, 1)[1], data.shape[1], device_id=data.device.index or 0).float()


class Evaluation:
    def __init__(self, userspace=None, closed_path=None, open_path=None, open_decoder=None, closed_decoder=None):
        """open_decoder / closed_decoder: optional pre-built SplineNets (instead of loading
        logs/pretrained_models/{open,closed}_spline.pth); they are frozen either way."""
        closed_path = closed_path or "logs/pretrained_models/closed_spline.pth"
        open_path = open_path or "logs/pretrained_models/open_spline.pth"
        self.res_loss = ResidualLoss()
        self.fitter = FittingModule(closed_path, open_path, open_decoder, closed_decoder)
        for net in (self.fitter.closed_control_decoder, self.fitter.open_control_decoder):
            for p in net.parameters():
                p.requires_grad = False
        self.ms = MeanShift()

    def guard_mean_shift(self, embedding, quantile, iterations, kernel_type="gaussian"):
        """grow the quantile by 1.2x until at most 49 clusters come out (reference :69-84)"""
        while True:
            _, center, bandwidth, cluster_ids = self.ms.mean_shift(embedding, 10000, quantile, iterations,
                                                                   kernel_type=kernel_type)
            if torch.unique(cluster_ids).shape[0] > 49:
                quantile *= 1.2
            else:
                break
        return center, bandwidth, cluster_ids

    def fitting_loss(self, embedding, points, normals, labels, primitives, primitives_log_prob, quantile=0.125,
                     iterations=5, lamb=1.0, debug=False, eval=False):
        """embedding (B,N,d), points/normals (B,N,3), labels/primitives numpy (B,N), log-probs (B,P,N).
        Returns ([loss, geometric mean, spline mean, seg IoU, type IoU] per shape, concatenated in order,
        [parameters, cluster ids, weights] of the LAST shape) — the reference is only ever called with B = 1.
        The mean-shift iterations of all shapes run as ONE batched launch sequence (per-shape bandwidths)."""
        if eval:
            raise NotImplementedError("Evaluation.fitting_loss(eval=True) is outside the hot path")
        from pnb200 import meanshift as _ms
        from pnb200.losses import l2_normalize
        B = embedding.shape[0]
        embedding = l2_normalize(embedding)
        prim_pred = torch.max(primitives_log_prob, 1)[1].data.cpu().numpy()
        with torch.no_grad():
            bws = torch.clamp(_ms.compute_bandwidth_batched(embedding, 10000, quantile), min=_ms.BW_FLOOR)
        shifted = _ms.mean_shift_iters(embedding, bws, iterations)
        with torch.no_grad():
            members = _ms.nearest_center_batched(embedding, shifted)
        out, parameters, cluster_ids, weights = [], None, None, None
        for b in range(B):
            with torch.no_grad():
                _, ids, cluster_ids = _ms.nms(shifted[b], embedding[b], bws[b], member=members[b])
            center, bandwidth = shifted[b][ids], bws[b]
            if torch.unique(cluster_ids).shape[0] > 49:      # rare: grow the quantile for this shape only (ref :76-83)
                center, bandwidth, cluster_ids = self.guard_mean_shift(embedding[b], quantile * 1.2, iterations)
            weights = center @ embedding[b].t()
            loss, parameters, _, rows, cols, distance = self.residual_train_mode(
                points[b], normals[b], labels[b], cluster_ids, primitives[b], weights, bandwidth, lamb=lamb)
            with torch.no_grad():
                s_iou, p_iou, _, _ = SIOU_matched_segments(labels[b], cluster_ids.data.cpu().numpy(), prim_pred[b],
                                                           primitives[b], weights.t())
            out = out + loss + [s_iou, p_iou]
        return out, [parameters, cluster_ids.data.cpu().numpy(), weights]

    def residual_train_mode(self, points, normals, labels, cluster_ids, primitives, weights, bw, lamb=1.0):
        if not isinstance(cluster_ids, np.ndarray):
            cluster_ids = cluster_ids.data.cpu().numpy()
        rows, cols, unique_target, unique_pred = match(labels, cluster_ids)
        data = []
        for index, i in enumerate(unique_pred):
            gt_i = labels == cols[i]
            if gt_i.sum() == 0 or (cluster_ids == i).sum() == 0:
                continue
            l = stats.mode(primitives[gt_i])[0]
            gt_idx = torch.from_numpy(np.nonzero(gt_i)[0]).to(points.device)
            data.append([points, normals, l, points[gt_idx], None, (index, i)])
        w = weights_normalize(weights, float(bw)).t()
        gt_points, _ = fit_one_shape_torch(data, self.fitter, w, bw, eval=False)
        distance = self.res_loss.residual_loss(gt_points, self.fitter.fitting.parameters)
        return self.separate_losses(distance, gt_points, lamb=lamb), self.fitter.fitting.parameters, None, rows, \
            cols, distance

    def separate_losses(self, distance, gt_points, lamb=1.0):
        """mean residual over fitted segments (spline terms weighted by lamb); residuals > 1 are treated as degenerate
        and replaced by the constant 0.1 (reference :333-378)"""
        terms, geo, spl = [], [], []
        for v in sorted(gt_points.keys()):
            if gt_points[v] is None:
                continue
            if distance[v][1] > 1:
                distance[v][1] = torch.ones(1, device=distance[v][1].device)[0] * 0.1
            if distance[v][0] in ("closed-spline", "open-spline"):
                spl.append(distance[v][1].item())
                terms.append(distance[v][1] * lamb)
            else:
                geo.append(distance[v][1].item())
                terms.append(distance[v][1])
        dev = next(iter(distance.values()))[1].device if distance else "cuda"
        loss = torch.mean(torch.stack(terms)) if terms else torch.zeros(1, device=dev)
        return [loss, float(np.mean(geo)) if geo else None, float(np.mean(spl)) if spl else None]
