"""Drop-in for reference src/segment_loss.py: EmbeddingLoss.triplet_loss (:31-124), evaluate_miou (:127),
primitive_loss (:151).  Sampling stays on the host (numpy, same draw order); distances / hinge / gradients run
in the pn_triplet_* kernels."""
import numpy as np
import torch

from pnb200.losses import NllFn, TripletFn, triplet_sample


class EmbeddingLoss:
    def __init__(self, margin=1.0, if_mean_shift=False):
        self.margin, self.if_mean_shift = margin, if_mean_shift

    def triplet_loss(self, output, labels, iterations=5):
        """output (B,D,N) cuda embedding, labels (B,N) numpy -> (1,) loss."""
        emb = output.permute(0, 2, 1)                                 # (B,N,D)
        if self.if_mean_shift:
            from src.mean_shift import MeanShift
            ms = MeanShift()
            emb = torch.nn.functional.normalize(emb, p=2, dim=2)
            emb = torch.stack([ms.mean_shift(emb[b], 4000, 0.015, iterations=iterations, nms=False)[0]
                               for b in range(emb.shape[0])], 0)
        groups = triplet_sample(np.asarray(labels), emb.shape[1])
        if not groups:
            return torch.zeros(1, device=output.device, requires_grad=True) * emb.sum() * 0
        return TripletFn.apply(emb, groups, float(self.margin))


def evaluate_miou(gt_labels, pred_labels):
    """host metric: gt (B,N) ints, pred (B,N,C) scores -> mean IoU over C classes (reference :127-148)."""
    pred = np.argmax(pred_labels, 2)
    C = pred_labels.shape[2]
    eps = np.finfo(np.float32).eps
    per_shape = []
    for g, p in zip(gt_labels, pred):
        cls = np.arange(C)[:, None]
        inter = np.logical_and(g[None] == cls, p[None] == cls).sum(1) + eps
        union = np.logical_or(g[None] == cls, p[None] == cls).sum(1) + eps
        per_shape.append(float(np.mean(inter / union)))
    return float(np.mean(per_shape))


def primitive_loss(pred, gt):
    """NLL of (B,P,N) log-probabilities against (B,N) integer types (reference :151)."""
    return NllFn.apply(pred, gt)


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
