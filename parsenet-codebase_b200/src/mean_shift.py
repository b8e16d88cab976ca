"""Drop-in for reference src/mean_shift.py (class MeanShift, :9-185) on the fused sm_100a mean-shift kernels.

Same methods, arguments and return tuples.  Only the gaussian kernel is implemented on the device (the reference's
callers never pass another kernel_type); `pdist` / `kernel` are small utilities kept for API completeness."""
import numpy as np
import torch

from pnb200 import meanshift as _ms


class MeanShift:
    def __init__(self):
        pass

    def mean_shift(self, X, num_samples, quantile, iterations, kernel_type="gaussian", bw=None, nms=True):
        """X (N,d) unit rows -> (new_X, center, bw, labels) or (new_X, bw) when nms=False   [ref :19-43]"""
        if bw is None:
            with torch.no_grad():
                bw = torch.clamp(self.compute_bandwidth(X, num_samples, quantile), min=_ms.BW_FLOOR)
        new_X, _ = self.mean_shift_(X, b=bw, iterations=iterations, kernel_type=kernel_type)
        if not nms:
            return new_X, bw
        with torch.no_grad():
            _, indices, new_labels = self.nms(new_X, X, b=bw)
        center = new_X[indices]
        return new_X, center, bw, new_labels

    def mean_shift_(self, X, b, iterations=10, kernel_type="gaussian"):
        if kernel_type != "gaussian":
            raise NotImplementedError("only the gaussian kernel has a device implementation")
        bw = torch.as_tensor(b, dtype=torch.float32, device=X.device).reshape(1)
        new_X = _ms.mean_shift_iters(X.unsqueeze(0), bw, iterations)[0]
        return new_X, X

    def guard_mean_shift(self, embedding, quantile, iterations, kernel_type="gaussian"):
        """retry with a doubled quantile while more than 49 clusters come out   [ref :81-96]"""
        while True:
            _, center, bandwidth, cluster_ids = self.mean_shift(embedding, 5000, quantile, iterations,
                                                                kernel_type=kernel_type)
            if torch.unique(cluster_ids).shape[0] > 49:
                quantile *= 2
            else:
                break
        return center, bandwidth, cluster_ids

    def kernel(self, X, kernel_type, bw):
        dist = 2.0 - 2.0 * X @ X.t()
        if kernel_type == "gaussian":
            return torch.exp(torch.clamp(-dist / (bw ** 2) / 2, min=-75.0, max=75.0))
        return torch.nn.functional.relu(3 / 4 * (1 - dist / (bw ** 2)))

    def compute_bandwidth(self, X, num_samples, quantile):
        return _ms.compute_bandwidth(X, num_samples, quantile)

    def nms(self, centers, X, b):
        return _ms.nms(centers, X, b)

    def pdist(self, x, y):
        return ((x.unsqueeze(1) - y.unsqueeze(0)) ** 2).sum(2)


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
