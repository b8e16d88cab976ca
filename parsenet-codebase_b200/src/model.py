"""Drop-in for reference src/model.py: SplineNet control-point regressor DGCNNControlPoints (:56-180), knn (:9),
get_graph_feature (:25).  Parameters / buffers keep the reference's names (conv1.0.weight, bn1.*, conv5.0.weight,
conv6..8, bn6, bn7, incl. the aliased convX.1 == bnX keys) so open_spline.pth / closed_spline.pth load unchanged.
The math runs in pnb200.splinenet.SplineNetFn on the sm_100a kernels."""
import torch
import torch.nn as nn

from pnb200 import ops
from pnb200.splinenet import SplineNetFn


def knn(x, k):
    """x (B,C,N) -> (B,N,k) int64 neighbour indices, nearest first (reference src/model.py:9-22)"""
    return ops.knn_graph(x.permute(0, 2, 1).contiguous(), k, 0, out_dtype=torch.int64)


def get_graph_feature(x, k=20, idx=None):
    """(B,C,N) -> (B,2C,N,k) edge features [x_j - x_i ; x_i] (reference src/model.py:25-53).  Kept for API parity and
    debugging only: the network itself never materialises this tensor."""
    B, C, N = x.shape
    if idx is None:
        idx = knn(x, k)
    xt = x.permute(0, 2, 1)
    nb = torch.gather(xt.unsqueeze(1).expand(B, N, N, C), 2, idx.unsqueeze(-1).expand(B, N, k, C))
    ctr = xt.unsqueeze(2).expand(B, N, k, C)
    return torch.cat([nb - ctr, ctr], 3).permute(0, 3, 1, 2)


def _block(cin, cout, bn, conv):
    return nn.Sequential(conv(cin, cout, kernel_size=1, bias=False), bn, nn.LeakyReLU(negative_slope=0.2))


class DGCNNControlPoints(nn.Module):
    WIDTHS = {0: (64, 64, 128, 256), 1: (128, 256, 256, 512)}

    def __init__(self, num_control_points, num_points=40, mode=0):
        super().__init__()
        if mode not in self.WIDTHS:
            raise NotImplementedError("DGCNNControlPoints: mode must be 0 or 1")
        self.k, self.mode, self.drop = num_points, mode, 0.0
        w = self.WIDTHS[mode]
        self.widths = w
        self.bn1, self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(c) for c in w)
        self.bn5 = nn.BatchNorm1d(1024)
        cins = (6, 2 * w[0], 2 * w[1], 2 * w[2])
        self.conv1 = _block(cins[0], w[0], self.bn1, nn.Conv2d)
        self.conv2 = _block(cins[1], w[1], self.bn2, nn.Conv2d)
        self.conv3 = _block(cins[2], w[2], self.bn3, nn.Conv2d)
        self.conv4 = _block(cins[3], w[3], self.bn4, nn.Conv2d)
        cat = sum(w) if mode == 0 else 1024 + 128
        self.conv5 = _block(cat, 1024, self.bn5, nn.Conv1d)
        self.controlpoints = num_control_points
        self.conv6 = nn.Conv1d(1024, 1024, 1)
        self.conv7 = nn.Conv1d(1024, 1024, 1)
        self.conv8 = nn.Conv1d(1024, 3 * (num_control_points ** 2), 1)
        self.bn6 = nn.BatchNorm1d(1024)
        self.bn7 = nn.BatchNorm1d(1024)
        self.tanh = nn.Tanh()

    def _params(self):
        return (self.conv1[0].weight, self.bn1.weight, self.bn1.bias, self.conv2[0].weight, self.bn2.weight,
                self.bn2.bias, self.conv3[0].weight, self.bn3.weight, self.bn3.bias, self.conv4[0].weight,
                self.bn4.weight, self.bn4.bias, self.conv5[0].weight, self.bn5.weight, self.bn5.bias,
                self.conv6.weight, self.conv6.bias, self.bn6.weight, self.bn6.bias,
                self.conv7.weight, self.conv7.bias, self.bn7.weight, self.bn7.bias, self.conv8.weight, self.conv8.bias)

    def _bns(self):
        return (self.bn1, self.bn2, self.bn3, self.bn4, self.bn5, self.bn6, self.bn7)

    def forward(self, x, weights=None):
        """x (B,3,N); weights (N,1)/(1,N)/(B,N) membership of every point or None -> (B, ncp^2, 3) in (-1,1)"""
        B = x.shape[0]
        x0 = x.permute(0, 2, 1)
        w = None
        if isinstance(weights, torch.Tensor):
            w = weights.reshape(1, -1).expand(B, -1) if weights.numel() == x.shape[2] else weights.reshape(B, -1)
        running = [(bn.running_mean, bn.running_var) for bn in self._bns()]
        stats = [] if self.training else None
        out = SplineNetFn.apply(x0, w, self.training, self.k, self.widths, running, stats, *self._params())
        if self.training:
            with torch.no_grad():                 # BatchNorm running statistics, momentum 0.1, unbiased variance
                for bn, (mean, sums, n) in zip(self._bns(), stats):
                    var_unb = (sums[:, 1] - sums[:, 0] ** 2 / n) / max(n - 1.0, 1.0)
                    bn.running_mean.mul_(0.9).add_(0.1 * mean)
                    bn.running_var.mul_(0.9).add_(0.1 * var_unb.float())
                    bn.num_batches_tracked += 1
        return out.view(B, self.controlpoints * self.controlpoints, 3)


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
