"""Drop-in for reference src/PointNet.py: DGCNN edge-conv backbone + per-point segmentation head.

Same class names, constructor arguments, forward signatures, return shapes and state-dict keys as the reference
(DGCNNEncoderGn src/PointNet.py:143, PrimitivesEmbeddingDGCNGn :223; knn :9, knn_points_normals :29), so
`parsenet_with_normals.pth`-style checkpoints (incl. the aliased `encoder.bnX` == `encoder.convX.1` keys and the
unused bn4/bn5) load unchanged.  The torch modules below only own parameters; the math runs in
pnb200.segnet.{EncoderFn,HeadFn} on hand-written sm_100a kernels (no eager fallback).
"""
import torch
import torch.nn as nn

from pnb200 import ops
from pnb200.segnet import EncoderFn, HeadFn


def knn(x, k1, k2):
    """x (B,C,N) -> idx (B,N,k1) int64, nearest first (reference src/PointNet.py:9-26; k1 == k2 at every call)."""
    if k2 % k1 != 0:
        raise ValueError("k2 must be a multiple of k1")
    idx = ops.knn_graph(x.permute(0, 2, 1).contiguous(), k2, 0, out_dtype=torch.int64)
    return idx[:, :, ::k2 // k1]


def knn_points_normals(x, k1, k2):
    """x (B,6,N) positions+normals metric (reference src/PointNet.py:29-69)."""
    if k2 % k1 != 0:
        raise ValueError("k2 must be a multiple of k1")
    idx = ops.knn_graph(x.permute(0, 2, 1).contiguous(), k2, 1, out_dtype=torch.int64)
    return idx[:, :, ::k2 // k1]


def _edge_block(cin, cout, norm):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), norm, nn.LeakyReLU(negative_slope=0.2))


class DGCNNEncoderGn(nn.Module):
    def __init__(self, mode=0, input_channels=3, nn_nb=80):
        super().__init__()
        if mode not in (0, 5):
            raise NotImplementedError("DGCNNEncoderGn: the reference defines layers only for mode 0 and 5")
        self.k, self.mode, self.dilation_factor, self.drop = nn_nb, mode, 1, 0.0
        widths = [(2, 64), (2, 64), (2, 128), (4, 256), (8, 1024)]
        for i, (g, c) in enumerate(widths, 1):
            setattr(self, "bn%d" % i, nn.GroupNorm(g, c))           # bn4 / bn5 exist but are unused (ref :154-155)
        self.conv1 = _edge_block(input_channels * 2, 64, self.bn1)
        self.conv2 = _edge_block(128, 64, self.bn2)
        self.conv3 = _edge_block(128, 128, self.bn3)
        self.mlp1 = nn.Conv1d(256, 1024, 1)
        self.bnmlp1 = nn.GroupNorm(8, 1024)
        self.last_idx = None

    def _params(self):
        return (self.conv1[0].weight, self.bn1.weight, self.bn1.bias,
                self.conv2[0].weight, self.bn2.weight, self.bn2.bias,
                self.conv3[0].weight, self.bn3.weight, self.bn3.bias,
                self.mlp1.weight, self.mlp1.bias, self.bnmlp1.weight, self.bnmlp1.bias)

    def forward(self, x, idx_override=None):
        """x (B,C,N) fp32 cuda -> (x4 (B,1024), x_features (B,256,N)).  `idx_override` (list of three (B,N,k)
        index tensors) is a test hook to feed a fixed graph."""
        x0 = x.permute(0, 2, 1)
        x4, xf = EncoderFn.apply(x0, self.k, self.mode, idx_override, *self._params())
        return x4, xf.permute(0, 2, 1)


DGCNNEncoder = DGCNNEncoderGn   # spelling used by BASELINE.json's north_star


class PrimitivesEmbeddingDGCNGn(nn.Module):
    """Per-point embedding + primitive-type log-probabilities; the embedding loss is evaluated inside forward
    exactly like the reference (src/PointNet.py:223-289)."""

    def __init__(self, emb_size=50, num_primitives=8, primitives=False, embedding=False, mode=0, num_channels=3,
                 loss_function=None, nn_nb=80):
        super().__init__()
        if not (primitives and embedding):
            raise NotImplementedError("the reference forward needs both heads (embedding=True, primitives=True)")
        self.mode, self.drop, self.loss_function = mode, 0.0, loss_function
        self.emb_size, self.primitives, self.embedding = emb_size, primitives, embedding
        self.encoder = DGCNNEncoderGn(mode=mode, input_channels=num_channels, nn_nb=nn_nb)
        self.conv1 = nn.Conv1d(1024 + 256, 512, 1)
        self.bn1 = nn.GroupNorm(8, 512)
        self.conv2 = nn.Conv1d(512, 256, 1)
        self.bn2 = nn.GroupNorm(4, 256)
        self.softmax, self.logsoftmax, self.tanh = nn.Softmax(dim=1), nn.LogSoftmax(dim=1), nn.Tanh()
        self.mlp_seg_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_seg_prob2 = nn.Conv1d(256, emb_size, 1)
        self.bn_seg_prob1 = nn.GroupNorm(4, 256)
        self.mlp_prim_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_prim_prob2 = nn.Conv1d(256, num_primitives, 1)
        self.bn_prim_prob1 = nn.GroupNorm(4, 256)

    def _head_params(self):
        return (self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias,
                self.conv2.weight, self.conv2.bias, self.bn2.weight, self.bn2.bias,
                self.mlp_seg_prob1.weight, self.mlp_seg_prob1.bias, self.bn_seg_prob1.weight, self.bn_seg_prob1.bias,
                self.mlp_seg_prob2.weight, self.mlp_seg_prob2.bias,
                self.mlp_prim_prob1.weight, self.mlp_prim_prob1.bias, self.bn_prim_prob1.weight,
                self.bn_prim_prob1.bias, self.mlp_prim_prob2.weight, self.mlp_prim_prob2.bias)

    def forward(self, points, labels, compute_loss=True, idx_override=None):
        # the label copy blocks the host; do it BEFORE the network is enqueued so the host-side triplet sampling of
        # loss_function overlaps the kernels (the reference copies after the forward, src/PointNet.py:286)
        labels_np = labels.data.cpu().numpy() if compute_loss else None
        x4, first_layer_features = self.encoder(points, idx_override)
        emb_pm, primitives_log_prob = HeadFn.apply(x4, first_layer_features.permute(0, 2, 1), *self._head_params())
        embedding = emb_pm.permute(0, 2, 1)                          # (B, emb, N) view, as the reference returns
        if compute_loss:
            embed_loss = self.loss_function(embedding, labels_np)
        else:
            embed_loss = torch.zeros(1, device=points.device)
        return embedding, primitives_log_prob, embed_loss


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
