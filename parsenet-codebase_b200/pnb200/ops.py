"""torch-facing ops: thin wrappers + autograd.Functions over the C-ABI kernels (no eager fallback)."""
import torch

from . import cabi
from .cabi import lib, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise cabi.PnError("parsenet_b200 ops need CUDA tensors (there is no CPU fallback)")


# --------------------------------------------------------------------------------------------- kNN
def knn_graph(x_bnc, k, metric=0, out_dtype=torch.int32, return_dist=False):
    """x_bnc: (B,N,C) fp32 point-major, possibly a channel slice of a wider buffer (stride(1) = row pitch).
    Returns idx (B,N,k) sorted best-first.  metric 0: feature space (src/PointNet.py:9), 1: positions+normals,
    C == 6 (src/PointNet.py:29)."""
    _need_cuda(x_bnc)
    assert x_bnc.dtype == torch.float32 and x_bnc.dim() == 3
    B, N, C = x_bnc.shape
    assert x_bnc.stride(2) == 1 and x_bnc.stride(0) == N * x_bnc.stride(1), "rows must be contiguous per shape"
    ld = x_bnc.stride(1)
    idx = torch.empty((B, N, k), dtype=out_dtype, device=x_bnc.device)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=x_bnc.device) if return_dist else None
    ws = torch.empty((B * N,), dtype=torch.float32, device=x_bnc.device)
    with torch.cuda.device(x_bnc.device):
        check(lib.pn_knn(_ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), 1 if out_dtype == torch.int64 else 0,
                         _ptr(dist), _ptr(ws), _stream()), "pn_knn")
    return (idx, dist) if return_dist else idx
