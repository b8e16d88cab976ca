"""Mean-shift clustering on the unit hypersphere over the fused kernels in csrc/meanshift.cu.

Schedules for MeanShift.mean_shift_ / compute_bandwidth / nms (reference src/mean_shift.py:45-179).  All functions
take a batch of shapes (B,N,d); the drop-in class in src/mean_shift.py calls them with B = 1.
"""
import numpy as np
import torch

from .cabi import call
from .ops import _need_cuda, _ptr, _stream

import os

BW_FLOOR = 0.003      # mean_shift.py:34
# forward iteration kernel: "tc" = tcgen05 split-TF32 tensor-core kernel (csrc/meanshift_tc.cu), "simt" = fp32 FMA kernel
FWD_IMPL = os.environ.get("PN_MS_FWD", "tc")
BWD_IMPL = os.environ.get("PN_MS_BWD", "tc")
KTH_IMPL = os.environ.get("PN_MS_KTH", "tc")
_KTH = {"tc": "pn_ms_kth_dist_tc", "simt": "pn_ms_kth_dist"}
SQRT_FLOOR = 1e-6     # guard_sqrt(top_k, 1e-6), mean_shift.py:135


class MeanShiftItersFn(torch.autograd.Function):
    """Y_it = shift(Y_{it-1}; X, b) for it = 1..iterations with Y_0 = X; returns Y_iterations.
    Saves the Y iterates + two (B,N) vectors per iteration; the N x N kernel matrices are recomputed in backward."""

    @staticmethod
    def forward(ctx, X, cinv, iterations):
        X = X.detach().contiguous()
        _need_cuda(X, cinv)
        B, N, d = X.shape
        cinv = cinv.detach().to(torch.float32).contiguous()
        Ys, dens, norms = [X], [], []
        for _ in range(iterations):
            Yn = torch.empty_like(X)
            den = torch.empty((B, N), dtype=torch.float32, device=X.device)
            un = torch.empty((B, N), dtype=torch.float32, device=X.device)
            call("pn_ms_iter_fwd_tc" if FWD_IMPL == "tc" else "pn_ms_iter_fwd", _ptr(Ys[-1]), _ptr(X), B, N, d,
                 _ptr(cinv), _ptr(Yn), _ptr(den), _ptr(un), _stream())
            Ys.append(Yn); dens.append(den); norms.append(un)
        ctx.saved = (X, cinv, Ys, dens, norms)
        return Ys[-1].clone() if iterations == 0 else Ys[-1]

    @staticmethod
    def backward(ctx, g):
        X, cinv, Ys, dens, norms = ctx.saved
        B, N, d = X.shape
        g = g.contiguous()
        gX = torch.zeros_like(X)
        Gn = torch.empty_like(X)
        gd = torch.empty((B, N), dtype=torch.float32, device=X.device)
        for it in range(len(dens) - 1, -1, -1):
            gprev = torch.empty_like(X)
            call("pn_ms_iter_bwd_tc" if BWD_IMPL == "tc" else "pn_ms_iter_bwd", _ptr(g), _ptr(Ys[it + 1]), _ptr(Ys[it]), _ptr(X), _ptr(dens[it]),
                 _ptr(norms[it]), B, N, d, _ptr(cinv), _ptr(Gn), _ptr(gd), _ptr(gprev), _ptr(gX), 1, _stream())
            g = gprev
        gX += g            # Y_0 = X.clone()
        return gX, None, None


def mean_shift_iters(X_bnd, bw_b, iterations):
    """X (B,N,d) unit rows, bw (B,) bandwidths -> shifted points (B,N,d)   [mean_shift.py:45-79, gaussian kernel]"""
    bw = bw_b.detach().to(torch.float32)
    cinv = 1.0 / (bw * bw)
    return MeanShiftItersFn.apply(X_bnd, cinv, int(iterations))


def compute_bandwidth(X_nd, num_samples, quantile, rng=np.random):
    """mean over sampled rows of sqrt(K-th smallest of 2 - 2 X X^T), K = int(quantile * num_samples)
    [mean_shift.py:115-137].  Consumes np.random.shuffle exactly like the reference."""
    _need_cuda(X_nd)
    N, d = X_nd.shape
    L = np.arange(N)
    rng.shuffle(L)
    S = min(N, int(num_samples))
    K = int(quantile * num_samples)
    X = X_nd.detach().contiguous()
    rows = None
    if S < N:                                   # a strict subset: the choice of rows matters
        rows = torch.from_numpy(L[:S].astype(np.int32)).to(X.device)
    kth = torch.empty((S,), dtype=torch.float32, device=X.device)
    call(_KTH[KTH_IMPL], _ptr(X), _ptr(rows), 1, S, N * d, d, K, _ptr(kth), _stream())
    return torch.sqrt(torch.clamp(kth, min=SQRT_FLOOR)).mean()


def compute_bandwidth_batched(X_bnd, num_samples, quantile, rng=np.random):
    """compute_bandwidth for a batch of shapes in ONE launch (N <= num_samples, i.e. every row is used; the host RNG
    is still consumed once per shape like the reference would).  Returns (B,) bandwidths (before the 0.003 floor)."""
    _need_cuda(X_bnd)
    B, N, d = X_bnd.shape
    if N > int(num_samples):
        return torch.stack([compute_bandwidth(X_bnd[b], num_samples, quantile, rng) for b in range(B)])
    for _ in range(B):
        rng.shuffle(np.arange(N))
    K = int(quantile * num_samples)
    X = X_bnd.detach().contiguous()
    kth = torch.empty((B, N), dtype=torch.float32, device=X.device)
    call(_KTH[KTH_IMPL], _ptr(X), None, B, N, N * d, d, K, _ptr(kth), _stream())
    return torch.sqrt(torch.clamp(kth, min=SQRT_FLOOR)).mean(1)


def nearest_center_batched(X_bnd, Y_bnd):
    """membership of every point to its nearest shifted point (nms step 1) for a batch of shapes in one launch"""
    B, N, d = X_bnd.shape
    X, Y = X_bnd.detach().contiguous(), Y_bnd.detach().contiguous()
    out = torch.empty((B, N), dtype=torch.int32, device=X.device)
    call("pn_ms_argsel", 0, _ptr(X), N * d, N, _ptr(Y), N * d, N, B, d, None, None, _ptr(out), _stream())
    return out


def _argsel(mode, A, Bm, cnt=None, thr=None):
    Ma, d = A.shape
    Nb = Bm.shape[0]
    out = torch.empty((Ma,), dtype=torch.int32, device=A.device)
    call("pn_ms_argsel", mode, _ptr(A), 0, Ma, _ptr(Bm), 0, Nb, 1, d, _ptr(cnt), _ptr(thr), _ptr(out), _stream())
    return out


def nms(centers_nd, X_nd, b, member=None):
    """non-max suppression of the shifted points [mean_shift.py:139-179] -> (kept centres, their ids, labels int64).
    `member` (N,) int32: precomputed nearest-centre ids (from nearest_center_batched)."""
    centers = centers_nd.detach().contiguous()
    X = X_nd.detach().contiguous()
    N = X.shape[0]
    if member is None:
        member = _argsel(0, X, centers)                              # nearest shifted centre per point
    counts = torch.bincount(member.long(), minlength=centers.shape[0]).to(torch.float32)
    uniq = torch.nonzero(counts > 0).flatten()                        # sorted, like np.unique
    thr = torch.as_tensor(b, dtype=torch.float32, device=X.device).reshape(1)
    nbr = _argsel(1, centers[uniq].contiguous(), centers, counts, thr)
    ids = torch.unique(nbr.long())
    kept = centers[ids].contiguous()
    labels = _argsel(2, X, kept).long()
    return kept, ids, labels
