"""torch-facing ops: thin wrappers + autograd.Functions over the C-ABI kernels (no eager fallback)."""
import os

import torch

from . import cabi
from .cabi import lib, check, call


def _ptr(t):
    """raw device pointer of a tensor argument; remembers the tensor's device for the launch that follows"""
    if t is None:
        return None
    cabi.note_device(t.device)
    return t.data_ptr()


def _stream():
    """stream of the launch being assembled = torch's current stream ON THE DEVICE OF ITS POINTER ARGUMENTS (not of
    torch.cuda.current_device(): the reference's scripts keep tensors on cuda:alt_gpu while the current device is 0,
    train_parsenet_e2e.py:58).  Must be the last argument evaluated: it closes the argument list of one launch; cabi.call
    then runs the entry point under that device.  Raises when the pointers of one launch live on different devices."""
    dev = cabi.close_device_set()
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise cabi.PnError("parsenet_b200 ops need CUDA tensors (there is no CPU fallback)")


# --------------------------------------------------------------------------------------------- kNN
KNN_IMPL = os.environ.get("PN_KNN", "tc")           # "tc" | "tma" | "simt" (A/B tests; all three give the same graph)
_KNN_WS = {}


def knn_tc_plan(N, k, single_list=False):
    """(stride, b, cap) of the tensor-core-filtered kNN (csrc/knn_tc.cu), or None when its lists cannot be expected to hold the
    answer: the sample {0, stride, ...} has m <= 1024 columns, of which a hypergeometric number with mean mu = k m / N belongs
    to the k nearest; the b = mu + 7 sigma + 2 -th smallest upper bound of the sample is the admission bracket, which keeps
    about b N / m entries per row (one list of cap / 2 per half of the tile columns)."""
    if N < 2048 or N >= 65536:
        return None
    stride = -(-N // 1024)
    m = -(-N // stride)
    mu = k * m / N
    sigma = (mu * (1.0 - k / N) * (N - m) / max(N - 1, 1)) ** 0.5
    b = int(mu + 7.0 * sigma + 2.0) + 1
    if b > m:
        return None
    length = (b + 6.0 * b ** 0.5) * N / m
    for cap in (1024, 2048):
        # knn_tc.cu keeps one list of cap / 2 per half of the tile columns; knn_lowdim.cu one list of cap / 2 per row
        if (length if single_list else length / 2 + 3.0 * length ** 0.5) <= cap / 2:
            return stride, b, cap
    return None


def _knn_tc_workspace(dev, B, N, C, cap, stride):
    """workspaces of pn_knn_tc / pn_knn_lowdim, kept per (device, stream, list capacity) and grown on demand (6 bytes x cap
    per row): the layers of one step alternate between the two capacities, and a block of 1 - 2 GB that is freed and carved up
    between calls would make the caching allocator churn"""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream, cap)
    rows = B * N
    Np, mp = (N + 63) // 64 * 64, (-(-N // stride) + 63) // 64 * 64
    n_xs, n_colc = rows * C, B * (Np + mp)
    ws = _KNN_WS.get(key)
    if ws is None or ws["rows"] < rows:
        ws = {"rows": rows,
              "val": torch.empty((rows, cap), dtype=torch.int32, device=dev),
              "col": torch.empty((rows, cap), dtype=torch.int16, device=dev),
              "cnt": torch.empty((rows, 2), dtype=torch.int32, device=dev),
              "T": torch.empty((rows,), dtype=torch.float32, device=dev),
              "flags": torch.empty((rows,), dtype=torch.int32, device=dev),
              "xs": torch.empty((n_xs,), dtype=torch.float32, device=dev),
              "colc": torch.empty((n_colc,), dtype=torch.float32, device=dev)}
        _KNN_WS[key] = ws
    if ws["xs"].numel() < n_xs:
        ws["xs"] = torch.empty((n_xs,), dtype=torch.float32, device=dev)
    if ws["colc"].numel() < n_colc:
        ws["colc"] = torch.empty((n_colc,), dtype=torch.float32, device=dev)
    return ws


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
    # feature spaces with whole 32-channel chunks go through the TMA-staged kernel (csrc/knn_tma.cu); positions (+ normals),
    # very large k or N >= 65536 through the loader-thread kernel (csrc/knn.cu).  Same results (both bit-exact vs the oracle).
    entry = "pn_knn"
    if KNN_IMPL in ("tma", "tc") and lib.pn_knn_tma_supported(x_bnc.data_ptr(), N, C, ld, k, metric):
        entry = "pn_knn_tma"
    plan = knn_tc_plan(N, k) if (KNN_IMPL == "tc" and entry == "pn_knn_tma"
                                 and lib.pn_knn_tc_supported(x_bnc.data_ptr(), N, C, ld, k, metric)) else None
    if plan is not None:
        # tensor-core filter + exact refinement of the survivors (csrc/knn_tc.cu); rows it could not decide are flagged and
        # redone by the TMA kernel (tiles without a flagged row exit at once): same graph, no host round trip
        stride, b, cap = plan
        w = _knn_tc_workspace(x_bnc.device, B, N, C, cap, stride)
        i64 = 1 if out_dtype == torch.int64 else 0
        with torch.cuda.device(x_bnc.device):
            call("pn_knn_tc", _ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), i64, _ptr(dist), stride, b, _ptr(ws), _ptr(w["xs"]),
                 _ptr(w["colc"]), _ptr(w["T"]), _ptr(w["val"]), _ptr(w["col"]), _ptr(w["cnt"]), cap, _ptr(w["flags"]), _stream())
            call("pn_knn_tma_flagged", _ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), i64, _ptr(dist), _ptr(ws), _ptr(w["flags"]),
                 _stream())
        return (idx, dist) if return_dist else idx
    plan = knn_tc_plan(N, k, single_list=True) if (KNN_IMPL == "tc" and entry == "pn_knn"
                                                   and lib.pn_knn_lowdim_supported(N, C, k, metric)) else None
    if plan is not None:
        # positions (+ normals): exact costs for every pair, one-pass bracketed selection (csrc/knn_lowdim.cu), flagged rows
        # redone by the loader-thread kernel
        stride, b, cap = plan
        w = _knn_tc_workspace(x_bnc.device, B, N, 1, cap, stride)
        i64 = 1 if out_dtype == torch.int64 else 0
        with torch.cuda.device(x_bnc.device):
            call("pn_knn_lowdim", _ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), i64, _ptr(dist), stride, b, _ptr(ws), _ptr(w["T"]),
                 _ptr(w["val"]), _ptr(w["col"]), _ptr(w["cnt"]), cap, _ptr(w["flags"]), _stream())
            call("pn_knn_flagged", _ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), i64, _ptr(dist), _ptr(ws), _ptr(w["flags"]),
                 _stream())
        return (idx, dist) if return_dist else idx
    with torch.cuda.device(x_bnc.device):
        call(entry, _ptr(x_bnc), B, N, C, ld, k, metric, _ptr(idx), 1 if out_dtype == torch.int64 else 0,
                         _ptr(dist), _ptr(ws), _stream())
    return (idx, dist) if return_dist else idx


# --------------------------------------------------------------------------------------------- low-level wrappers
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
GN_EPS = 1e-5
LINEAR_IMPL = os.environ.get("PN_LINEAR", "tc")
LINEAR_BWD_IMPL = os.environ.get("PN_LINEAR_BWD", LINEAR_IMPL)       # "tc" | "simt" for the two backward GEMMs


def _pitch(t):
    """row pitch (floats) of a (B,Np,C) tensor whose rows are contiguous and shapes are back-to-back"""
    assert t.dim() == 3 and t.stride(2) == 1 and t.stride(0) == t.shape[1] * t.stride(1), (t.shape, t.stride())
    return t.stride(1)


class Norm:
    """finalised normalisation of one activation tensor: per-(shape,channel) scale/shift + per-group mean/rstd"""
    __slots__ = ("mean_rstd", "scale", "shift", "act", "G", "count", "per_shape", "gamma", "beta")

    def __init__(self, mean_rstd, scale, shift, act, G, count, per_shape, gamma, beta):
        self.mean_rstd, self.scale, self.shift, self.act = mean_rstd, scale, shift, act
        self.G, self.count, self.per_shape, self.gamma, self.beta = G, count, per_shape, gamma, beta


def linear_fwd(A, W, bias=None, sbias=None, in_norm=None, stats_groups=0, per_shape=True):
    """Y[b,n,:] = act(norm(A[b,n,:])) @ W.T + bias + sbias[b]; returns (Y, stats|None)"""
    _need_cuda(A, W)
    B, Np, K = A.shape
    Nout = W.shape[0]
    assert W.shape[1] == K and W.stride(1) == 1
    Y = torch.empty((B, Np, Nout), dtype=torch.float32, device=A.device)
    stats = None
    if stats_groups:
        stats = torch.zeros((B if per_shape else 1, stats_groups, 2), dtype=torch.float64, device=A.device)
    sc = sh = None
    act = ACT_NONE
    if in_norm is not None:
        sc, sh, act = in_norm.scale, in_norm.shift, in_norm.act
    # dense per-point MLP GEMMs run on the tcgen05 tensor cores (split-TF32) whenever the tile constraints hold;
    # tiny / oddly shaped problems (K = 6 first edge-conv, 10 logits, per-channel BatchNorm statistics, M < 128) stay
    # on the FP32-pipe kernel.  PN_LINEAR=simt forces the latter (used by the A/B parity test).
    entry = "pn_linear_fwd"
    if LINEAR_IMPL == "tc" and Np >= 128 and lib.pn_linear_fwd_tc_supported(
            _ptr(A), _pitch(A), _ptr(W), W.stride(0), _ptr(Y), Nout, Np, K, Nout, max(stats_groups, 1),
            1 if stats_groups else 0):
        entry = "pn_linear_fwd_tc"
    call(entry, _ptr(A), _pitch(A), _ptr(W), W.stride(0), _ptr(bias), _ptr(sbias), _ptr(sc), _ptr(sh),
                            act, _ptr(Y), Nout, _ptr(stats), B, Np, K, Nout, max(stats_groups, 1),
                            1 if per_shape else 0, _stream())
    return Y, stats


def norm_finalize(stats, gamma, beta, B, C, count, act, per_shape=True, eps=GN_EPS):
    S, G = stats.shape[0], stats.shape[1]
    mr = torch.empty((S, G, 2), dtype=torch.float32, device=stats.device)
    scale = torch.empty((S, C), dtype=torch.float32, device=stats.device)
    shift = torch.empty((S, C), dtype=torch.float32, device=stats.device)
    call("pn_norm_finalize", _ptr(stats), _ptr(gamma), _ptr(beta), S, G, C, float(count), float(eps), _ptr(mr),
                               _ptr(scale), _ptr(shift), _stream())
    if S != B:   # batch-wide statistics: kernels index scale/shift per shape
        scale = scale.expand(B, C).contiguous()
        shift = shift.expand(B, C).contiguous()
    return Norm(mr, scale, shift, act, G, float(count), per_shape, gamma, beta)


def linear_bwd_data(dY, W, dZ=None, accumulate=False, fin_A=None, fin_norm=None, fin_act=None, K=None):
    """dZ (+)= dY @ W ; with fin_A: multiplies by act'(pre) of the layer input and accumulates the GN-backward
    sums (returned).  dZ may be a strided (B,Np,K) view (e.g. a channel slice of the concat-gradient buffer)."""
    B, Np, Nout = dY.shape
    K = W.shape[1]
    if dZ is None:
        dZ = torch.empty((B, Np, K), dtype=torch.float32, device=dY.device)
    gsum = None
    finalize = fin_A is not None
    sc = sh = gamma = mr = None
    act = ACT_NONE
    G, per_shape = 1, 1
    if finalize:
        if fin_norm is not None:
            sc, sh, act, gamma, mr = fin_norm.scale, fin_norm.shift, fin_norm.act, fin_norm.gamma, fin_norm.mean_rstd
            G, per_shape = fin_norm.G, 1 if fin_norm.per_shape else 0
            gsum = torch.zeros((mr.shape[0], G, 2), dtype=torch.float64, device=dY.device)
        else:
            act = fin_act if fin_act is not None else ACT_NONE
    # tensor-core path (split-TF32, csrc/linear_tc.cu): the forward GEMM kernel on (dY, W^T) with the finalize epilogue;
    # shapes it does not take (narrow layers, per-channel BatchNorm groups, M < 128) stay on the FP32-pipe kernel
    if LINEAR_BWD_IMPL == "tc" and Np >= 128 and W.stride(1) == 1:
        Wt = W.t().contiguous()
        if lib.pn_linear_bwd_data_tc_supported(_ptr(dY), _pitch(dY), _ptr(Wt), Wt.stride(0), _ptr(dZ), _pitch(dZ),
                                               _ptr(fin_A), _pitch(fin_A) if finalize else 0, Np, K, Nout, G,
                                               1 if (finalize and gamma is not None) else 0):
            call("pn_linear_bwd_data_tc", _ptr(dY), _pitch(dY), _ptr(Wt), Wt.stride(0), _ptr(dZ), _pitch(dZ),
                 1 if accumulate else 0, 1 if finalize else 0, _ptr(fin_A), _pitch(fin_A) if finalize else 0, _ptr(sc),
                 _ptr(sh), act, _ptr(gamma), _ptr(mr), _ptr(gsum), B, Np, K, Nout, G, per_shape, _stream())
            return dZ, gsum
    call("pn_linear_bwd_data", _ptr(dY), _pitch(dY), _ptr(W), W.stride(0), _ptr(dZ), _pitch(dZ),
                                 1 if accumulate else 0, 1 if finalize else 0, _ptr(fin_A),
                                 _pitch(fin_A) if finalize else 0, _ptr(sc), _ptr(sh), act, _ptr(gamma), _ptr(mr),
                                 _ptr(gsum), B, Np, K, Nout, G, per_shape, _stream())
    return dZ, gsum


def norm_bwd_apply(dZ, A, norm, gsum, want_affine_grads=True):
    """in place: dZ (masked grad wrt norm output) -> grad wrt the pre-norm tensor A; returns (dgamma, dbeta)"""
    B, Np, C = dZ.shape
    dg = db = None
    if want_affine_grads:
        dg = torch.zeros((C,), dtype=torch.float32, device=dZ.device)
        db = torch.zeros((C,), dtype=torch.float32, device=dZ.device)
    call("pn_norm_bwd_apply", _ptr(dZ), _pitch(dZ), _ptr(A), _pitch(A), _ptr(norm.gamma), _ptr(norm.mean_rstd),
                                _ptr(gsum), B, Np, C, norm.G, 1 if norm.per_shape else 0, norm.count, _ptr(dg),
                                _ptr(db), _stream())
    return dg, db


def linear_bwd_weight(dY, A, in_norm=None, dW=None, want_bias=True, want_sbias=False):
    """dW += dY^T @ act(norm(A)); returns (dW, db, dsb).  dW may be a strided column view of a larger weight grad."""
    B, Np, Nout = dY.shape
    K = A.shape[2]
    if dW is None:
        dW = torch.zeros((Nout, K), dtype=torch.float32, device=dY.device)
    db = torch.zeros((Nout,), dtype=torch.float32, device=dY.device) if want_bias else None
    dsb = torch.zeros((B, Nout), dtype=torch.float32, device=dY.device) if want_sbias else None
    sc = sh = None
    act = ACT_NONE
    if in_norm is not None:
        sc, sh, act = in_norm.scale, in_norm.shift, in_norm.act
    assert dW.stride(1) == 1
    entry = "pn_linear_bwd_weight"
    if LINEAR_BWD_IMPL == "tc" and lib.pn_linear_bwd_weight_tc_supported(_ptr(dY), _pitch(dY), _ptr(A), _pitch(A), Np, K, Nout) \
            and (sc is None or (sc.data_ptr() % 16 == 0 and sh.data_ptr() % 16 == 0)):
        entry = "pn_linear_bwd_weight_tc"
    call(entry, _ptr(dY), _pitch(dY), _ptr(A), _pitch(A), _ptr(sc), _ptr(sh), act, _ptr(dW),
                                   dW.stride(0), _ptr(db), _ptr(dsb), B, Np, K, Nout, _stream())
    return dW, db, dsb


def edge_gather_fwd(PQ, idx, gamma, G, per_shape=True, want_stats=True):
    B, N, C2 = PQ.shape
    Cout = C2 // 2
    k = idx.shape[2]
    dev = PQ.device
    esel = torch.empty((B, N, Cout), dtype=torch.float32, device=dev)
    jsel = torch.empty((B, N, Cout), dtype=torch.int32, device=dev)
    esum = torch.empty((B, N, Cout), dtype=torch.float32, device=dev)
    stats = torch.zeros((B if per_shape else 1, G, 2), dtype=torch.float64, device=dev) if want_stats else None
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    call("pn_edge_gather_fwd", _ptr(PQ), _pitch(PQ), _ptr(idx), B, N, k, Cout, _ptr(gamma), _ptr(esel),
                                 _ptr(jsel), _ptr(esum), _ptr(stats), G, 1 if per_shape else 0, _stream())
    return esel, jsel, esum, stats


def edge_apply(esel, norm, out):
    B, N, Cout = esel.shape
    call("pn_edge_apply", _ptr(esel), _ptr(norm.scale), _ptr(norm.shift), _ptr(out), _pitch(out), B, N, Cout,
                            _stream())
    return out


def knn_csr_transpose(idx):
    B, N, k = idx.shape
    dev = idx.device
    cnt = torch.zeros((B, N), dtype=torch.int32, device=dev)
    off = torch.empty((B, N + 1), dtype=torch.int32, device=dev)
    cursor = torch.empty((B, N), dtype=torch.int32, device=dev)
    rev = torch.empty((B, N * k), dtype=torch.int32, device=dev)
    call("pn_knn_csr_transpose", _ptr(idx), B, N, k, _ptr(cnt), _ptr(off), _ptr(cursor), _ptr(rev), _stream())
    return off, rev


def edge_bwd(g, PQ, idx, esel, jsel, esum, norm, dense=True, want_affine_grads=True):
    """g: grad wrt the activated edge-conv output (B,N,Cout) (may be a strided slice). Returns (dPQ, dgamma, dbeta)."""
    B, N, C2 = PQ.shape
    Cout = C2 // 2
    k = idx.shape[2]
    dev = PQ.device
    dy = torch.empty((B, N, Cout), dtype=torch.float32, device=dev)
    S = norm.mean_rstd.shape[0]
    gsum = torch.zeros((S, norm.G, 2), dtype=torch.float64, device=dev) if dense else None
    dg = torch.zeros((Cout,), dtype=torch.float32, device=dev) if want_affine_grads else None
    db = torch.zeros((Cout,), dtype=torch.float32, device=dev) if want_affine_grads else None
    call("pn_edge_bwd_prep", _ptr(g), _pitch(g), _ptr(esel), _ptr(norm.scale), _ptr(norm.shift),
                               _ptr(norm.mean_rstd), _ptr(norm.gamma), B, N, Cout, norm.G,
                               1 if norm.per_shape else 0, _ptr(dy), _ptr(gsum), _ptr(dg), _ptr(db), _stream())
    off = rev = None
    if dense:
        off, rev = knn_csr_transpose(idx)
    dPQ = torch.empty((B, N, C2), dtype=torch.float32, device=dev)
    call("pn_edge_bwd", _ptr(PQ), _pitch(PQ), _ptr(dy), _ptr(esum), _ptr(jsel), _ptr(off), _ptr(rev),
                          _ptr(norm.mean_rstd), _ptr(gsum), _ptr(norm.scale), B, N, k, Cout, norm.G,
                          1 if norm.per_shape else 0, norm.count, 1 if dense else 0, _ptr(dPQ), C2, _stream())
    return dPQ, dg, db


def colmax_norm(Y, norm, wts=None):
    """out[b,c] = max_n act(norm(Y[b,n,c])) * (wts[b,n] if given); returns (out, arg)"""
    B, N, C = Y.shape
    out = torch.empty((B, C), dtype=torch.float32, device=Y.device)
    arg = torch.empty((B, C), dtype=torch.int32, device=Y.device)
    call("pn_colmax_norm", _ptr(Y), _pitch(Y), B, N, C, _ptr(norm.scale), _ptr(norm.shift), norm.act, _ptr(wts),
         _ptr(out), _ptr(arg), _stream())
    return out, arg


def colmax_bwd_fill(Y, gt, arg, norm, gsum, dense=True):
    B, N, C = Y.shape
    dY = torch.empty((B, N, C), dtype=torch.float32, device=Y.device)
    call("pn_colmax_bwd_fill", _ptr(Y), _pitch(Y), _ptr(gt), _ptr(arg), _ptr(norm.gamma), _ptr(norm.mean_rstd),
                                 _ptr(gsum), B, N, C, norm.G, 1 if norm.per_shape else 0, norm.count,
                                 1 if dense else 0, _ptr(dY), C, _stream())
    return dY


def logsoftmax_fwd(logits):
    B, N, P = logits.shape
    logp = torch.empty((B, P, N), dtype=torch.float32, device=logits.device)
    call("pn_logsoftmax_fwd", _ptr(logits), _pitch(logits), B, N, P, _ptr(logp), _stream())
    return logp


def logsoftmax_bwd(logp, dlp):
    B, P, N = logp.shape
    dl = torch.empty((B, N, P), dtype=torch.float32, device=logp.device)
    call("pn_logsoftmax_bwd", _ptr(logp), _ptr(dlp), B, N, P, _ptr(dl), P, _stream())
    return dl


def l2norm_fwd(x2d, eps=1e-12):
    rows, D = x2d.shape
    assert x2d.stride(1) == 1
    y = torch.empty((rows, D), dtype=torch.float32, device=x2d.device)
    norms = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    call("pn_l2norm_fwd", _ptr(x2d), x2d.stride(0), rows, D, float(eps), _ptr(y), D, _ptr(norms), _stream())
    return y, norms


def l2norm_bwd(y, dy, norms):
    rows, D = y.shape
    dx = torch.empty((rows, D), dtype=torch.float32, device=y.device)
    call("pn_l2norm_bwd", _ptr(y), y.stride(0), _ptr(dy), dy.stride(0), _ptr(norms), rows, D, _ptr(dx), D, 0,
                            _stream())
    return dx
