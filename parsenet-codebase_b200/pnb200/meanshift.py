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
ARGSEL_IMPL = os.environ.get("PN_MS_ARGSEL", "tc")      # modes 0 / 1 of the nms arg-selects (mode 2 is always simt)
# operand tiles of the forward / backward kernels fetched by TMA (csrc/meanshift_tma.cu): bit-identical to the
# loader-warp kernels and 1.12x (forward) / 1.17x (dense backward) faster at B = 16, N = 10^4 (first GPU call of round 2,
# profiles/r02_first_call.md); default since then.  PN_MS_TMA=0 selects the loader-warp kernels (A/B tests).
USE_TMA = os.environ.get("PN_MS_TMA", "1") == "1"
_KTH = {"tc": "pn_ms_kth_dist_tc", "simt": "pn_ms_kth_dist"}


WIDTH = 128


def _check_width(d):
    """all mean-shift kernels are built for the embedding width of every reference config (emb_size = 128,
    configs/config_parsenet*.yml); say so up front instead of failing inside a launch"""
    if d != WIDTH:
        raise ValueError(f"mean-shift kernels are built for embedding width {WIDTH}, got {d} "
                         "(construct PrimitivesEmbeddingDGCNGn with emb_size=128)")


def _argsel_entry(mode, d):
    return "pn_ms_argsel_tc" if (ARGSEL_IMPL == "tc" and mode in (0, 1) and d == 128) else "pn_ms_argsel"


def _launch_argsel(mode, A, a_stride, Ma, Bm, b_stride, Nb, B, d, cnt, thr, out):
    """nms arg-select (modes 0 / 1 / 2): the TMA-fed tcgen05 kernel when it applies (csrc/meanshift_tma.cu: 64-column tiles;
    needs the small split part of Bm, one extra launch), else the loader-warp tcgen05 kernel, else the FP32-pipe kernel.
    Same picks from all three up to tensor-core near-ties."""
    from .cabi import lib
    if (USE_TMA and ARGSEL_IMPL == "tc" and mode in (0, 1)
            and lib.pn_ms_argsel_tma_supported(Bm.data_ptr(), b_stride, Nb, d)):
        ws = torch.empty((B * Nb * d,), dtype=torch.float32, device=Bm.device)
        call("pn_ms_argsel_tma", mode, _ptr(A), a_stride, Ma, _ptr(Bm), b_stride, Nb, B, d, _ptr(cnt), _ptr(thr), _ptr(ws),
             _ptr(out), _stream())
        return
    call(_argsel_entry(mode, d), mode, _ptr(A), a_stride, Ma, _ptr(Bm), b_stride, Nb, B, d, _ptr(cnt), _ptr(thr), _ptr(out),
         _stream())
SQRT_FLOOR = 1e-6     # guard_sqrt(top_k, 1e-6), mean_shift.py:135


def _operand_forms(X):
    """TMA operand forms of the constant matrix X (B,N,128) of one mean_shift call: Xs = X - tf32_hi(X) and the
    transposes Xt, Xst (B,128,Np), Np = N rounded up to 32 (csrc/meanshift_tma.cu)"""
    B, N, d = X.shape
    Np = (N + 31) // 32 * 32
    Xs = torch.empty_like(X)
    Xt = torch.empty((B, d, Np), dtype=torch.float32, device=X.device)
    Xst = torch.empty((B, d, Np), dtype=torch.float32, device=X.device)
    call("pn_ms_prepare_operands", _ptr(X), B, N, d, Np, _ptr(Xs), _ptr(Xt), _ptr(Xst), _stream())
    return Xs, Xt, Xst, Np


def _forward_iterations(X, cinv, iterations):
    """the `iterations` fused mean-shift iterations starting from Y_0 = X (mean_shift.py:58-77); returns the iterates
    [Y_0 .. Y_it] and the per-iteration kernel row sums / pre-normalisation norms the backward kernels need"""
    B, N, d = X.shape
    Ys, dens, norms = [X], [], []
    forms = _operand_forms(X) if (USE_TMA and FWD_IMPL == "tc" and d == 128 and iterations > 0) else None
    for _ in range(iterations):
        Yn = torch.empty_like(X)
        den = torch.empty((B, N), dtype=torch.float32, device=X.device)
        un = torch.empty((B, N), dtype=torch.float32, device=X.device)
        if forms is not None:
            Xs, Xt, Xst, Np = forms
            call("pn_ms_iter_fwd_tma", _ptr(Ys[-1]), _ptr(X), _ptr(Xs), _ptr(Xt), _ptr(Xst), B, N, d, Np,
                 _ptr(cinv), _ptr(Yn), _ptr(den), _ptr(un), _stream())
        else:
            call("pn_ms_iter_fwd_tc" if (FWD_IMPL == "tc" and d == 128) else "pn_ms_iter_fwd", _ptr(Ys[-1]), _ptr(X),
                 B, N, d, _ptr(cinv), _ptr(Yn), _ptr(den), _ptr(un), _stream())
        Ys.append(Yn); dens.append(den); norms.append(un)
    return Ys, dens, norms


class MeanShiftItersFn(torch.autograd.Function):
    """Y_it = shift(Y_{it-1}; X, b) for it = 1..iterations with Y_0 = X; returns Y_iterations.
    Saves the Y iterates + two (B,N) vectors per iteration; the N x N kernel matrices are recomputed in backward."""

    @staticmethod
    def forward(ctx, X, cinv, iterations):
        X = X.detach().contiguous()
        _need_cuda(X, cinv)
        cinv = cinv.detach().to(torch.float32).contiguous()
        Ys, dens, norms = _forward_iterations(X, cinv, iterations)
        ctx.saved = (X, cinv, Ys, dens, norms)
        # (fresh view: an output object kept in ctx would form a reference cycle, see segnet.EncoderFn.forward)
        return Ys[-1].clone() if iterations == 0 else Ys[-1].view_as(Ys[-1])

    @staticmethod
    def backward(ctx, g):
        X, cinv, Ys, dens, norms = ctx.saved
        B, N, d = X.shape
        g = g.contiguous()
        gX = torch.zeros_like(X)
        Gn = torch.empty_like(X)
        gd = torch.empty((B, N), dtype=torch.float32, device=X.device)
        # (the operand forms are recomputed here rather than kept alive between forward and backward: one 15 MB / shape
        # pass against 10 iterations of N^2 work)
        forms = _operand_forms(X) if (USE_TMA and BWD_IMPL == "tc" and d == 128 and len(dens) > 0) else None
        ws_C = torch.empty((4, B, 2 * ((N + 15) // 16 * 16), d), dtype=torch.float32, device=X.device) if forms else None
        for it in range(len(dens) - 1, -1, -1):
            gprev = torch.empty_like(X)
            if forms is not None:
                Xs, Xt, Xst, Np = forms
                call("pn_ms_iter_bwd_tma", _ptr(g), _ptr(Ys[it + 1]), _ptr(Ys[it]), _ptr(X), _ptr(Xs), _ptr(Xt), _ptr(Xst),
                     _ptr(dens[it]), _ptr(norms[it]), B, N, d, Np, _ptr(cinv), _ptr(Gn), _ptr(gd), _ptr(ws_C), _ptr(gprev),
                     _ptr(gX), 1, _stream())
            else:
                call("pn_ms_iter_bwd_tc" if BWD_IMPL == "tc" else "pn_ms_iter_bwd", _ptr(g), _ptr(Ys[it + 1]), _ptr(Ys[it]), _ptr(X), _ptr(dens[it]),
                     _ptr(norms[it]), B, N, d, _ptr(cinv), _ptr(Gn), _ptr(gd), _ptr(gprev), _ptr(gX), 1, _stream())
            g = gprev
        gX += g            # Y_0 = X.clone()
        return gX, None, None


# ------------------------------------------------------------------------------------------------ sparse-row backward
# Default in Evaluation.fitting_loss since round 2 (csrc/meanshift.cu::ms_bwd_sparse_kernel; equals the dense backward to
# 1e-5 in tests/test_gpu_zz_fresh_inputs.py::test_sparse_row_backward_equals_dense_backward, step 377 -> 226 ms).  The
# dense kernels stay for callers that consume all of new_X (MeanShift.mean_shift with nms=False, segment_loss.py:50).
# PN_MS_SPARSE_BWD=0 forces the dense backward everywhere.
SPARSE_BWD = os.environ.get("PN_MS_SPARSE_BWD", "1") == "1"
SPARSE_ROWS = 64          # the compact row set is padded to exactly this many rows (>= the 49 clusters the guards allow)


def mean_shift_iters_keep(X_bnd, bw_b, iterations):
    """the forward iterations WITHOUT an autograd node: returns (Y_final, state); state feeds centers_sparse"""
    _check_width(X_bnd.shape[-1])
    X = X_bnd.detach().contiguous()
    _need_cuda(X)
    B, N, d = X.shape
    bw = bw_b.detach().to(torch.float32)
    cinv = (1.0 / (bw * bw)).contiguous()
    Ys, dens, norms = _forward_iterations(X, cinv, int(iterations))
    return Ys[-1], (cinv, Ys, dens, norms)


class MeanShiftCentersFn(torch.autograd.Function):
    """rows `ids` (B,R) of the last mean-shift iterate as a differentiable function of X: the forward is a gather from
    the iterates computed by mean_shift_iters_keep, the backward runs every iteration for those R rows only (the
    gradient of a loss that sees the shifted points through `center = new_X[indices]` never leaves these rows)"""

    @staticmethod
    def forward(ctx, X, ids, state):
        cinv, Ys, dens, norms = state
        B, N, d = Ys[0].shape
        idx3 = ids.unsqueeze(2).expand(B, ids.shape[1], d)
        ctx.saved = (X.detach().contiguous(), ids, idx3, cinv, Ys, dens, norms)
        return torch.gather(Ys[-1], 1, idx3)

    @staticmethod
    def backward(ctx, g):
        X, ids, idx3, cinv, Ys, dens, norms = ctx.saved
        B, N, d = X.shape
        R = ids.shape[1]
        dev = X.device
        g = g.contiguous()
        gX = torch.zeros_like(X)
        nblk = (N + 63) // 64
        Gn = torch.empty((B, R, d), dtype=torch.float32, device=dev)
        gd = torch.empty((B, R), dtype=torch.float32, device=dev)
        part = torch.empty((B, nblk, R, d), dtype=torch.float32, device=dev)
        for it in range(len(dens) - 1, -1, -1):
            y_new = torch.gather(Ys[it + 1], 1, idx3)
            y_prev = torch.gather(Ys[it], 1, idx3)
            den_r = torch.gather(dens[it], 1, ids)
            un_r = torch.gather(norms[it], 1, ids)
            gprev = torch.empty_like(g)
            call("pn_ms_rows_bwd", _ptr(g), _ptr(y_new), _ptr(y_prev), _ptr(den_r), _ptr(un_r), _ptr(X), B, R, N, d,
                 _ptr(cinv), _ptr(Gn), _ptr(gd), _ptr(part), _ptr(gprev), _ptr(gX), _stream())
            g = gprev
        gX.scatter_add_(1, idx3, g)            # Y_0 = X: the remaining gradient belongs to the same rows of X
        return gX, None, None


def _padded_ids(ids_list, R):
    rows = []
    for ids_b in ids_list:
        k = int(ids_b.shape[0])
        # (more than R centres only happens when the caller is about to re-cluster this shape with a larger quantile --
        # Evaluation.fitting_loss does so above 49 -- and discards these centres: truncate instead of failing)
        rows.append(torch.cat([ids_b, ids_b[:1].expand(R - k)]) if k < R else ids_b[:R])
    return torch.stack(rows, 0).contiguous()


def centers_padded(X_bnd, state, ids_list, shifted=None):
    """kept centres of every shape as ONE (B, SPARSE_ROWS, d) tensor (slots beyond K_b repeat the shape's first centre):
    through the sparse-row backward when `state` is given, else gathered from the dense `shifted` (B,N,d)"""
    ids = _padded_ids(ids_list, SPARSE_ROWS)
    if state is not None:
        return MeanShiftCentersFn.apply(X_bnd, ids, state)
    return torch.gather(shifted, 1, ids.unsqueeze(2).expand(ids.shape[0], ids.shape[1], shifted.shape[2]))


def centers_sparse(X_bnd, state, ids_list):
    """ids_list: per shape a (K_b,) int64 device tensor of kept rows (nms_batched) -> list of (K_b, d) centre tensors,
    differentiable w.r.t. X through the sparse-row backward.  Slots beyond K_b repeat the shape's first id and are
    sliced away, i.e. receive a zero gradient and contribute exactly zero."""
    B = X_bnd.shape[0]
    R = SPARSE_ROWS
    ids = _padded_ids(ids_list, R)
    centers = MeanShiftCentersFn.apply(X_bnd, ids, state)
    return [centers[b, :min(int(ids_list[b].shape[0]), R)] for b in range(B)]


def mean_shift_iters(X_bnd, bw_b, iterations):
    """X (B,N,d) unit rows, bw (B,) bandwidths -> shifted points (B,N,d)   [mean_shift.py:45-79, gaussian kernel]"""
    _check_width(X_bnd.shape[-1])
    bw = bw_b.detach().to(torch.float32)
    cinv = 1.0 / (bw * bw)
    return MeanShiftItersFn.apply(X_bnd, cinv, int(iterations))


KTH_CAPS = (1024, 2048)   # entries per row of the candidate lists of pn_ms_kth_dist_tma (the smallest that fits is used)
KTH_SAMPLE = 1024         # at most this many sample columns
# one-pass bracketed K-th distance (csrc/meanshift_tma.cu) when every row is used; PN_MS_KTH_BRACKET=0: four radix passes
KTH_BRACKET = os.environ.get("PN_MS_KTH_BRACKET", "1") == "1"


def kth_bracket_plan(N, K):
    """(stride, b, cap) of the bracketed selection, or None when lists of 2048 entries cannot be expected to hold the answer:
    the column sample {0, stride, ...} has m <= 1024 points; the number of them below the K-th smallest of the full row is
    hypergeometric with mean mu = K m / N, so the b = mu + 7 sigma + 2 -th smallest of the sample lies above it (else the row
    is flagged and redone exactly); the one full pass then keeps about b N / m entries per row."""
    if N < 2048 or N >= 65536 or K < 1:
        return None
    stride = -(-N // KTH_SAMPLE)
    m = -(-N // stride)
    if m % 4 != 0 and m + 3 >= KTH_SAMPLE:
        return None
    mu = K * m / N
    sigma = (mu * (1.0 - K / N) * (N - m) / max(N - 1, 1)) ** 0.5
    b = int(mu + 7.0 * sigma + 2.0) + 1
    if b > m:
        return None
    # expected list length + 6 sigma of the bracket's rank in the full row; the kernel keeps one list of cap / 2 entries per
    # half of the tile columns, so leave the binomial split of the entries over the halves its 6 sigma too
    length = (b + 6.0 * b ** 0.5) * N / m
    for cap in KTH_CAPS:
        if length / 2 + 3.0 * length ** 0.5 <= cap / 2:
            return stride, b, cap
    return None


_KTH_WS = {}


def _kth_workspace(dev, rows, cap):
    """candidate lists of the bracketed selection: 6 bytes x cap per row (1 - 2 GB at 16 x 10^4 rows).  Kept between calls
    (one per device and stream order: the kernels of a call are enqueued on the caller's stream, a later call on the same
    stream reuses the lists only after them) so that the caching allocator does not carve the block up in between."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _KTH_WS.get(key)
    if ws is None or ws[0].shape[0] < rows or ws[0].shape[1] != cap:
        ws = (torch.empty((rows, cap), dtype=torch.int32, device=dev), torch.empty((rows, cap), dtype=torch.int16, device=dev),
              torch.empty((rows, 2), dtype=torch.int32, device=dev), torch.empty((rows,), dtype=torch.float32, device=dev),
              torch.empty((rows,), dtype=torch.int32, device=dev))
        _KTH_WS[key] = ws
    return ws


def _kth_all_rows(X, K):
    """K-th smallest of 2 - 2 X X^T per row, X (B,N,128) contiguous, every row against all N points of its shape"""
    B, N, d = X.shape
    kth = torch.empty((B, N), dtype=torch.float32, device=X.device)
    plan = kth_bracket_plan(N, K) if (KTH_BRACKET and KTH_IMPL == "tc" and USE_TMA and d == 128) else None
    if plan is None:
        call(_KTH[KTH_IMPL], _ptr(X), None, B, N, N * d, d, K, _ptr(kth), _stream())
        return kth
    stride, b, cap = plan
    Xs = _operand_forms(X)[0]
    ws_key, ws_col, ws_cnt, ws_hi, flags = _kth_workspace(X.device, B * N, cap)
    call("pn_ms_kth_dist_tma", _ptr(X), _ptr(Xs), B, N, d, K, stride, b, _ptr(ws_key), _ptr(ws_col), _ptr(ws_cnt), _ptr(ws_hi),
         cap, _ptr(flags), _ptr(kth), _stream())
    call("pn_ms_kth_dist_tc_flagged", _ptr(X), None, B, N, N * d, d, K, _ptr(flags), _ptr(kth), _stream())
    return kth


def compute_bandwidth(X_nd, num_samples, quantile, rng=np.random):
    """mean over sampled rows of sqrt(K-th smallest of 2 - 2 X X^T), K = int(quantile * num_samples)
    [mean_shift.py:115-137].  Consumes np.random.shuffle exactly like the reference."""
    _check_width(X_nd.shape[-1])
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
    if rows is None:
        kth = _kth_all_rows(X.unsqueeze(0), K)[0]
    else:
        kth = torch.empty((S,), dtype=torch.float32, device=X.device)
        call(_KTH[KTH_IMPL], _ptr(X), _ptr(rows), 1, S, N * d, d, K, _ptr(kth), _stream())
    return torch.sqrt(torch.clamp(kth, min=SQRT_FLOOR)).mean()


def compute_bandwidth_batched(X_bnd, num_samples, quantile, rng=np.random):
    """compute_bandwidth for a batch of shapes in ONE launch (N <= num_samples, i.e. every row is used; the host RNG
    is still consumed once per shape like the reference would).  Returns (B,) bandwidths (before the 0.003 floor)."""
    _check_width(X_bnd.shape[-1])
    _need_cuda(X_bnd)
    B, N, d = X_bnd.shape
    if N > int(num_samples):
        return torch.stack([compute_bandwidth(X_bnd[b], num_samples, quantile, rng) for b in range(B)])
    for _ in range(B):
        rng.shuffle(np.arange(N))
    K = int(quantile * num_samples)
    X = X_bnd.detach().contiguous()
    kth = _kth_all_rows(X, K)
    return torch.sqrt(torch.clamp(kth, min=SQRT_FLOOR)).mean(1)


def nearest_center_batched(X_bnd, Y_bnd):
    """membership of every point to its nearest shifted point (nms step 1) for a batch of shapes in one launch"""
    B, N, d = X_bnd.shape
    X, Y = X_bnd.detach().contiguous(), Y_bnd.detach().contiguous()
    out = torch.empty((B, N), dtype=torch.int32, device=X.device)
    _launch_argsel(0, X, N * d, N, Y, N * d, N, B, d, None, None, out)
    return out


def _argsel(mode, A, Bm, cnt=None, thr=None):
    Ma, d = A.shape
    Nb = Bm.shape[0]
    out = torch.empty((Ma,), dtype=torch.int32, device=A.device)
    call(_argsel_entry(mode, d), mode, _ptr(A), 0, Ma, _ptr(Bm), 0, Nb, 1, d, _ptr(cnt), _ptr(thr), _ptr(out),
         _stream())
    return out


def nms(centers_nd, X_nd, b, member=None):
    """non-max suppression of the shifted points [mean_shift.py:139-179] -> (kept centres, their ids, labels int64).
    `member` (N,) int32: precomputed nearest-centre ids (from nearest_center_batched)."""
    _check_width(centers_nd.shape[-1])
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


NMS_WIDTH = 64        # kept centres per shape handled without a second read-back (the fit stage allows at most 49)


def nms_batched(Y_bnd, X_bnd, bw_b, member=None, also=None):
    """nms (mean_shift.py:139-179) for a batch of shapes with ONE blocking read-back instead of ~5 per shape.
    Returns (ids, labels, K): ids = list of (K_b,) int64 device tensors (kept centre rows of Y, ascending),
    labels (B,N) int64 device tensor (position of the nearest kept centre), K = list of K_b (host ints).
    Same arithmetic as nms(): the occupied-centre rows of the neighbour arg-select are a subset of the all-rows
    launch used here, and padding the kept list with a repeat of its first entry cannot win an arg-max tie
    (first occurrence wins).
    The kept rows are selected on the device into a fixed NMS_WIDTH-wide table (ascending ids, padded), so the label
    arg-select is enqueued without waiting for the host; the single read-back at the end brings K, the labels and the
    float tensors listed in `also` to the host together: with `also` given the return value is
    (ids, labels, K, labels_host (B,N) int32 numpy, [numpy copies of also])."""
    _check_width(Y_bnd.shape[-1])
    Y = Y_bnd.detach().contiguous()
    X = X_bnd.detach().contiguous()
    B, N, d = X.shape
    dev = X.device
    if member is None:
        member = nearest_center_batched(X, Y)
    counts = torch.zeros((B, N), dtype=torch.float32, device=dev)
    counts.scatter_add_(1, member.long(), torch.ones((B, N), dtype=torch.float32, device=dev))
    thr = bw_b.detach().to(torch.float32).reshape(B).contiguous()
    nbr = torch.empty((B, N), dtype=torch.int32, device=dev)
    _launch_argsel(1, Y, N * d, N, Y, N * d, N, B, d, counts, thr, nbr)
    # neighbours chosen by OCCUPIED centres are kept; unoccupied rows scatter into a dump column
    tgt = torch.where(counts > 0, nbr.long(), torch.full_like(nbr, N, dtype=torch.int64))
    mark = torch.zeros((B, N + 1), dtype=torch.bool, device=dev)
    mark.scatter_(1, tgt, torch.ones((B, N), dtype=torch.bool, device=dev))
    W = min(NMS_WIDTH, N)
    ar = torch.arange(N, device=dev, dtype=torch.int32)
    cand = torch.where(mark[:, :N], ar, torch.full((), N, dtype=torch.int32, device=dev))           # (B,N): id or N
    table = torch.topk(cand, W, dim=1, largest=False, sorted=True)[0]                                # ascending ids, N = empty
    Kd = mark[:, :N].sum(1, dtype=torch.int32)                                                       # (B,) kept rows per shape
    pad_d = torch.where(table < N, table, table[:, :1]).long()                                       # padded with the first id
    centres = torch.gather(Y, 1, pad_d.unsqueeze(2).expand(B, W, d)).contiguous()
    lab = torch.empty((B, N), dtype=torch.int32, device=dev)
    call("pn_ms_argsel", 2, _ptr(X), N * d, N, _ptr(centres), W * d, W, B, d, None, None, _ptr(lab), _stream())
    extra = [t.detach().to(torch.float32).reshape(-1) for t in (also or [])]
    pack = torch.cat([Kd, lab.reshape(-1)] + [t.view(torch.int32) for t in extra]).cpu().numpy()     # THE read-back
    K = pack[:B].astype(np.int64)
    if int(K.max()) > W:                                # more kept centres than the table holds: redo the tail on the host
        kept = mark[:, :N].nonzero().cpu().numpy()
        Kmax = int(K.max())
        starts = np.concatenate([[0], np.cumsum(K)])
        pad = np.empty((B, Kmax), dtype=np.int64)
        for b in range(B):
            ids_b = kept[starts[b]:starts[b + 1], 1]
            pad[b, :K[b]] = ids_b
            pad[b, K[b]:] = ids_b[0]
        from .staging import arena
        stage = arena("nms", dev)
        stage.reset()
        pad_d = stage.upload(pad, dev)
        centres = torch.gather(Y, 1, pad_d.unsqueeze(2).expand(B, Kmax, d)).contiguous()
        call("pn_ms_argsel", 2, _ptr(X), N * d, N, _ptr(centres), Kmax * d, Kmax, B, d, None, None, _ptr(lab), _stream())
        lab_host = lab.cpu().numpy()
    else:
        lab_host = pack[B:B + B * N].reshape(B, N)
    ids = [pad_d[b, :int(K[b])] for b in range(B)]
    if also is None:
        return ids, lab.long(), [int(k) for k in K]
    outs, o = [], B + B * N
    for t, src in zip(extra, also):
        outs.append(pack[o:o + t.numel()].view(np.float32).reshape(tuple(src.shape)))
        o += t.numel()
    return ids, lab.long(), [int(k) for k in K], lab_host, outs
