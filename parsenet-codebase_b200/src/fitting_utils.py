"""Drop-in for the hot-path part of reference src/fitting_utils.py: LeastSquares.lstsq (:36), best_lambda (:68),
weights_normalize (:306), match (:362), customsvd (:420), standardize_point[s]_torch (:493-553), pca_torch (:585),
rotation_matrix_a_to_b (:556), sample_points_from_control_points_ (:609), up_sample_points_torch[_in_range] (:150-237).  Mesh / open3d / visualisation helpers of
that file are out of scope."""
import numpy as np
import torch

from pnb200 import fitting as _f
from pnb200.fitting import spline_eval
from src.guard import guard_exp
from src.segment_utils import iou_cost_host, relaxed_iou_fast, solve_dense, to_one_hot

EPS = float(np.finfo(np.float32).eps)


class LeastSquares:
    def lstsq(self, A, Y, lamb=0.0):
        """x = argmin |A x - Y| for a tall A (m,3); rank-deficient systems use the Tikhonov rule of the reference."""
        if A.dim() != 2 or Y.dim() != 2 or Y.shape[0] != A.shape[0]:
            raise ValueError(f"LeastSquares.lstsq: expected A (m,n) and Y (m,c), got {tuple(A.shape)} and {tuple(Y.shape)}")
        if not torch.isfinite(A).all():
            raise FloatingPointError("LeastSquares.lstsq: non-finite entries in A")
        if A.shape[1] != 3 or Y.shape[1] != 1:
            # every fit on the hot path solves for 3 unknowns (sphere / cylinder centre, cone apex); other widths take
            # the reference's algorithm literally (fitting_utils.py:36-65: QR when A has full column rank, otherwise the
            # normal equations regularised with best_lambda) as plain torch.linalg calls on the caller's device.
            # `lamb` is ignored there as well (the reference overwrites it before use).
            n = A.shape[1]
            if n == int(torch.linalg.matrix_rank(A)):
                q, r = torch.linalg.qr(A)
                return torch.inverse(r) @ q.t() @ Y
            AtA = A.t() @ A
            with torch.no_grad():
                lam = best_lambda(AtA)
            return self.lstsq(AtA + lam * torch.eye(n, device=A.device, dtype=A.dtype), A.t() @ Y, 1)
        Ad, Yd = A.double(), Y.double()
        x = _f.solve_normal((Ad.t() @ Ad).unsqueeze(0), (Ad.t() @ Yd).reshape(1, -1), A.shape[0])
        return x.reshape(-1, 1).to(A.dtype)


def best_lambda(A):
    lamb = 1e-6
    n = A.shape[0]
    for _ in range(7):
        if n == int(torch.linalg.matrix_rank(A + lamb * torch.eye(n, device=A.device, dtype=A.dtype))):
            break
        lamb *= 10
    return lamb


class CustomSVD(torch.autograd.Function):
    """SVD of a tall (m,3) matrix with the reference's guarded gradient (only grad_V flows, K floored at 1e-6)."""

    @staticmethod
    def forward(ctx, inp):
        U, S, Vh = torch.linalg.svd(inp, full_matrices=False)
        V = Vh.transpose(-2, -1)
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, gU, gS, gV):
        U, S, V = ctx.saved_tensors
        n = S.shape[0]
        diff = S.view(n, 1) - S.view(1, n)
        eye = torch.eye(n, device=S.device, dtype=S.dtype)
        kneg = torch.sign(diff) * torch.clamp(diff.abs(), min=1e-6)
        kneg = kneg * (1 - eye) + 1e-6 * eye
        K = (1 / kneg) * (1 / (S.view(n, 1) + S.view(1, n))) * (1 - eye)
        inner = K.t() * (V.t() @ gV)
        inner = (inner + inner.t()) / 2.0
        return 2 * U @ torch.diag(S) @ inner @ V.t()


customsvd = CustomSVD.apply


def weights_normalize(weights, bw):
    """(K,N) centre-point similarities -> per-point cluster probabilities, then per-cluster min-max to [0,1]
    (reference :306-325).  Up to 64 clusters (the guards of the path allow 49) this is csrc/weights.cu through the batched
    entry point with B = 1; wider inputs, which no caller on the path produces, take the same formula as torch expressions."""
    K, N = weights.shape
    if weights.is_cuda and K <= 64:
        raw = torch.zeros((1, N, 64), dtype=torch.float32, device=weights.device)
        raw[0, :, :K] = weights.t()
        bw2 = torch.as_tensor(float(bw) ** 2, dtype=torch.float32, device=weights.device).reshape(1)
        Kd = torch.full((1,), K, dtype=torch.int32, device=weights.device)
        return _f.WeightsNormalizeFn.apply(raw, bw2, Kd)[0, :, :K].t()
    prob = guard_exp(weights / (bw ** 2) / 2)
    prob = prob / prob.sum(0, keepdim=True)
    if weights.shape[0] == 1:
        return prob
    prob = prob - prob.min(1, keepdim=True)[0]
    return prob / (prob.max(1, keepdim=True)[0] + EPS)


def match(target, pred_labels):
    """Hungarian matching of predicted clusters to gt segments on 1 - relaxed IoU (50x50)"""
    rids, cids = solve_dense(iou_cost_host(pred_labels, target))
    return rids, cids, _unique_labels(target), _unique_labels(pred_labels)


def _unique_labels(a):
    """np.unique for a vector of small non-negative integer labels (sorted distinct values, same dtype) without the sort"""
    a = np.asarray(a)
    if a.dtype.kind not in "iu" or a.size == 0 or a.min() < 0:
        return np.unique(a)
    return np.nonzero(np.bincount(a.reshape(-1)))[0].astype(a.dtype)


def rotation_matrix_a_to_b(A, B):
    """3x3 rotation taking unit vector A onto B (numpy, float64)"""
    cos, sin = np.dot(A, B), np.linalg.norm(np.cross(B, A))
    v = B - np.dot(A, B) * A
    v = v / (np.linalg.norm(v) + EPS)
    w = np.cross(B, A)
    w = w / (np.linalg.norm(w) + EPS)
    Fm = np.stack([A, v, w], 1)
    G = np.array([[cos, -sin, 0], [sin, cos, 0], [0, 0, 1]])
    try:
        return Fm @ G @ np.linalg.inv(Fm)
    except np.linalg.LinAlgError:
        return np.eye(3, dtype=np.float32)


def pca_torch(X):
    """eigen-decomposition of X^T X as (eigenvalues (3,2) real/imag, eigenvectors (3,3)); LAPACK geev on the host so
    that eigenvector signs are the reference's (torch.eig on a 3x3 runs on the CPU there as well)."""
    Xd = X.detach().double()
    cov = (Xd.t() @ Xd).float().cpu()       # (float64 accumulation, rounded once: see pnb200.fitstage.standardize_batched)
    w, v = torch.linalg.eig(cov)
    return torch.stack([w.real, w.imag], 1), v.real


def standardize_point_torch(point, weights):
    """centre on the weighted mean of the confident points, rotate the minor PCA axis onto x, scale every axis by the
    extent of the (weighted) confident points.  Returns (points, std (1,3), mean (3,), R (3,3)).
    The confident subset (w > 0.8, or the top N/4 | N/2 weights when fewer than 400 pass, reference :516-522) is kept
    as a device-side MASK instead of an index list: sums / extrema over the subset become masked reductions, so the
    only host round trip left is the 3x3 covariance needed by LAPACK (eigenvector signs must be the reference's)."""
    from pnb200.staging import arena
    N = weights.shape[0]
    w0 = weights[:, 0].detach()
    thr = w0 > 0.8
    kk = N // 4 if N >= 7500 else N // 2
    top = torch.zeros(N, dtype=torch.bool, device=point.device).scatter_(
        0, torch.topk(w0, kk)[1], torch.ones(kk, dtype=torch.bool, device=point.device))
    mask = torch.where(thr.sum() < 400, top, thr).unsqueeze(1)                 # (N,1) bool
    m = mask.to(point.dtype)
    wm = weights * m
    mean = (point * wm).sum(0) / (wm.sum() + EPS)
    point = point - mean
    Xm = (point * m).detach()
    S, U = pca_torch(Xm)
    smallest = U[:, int(torch.min(S[:, 0], 0)[1])].numpy()
    R_np = rotation_matrix_a_to_b(smallest, np.array([1, 0, 0])).astype(np.float32)
    R = arena("fit", point.device).upload(R_np, point.device)
    # the inverse used when mapping the surface back (torch.inverse on the device would block on its info read-back)
    R._pn_inv = arena("fit", point.device).upload(np.linalg.inv(R_np).astype(np.float32), point.device)
    point = (R @ point.t()).t()
    wp = point * weights
    inf = torch.full_like(wp, float("inf"))
    std = (torch.where(mask, wp, -inf).max(0)[0] - torch.where(mask, wp, inf).min(0)[0]).abs().reshape(1, 3).detach()
    return point / (std + EPS), std, mean, R


def standardize_points_torch(points, weights):
    outs = [standardize_point_torch(points[i], weights) for i in range(points.shape[0])]
    return torch.stack([o[0] for o in outs], 0), [o[1] for o in outs], [o[2] for o in outs], [o[3] for o in outs]


# ------------------------------------------------------------------------------------------------ up-sampling
def _nearest5(points):
    """(N,3) -> (N,5) indices of the 5 smallest squared distances sum((p_i - p_j)**2), nearest (the point itself) first:
    the kNN kernel with the squared-difference metric (csrc/knn.cu metric 2; ties to the lower index)"""
    from pnb200 import ops
    if points.shape[0] < 5:
        raise ValueError("up-sampling needs at least 5 points")
    return ops.knn_graph(points.detach().float().contiguous().unsqueeze(0), 5, 2, out_dtype=torch.int64)[0]


def up_sample_points_torch(points, times=1):
    """append, per point, the centroid of its 4 nearest other points (reference :150-163); N -> 2^times N points.
    The reference materialises the (N,N,3) difference tensor and a full-row topk; here the graph comes from the kNN kernel."""
    for _ in range(times):
        idx = _nearest5(points)
        points = torch.cat([points, torch.mean(points[idx[:, 1:]], 1)])
    return points


def up_sample_points_torch_memory_efficient(points, times=1):
    """reference :166-189: centroid over all 5 nearest (the point included); rows beyond the last whole block of
    min(N, 100) rows get no new point (the reference's block loop drops them)"""
    for _ in range(times):
        n = points.shape[0]
        blk = min(n, 100)
        m = (n // blk) * blk
        idx = _nearest5(points)[:m]
        points = torch.cat([points, torch.mean(points[idx], 1)])
    return points


def up_sample_points_in_range(points, weights, a_min, a_max):
    """reference :202-219 (np.random.choice consumed in the same order)"""
    N = points.shape[0]
    if N > a_max:
        L = np.random.choice(np.arange(N), a_max, replace=False)
        return points[L], weights[L]
    while True:
        points = up_sample_points_torch(points)
        weights = torch.cat([weights, weights], 0)
        if points.shape[0] >= a_max:
            break
    L = np.random.choice(np.arange(points.shape[0]), a_max, replace=False)
    return points[L], weights[L]


def up_sample_points_torch_in_range(points, a_min, a_max):
    """reference :222-237"""
    N = points.shape[0]
    if N > a_max:
        L = np.random.choice(np.arange(N), a_max, replace=False)
        return points[L]
    while True:
        points = up_sample_points_torch(points)
        if points.shape[0] >= a_max:
            break
    L = np.random.choice(np.arange(points.shape[0]), a_max, replace=False)
    return points[L]


def sample_points_from_control_points_(nu, nv, outputs, batch_size, input_size_u=20, input_size_v=20):
    """(B, cu*cv, 3) control points -> (B, g*g, 3) surface samples Nu P Nv^T"""
    P = outputs.reshape(outputs.shape[0], input_size_u, input_size_v, 3)
    return spline_eval(P, nu.to(P.device), nv.to(P.device))


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
