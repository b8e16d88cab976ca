"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of the differentiable mean-shift clustering.

Follows (reference): MeanShift.compute_bandwidth src/mean_shift.py:115-137, mean_shift_ :45-79, nms :139-179,
mean_shift :19-43; guard loops :81-96 and src/residual_utils.py:69-84.  Host RNG (np.random.shuffle) is consumed in
the reference's order."""
import numpy as np
import torch


def _gexp(x):
    return torch.exp(torch.clamp(x, min=-75.0, max=75.0))


def compute_bandwidth(X, num_samples, quantile, rng=np.random):
    N = X.shape[0]
    L = np.arange(N)
    rng.shuffle(L)
    Xs = X[L[0:num_samples]]
    dist = 2 - 2 * Xs @ Xs.t()
    K = int(quantile * num_samples)
    kth = torch.topk(dist, k=K, dim=1, largest=False)[0][:, -1]
    return torch.sqrt(torch.clamp(kth, min=1e-6)).mean()


def mean_shift_iters(X, b, iterations):
    Y = X.clone()
    for _ in range(iterations):
        dist = 2.0 - 2.0 * Y @ X.t()
        K = _gexp(-dist / (b ** 2) / 2)
        Dinv = 1 / K.sum(1, keepdim=True)
        M = (K @ X) * Dinv - Y
        Y = Y + M
        Y = Y / torch.norm(Y, dim=1, p=2, keepdim=True)
    return Y


def nms(centers, X, b):
    member = torch.min(2.0 - 2.0 * centers @ X.t(), 0)[1]
    uniq, counts = np.unique(member.numpy(), return_counts=True)
    num = torch.zeros(X.shape[0])
    num[uniq] = torch.from_numpy(counts.astype(np.float32))
    dist = 2.0 - 2.0 * centers @ centers.t()
    nbrs = (dist < b).float()
    ids = torch.unique(torch.max(nbrs[uniq] * num.reshape(1, -1), 1)[1])
    kept = centers[ids]
    labels = torch.max(kept @ X.t(), 0)[1]
    return kept, ids, labels


def mean_shift(X, num_samples, quantile, iterations, bw=None, do_nms=True, rng=np.random):
    if bw is None:
        with torch.no_grad():
            bw = torch.clamp(compute_bandwidth(X, num_samples, quantile, rng), min=0.003)
    Y = mean_shift_iters(X, bw, iterations)
    if not do_nms:
        return Y, bw
    with torch.no_grad():
        _, ids, labels = nms(Y, X, bw)
    return Y, Y[ids], bw, labels


def guard_mean_shift(X, quantile, iterations, num_samples=10000, growth=1.2, rng=np.random):
    """Evaluation.guard_mean_shift (residual_utils.py:69-84): growth 1.2, 10000 samples;
    MeanShift.guard_mean_shift (mean_shift.py:81-96): growth 2, 5000 samples."""
    while True:
        Y, center, bw, labels = mean_shift(X, num_samples, quantile, iterations, rng=rng)
        if torch.unique(labels).shape[0] > 49:
            quantile *= growth
        else:
            break
    return center, bw, labels


def sparse_rows_backward(g_R, Ynew_R, Yprev_R, den_R, unorm_R, X, b):
    """Closed-form backward of ONE mean-shift iteration for a subset R of the rows (the math of the product's
    ms_bwd_sparse_kernel, written out with torch on the CPU; checked against autograd through mean_shift_iters in
    tests/test_oracle_golden.py).  Row i of Y_t depends on row i of Y_{t-1} and on X only, so a gradient that is non-zero
    in rows R stays in rows R.  Inputs: gradient w.r.t. rows R of Y_t, rows R of Y_t and Y_{t-1}, their kernel row sums
    `den` and pre-normalisation norms `unorm`, all of X, bandwidth b.  Returns (grad w.r.t. rows R of Y_{t-1}, grad w.r.t. X).
    u = K X / den ; Y_t = u / |u| ; K = exp(clamp((S - 1) / b^2)), S = Y_{t-1} X^T"""
    c = 1.0 / (b * b)
    g_u = (g_R - Ynew_R * (Ynew_R * g_R).sum(1, keepdim=True)) / unorm_R[:, None]
    Gn = g_u / den_R[:, None]
    gd = -((g_u * Ynew_R).sum(1) * unorm_R) / den_R
    S = Yprev_R @ X.t()
    e = (S - 1.0) * c
    clamped = (e > 75.0) | (e < -75.0)
    K = torch.exp(torch.clamp(e, -75.0, 75.0))
    gS = torch.where(clamped, torch.zeros_like(S), (Gn @ X.t() + gd[:, None]) * K * c)
    return gS @ X, gS.t() @ Yprev_R + K.t() @ Gn
