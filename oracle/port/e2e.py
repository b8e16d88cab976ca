"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of the fit half of the reference's end-to-end path.

Covers SURVEY.md §8 rows a6 (SplineNet `DGCNNControlPoints`), a27/a28 (input standardisation, open / closed spline
forward passes), a29 (`fit_one_shape_torch`, training mode), a31 (`Evaluation.fitting_loss` / `residual_train_mode` /
`separate_losses`) and a32 (`match`, `relaxed_iou_fast`, `to_one_hot`); the analytic fits, residuals and Chamfer come from
`oracle/port/fitting.py`, the clustering from `oracle/port/meanshift.py`.  Each function cites the reference lines it
restates.  Pinned in `tests/test_oracle_golden.py` against vectors dumped from the UNMODIFIED reference
(`tests/golden/splinenet.npz`, `e2e.npz`, `e2e_nocyl.npz`).  Nothing in the product package imports this file.

Training-mode path only (`eval=False`, `if_optimize=False`): that is what `train_parsenet_e2e.py:230` calls.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import fitting as fit
from . import meanshift as pms

EPS = float(np.finfo(np.float32).eps)
CLOSED_IDS, OPEN_IDS = (0, 6, 7, 9), (2, 8)


# ------------------------------------------------------------------------------------------------------ SplineNet (a6)
def knn_feature_space(x, k):
    """src/model.py:9-22: per shape D = -|x_i|^2 + 2 x_i.x_j - |x_j|^2 (this association), top-k largest; x (B,C,N)"""
    out = []
    for b in range(x.shape[0]):
        xb = x[b:b + 1]
        inner = -2 * torch.matmul(xb.transpose(2, 1), xb)
        sq = torch.sum(xb ** 2, dim=1, keepdim=True)
        out.append((-sq - inner - sq.transpose(2, 1))[0])
    return torch.stack(out, 0).topk(k=k, dim=-1)[1]


def edge_features(x, k):
    """src/model.py:25-53: (B,C,N) -> (B,2C,N,k) = [x_j - x_i ; x_i] over the k nearest neighbours in x's own space"""
    B, C, N = x.shape
    with torch.no_grad():
        idx = knn_feature_space(x, k)
    xt = x.transpose(2, 1)                                                    # (B,N,C)
    nb = xt.reshape(B * N, C)[(idx + torch.arange(B).view(-1, 1, 1) * N).reshape(-1)].reshape(B, N, k, C)
    ctr = xt.unsqueeze(2).expand(B, N, k, C)
    return torch.cat([nb - ctr, ctr], 3).permute(0, 3, 1, 2)


def _bn(x, sd, name, train):
    return F.batch_norm(x, sd[name + ".running_mean"].clone(), sd[name + ".running_var"].clone(), sd[name + ".weight"],
                        sd[name + ".bias"], training=train, momentum=0.1, eps=1e-5)


def splinenet_fwd(sd, x, k=10, weights=None, train=False):
    """DGCNNControlPoints.forward (src/model.py:138-180) as a function of a state-dict with the reference's keys:
    four edge-convs (Conv2d 1x1 no bias + BatchNorm + LeakyReLU(0.2) + max over neighbours), concat, Conv1d + BN + LReLU,
    optional per-point weights, global max, two Conv1d + BN + ReLU, Conv1d, tanh -> (B, 400, 3).  mode 0 / 1 differ only
    in the layer widths, i.e. in the state-dict."""
    B = x.shape[0]
    feats, cur = [], x
    for i in (1, 2, 3, 4):
        y = F.conv2d(edge_features(cur, k), sd[f"conv{i}.0.weight"])
        cur = F.leaky_relu(_bn(y, sd, f"bn{i}", train), 0.2).max(dim=-1)[0]
        feats.append(cur)
    y = F.leaky_relu(_bn(F.conv1d(torch.cat(feats, 1), sd["conv5.0.weight"]), sd, "bn5", train), 0.2)
    if weights is not None:
        y = y * weights.reshape(1, 1, -1)
    g = F.adaptive_max_pool1d(y, 1).view(B, -1, 1)
    g = F.relu(_bn(F.conv1d(g, sd["conv6.weight"], sd["conv6.bias"]), sd, "bn6", train))
    g = F.relu(_bn(F.conv1d(g, sd["conv7.weight"], sd["conv7.bias"]), sd, "bn7", train))
    g = torch.tanh(F.conv1d(g, sd["conv8.weight"], sd["conv8.bias"])[:, :, 0])
    return g.view(B, -1, 3)


# ------------------------------------------------------------------------------------------------------ a27 / a28
def rotation_a_to_b(a, b):
    """fitting_utils.py:556-579 (numpy, float64): rotation with R a = b built from the frame (a, b_perp, b x a)"""
    cos, sin = np.dot(a, b), np.linalg.norm(np.cross(b, a))
    v = b - np.dot(a, b) * a
    v = v / (np.linalg.norm(v) + EPS)
    w = np.cross(b, a)
    w = w / (np.linalg.norm(w) + EPS)
    Fm = np.stack([a, v, w], 1)
    G = np.array([[cos, -sin, 0], [sin, cos, 0], [0, 0, 1]])
    try:
        return Fm @ G @ np.linalg.inv(Fm)
    except np.linalg.LinAlgError:
        return np.eye(3, dtype=np.float32)


def standardize_point(point, weights):
    """fitting_utils.py:512-553: confident points (w > 0.8, else the top quarter / half), weighted mean, PCA of those
    points (general `eig` of X^T X like the reference), smallest-eigenvalue direction rotated onto x, anisotropic scale
    by the weighted extent"""
    sel = weights[:, 0] > 0.8
    if sel.sum() < 400:
        n = weights.shape[0]
        sel = torch.topk(weights[:, 0], n // 4 if n >= 7500 else n // 2)[1]
    mean = (point[sel] * weights[sel]).sum(0) / (weights[sel].sum() + EPS)
    point = point - mean
    Xc = point[sel]
    ev, U = torch.linalg.eig(Xc.t() @ Xc)
    direction = U.real[:, torch.min(ev.real, 0)[1]].detach().numpy()
    R = torch.from_numpy(rotation_a_to_b(direction, np.array([1, 0, 0])).astype(np.float32))
    point = (R @ point.t()).t()
    wp = point[sel] * weights[sel]
    std = (wp.max(0)[0] - wp.min(0)[0]).abs().reshape(1, 3).detach()
    return point / (std + EPS), std, mean, R


def _unstandardize(x, std, R, mean):
    return (torch.inverse(R) @ (x * std.reshape(1, 3)).t()).t() + mean


def forward_open_spline(points, weights, sd, nu, nv):
    """primitive_forward.py:34-85 with if_optimize=False: (m,3),(m,1) -> reconstructed surface (1,900,3)"""
    with torch.no_grad():
        p, std, mean, R = standardize_point(points, weights)
    cp = splinenet_fwd(sd, p.t().unsqueeze(0), 10, weights.t())
    rec = fit.sample_points_from_control_points_(nu, nv, cp, 1)
    return _unstandardize(rec[0], std, R, mean).unsqueeze(0)


def forward_closed_spline(points, weights, sd, nu, nv):
    """primitive_forward.py:347-397 with if_optimize=False: the first row of the 30 x 30 grid is appended -> (1,930,3)"""
    with torch.no_grad():
        p, std, mean, R = standardize_point(points, weights)
    cp = splinenet_fwd(sd, p.t().unsqueeze(0), 10, weights.t())
    rec = fit.sample_points_from_control_points_(nu, nv, cp, 1)
    grid = _unstandardize(rec[0], std, R, mean).reshape(30, 30, 3)
    return torch.cat([grid, grid[0:1]], 0).reshape(1, 930, 3)


# ------------------------------------------------------------------------------------------------------ a32
def one_hot(labels, width=50):
    """segment_utils.py:283-292"""
    t = torch.as_tensor(np.asarray(labels).astype(np.int64))
    return torch.zeros(t.shape[0], width).scatter_(1, t.unsqueeze(1), 1)


def match(target, pred_labels):
    """fitting_utils.py:362-376 + segment_utils.py:356-374: Hungarian matching on 1 - relaxed IoU of the two one-hot
    labelings (50 x 50); lapsolver.solve_dense -> scipy.optimize.linear_sum_assignment (same optimum)"""
    p, g = one_hot(pred_labels), one_hot(target)
    dots = p.t() @ g
    iou = dots / (p.sum(0).unsqueeze(1) + g.sum(0).unsqueeze(0) - dots + 1e-7)
    rows, cols = linear_sum_assignment(1.0 - iou.numpy())
    return rows, cols, np.unique(target), np.unique(pred_labels)


# ------------------------------------------------------------------------------------------------------ a29
def fit_one_shape(data, weights, nets, nu, nv):
    """primitive_forward.py:925-1047, training mode.  data: [points, normals, primitive id, gt points, _, (column, key)];
    every point set is halved, analytic primitives are halved again, at most 4 spline segments, small segments dropped.
    Returns ({key: gt points | None}, {key: [kind, params...] | None})"""
    gt_points, parameters, splines = {}, {}, 0
    for points, normals, prim, gpoints, _, (col, key) in data:
        prim = int(np.asarray(prim).reshape(-1)[0])
        w = weights[:, col:col + 1] + EPS
        points, normals, w = points[0::2], normals[0::2], w[0::2]
        skip = False
        if prim in CLOSED_IDS + OPEN_IDS:
            splines += 1
            skip = splines > 4
        else:
            points, normals, w = points[0::2], normals[0::2], w[0::2]
        if skip or points.shape[0] < 20 or (prim in CLOSED_IDS + OPEN_IDS and points.shape[0] < 100):
            gt_points[key], parameters[key] = None, None
            continue
        if prim in CLOSED_IDS:
            parameters[key] = ["closed-spline", forward_closed_spline(points, w, nets["closed"], nu, nv)]
        elif prim in OPEN_IDS:
            parameters[key] = ["open-spline", forward_open_spline(points, w, nets["open"], nu, nv)]
        elif prim == 1:
            a, d = fit.fit_plane(points, w)
            parameters[key] = ["plane", a.reshape(3, 1), d]
        elif prim == 3:
            apex, a, th = fit.fit_cone(points, normals, w)
            parameters[key] = ["cone", apex.reshape(1, 3), a.reshape(3, 1), th]
        elif prim == 4:
            parameters[key] = ["cylinder"] + list(fit.fit_cylinder(points, normals, w))
        elif prim == 5:
            parameters[key] = ["sphere"] + list(fit.fit_sphere(points, w))
        gt_points[key] = gpoints
    return gt_points, parameters


# ------------------------------------------------------------------------------------------------------ a31
def separate_losses(distance, gt_points, lamb):
    """residual_utils.py:333-378: mean over fitted segments, spline terms weighted by lamb, residuals > 1 -> 0.1"""
    terms, geo, spl = [], [], []
    for key in sorted(gt_points.keys()):
        if gt_points[key] is None:
            continue
        kind, d = distance[key]
        if d > 1:
            d = torch.ones(1)[0] * 0.1
            distance[key][1] = d
        if kind in ("closed-spline", "open-spline"):
            spl.append(d.item()); terms.append(d * lamb)
        else:
            geo.append(d.item()); terms.append(d)
    loss = torch.mean(torch.stack(terms)) if terms else torch.zeros(1)
    return [loss, float(np.mean(geo)) if geo else None, float(np.mean(spl)) if spl else None]


def residual_train_mode(points, normals, labels, cluster_ids, primitives, weights, bw, nets, nu, nv, lamb=1.0):
    """residual_utils.py:152-208: match clusters to gt segments, fit each matched segment from its membership weights
    (majority primitive type of the gt segment), residual of the gt points to the fit"""
    rows, cols, _, unique_pred = match(labels, cluster_ids)
    data = []
    for index, i in enumerate(unique_pred):
        gt_i = labels == cols[i]
        if gt_i.sum() == 0 or (cluster_ids == i).sum() == 0:
            continue
        kind = np.bincount(primitives[gt_i]).argmax()                 # scipy.stats.mode: smallest of the most frequent
        data.append([points, normals, kind, points[torch.from_numpy(gt_i)], None, (index, i)])
    w = fit.weights_normalize(weights, float(bw)).t()
    gt_points, parameters = fit_one_shape(data, w, nets, nu, nv)
    distance = {k: [v[0], fit.DISTANCES[v[0]](gt_points[k], v[1:])] for k, v in parameters.items() if v is not None}
    return separate_losses(distance, gt_points, lamb), parameters, distance


def fitting_loss(embedding, points, normals, labels, primitives, nets, quantile, iterations, lamb):
    """Evaluation.fitting_loss (residual_utils.py:86-150) for ONE shape: embedding (N,d) -> mean-shift clusters (with the
    1.2x quantile retry while more than 49 clusters come out, :69-84), membership weights = centre . embedding,
    residual_train_mode.  Returns ([loss, geometric mean, spline mean], parameters, distance, cluster ids)"""
    emb = F.normalize(embedding, p=2, dim=1)
    while True:
        _, center, bw, cluster_ids = pms.mean_shift(emb, 10000, quantile, iterations)
        if torch.unique(cluster_ids).shape[0] > 49:
            quantile *= 1.2
        else:
            break
    weights = center @ emb.t()
    nu, nv = fit.uniform_knot_bspline(20, 20, 3, 3, 30)
    nu, nv = torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32))
    loss, parameters, distance = residual_train_mode(points, normals, labels, cluster_ids.numpy(), primitives, weights, bw,
                                                     nets, nu, nv, lamb)
    return loss, parameters, distance, cluster_ids.numpy()
