"""The fit half of Evaluation.fitting_loss for a whole batch of shapes at once.

Reference call sequence per shape (src/residual_utils.py:86-208 -> residual_train_mode :150 -> fit_one_shape_torch
src/primitive_forward.py:925-1047 -> ResidualLoss.residual_loss src/primitives.py:36 -> separate_losses :333):
membership weights of every cluster, Hungarian match to the gt segments (host), per matched segment a primitive or
SplineNet fit from the weights, residual of the gt points, mean over the segments.

The per-shape python loop issued ~1 400 sub-10-us launches per shape (22 900 device activities per step at B = 16 with 8
segments per shape; profiles/r02_run2_step_kernels.md).  Here every tensor is indexed by (shape, slot) with a fixed number
of SLOTS = 64 weight columns per shape (the guards allow at most 49 clusters), and each stage is one launch or one small
batched torch expression for ALL shapes:

  weights      (B,N,64)   = embedding @ centres^T (one bmm), exp / normalise / min-max (fitting_utils.py:306-325)
  moments      (B,64,55)  one launch                     csrc/fit.cu
  solve        (B*64,8)   one launch, with Jacobians      csrc/fitsolve.cu  (plane / sphere / cylinder / cone)
  cone angle   (cones,)   batched torch expression        primitive_forward.py:833-842
  residuals    (B,64)     one launch                      csrc/primitives.cu
  splines      standardisation batched over all spline segments of the step (ONE read-back of the 3x3 covariances for the
               LAPACK eigenvectors), ONE SplineNet forward per decoder, one spline evaluation, padded two-sided Chamfer
  losses       one (B, terms) matrix product; every statistic of the step comes back in ONE transfer

Only the host part (Hungarian matching, majority type of a gt segment, size rules) remains a loop over shapes.
"""
import numpy as np
import torch

from . import fitting as F
from .staging import arena

EPS = F.EPS
SLOTS = 64
CLOSED_IDS, OPEN_IDS = (0, 6, 7, 9), (2, 8)
ANALYTIC_KIND = {1: 0, 5: 1, 4: 2, 3: 3}               # reference primitive id -> kernel kind (plane, sphere, cylinder, cone)
KIND_NAME = {0: "plane", 1: "sphere", 2: "cylinder", 3: "cone"}
STATS = {"analytic_fits": 0, "open_spline_fits": 0, "closed_spline_fits": 0}


# ------------------------------------------------------------------------------------------------ host plan
class Plan:
    """what the host decides from the matching: which slot of which shape is fitted as what, and from which gt points"""

    def __init__(self, B, N):
        self.kind = np.full((B, SLOTS), -1, np.int32)          # analytic kernel kind per slot
        self.seg = np.full((B, N), -1, np.int32)               # slot whose residual a gt point enters (analytic only)
        self.splines = []                                      # (shape, slot, key, closed, gt point indices)
        self.keys = [dict() for _ in range(B)]                 # per shape: key (cluster id) -> ("kind"|"open"|"closed"|None, slot)
        self.matching = []


def make_plan(labels, cluster_np, primitives, N, match_fn):
    """fit_one_shape_torch's segment rules (training mode) for every shape; host only.  Per shape the point counts of every gt
    segment / cluster and the majority primitive of every gt segment come from three bincounts (not from one boolean pass over
    the points per matched pair), the per-point slot table from one look-up."""
    B = labels.shape[0]
    plan = Plan(B, N)
    n_half = (N + 1) // 2
    n_quarter = (n_half + 1) // 2
    for b in range(B):
        rows, cols, _, unique_pred = match_fn(labels[b], cluster_np[b])
        plan.matching.append((rows, cols))
        lab = np.asarray(labels[b]).astype(np.int64)
        cl = np.asarray(cluster_np[b]).astype(np.int64)
        pr = np.asarray(primitives[b]).astype(np.int64)
        L = int(max(lab.max(), cl.max(), np.max(cols), np.max(unique_pred))) + 1
        n_lab = np.bincount(lab, minlength=L)
        n_cl = np.bincount(cl, minlength=L)
        P = int(pr.max()) + 1
        # majority primitive id of every gt segment; argmax = smallest of the most frequent, like scipy.stats.mode
        prim_of = np.bincount(lab * P + pr, minlength=L * P).reshape(L, P).argmax(1)
        slot_of_label = np.full(L, -1, np.int32)
        spline_count = 0
        for index, i in enumerate(unique_pred):
            c = int(cols[i])
            if n_lab[c] == 0 or n_cl[int(i)] == 0:
                continue
            prim = int(prim_of[c])
            key = int(i)
            if prim in CLOSED_IDS + OPEN_IDS:
                spline_count += 1
                if spline_count > 4 or n_half < 20 or n_half < 100:
                    plan.keys[b][key] = (None, index)
                    continue
                plan.splines.append((b, index, key, prim in CLOSED_IDS, np.nonzero(lab == c)[0]))
                plan.keys[b][key] = ("closed" if prim in CLOSED_IDS else "open", index)
            elif prim in ANALYTIC_KIND:
                if n_quarter < 20:
                    plan.keys[b][key] = (None, index)
                    continue
                plan.kind[b, index] = ANALYTIC_KIND[prim]
                slot_of_label[c] = index
                plan.keys[b][key] = (KIND_NAME[ANALYTIC_KIND[prim]], index)
            else:
                raise ValueError(f"unknown primitive id {prim} (the reference handles 0-9 except torus-like ids)")
        plan.seg[b] = slot_of_label[lab]
    return plan


# ------------------------------------------------------------------------------------------------ device stages
def normalized_weights(raw, bws, K, stage):
    """weights_normalize (fitting_utils.py:306-325) for every shape: raw (B,N,SLOTS) centre . point similarities, columns
    >= K[b] are padding -> exp(clamp(raw / bw^2 / 2)), normalised over the clusters of a point, then min-max over the
    points of a cluster (skipped for single-cluster shapes, :318-319).  Padded columns come out as exact zeros.
    On the device this is csrc/weights.cu (2 launches forward, 2 backward); the torch expression below is its reference
    (tests) and the path taken by CPU tensors (host-logic tests only)."""
    if raw.is_cuda:
        bw2 = (bws.detach().double() ** 2).float().contiguous()
        Kd = stage.upload(np.asarray(K, np.int32), raw.device)
        return F.WeightsNormalizeFn.apply(raw, bw2, Kd)
    return normalized_weights_torch(raw, bws, K, stage)


def normalized_weights_torch(raw, bws, K, stage):
    """the same as ~12 torch expressions over (B,N,SLOTS) (+ ~25 autograd kernels)"""
    B, N, S = raw.shape
    dev = raw.device
    Kh = np.asarray(K, np.int64)
    colmask = stage.upload((np.arange(S)[None, :] < Kh[:, None]).astype(np.float32).reshape(B, 1, S), dev)
    single = stage.upload((Kh == 1).reshape(B, 1, 1), dev)
    bw2 = (bws.detach().double() ** 2).float().view(B, 1, 1)
    prob = torch.exp(torch.clamp(raw / bw2 / 2, min=-75.0, max=75.0)) * colmask
    prob = prob / prob.sum(2, keepdim=True)
    mm = prob - prob.min(1, keepdim=True)[0]
    mm = mm / (mm.max(1, keepdim=True)[0] + EPS)
    return torch.where(single, prob, mm)


def cone_angles(par32, bad, Wn, points, cones, n_quarter, stage):
    """weighted mean opening angle of every cone slot (primitive_forward.py:833-842) -> (len(cones),)"""
    dev = par32.device
    B = points.shape[0]
    flat = stage.upload(np.array([b * SLOTS + c for b, c in cones], np.int64), dev)
    bi = stage.upload(np.array([b for b, _ in cones], np.int64), dev)
    ci = stage.upload(np.array([c for _, c in cones], np.int64), dev)
    pq = points[:, 0::4][bi]                                               # (C, nq, 3)
    wq = Wn[:, 0::4][bi, :, ci] + EPS                                      # (C, nq)
    apex, axis = par32[flat, 0:3], par32[flat, 3:6]
    diff = torch.nn.functional.normalize(pq - apex.unsqueeze(1), p=2, dim=2)
    v = torch.clamp((diff * axis.unsqueeze(1)).sum(2).abs(), max=0.999)
    theta = (wq * torch.acos(v)).sum(1) / (wq.sum(1) + EPS)
    theta = torch.clamp(theta, min=1e-3, max=3.142 / 2 - 1e-3)
    return flat, theta * (1.0 - bad[flat])


def standardize_batched(P, w, rotation_fn, stage):
    """standardize_point_torch (fitting_utils.py:512-553) for E point sets at once: P (E,n,3), w (E,n) membership weights.
    -> standardised points (E,n,3), extents (E,3), means (E,3), R (E,3,3), R^-1 (E,3,3).  One host round trip: the 3x3
    covariances go to LAPACK geev (eigenvector SIGNS must be the reference's: the SplineNet is not rotation invariant)."""
    E, n, _ = P.shape
    dev = P.device
    w0 = w.detach()
    thr = w0 > 0.8
    kk = n // 4 if n >= 7500 else n // 2
    top = torch.zeros((E, n), dtype=torch.bool, device=dev).scatter_(
        1, torch.topk(w0, kk, dim=1)[1], torch.ones((E, kk), dtype=torch.bool, device=dev))
    mask = torch.where(thr.sum(1, keepdim=True) < 400, top, thr).unsqueeze(2)        # (E,n,1)
    m = mask.to(P.dtype)
    w3 = w0.unsqueeze(2)
    wm = w3 * m
    mean = (P * wm).sum(1) / (wm.sum(1) + EPS)                                        # (E,3)
    Pc = P - mean.unsqueeze(1)
    Xm = Pc * m
    # covariance accumulated in float64 and rounded once: its fp32 value then does not depend on how many point sets share
    # the launch (a batched product sums in another order than a single one), so LAPACK sees the same matrix -- hence returns
    # the same eigenvector SIGNS -- whether a segment is standardised alone or in a batch
    Xd = Xm.double()
    cov = (Xd.transpose(1, 2) @ Xd).float().cpu()                                     # THE read-back of the spline stage
    evals, evecs = torch.linalg.eig(cov)
    R_np = np.empty((E, 3, 3), np.float32)
    Rinv_np = np.empty((E, 3, 3), np.float32)
    for e in range(E):
        smallest = evecs[e].real[:, int(torch.min(evals[e].real, 0)[1])].numpy()
        R_np[e] = rotation_fn(smallest, np.array([1, 0, 0])).astype(np.float32)
        Rinv_np[e] = np.linalg.inv(R_np[e]).astype(np.float32)
    R, Rinv = stage.upload(R_np, dev), stage.upload(Rinv_np, dev)
    Pr = torch.bmm(R, Pc.transpose(1, 2)).transpose(1, 2)
    wp = Pr * w3
    inf = torch.full_like(wp, float("inf"))
    std = (torch.where(mask, wp, -inf).max(1)[0] - torch.where(mask, wp, inf).min(1)[0]).abs()   # (E,3)
    return Pr / (std.unsqueeze(1) + EPS), std, mean, R, Rinv


def spline_stage(plan, fitter, Wn, points, rotation_fn, stage):
    """every spline segment of the step: SplineNet control points from the standardised half-decimated shape + weights,
    surface samples mapped back, two-sided Chamfer to the segment's gt points (distance_from_bspline, primitives.py:197).
    Returns (distances (E,), list of reconstructed sample tensors (1, 900 | 930, 3) in plan.splines order)."""
    dev = points.device
    B, N, _ = points.shape
    E = len(plan.splines)
    bi = stage.upload(np.array([s[0] for s in plan.splines], np.int64), dev)
    ci = stage.upload(np.array([s[1] for s in plan.splines], np.int64), dev)
    P = points[:, 0::2][bi]                                                 # (E, nh, 3) (points are constants here)
    w = Wn[:, 0::2][bi, :, ci] + EPS                                        # (E, nh), differentiable
    with torch.no_grad():
        Ps, std, mean, R, Rinv = standardize_batched(P, w, rotation_fn, stage)
    fitter._basis_on(dev)
    nu, nv = fitter.nu, fitter.nv
    g = nu.shape[0]
    closed_flags = np.array([s[3] for s in plan.splines])
    recs = [None] * E
    dists = [None] * E
    for closed in (False, True):
        sel = np.nonzero(closed_flags == closed)[0]
        if sel.size == 0:
            continue
        net = fitter.closed_control_decoder if closed else fitter.open_control_decoder
        sel_d = stage.upload(sel.astype(np.int64), dev)
        Eg = int(sel.size)
        out = net(Ps[sel_d].permute(0, 2, 1), w[sel_d])                     # (Eg, 400, 3)
        rec = F.spline_eval(out.reshape(Eg, 20, 20, 3), nu, nv)            # (Eg, g*g, 3)
        rec = torch.bmm(rec * std[sel_d].unsqueeze(1), Rinv[sel_d].transpose(1, 2)) + mean[sel_d].unsqueeze(1)
        if closed:
            rec = rec.reshape(Eg, g, g, 3)
            rec = torch.cat([rec, rec[:, 0:1]], 1).reshape(Eg, (g + 1) * g, 3)
        # gt points of every segment, padded to the longest with repeats of the segment's first point (a repeat cannot
        # change a minimum; the padded tail is masked out of the mean)
        lens = np.array([plan.splines[e][4].shape[0] for e in sel])
        Mmax = int(lens.max())
        gidx = np.empty((Eg, Mmax), np.int64)
        for r, e in enumerate(sel):
            b, idx = plan.splines[e][0], plan.splines[e][4]
            gidx[r, :idx.shape[0]] = b * N + idx
            gidx[r, idx.shape[0]:] = b * N + idx[0]
        gt = points.reshape(B * N, 3)[stage.upload(gidx.reshape(-1), dev)].reshape(Eg, Mmax, 3)
        valid = stage.upload((np.arange(Mmax)[None, :] < lens[:, None]).astype(np.float32), dev)
        inv_len = stage.upload((1.0 / lens).astype(np.float32), dev)
        d_gt = (F.nearest_sqdist(gt, rec) * valid).sum(1) * inv_len          # every gt point -> nearest sample
        d_pred = F.nearest_sqdist(rec, gt).mean(1)                          # every sample -> nearest gt point
        d = (d_pred + d_gt) / 2.0
        for r, e in enumerate(sel):
            recs[e] = rec[r:r + 1]
            dists[e] = d[r]
        STATS["closed_spline_fits" if closed else "open_spline_fits"] += Eg
    return torch.stack(dists), recs


def run(evaluation, embedding, centers, K, bws, points, normals, labels, primitives, cluster_np, lamb, match_fn,
        rotation_fn):
    """embedding (B,N,d) unit rows; centers (B,SLOTS,d) kept mean-shift centres of every shape (columns >= K[b] are padding);
    bws (B,) bandwidths; points / normals (B,N,3); labels / primitives / cluster_np numpy (B,N).
    Returns dict(loss (B,) tensor, stats (B,2) float64 tensor [geometric mean, spline mean] with NaN where a shape has no
    such segment, has_terms (B,) numpy bool, plan, raw (B,N,SLOTS) similarities, parameters of the last shape)."""
    B, N, d = embedding.shape
    dev = embedding.device
    stage = arena("fitstage", dev)
    stage.reset()
    # (the similarities and membership weights need nothing from the plan: enqueued first, they run while the host plans)
    raw = torch.bmm(embedding, centers.transpose(1, 2))                                   # (B,N,SLOTS)
    Wn = normalized_weights(raw, bws, K, stage)
    plan = make_plan(labels, cluster_np, primitives, N, match_fn)
    n_half = (N + 1) // 2
    n_quarter = (n_half + 1) // 2
    kind = stage.upload(plan.kind, dev)
    terms = []                       # (shape, value index, weight, is_spline)
    values = []
    par32 = None
    n_analytic = int((plan.kind >= 0).sum())
    if n_analytic:
        seg = stage.upload(plan.seg, dev)
        mom = F.MomentsBatchedFn.apply(Wn, points.contiguous().float(), normals.contiguous().float(), 0, 4, n_quarter, EPS)
        par, bad = F.FitSolveFn.apply(mom.view(B * SLOTS, F.NM), kind.view(-1), n_quarter)
        par32 = par.float()
        cones = [(int(b), int(c)) for b, c in zip(*np.nonzero(plan.kind == 3))]
        if cones:
            flat, theta = cone_angles(par32, bad, Wn, points, cones, n_quarter, stage)
            col6 = torch.full_like(flat, 6)
            par32 = par32 + torch.zeros_like(par32).index_put((flat, col6), theta)
        dist = F.ResidualBatchedFn.apply(par32.view(B, SLOTS, 8), points.contiguous().float(), seg, kind)
        values.append(dist.reshape(-1))
        for b, c in zip(*np.nonzero(plan.kind >= 0)):
            terms.append((int(b), int(b) * SLOTS + int(c), 1.0, False))
        STATS["analytic_fits"] += n_analytic
    recs = []
    if plan.splines:
        off = B * SLOTS if n_analytic else 0
        d_spl, recs = spline_stage(plan, evaluation.fitter, Wn, points, rotation_fn, stage)
        values.append(d_spl)
        for e, s in enumerate(plan.splines):
            terms.append((s[0], off + e, float(lamb), True))
    out = {"plan": plan, "raw": raw, "recs": recs, "par32": par32, "Wn": Wn}
    if not terms:
        out.update(loss=torch.zeros(B, device=dev), stats=torch.full((B, 2), float("nan"), dtype=torch.float64, device=dev),
                   has_terms=np.zeros(B, bool), D=None, terms=terms)
        return out
    D = torch.cat(values)
    D = torch.where(D > 1, torch.full_like(D, 0.1), D)          # degenerate fits count as the constant 0.1 (:343-346)
    T = D.shape[0]
    A = np.zeros((B, T), np.float32)
    G = np.zeros((2 * B, T), np.float64)
    cnt = np.zeros(B)
    cg, cs = np.zeros(B), np.zeros(B)
    for b, j, wt, is_spl in terms:
        cnt[b] += 1
        (cs if is_spl else cg)[b] += 1
    for b, j, wt, is_spl in terms:
        A[b, j] = wt / cnt[b]
        if is_spl:
            G[B + b, j] = 1.0 / cs[b]
        else:
            G[b, j] = 1.0 / cg[b]
    loss = torch.mv(stage.upload(A, dev), D)
    stats = torch.mv(stage.upload(G, dev), D.detach().double()).view(2, B).t()
    nan = np.full((B, 2), 0.0)
    nan[cg == 0, 0] = np.nan
    nan[cs == 0, 1] = np.nan
    out.update(loss=loss, stats=stats + stage.upload(nan, dev), has_terms=cnt > 0, D=D, terms=terms)
    return out


def parameters_of_shape(out, b):
    """the reference's `fitter.fitting.parameters` dictionary (key = cluster id) of shape b from the batched tables"""
    plan, par32, recs = out["plan"], out["par32"], out["recs"]
    params = {}
    spl = {(s[0], s[2]): e for e, s in enumerate(plan.splines)}
    for key, (what, slot) in plan.keys[b].items():
        if what is None:
            params[key] = None
        elif what in ("open", "closed"):
            params[key] = [what + "-spline", recs[spl[(b, key)]]]
        else:
            q = par32[b * SLOTS + slot]
            if what == "plane":
                params[key] = ["plane", q[0:3].reshape(3, 1), q[3]]
            elif what == "sphere":
                params[key] = ["sphere", q[0:3].reshape(1, 3), q[3]]
            elif what == "cylinder":
                params[key] = ["cylinder", q[0:3].reshape(3, 1), q[3:6].reshape(1, 3), q[6]]
            else:
                params[key] = ["cone", q[0:3].reshape(1, 3), q[3:6].reshape(3, 1), q[6]]
    return params


def segment_distances(out, b):
    """{key: (kind name, residual tensor after the degenerate rule)} of shape b (debugging / tests; reads nothing back)"""
    plan, D = out["plan"], out["D"]
    n_analytic = int((plan.kind >= 0).sum())
    off = plan.kind.shape[0] * SLOTS if n_analytic else 0
    res = {}
    spl = {(s[0], s[2]): e for e, s in enumerate(plan.splines)}
    for key, (what, slot) in plan.keys[b].items():
        if what is None:
            continue
        if what in ("open", "closed"):
            res[key] = (what + "-spline", D[off + spl[(b, key)]])
        else:
            res[key] = (what, D[b * SLOTS + slot])
    return res
