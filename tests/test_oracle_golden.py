"""CPU tests: the oracle (oracle/c + oracle/port) against golden vectors dumped from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU tests then check CUDA against the oracle."""
import os

import numpy as np
import pytest
import torch


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("case", ["c3", "c64", "pn"])
def test_knn_oracle_vs_reference_with_ambiguity_audit(golden_dir, case):
    """Reference kNN (torch matmul + topk, src/PointNet.py:9-69) vs the pinned-order C oracle: they may only
    disagree where the reference's own distance values are within a few ulp of each other (near ties)."""
    from oracle import knn as oknn
    g = _load(golden_dir, "knn.npz")
    x = g[case + "_x"]; ref_idx = g[case + "_idx"].astype(np.int64)
    N, C, k, metric = g[case + "_meta"]
    xt = np.ascontiguousarray(x.transpose(0, 2, 1))
    mine = oknn.knn(xt, int(k), int(metric))
    agree = (mine == ref_idx)
    assert agree.mean() > 0.999
    bad = np.argwhere(~agree.all(-1))
    for b, i in bad:
        row = oknn.knn_row(xt[b], int(i), int(metric))
        scale = np.abs(row[ref_idx[b, i]]).max() + (xt[b] ** 2).sum(1).max()
        d_ref = row[ref_idx[b, i]]; d_mine = row[mine[b, i]]
        pos = np.where(mine[b, i] != ref_idx[b, i])[0]
        assert np.all(np.abs(d_ref[pos] - d_mine[pos]) <= 8 * np.finfo(np.float32).eps * scale), (b, i)


def _segnet_inputs(g):
    from oracle.port import common
    shapes = {k: eval(s) for k, s in zip(g["state_keys"], g["state_shapes"])}
    B, N, k, wseed, rseed = g["meta"]
    sd = common.seeded_state_dict(shapes, seed=int(wseed))
    for i in (1, 2, 3):
        for s in ("weight", "bias"):
            sd[f"encoder.conv{i}.1.{s}"] = sd[f"encoder.bn{i}.{s}"]
    return sd, int(k), int(rseed)


def test_segnet_port_matches_reference_forward_backward(golden_dir):
    from oracle.port import segnet as port
    g = _load(golden_dir, "segnet.npz")
    sd, k, rseed = _segnet_inputs(g)
    sd = {n: v.clone().requires_grad_(v.is_floating_point()) for n, v in sd.items()}
    x = torch.from_numpy(g["points"])
    idxs = [torch.from_numpy(g[f"idx{i}"].astype(np.int64)) for i in (1, 2, 3)]
    emb, lp, _, _, _ = port.segnet_fwd(sd, x, k, 5, idx_list=idxs)
    np.testing.assert_allclose(emb.detach().numpy(), g["embedding"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(lp.detach().numpy(), g["logprob"], rtol=1e-4, atol=2e-5)
    rng = np.random.RandomState(); np.random.seed(rseed)
    el = port.triplet_loss(emb, g["labels"], 1.0)
    nll = torch.nn.functional.nll_loss(lp, torch.from_numpy(g["prims"]))
    np.testing.assert_allclose(el.detach().numpy(), g["embed_loss"], rtol=1e-4)
    np.testing.assert_allclose(nll.detach().numpy(), g["nll"], rtol=1e-5)
    (el.sum() + nll).backward()
    checked = 0
    for key in g.files:
        if not key.startswith("grad:"):
            continue
        name = key[5:]
        if name.startswith("encoder.conv") and ".1." in name:
            continue                                # alias of encoder.bnX
        gr = sd[name].grad
        assert gr is not None, name
        t = gr.reshape(-1).double()
        got = np.array([t.sum().item(), t.norm().item()] + t[:14].tolist())
        np.testing.assert_allclose(got[1:], g[key][1:], rtol=2e-3, atol=1e-6, err_msg=name)
        checked += 1
    assert checked >= 30


def test_segnet_port_free_running_knn_close_to_reference(golden_dir):
    """with the oracle's own (pinned-order) kNN instead of the recorded reference graph: neighbour lists agree
    except at near ties, outputs agree for all but a handful of points."""
    from oracle.port import segnet as port
    g = _load(golden_dir, "segnet.npz")
    sd, k, _ = _segnet_inputs(g)
    x = torch.from_numpy(g["points"])
    with torch.no_grad():
        emb, lp, idxs, _, _ = port.segnet_fwd(sd, x, k, 5)
    for i, idx in enumerate(idxs, 1):
        assert (idx.numpy() == g[f"idx{i}"]).mean() > 0.995
    err = np.abs(emb.numpy() - g["embedding"]) / (np.abs(g["embedding"]) + 1e-2)
    assert (err < 1e-3).mean() > 0.98


@pytest.mark.parametrize("case", ["a", "b"])
def test_meanshift_port_matches_reference(golden_dir, case):
    from oracle.port import meanshift as port
    g = _load(golden_dir, "meanshift.npz")
    N, ncl, seed, it = [int(v) for v in g[case + "_meta"]]
    X = torch.from_numpy(g[case + "_X"]).requires_grad_()
    np.random.seed(seed)
    newX, center, bw, labels = port.mean_shift(X, N, float(g[case + "_q"]), it)
    np.testing.assert_allclose(bw.numpy(), g[case + "_bw"], rtol=1e-6)
    np.testing.assert_array_equal(labels.numpy(), g[case + "_labels"])
    np.testing.assert_allclose(newX.detach().numpy(), g[case + "_newX"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(center.detach().numpy(), g[case + "_center"], rtol=1e-5, atol=1e-6)
    gen = torch.Generator().manual_seed(seed + 100)
    w = torch.randn(center.shape, generator=gen); w2 = torch.randn(newX.shape, generator=gen) * 0.01
    ((center * w).sum() + (newX * w2).sum()).backward()
    np.testing.assert_allclose(X.grad.numpy(), g[case + "_gradX"], rtol=1e-4, atol=1e-6)


def test_cfg1_control_point_solve_restatement_vs_reference_golden(golden_dir):
    """BASELINE config 1 on the CPU: the pseudo-inverse restatement P = Nu^+ S (Nv^+)^T of the reference's gridded
    solve (src/approximation.py:308-334) reproduces the reference's float64 output and recovers the control grid."""
    import numpy as np
    import os
    g = np.load(os.path.join(golden_dir, "cfg1.npz"))
    nu, nv = g["nu"], g["nv"]
    pu, pv = np.linalg.inv(nu.T @ nu) @ nu.T, np.linalg.inv(nv.T @ nv) @ nv.T
    for S, want in ((g["S"], g["rec"]), (g["Sn"], g["rec_n"])):
        rec = np.einsum("iu,buvc,jv->bijc", pu, S.reshape(2, 30, 30, 3), pv)
        assert np.abs(rec - want).max() < 1e-10
    assert np.abs(g["rec"] - g["cp"]).max() < 1e-10 and np.abs(g["rec_k"] - g["cp"][0]).max() < 1e-10
    assert abs(np.linalg.cond(nu) - 162.52) < 0.01


# ------------------------------------------------------------------------------------------ fitting / loss stage
def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a))
    return t.requires_grad_() if grad else t


def _rel(got, want, rtol, name):
    got = got.detach().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    err = np.abs(got.reshape(want.shape) - want).max()
    assert err <= rtol * (np.abs(want).max() + 1e-30) + 1e-12, f"{name}: err {err:.3e} vs scale {np.abs(want).max():.3e}"


@pytest.mark.parametrize("kind", ["plane", "sphere", "cylinder", "cone"])
def test_fitting_port_matches_reference_fits_and_residuals(golden_dir, kind):
    """oracle/port/fitting.py vs Fit.fit_*_torch / ComputePrimitiveDistance of the unmodified reference (same torch, same
    CPU): outputs, gradient w.r.t. the membership weights through the custom SVD backward / regularised lstsq, residual
    distances and their parameter gradients"""
    from oracle.port import fitting as F
    g = _load(golden_dir, "fits.npz")
    P, Nn, W = _t(g[kind + "_p"]), _t(g[kind + "_n"]), _t(g[kind + "_w"], True)
    res = {"plane": lambda: F.fit_plane(P, W), "sphere": lambda: F.fit_sphere(P, W),
           "cylinder": lambda: F.fit_cylinder(P, Nn, W), "cone": lambda: F.fit_cone(P, Nn, W)}[kind]()
    loss = 0
    for i, r in enumerate(res):
        _rel(r, g[f"{kind}_out{i}"], 2e-5, f"{kind} out{i}")
        loss = loss + (r * _t(g[f"{kind}_coef{i}"]).reshape(r.shape)).sum()
    loss.backward()
    _rel(W.grad, g[kind + "_gw"], 1e-3, f"{kind} d/dweights")
    params = [_t(g[f"{kind}_par{i}"], True) for i in range(len(res))]
    d = F.DISTANCES[kind](_t(g[kind + "_q"]), params)
    _rel(d, g[kind + "_dist"], 1e-5, f"{kind} residual")
    d.backward()
    for i, p in enumerate(params):
        _rel(p.grad, g[f"{kind}_gpar{i}"], 1e-4, f"{kind} dpar{i}")


def test_fitting_port_matches_reference_chamfer_spline_losses(golden_dir):
    from oracle.port import fitting as F
    g = _load(golden_dir, "losses.npz")
    pred, gt = _t(g["pred"], True), _t(g["gt"], True)
    vals = {"cd": F.chamfer_distance(pred, gt), "cds": F.chamfer_distance(pred, gt, sqrt=True),
            "c0": F.chamfer_distance_one_side(pred, gt, 0), "c1": F.chamfer_distance_one_side(pred, gt, 1),
            "s1": F.chamfer_distance_single_shape(pred[0], gt[0]),
            "s2": F.chamfer_distance_single_shape(pred[0], gt[0], one_side=True)}
    for k, v in vals.items():
        assert abs(v.item() - float(g[k])) <= 1e-6 * abs(float(g[k])), k
    (vals["cd"] + 2 * vals["cds"] + 3 * vals["c0"] + 4 * vals["c1"] + 5 * vals["s1"] + 6 * vals["s2"]).backward()
    _rel(pred.grad, g["gpred"], 1e-5, "chamfer dpred"); _rel(gt.grad, g["ggt"], 1e-5, "chamfer dgt")
    # basis matrices: a different evaluation scheme (in-place Cox-de Boor) than the reference's triangular table
    nu, nv = F.uniform_knot_bspline(20, 20, 3, 3, 30)
    np.testing.assert_allclose(nu, g["nu"], rtol=0, atol=1e-13); np.testing.assert_allclose(nv, g["nv"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(F.uniform_knot_bspline(20, 20, 3, 3, 40)[0], g["nu40"], rtol=0, atol=1e-13)
    cpts = _t(g["cpts"], True)
    rec = F.sample_points_from_control_points_(_t(nu.astype(np.float32)), _t(nv.astype(np.float32)), cpts, 2)
    _rel(rec, g["rec"], 1e-5, "spline evaluation")
    (rec * _t(g["recw"])).sum().backward()
    _rel(cpts.grad, g["gcpts"], 1e-5, "spline evaluation grad")
    # open / closed spline training losses (train_open_splines.py:158-178)
    nu4 = _t(g["nu40"].astype(np.float32))
    outp = _t(g["tl_out"], True)
    cd1, _ = F.spline_reconstruction_loss_one_sided(nu4, nu4, outp, _t(g["tl_pts"]), 2, 20)
    lreg, perm = F.control_points_permute_reg_loss(outp, _t(g["tl_gtcp"]), 20)
    lap = F.laplacian_loss(outp.reshape(2, 20, 20, 3), perm)
    lclosed, _ = F.control_points_permute_closed_reg_loss(outp, _t(g["tl_gtcp"]), 20, 20)
    for k, v in [("tl_cd", cd1), ("tl_reg", lreg), ("tl_lap", lap), ("tl_closed", lclosed)]:
        assert abs(v.item() - float(g[k])) <= 2e-6 * abs(float(g[k])), (k, v.item(), float(g[k]))
    (0.9 * lreg + 0.1 * (cd1 + lap) + 0.5 * lclosed).backward()
    _rel(outp.grad, g["tl_gout"], 1e-5, "spline loss grads")
    wts = _t(g["wn_in"], True)
    wn = F.weights_normalize(wts, 0.8)
    _rel(wn, g["wn_out"], 1e-6, "weights_normalize")
    (wn * _t(g["wn_w"])).sum().backward()
    _rel(wts.grad, g["wn_g"], 1e-5, "weights_normalize grad")


def test_fitting_port_known_answers_of_reference_test_sketches():
    """the analytic refits sketched in the reference's src/test_fitting_utils.py:5-57 as real assertions: unit sphere at
    the origin, cylinder r = 1 about (1,2,0)/sqrt 5, cone with half-angle pi/3 about (1,1,0)/sqrt 2; rank-deficient
    lstsq takes the Tikhonov branch (fitting_utils.py:52-64)"""
    from oracle.port import fitting as F
    rng = np.random.RandomState(0)
    n = rng.randn(2000, 3); n /= np.linalg.norm(n, axis=1, keepdims=True)
    w = torch.ones(2000, 1)
    c, r = F.fit_sphere(_t(n.astype(np.float32)), w)
    assert c.abs().max() < 1e-3 and abs(float(r) - 1.0) < 1e-3
    ax = np.array([1.0, 2.0, 0.0]) / np.sqrt(5.0)
    e1 = np.cross(ax, [0, 0, 1.0]); e1 /= np.linalg.norm(e1); e2 = np.cross(ax, e1)
    th, h = rng.rand(2000) * 2 * np.pi, rng.rand(2000) * 2 - 1
    nrm = np.cos(th)[:, None] * e1 + np.sin(th)[:, None] * e2
    pts = nrm + h[:, None] * ax
    a, c, r = F.fit_cylinder(_t(pts.astype(np.float32)), _t(nrm.astype(np.float32)), w)
    assert abs(abs(float((a.reshape(3) * _t(ax.astype(np.float32))).sum())) - 1.0) < 1e-4 and abs(float(r) - 1.0) < 2e-3
    ang = np.pi / 3
    axc = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    f1 = np.cross(axc, [0, 0, 1.0]); f1 /= np.linalg.norm(f1); f2 = np.cross(axc, f1)
    hh = 0.2 + rng.rand(2000)
    radial = np.cos(th)[:, None] * f1 + np.sin(th)[:, None] * f2
    ptc = hh[:, None] * (axc + np.tan(ang) * radial)
    nc = np.cos(ang) * radial - np.sin(ang) * axc
    apex, a, theta = F.fit_cone(_t(ptc.astype(np.float32)), _t(nc.astype(np.float32)), w)
    assert apex.abs().max() < 1e-3 and abs(float(theta) - ang) < 1e-3
    assert abs(float((a.reshape(3) * _t(axc.astype(np.float32))).sum()) - 1.0) < 1e-4          # axis points into the cone
    # rank-2 system: solution of the regularised normal equations, finite and consistent in the row space
    A = torch.randn(50, 2, generator=torch.Generator().manual_seed(0)) @ torch.tensor([[1.0, 0.0, 1.0], [0.0, 1.0, 1.0]])
    x_true = torch.tensor([[0.3], [-0.2], [0.1]])
    x = F.lstsq(A, A @ x_true)
    assert torch.isfinite(x).all() and (A @ x - A @ x_true).abs().max() < 1e-2


# ------------------------------------------------------------------------------------------ SplineNet / end-to-end fit
def _spline_sd(shapes, seed):
    from oracle.port import common
    sd = common.seeded_state_dict(shapes, seed=seed)
    for i in (1, 2, 3, 4, 5):
        for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    return sd


@pytest.mark.parametrize("mode", [0, 1])
def test_splinenet_port_matches_reference_eval_mode(golden_dir, mode):
    """oracle/port/e2e.py::splinenet_fwd vs DGCNNControlPoints (src/model.py:56-180) in eval mode with per-point weights:
    control points and the gradient w.r.t. the weights"""
    from oracle.port import e2e
    g = _load(golden_dir, "splinenet.npz")
    shapes = {k: eval(s) for k, s in zip(g[f"m{mode}_keys"], g[f"m{mode}_shapes"])}
    sd = _spline_sd(shapes, 30 + mode)
    w = _t(g[f"m{mode}_w"], True)
    out = e2e.splinenet_fwd(sd, _t(g[f"m{mode}_x"]), 10, w.t())
    _rel(out, g[f"m{mode}_out"], 1e-5, "control points")
    (out * _t(g[f"m{mode}_c"])).sum().backward()
    _rel(w.grad, g[f"m{mode}_gw"], 1e-4, "d/dweights")


def test_splinenet_port_matches_reference_train_mode(golden_dir):
    from oracle.port import e2e
    g = _load(golden_dir, "splinenet.npz")
    shapes = {k: eval(s) for k, s in zip(g["m0_keys"], g["m0_shapes"])}
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v)
          for k, v in _spline_sd(shapes, 33).items()}
    out = e2e.splinenet_fwd(sd, _t(g["tr_x"]), 10, None, train=True)
    _rel(out, g["tr_out"], 1e-5, "train-mode control points")
    (out * _t(g["tr_c"])).sum().backward()
    for key in ("conv8.weight", "conv5.0.weight", "bn3.weight", "conv1.0.weight"):
        t = sd[key].grad.reshape(-1).double()
        got = np.array([t.sum().item(), t.norm().item()] + t[:14].tolist())
        want = g["trgrad:" + key]
        assert abs(got[1] - want[1]) <= 1e-3 * want[1] + 1e-9, (key, got[1], want[1])


@pytest.mark.parametrize("variant", ["e2e", "e2e_nocyl"])
def test_e2e_port_matches_reference_fitting_loss(golden_dir, variant):
    """the whole fit half of the path (mean-shift -> match -> per-segment fit incl. both SplineNets -> residuals ->
    loss -> gradient w.r.t. the embedding) restated in oracle/port vs Evaluation.fitting_loss of the unmodified
    reference on the same inputs"""
    from oracle.make_golden_helpers import e2e_inputs
    from oracle.port import e2e
    from src.model import DGCNNControlPoints       # only for the parameter shapes (a torch.nn.Module definition, CPU)
    g = _load(golden_dir, variant + ".npz")
    N = int(g["N"])
    pts, nrm, lab, prim, emb, logp = e2e_inputs(N, 77, variant == "e2e_nocyl")
    nets = {}
    for name, mode, seed in (("open", 0, 41), ("closed", 1, 42)):
        shapes = {k: tuple(v.shape) for k, v in DGCNNControlPoints(20, num_points=10, mode=mode).state_dict().items()}
        nets[name] = _spline_sd(shapes, seed)
    E = emb[0].clone().requires_grad_()
    np.random.seed(5)
    loss, params, distance, cluster_ids = e2e.fitting_loss(E, torch.from_numpy(pts[0]), torch.from_numpy(nrm[0]), lab[0],
                                                           prim[0].copy(), nets, 0.015, 10, 0.1)
    np.testing.assert_array_equal(cluster_ids, g["cluster_ids"])
    kinds = [f"{k}:{v[0] if v is not None else 'none'}" for k, v in sorted(params.items())]
    assert kinds == list(g["kinds"])
    got = sorted((v[0], float(v[1])) for v in distance.values())
    want = sorted(zip(g["seg_kind"], g["seg_dist"]))
    for (k1, d1), (k2, d2) in zip(got, want):
        assert k1 == k2
        # the reference's cylinder fit amplifies fp32 noise 1e4-fold (rank-deficient regularised solve)
        assert abs(d1 - d2) <= (5e-2 if k1 == "cylinder" else 1e-4) * d2, (k1, d1, d2)
    tol = 3e-2 if variant == "e2e" else 1e-4
    assert abs(loss[0].item() - float(g["loss"])) <= tol * abs(float(g["loss"]))
    assert abs(loss[2] - float(g["spl"])) <= 1e-4 * float(g["spl"])
    loss[0].backward()
    if variant == "e2e_nocyl":
        _rel(E.grad, g["gradE"][0], 1e-3, "d loss / d embedding")


def test_sparse_row_backward_of_mean_shift_equals_dense_autograd():
    """the argument behind the product's sparse-row backward (PN_MS_SPARSE_BWD): when the loss depends on K rows of the
    last iterate only (`center = new_X[indices]`), autograd through ALL N x N kernel matrices gives exactly what the
    closed-form per-iteration backward restricted to those K rows gives -- and the gradient never leaves those rows"""
    from oracle.make_golden_helpers import clustered_embedding
    from oracle.port import meanshift as oms
    N, K, its, bw = 600, 9, 5, 0.35
    X0, _ = clustered_embedding(N, 128, 5, 3)
    rows = torch.tensor([3, 17, 99, 100, 101, 256, 400, 577, 599])
    w = torch.randn(K, 128, generator=torch.Generator().manual_seed(1))
    # dense autograd through every iterate (the iteration of oracle/port/meanshift.py::mean_shift_iters, unrolled so that
    # the intermediate iterates and their gradients can be inspected)
    X = X0.clone().requires_grad_()
    Y = X.clone()
    iterates, dens, norms = [Y], [], []
    for _ in range(its):
        Kmat = torch.exp(torch.clamp(-(2.0 - 2.0 * Y @ X.t()) / (bw ** 2) / 2, -75.0, 75.0))
        den = Kmat.sum(1, keepdim=True)
        u = Y + ((Kmat @ X) / den - Y)
        nrm = torch.norm(u, dim=1, p=2, keepdim=True)
        Y = u / nrm
        Y.retain_grad()
        iterates.append(Y); dens.append(den[:, 0].detach()); norms.append(nrm[:, 0].detach())
    (iterates[-1][rows] * w).sum().backward()
    mask = torch.ones(N, dtype=torch.bool); mask[rows] = False
    for t in range(1, its + 1):
        assert iterates[t].grad[mask].abs().max() == 0.0          # the gradient stays confined to the selected rows
    # sparse: closed-form backward over the K rows only
    Xd = X0
    g = w.clone()
    gX = torch.zeros_like(Xd)
    for t in range(its, 0, -1):
        y_new, y_prev = iterates[t].detach()[rows], iterates[t - 1].detach()[rows]
        g, gx = oms.sparse_rows_backward(g, y_new, y_prev, dens[t - 1][rows], norms[t - 1][rows], Xd, bw)
        gX += gx
        want = iterates[t - 1].grad[rows] if t > 1 else None
        if want is not None:
            assert ((g - want).abs().max() / want.abs().max()).item() < 1e-4, t
    gX[rows] += g                                               # Y_0 = X
    assert ((gX - X.grad).abs().max() / X.grad.abs().max()).item() < 1e-4


def test_meanshift_port_retry_loops_match_reference(golden_dir):
    """more than 49 clusters -> the quantile grows (x1.2 with 10000 samples in Evaluation.guard_mean_shift,
    residual_utils.py:69-84; x2 with 5000 samples in MeanShift.guard_mean_shift, mean_shift.py:81-96): same number of
    attempts is implied by the same final bandwidth; labels and centres identical"""
    from oracle.make_golden_helpers import clustered_embedding
    from oracle.port import meanshift as oms
    g = _load(golden_dir, "guard.npz")
    X, _ = clustered_embedding(1500, 128, 60, 11, spread=0.05)
    assert abs(float(X.double().sum()) - float(g["x_checksum"])) < 1e-9
    for prefix, kw in (("ev", dict(num_samples=10000, growth=1.2)), ("ms", dict(num_samples=5000, growth=2))):
        np.random.seed(int(g["seed"]))
        c, bw, lab = oms.guard_mean_shift(X, float(g["quantile"]), int(g["iterations"]), **kw)
        assert abs(float(bw) - float(g[prefix + "_bw"])) <= 1e-6 * float(g[prefix + "_bw"])
        np.testing.assert_array_equal(lab.numpy(), g[prefix + "_labels"])
        _rel(c, g[prefix + "_center"], 1e-5, prefix + " centres")
        assert int(g[prefix + "_attempts"]) >= 2          # the fixture does exercise the retry


def test_upsample_port_and_c_oracle_vs_reference(golden_dir):
    """the port of up_sample_points_torch / _memory_efficient / _in_range reproduces the unmodified reference
    exactly (tests/golden/upsample.npz), and the C oracle's squared-difference metric returns the reference's 5-nearest
    indices (torch.topk of sum((p_i - p_j)**2, 2), largest=False)"""
    import numpy as np
    import torch
    from oracle import knn as oknn
    from oracle.port import fitting as OP
    g = np.load(os.path.join(golden_dir, "upsample.npz"))
    for name in ("a", "b"):
        p = torch.from_numpy(g[name + "_p"])
        np.testing.assert_array_equal(OP.up_sample_points(p.clone(), 2).numpy(), g[name + "_up2"])
        np.testing.assert_array_equal(OP.up_sample_points_memory_efficient(p.clone(), 1).numpy(), g[name + "_me1"])
        d = ((p.unsqueeze(1) - p.unsqueeze(0)) ** 2).sum(2)
        np.testing.assert_array_equal(oknn.knn(p.numpy()[None], 5, 2)[0], torch.topk(d, 5, 1, largest=False)[1].numpy())
    np.random.seed(3)
    mp, mw = OP.up_sample_points_in_range(torch.from_numpy(g["r_p"]), 1400, 1800, weights=torch.from_numpy(g["r_w"]))
    np.testing.assert_array_equal(mp.numpy(), g["r_out_p"])
    np.testing.assert_array_equal(mw.numpy(), g["r_out_w"])
