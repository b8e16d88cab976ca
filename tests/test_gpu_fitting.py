"""Fits, residual distances, Chamfer / spline losses and SplineNet on the GPU vs golden vectors of the reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _close(got, want, rtol=1e-4, name=""):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale = np.abs(want).max() + 1e-30
    err = np.abs(got.reshape(want.shape) - want).max()
    assert err <= rtol * scale + 1e-9, f"{name}: err {err:.3e} scale {scale:.3e}"


def _cu(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).cuda()
    return t.requires_grad_() if grad else t


# ------------------------------------------------------------------------------------------------ primitive fits
@pytest.mark.parametrize("kind", ["plane", "sphere", "cylinder", "cone"])
def test_fit_and_residual_vs_reference(golden_dir, kind):
    from src.primitive_forward import Fit
    from src.primitives import ComputePrimitiveDistance
    g = _g(golden_dir, "fits.npz")
    P, Nn, W = _cu(g[kind + "_p"]), _cu(g[kind + "_n"]), _cu(g[kind + "_w"], True)
    res = getattr(Fit(), f"fit_{kind}_torch")(P, Nn, W)
    nout = len(res)
    # eigenvector sign: the plane / cylinder axis is defined up to sign; align with the reference before comparing
    sign = 1.0
    axis_slot = {"plane": 0, "cylinder": 0}.get(kind)
    if axis_slot is not None:
        want_axis = g[f"{kind}_out{axis_slot}"].reshape(-1)
        sign = float(np.sign((res[axis_slot].detach().cpu().numpy().reshape(-1) * want_axis).sum()))
    loss = 0
    for i in range(nout):
        want = g[f"{kind}_out{i}"]
        s = sign if (i == axis_slot or (kind == "plane" and i == 1)) else 1.0
        if kind == "cylinder" and i == 1:
            # the centre of a cylinder is only defined up to a shift along the axis (the projected system is rank 2
            # and the reference's regularised solve amplifies fp32 noise along it): compare the component _|_ axis
            ax = torch.from_numpy(g["cylinder_out0"].reshape(3)).double()
            perp = lambda c: c - (c * ax).sum() * ax
            got_c = res[i].detach().cpu().double().reshape(3); want_c = torch.from_numpy(want).double().reshape(3)
            _close(perp(got_c), perp(want_c).numpy(), rtol=3e-2, name="cylinder centre (perpendicular part)")
            # the reference radius contains its own (fp32-noise) along-axis centre offset: r_ref^2 = r^2 + offset^2
            off = float(((want_c - got_c) * ax).sum())
            r_got, r_ref = float(res[2].detach()), float(g["cylinder_out2"])
            assert abs(np.sqrt(r_got ** 2 + off ** 2) - r_ref) < 5e-3 * r_ref
            continue
        if kind == "cylinder" and i == 2:
            continue
        _close(res[i] * s, want, rtol=2e-4, name=f"{kind} out{i}")
        loss = loss + (res[i] * s * _cu(g[f"{kind}_coef{i}"]).reshape(res[i].shape)).sum()
    loss.backward()
    if kind != "cylinder":      # (cylinder: the golden gradient includes the arbitrary along-axis centre component)
        _close(W.grad, g[kind + "_gw"], rtol=2e-3, name=f"{kind} grad wrt weights")
    # residual distance + parameter gradients
    n_par = {"plane": 2, "sphere": 2, "cylinder": 3, "cone": 3}[kind]
    params = [_cu(g[f"{kind}_par{i}"], True) for i in range(n_par)]
    d = getattr(ComputePrimitiveDistance(reduce=True), "distance_from_" + kind)(points=_cu(g[kind + "_q"]),
                                                                                  params=params, sqrt=False)
    _close(d, g[kind + "_dist"], name=f"{kind} residual")
    d.backward()
    for i in range(n_par):
        _close(params[i].grad, g[f"{kind}_gpar{i}"], rtol=1e-3, name=f"{kind} dpar{i}")


def test_batched_moment_fits_match_single_calls():
    """the batched path of fit_one_shape_torch (one moment pass for all segments) equals per-segment Fit calls"""
    from pnb200 import fitting as F
    from src.primitive_forward import Fit
    g = torch.Generator().manual_seed(0)
    N, K = 2000, 5
    P = (torch.randn(N, 3, generator=g) * 0.4).cuda()
    Nr = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1).cuda()
    W = torch.rand(N, K, generator=g).cuda()
    mom = F.MomentsFn.apply(W, P, Nr, 0, 4, (N + 3) // 4, F.EPS)
    a, d = F.fit_planes(mom)
    c, r = F.fit_spheres(mom, (N + 3) // 4)
    fit = Fit()
    for k in range(K):
        a1, d1 = fit.fit_plane_torch(P[0::4], None, W[0::4, k:k + 1] + F.EPS)
        s = float(torch.sign((a1.reshape(-1) * a[k].float()).sum()))
        _close(a[k] * s, a1.cpu().numpy(), name="plane a"); _close(d[k] * s, d1.cpu().numpy(), name="plane d")
        c1, r1 = fit.fit_sphere_torch(P[0::4], None, W[0::4, k:k + 1] + F.EPS)
        _close(c[k], c1.cpu().numpy(), rtol=1e-3, name="sphere c"); _close(r[k], r1.cpu().numpy(), rtol=1e-3, name="r")


# ------------------------------------------------------------------------------------------------ chamfer / spline losses
def test_chamfer_and_spline_losses_vs_reference(golden_dir):
    from src import loss as L, utils as U
    from src.fitting_utils import sample_points_from_control_points_, weights_normalize
    g = _g(golden_dir, "losses.npz")
    pred, gt = _cu(g["pred"], True), _cu(g["gt"], True)
    cd = U.chamfer_distance(pred, gt); cds = U.chamfer_distance(pred, gt, sqrt=True)
    c0 = U.chamfer_distance_one_side(pred, gt, 0); c1 = U.chamfer_distance_one_side(pred, gt, 1)
    s1 = U.chamfer_distance_single_shape(pred[0], gt[0]); s2 = U.chamfer_distance_single_shape(pred[0], gt[0], one_side=True)
    for name, v in [("cd", cd), ("cds", cds), ("c0", c0), ("c1", c1), ("s1", s1), ("s2", s2)]:
        assert abs(v.item() - float(g[name])) <= 1e-4 * abs(float(g[name])), name
    (cd + 2 * cds + 3 * c0 + 4 * c1 + 5 * s1 + 6 * s2).backward()
    _close(pred.grad, g["gpred"], rtol=1e-4, name="chamfer dpred"); _close(gt.grad, g["ggt"], rtol=1e-4, name="dgt")
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, 30)
    np.testing.assert_allclose(nu, g["nu"], rtol=0, atol=1e-15); np.testing.assert_allclose(nv, g["nv"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(L.uniform_knot_bspline(20, 20, 3, 3, 40)[0], g["nu40"], rtol=0, atol=1e-15)
    cpts = _cu(g["cpts"], True)
    rec = sample_points_from_control_points_(torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32)), cpts, 2)
    _close(rec, g["rec"], name="spline eval")
    (rec * _cu(g["recw"])).sum().backward()
    _close(cpts.grad, g["gcpts"], name="spline eval grad")

    class Cfg: batch_size = 2; grid_size = 20
    nu4 = torch.from_numpy(g["nu40"].astype(np.float32))
    outp = _cu(g["tl_out"], True)
    cd1, _ = L.spline_reconstruction_loss_one_sided(nu4, nu4, outp, _cu(g["tl_pts"]), Cfg)
    lreg, perm = L.control_points_permute_reg_loss(outp, _cu(g["tl_gtcp"]), 20)
    lap = L.laplacian_loss(outp.reshape(2, 20, 20, 3), perm)
    lclosed, _ = L.control_points_permute_closed_reg_loss(outp, _cu(g["tl_gtcp"]), 20, 20)
    for name, v in [("tl_cd", cd1), ("tl_reg", lreg), ("tl_lap", lap), ("tl_closed", lclosed)]:
        assert abs(v.item() - float(g[name])) <= 1e-4 * abs(float(g[name])), name
    (0.9 * lreg + 0.1 * (cd1 + lap) + 0.5 * lclosed).backward()
    _close(outp.grad, g["tl_gout"], rtol=1e-4, name="open-spline loss grads")
    wts = _cu(g["wn_in"], True)
    wn = weights_normalize(wts, 0.8)
    _close(wn, g["wn_out"], name="weights_normalize")
    (wn * _cu(g["wn_w"])).sum().backward()
    _close(wts.grad, g["wn_g"], rtol=1e-3, name="weights_normalize grad")


# ------------------------------------------------------------------------------------------------ SplineNet
def _spline_net(g, prefix, mode, seed):
    from oracle.port import common
    from src.model import DGCNNControlPoints
    net = DGCNNControlPoints(20, num_points=10, mode=mode)
    sd0 = net.state_dict()
    assert sorted(sd0.keys()) == list(g[prefix + "_keys"])
    shapes = {k: eval(s) for k, s in zip(g[prefix + "_keys"], g[prefix + "_shapes"])}
    for k in shapes:
        assert tuple(sd0[k].shape) == shapes[k], k
    sd = common.seeded_state_dict(shapes, seed=seed)
    for i in (1, 2, 3, 4, 5):
        for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    net.load_state_dict(sd)
    return net.cuda()


@pytest.mark.parametrize("mode", [0, 1])
def test_splinenet_eval_with_weights_vs_reference(golden_dir, mode):
    g = _g(golden_dir, "splinenet.npz")
    net = _spline_net(g, f"m{mode}", mode, 30 + mode).eval()
    w = _cu(g[f"m{mode}_w"], True)
    o = net(_cu(g[f"m{mode}_x"]), w.t())
    _close(o, g[f"m{mode}_out"], rtol=1e-5, name="control points")    # measured 0.9e-6 / 1.5e-6 (profiles/r02_parity_bounds.md)
    (o * _cu(g[f"m{mode}_c"])).sum().backward()
    _close(w.grad, g[f"m{mode}_gw"], rtol=2e-3, name="grad wrt weights")


def test_splinenet_train_mode_vs_reference(golden_dir):
    g = _g(golden_dir, "splinenet.npz")
    net = _spline_net(g, "m0", 0, 33).train()
    o = net(_cu(g["tr_x"]))
    _close(o, g["tr_out"], rtol=3e-4, name="train-mode output")
    (o * _cu(g["tr_c"])).sum().backward()
    _close(net.bn5.running_mean, g["tr_rm5"], rtol=1e-3, name="running mean 5")
    _close(net.bn5.running_var, g["tr_rv5"], rtol=1e-3, name="running var 5")
    _close(net.bn1.running_var, g["tr_rv1"], rtol=1e-3, name="running var 1")
    checked = 0
    bad = []
    for key in g.files:
        if not key.startswith("trgrad:") or (".1." in key and key.startswith("trgrad:conv")):
            continue
        p = dict(net.named_parameters())[key[7:]]
        t = p.grad.detach().cpu().reshape(-1).double()
        got = np.array([t.sum().item(), t.norm().item()] + t[:14].tolist())
        want = g[key]
        if key[7:] in ("conv6.bias", "conv7.bias"):
            continue        # a bias in front of BatchNorm has an exactly-zero gradient; both sides hold rounding noise
        if abs(got[1] - want[1]) > 5e-3 * want[1] + 1e-7:
            bad.append((key, got[1], want[1]))
        checked += 1
    assert not bad, bad
    assert checked >= 20


# ------------------------------------------------------------------------------------------------ end-to-end fitting loss
def _seeded_splinenet(mode, seed):
    from oracle.port import common
    from src.model import DGCNNControlPoints
    net = DGCNNControlPoints(20, num_points=10, mode=mode)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = common.seeded_state_dict(shapes, seed=seed)
    for i in (1, 2, 3, 4, 5):
        for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    net.load_state_dict(sd)
    return net.cuda().eval()


@pytest.mark.parametrize("stage", ["batched", "loop"])
@pytest.mark.parametrize("variant", ["e2e", "e2e_nocyl"])
def test_evaluation_fitting_loss_vs_reference(golden_dir, variant, stage, monkeypatch):
    """Evaluation.fitting_loss (mean-shift -> match -> fit -> residual) on one synthetic shape with all six segment
    kinds, against the reference run on the same inputs: loss within 1e-4 relative, same partition, same kinds.
    stage: the batched fit stage (default, pnb200/fitstage.py) and the per-shape loop of round 1."""
    import src.residual_utils as RU
    from oracle.make_golden_helpers import e2e_inputs
    from src.residual_utils import Evaluation
    monkeypatch.setattr(RU, "FIT_STAGE", stage)
    g = _g(golden_dir, variant + ".npz")
    N = int(g["N"])
    pts, nrm, lab, prim, emb, logp = e2e_inputs(N, 77, variant == "e2e_nocyl")
    ev = Evaluation(open_decoder=_seeded_splinenet(0, 41), closed_decoder=_seeded_splinenet(1, 42))
    E = emb.cuda().requires_grad_()
    np.random.seed(5)
    captured = {}
    orig_sep = ev.separate_losses

    def sep(distance, gt_points, lamb=1.0, **kw):
        captured.update({k: (v[0], float(v[1])) for k, v in distance.items()})
        return orig_sep(distance, gt_points, lamb=lamb, **kw)

    ev.separate_losses = sep
    res, extra = ev.fitting_loss(E, torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda(), lab, prim.copy(),
                                 logp.cuda(), quantile=0.015, iterations=10, lamb=0.1)
    params, cluster_ids, weights = extra
    if stage == "batched":
        from pnb200 import fitstage
        assert not captured, "the batched stage does not go through separate_losses"
        captured.update({k: (v[0], float(v[1])) for k, v in fitstage.segment_distances(ev.last_fit, 0).items()})
    # partition identical (cluster numbering is representative-point dependent, see test_gpu_meanshift)
    def canon(l):
        _, first = np.unique(l, return_index=True)
        order = l[np.sort(first)]
        m = {int(v): i for i, v in enumerate(order)}
        return np.array([m[int(v)] for v in l])
    np.testing.assert_array_equal(canon(cluster_ids), canon(g["cluster_ids"]))
    kinds = sorted(v[0] for v in params.values() if v is not None)
    assert kinds == sorted(k.split(":")[1] for k in g["kinds"] if not k.endswith("none"))
    assert abs(res[3] - float(g["s_iou"])) < 1e-6
    print("loss", res[0].item(), "ref", float(g["loss"]), "geo", res[1], float(g["geo"]), "spline", res[2], float(g["spl"]))
    # per-segment residuals by kind (cluster numbering differs, kinds are unique in this shape)
    mine = sorted(captured.values())
    ref = sorted(zip(g["seg_kind"], g["seg_dist"]))
    assert [k for k, _ in mine] == [k for k, _ in ref]
    for (kind, d_mine), (_, d_ref) in zip(mine, ref):
        # cylinder: the reference's own radius carries fp32 noise from its rank-deficient solve (measured and explained in
        # tests/test_cpu_fitsolve.py::test_cylinder_deviation_is_the_references_own_fp32_noise); everything else 1e-4
        tol = 5e-2 if kind == "cylinder" else 1e-4
        assert abs(d_mine - d_ref) <= tol * d_ref, (kind, d_mine, d_ref)
    assert abs(res[2] - float(g["spl"])) <= 1e-5 * float(g["spl"])
    assert abs(res[0].item() - float(g["loss"])) <= 3e-2 * abs(float(g["loss"]))
    res[0].backward()
    ge, gr = E.grad.cpu().double().numpy(), g["gradE"].astype(np.float64)
    rel = np.abs(ge - gr).max() / (np.abs(gr).max() + 1e-30)
    print("grad rel err", rel)
    if variant == "e2e_nocyl":
        # measured on B200 (round 2): loss 1.1e-6, gradient 4e-6 relative
        assert abs(res[0].item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
        assert rel < 1e-4
    # with a cylinder segment the reference gradient carries the 1e4-amplified fp32 noise of its rank-deficient
    # regularised solve (primitive_forward.py:803 -> fitting_utils.py:52-64); only the loss value is compared there


# ------------------------------------------------------------------------------------------------ BASELINE config 1
def test_cfg1_open_spline_control_point_solve_vs_reference(golden_dir):
    """open-spline fit only: 30x30 samples of a random 20x20 control grid -> control points.  Golden = the unmodified
    reference's approximation.fit_bezier_surface / fit_bezier_surface_fit_kronecker (float64, numpy).  The solve runs
    in float64 on the device like the reference: 1e-9 relative.  With fp32 SAMPLES the rounding of the input alone is
    amplified by ||Nu^+||_1 ||Nv^+||_1 ~ 8e4 to ~3e-4 (measured in float64 numpy on the rounded samples), so that case
    is checked against exact arithmetic on the same rounded samples (1e-6) and at 1e-3 against the golden."""
    from pnb200.fitting import fit_control_points_grid, spline_eval
    from src import approximation as AP
    g = _g(golden_dir, "cfg1.npz")
    nu, nv, cp, S = g["nu"], g["nv"], g["cp"], g["S"]
    pu = np.linalg.inv(nu.T @ nu) @ nu.T
    pv = np.linalg.inv(nv.T @ nv) @ nv.T
    Sd = torch.from_numpy(S.reshape(2, 30, 30, 3)).cuda().requires_grad_()                        # float64
    rec = fit_control_points_grid(Sd, nu, nv)
    assert rec.dtype == torch.float64
    _close(rec, g["rec"], rtol=1e-9, name="cfg1 control points vs reference")
    _close(rec, cp, rtol=1e-9, name="cfg1 control-point residual")
    # fp32 samples
    S32 = S.reshape(2, 30, 30, 3).astype(np.float32)
    rec32 = fit_control_points_grid(torch.from_numpy(S32).cuda(), nu, nv)
    assert rec32.dtype == torch.float32
    _close(rec32, np.einsum("iu,buvc,jv->bijc", pu, S32.astype(np.float64), pv), rtol=1e-6, name="cfg1 fp32 samples, exact arithmetic")
    _close(rec32, g["rec"], rtol=1e-3, name="cfg1 fp32 samples vs reference")
    # numpy-in / numpy-out drop-in and the noisy (genuinely least-squares) case
    rec_np = AP.fit_bezier_surface(g["Sn"][0].reshape(30, 30, 3), nu, nv)
    assert isinstance(rec_np, np.ndarray) and rec_np.shape == (20, 20, 3) and rec_np.dtype == np.float64
    _close(torch.from_numpy(rec_np), g["rec_n"][0], rtol=1e-9, name="cfg1 noisy samples vs reference")
    # scattered-sample (Kronecker) solve
    A_u, A_v = np.repeat(nu, 30, axis=0), np.tile(nv, (30, 1))
    rec_k = AP.fit_bezier_surface_fit_kronecker(S[0], A_u, A_v)
    _close(torch.from_numpy(rec_k), g["rec_k"], rtol=1e-6, name="cfg1 kronecker solve vs reference")
    # round trip through the evaluation kernel and gradient w.r.t. the samples (linear map: grad = pinv^T (.) pinv)
    back = spline_eval(rec, torch.from_numpy(nu).cuda(), torch.from_numpy(nv).cuda())
    assert back.dtype == torch.float64
    _close(back, S, rtol=1e-9, name="cfg1 re-evaluated surface")
    w = torch.randn(rec.shape, generator=torch.Generator().manual_seed(0), dtype=torch.float64).cuda()
    (rec * w).sum().backward()
    gref = np.einsum("iu,bijc,jv->buvc", pu, w.cpu().numpy(), pv)
    _close(Sd.grad, gref.reshape(2, 30, 30, 3), rtol=1e-9, name="cfg1 gradient w.r.t. samples")


# ---------------------------------------------------------------------------------------------- SURVEY 8f-1
def test_kronecker_fit_kernel_vs_reference(golden_dir):
    """csrc/kronfit.cu (normal equations + in-kernel Cholesky, one CTA per surface) against the unmodified reference's
    fit_bezier_surface_fit_kronecker (numpy lstsq) on the optimisers' problem size: 1600 scattered samples, 10 x 10 control
    points of degree 2 and 3 (tests/golden/kronecker.npz); batched launch; scattered evaluation kernel; rank-deficient flag"""
    from pnb200 import fitting as F
    from src import approximation as AP
    g = _g(golden_dir, "kronecker.npz")
    P = torch.from_numpy(np.stack([g["pts2"], g["pts3"]])).cuda()
    U = torch.from_numpy(np.stack([g["NU2"], g["NU3"]])).cuda()
    V = torch.from_numpy(np.stack([g["NV2"], g["NV3"]])).cuda()
    ctrl, flag = F.kron_fit(P, U, V)
    assert not flag.any()
    _close(ctrl[0], g["rec2"], rtol=1e-9, name="kronecker fit degree 2 vs reference")
    _close(ctrl[1], g["rec3"], rtol=1e-9, name="kronecker fit degree 3 vs reference")
    rec = AP.fit_bezier_surface_fit_kronecker(g["pts3"], g["NU3"], g["NV3"])          # numpy in / numpy out drop-in
    assert isinstance(rec, np.ndarray) and rec.dtype == np.float64
    _close(torch.from_numpy(rec), g["rec3"], rtol=1e-9, name="drop-in kronecker fit")
    ev = F.kron_eval(ctrl, U, V)
    want = np.stack([np.einsum("ia,ib,abc->ic", g[f"NU{d}"], g[f"NV{d}"], g[f"rec{d}"]) for d in (2, 3)])
    _close(ev, want, rtol=1e-12, name="scattered surface evaluation")
    # rank-deficient sampling (all samples in one knot span): flagged by the kernel, minimum-norm solution from the drop-in
    par = np.random.RandomState(0).random_sample((300, 2)) * 0.1
    nu, nv = AP.basis_rows(par, 10, 10, 2, 2)
    pts = np.random.RandomState(1).rand(300, 3)
    _, fl = F.kron_fit(torch.from_numpy(pts).cuda()[None], torch.from_numpy(nu).cuda()[None], torch.from_numpy(nv).cuda()[None])
    assert int(fl[0]) == 1
    from oracle.port import optimize as PO
    _close(torch.from_numpy(AP.fit_bezier_surface_fit_kronecker(pts, nu, nv)), PO.fit_bezier_surface_fit_kronecker(pts, nu, nv),
           rtol=1e-6, name="rank-deficient kronecker fit (minimum norm)")


@pytest.mark.parametrize("closed", [False, True])
def test_kronecker_optimisers_vs_port(closed):
    """optimize_open_spline_kronecker / optimize_close_spline_kronecker (deform=False) against the numpy restatement of the
    reference (oracle/port/optimize.py; PARITY UNPINNED for the geomdl evaluation and the lapsolver assignment, see its header):
    same np.random stream, same up-sampling, same optimal assignment -> re-fitted surface samples to 1e-4 of their scale"""
    from oracle.port import optimize as PO
    from src import primitive_forward as PF
    from tools.synth import open_spline_batch
    rs = np.random.RandomState(3 + closed)
    cu = 21 if closed else 20
    # a smooth predicted surface and a noisy, denser input cloud around it
    base = rs.rand(4, 4, 3) * 0.5
    grid = np.stack(np.meshgrid(np.linspace(0, 1, cu), np.linspace(0, 1, 20), indexing="ij"), -1)
    cp = np.concatenate([grid, 0.3 * np.sin(3 * grid[..., :1] + 2 * grid[..., 1:2])], -1) + 0.01 * rs.randn(cu, 20, 3)
    par = rs.random_sample((1300, 2))
    inp = (PO.evaluate_list(cp, par, 3, 3) + 0.003 * rs.randn(1300, 3)).astype(np.float32)
    cp32 = cp.astype(np.float32)
    np.random.seed(7)
    if closed:
        want, _ = PO.optimize_close_spline_kronecker(inp, cp32.astype(np.float64))
    else:
        want, _ = PO.optimize_open_spline_kronecker(inp, cp32.reshape(400, 3).astype(np.float64))
    np.random.seed(7)
    inp_d = torch.from_numpy(inp).cuda().unsqueeze(0)
    if closed:
        got = PF.optimize_close_spline_kronecker(None, inp_d, torch.from_numpy(cp32).cuda().unsqueeze(0), deform=False)
        assert got.shape == (1, 930, 3)
    else:
        got = PF.optimize_open_spline_kronecker(None, inp_d, torch.from_numpy(cp32.reshape(1, 400, 3)).cuda())
        assert got.shape == (1, 900, 3)
    _close(got[0], want, rtol=1e-4, name="re-fitted surface samples")
    # the re-fitted surface is closer to the input cloud than the prediction it started from is not guaranteed in general,
    # but it must stay a sensible surface: finite and within the cloud's bounding box (+ margin)
    assert torch.isfinite(got).all()
    with pytest.raises(NotImplementedError):
        PF.optimize_open_spline_kronecker(None, inp_d, torch.from_numpy(cp32.reshape(1, -1, 3)).cuda(), deform=True)
    # device-only matcher (nearest input point per surface sample): runs without the host assignment, lands near the same surface
    np.random.seed(7)
    if not closed:
        near = PF.optimize_open_spline_kronecker(None, inp_d, torch.from_numpy(cp32.reshape(1, 400, 3)).cuda(), matcher="nearest")
        assert (near[0].cpu() - torch.from_numpy(want)).abs().max().item() < 0.25        # (another matching: another, nearby, surface)
