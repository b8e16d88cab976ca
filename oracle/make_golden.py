"""TEST INFRASTRUCTURE ONLY — dumps golden vectors from the UNMODIFIED reference (/root/reference, CPU, torch 2.11)
into tests/golden/*.npz.  Run in the build container:  python -m oracle.make_golden [names...]

The fixtures pin the CPU restatement in oracle/port (tests/test_oracle_golden.py); the GPU parity tests then
compare the CUDA path with the port on fresh seeded inputs and with these fixtures directly.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as rl  # noqa: E402
from oracle.port import common  # noqa: E402
from oracle.make_golden_helpers import prim_cloud  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def grad_summary(t):
    t = t.detach().reshape(-1).double()
    return np.array([t.sum().item(), t.norm().item()] + t[:14].tolist(), dtype=np.float64)


def fix_aliases(sd):
    for i in (1, 2, 3):
        for s in ("weight", "bias"):
            a, b = f"encoder.bn{i}.{s}", f"encoder.conv{i}.1.{s}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    return sd


def gen_knn():
    PN = rl.ref("src.PointNet")
    out = {}
    for name, (N, C, k, metric, seed) in {"c3": (600, 3, 10, 0, 1), "c64": (700, 64, 80, 0, 2), "pn": (650, 6, 80, 1, 3)}.items():
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, C, N, generator=g) * 0.3
        if metric == 1:
            x[:, 3:] = torch.nn.functional.normalize(x[:, 3:], dim=1)
        idx = (PN.knn(x, k, k) if metric == 0 else PN.knn_points_normals(x, k, k)).numpy()
        out[name + "_x"] = x.numpy(); out[name + "_idx"] = idx.astype(np.int32)
        out[name + "_meta"] = np.array([N, C, k, metric])
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **out)


def gen_segnet():
    PN = rl.ref("src.PointNet"); SL = rl.ref("src.segment_loss")
    B, N, k = 2, 320, 20
    pts, nrm, lab, prim = common.synth_cloud(B, N, seed=5)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous()
    loss = SL.EmbeddingLoss(margin=1.0)
    m = PN.PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10,
                                     loss_function=loss.triplet_loss, mode=5, num_channels=6, nn_nb=k)
    shapes = {n: tuple(v.shape) for n, v in m.state_dict().items()}
    sd = fix_aliases(common.seeded_state_dict(shapes, seed=11))
    m.load_state_dict(sd)
    rec = []
    ok, okn = PN.knn, PN.knn_points_normals
    PN.knn = lambda *a, **kw: rec.append(ok(*a, **kw)) or rec[-1]
    PN.knn_points_normals = lambda *a, **kw: rec.append(okn(*a, **kw)) or rec[-1]
    np.random.seed(7)
    emb, lp, el = m(x, torch.from_numpy(lab), True)
    PN.knn, PN.knn_points_normals = ok, okn
    nll = SL.primitive_loss(lp, torch.from_numpy(prim))
    total = el.sum() + nll
    total.backward()
    out = dict(points=x.numpy(), labels=lab, prims=prim, meta=np.array([B, N, k, 11, 7]),
               embedding=emb.detach().numpy(), logprob=lp.detach().numpy(), embed_loss=el.detach().numpy(),
               nll=nll.detach().numpy(), idx1=rec[0].numpy().astype(np.int32), idx2=rec[1].numpy().astype(np.int32),
               idx3=rec[2].numpy().astype(np.int32))
    out["state_keys"] = np.array(sorted(shapes.keys()))
    out["state_shapes"] = np.array([str(shapes[k_]) for k_ in sorted(shapes.keys())])
    for n, p in m.named_parameters():
        if p.grad is not None:
            out["grad:" + n] = grad_summary(p.grad)
    np.savez_compressed(os.path.join(OUT, "segnet.npz"), **out)


from oracle.make_golden_helpers import clustered_embedding  # noqa: E402


def gen_meanshift():
    MS = rl.ref("src.mean_shift")
    ms = MS.MeanShift()
    out = {}
    for name, (N, ncl, seed, q, it) in {"a": (700, 6, 3, 0.05, 10), "b": (520, 3, 4, 0.1, 5)}.items():
        X, _ = clustered_embedding(N, 128, ncl, seed)
        Xr = X.clone().requires_grad_()
        np.random.seed(seed)
        newX, center, bw, labels = ms.mean_shift(Xr, N, q, it)
        g = torch.Generator().manual_seed(seed + 100)
        w = torch.randn(center.shape, generator=g)
        w2 = torch.randn(newX.shape, generator=g) * 0.01
        ((center * w).sum() + (newX * w2).sum()).backward()
        out[name + "_X"] = X.numpy(); out[name + "_newX"] = newX.detach().numpy()
        out[name + "_center"] = center.detach().numpy(); out[name + "_bw"] = bw.numpy()
        out[name + "_labels"] = labels.numpy(); out[name + "_gradX"] = Xr.grad.numpy()
        out[name + "_meta"] = np.array([N, ncl, seed, it], dtype=np.int64); out[name + "_q"] = np.array(q)
    np.savez_compressed(os.path.join(OUT, "meanshift.npz"), **out)


def gen_fits():
    PF = rl.ref("src.primitive_forward"); PR = rl.ref("src.primitives")
    fit = PF.Fit(); cp = PR.ComputePrimitiveDistance(reduce=True)
    out = {}
    for kind, seed in [("plane", 1), ("sphere", 2), ("cylinder", 3), ("cone", 4)]:
        p, n, w = prim_cloud(kind, 900, seed)
        P, Nn = torch.from_numpy(p), torch.from_numpy(n)
        W = torch.from_numpy(w).requires_grad_()
        res = getattr(fit, f"fit_{kind}_torch")(P, Nn, W)
        g = torch.Generator().manual_seed(seed)
        coef = [torch.randn(r.shape, generator=g) for r in res]
        sum((r * c).sum() for r, c in zip(res, coef)).backward()
        out[kind + "_p"] = p; out[kind + "_n"] = n; out[kind + "_w"] = w
        for i, r in enumerate(res):
            out[f"{kind}_out{i}"] = r.detach().numpy(); out[f"{kind}_coef{i}"] = coef[i].numpy()
        out[kind + "_gw"] = W.grad.numpy()
        # residual distance of a second sample of the same surface to the fitted primitive (+ grads wrt params)
        q = torch.from_numpy(prim_cloud(kind, 500, seed + 50)[0])
        if kind == "plane":
            params = [res[0].detach().reshape(3, 1), res[1].detach()]
        elif kind == "sphere":
            params = [res[0].detach(), res[1].detach()]
        elif kind == "cylinder":
            params = [res[0].detach(), res[1].detach(), res[2].detach()]
        else:
            params = [res[0].detach().reshape(1, 3), res[1].detach().reshape(3, 1), res[2].detach()]
        params = [(t + 0.02).clone().requires_grad_() for t in params]
        d = getattr(cp, "distance_from_" + kind)(points=q, params=params, sqrt=False)
        d.backward()
        out[kind + "_q"] = q.numpy(); out[kind + "_dist"] = d.detach().numpy()
        for i, t in enumerate(params):
            out[f"{kind}_par{i}"] = t.detach().numpy(); out[f"{kind}_gpar{i}"] = t.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "fits.npz"), **out)


def gen_losses():
    U = rl.ref("src.utils"); L = rl.ref("src.loss"); FU = rl.ref("src.fitting_utils")
    g = torch.Generator().manual_seed(0)
    out = {}
    pred = torch.randn(3, 150, 3, generator=g).requires_grad_(); gt = torch.randn(3, 220, 3, generator=g).requires_grad_()
    cd = U.chamfer_distance(pred, gt); cds = U.chamfer_distance(pred, gt, sqrt=True)
    c0 = U.chamfer_distance_one_side(pred, gt, 0); c1 = U.chamfer_distance_one_side(pred, gt, 1)
    s1 = U.chamfer_distance_single_shape(pred[0], gt[0]); s2 = U.chamfer_distance_single_shape(pred[0], gt[0], one_side=True)
    (cd + 2 * cds + 3 * c0 + 4 * c1 + 5 * s1 + 6 * s2).backward()
    out.update(pred=pred.detach().numpy(), gt=gt.detach().numpy(), cd=cd.item(), cds=cds.item(), c0=c0.item(),
               c1=c1.item(), s1=s1.item(), s2=s2.item(), gpred=pred.grad.numpy(), ggt=gt.grad.numpy())
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, 30)
    nu40, nv40 = L.uniform_knot_bspline(20, 20, 3, 3, 40)
    out.update(nu=nu, nv=nv, nu40=nu40)
    cpts = (torch.rand(2, 400, 3, generator=g) - 0.5).requires_grad_()
    rec = FU.sample_points_from_control_points_(torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32)), cpts, 2)
    w = torch.randn(rec.shape, generator=g)
    (rec * w).sum().backward()
    out.update(cpts=cpts.detach().numpy(), rec=rec.detach().numpy(), recw=w.numpy(), gcpts=cpts.grad.numpy())
    # open-spline training losses (train_open_splines.py:158-178)
    class Cfg: batch_size = 2; grid_size = 20
    pts = torch.randn(2, 3, 300, generator=g) * 0.3
    outp = (torch.rand(2, 400, 3, generator=g) - 0.5).requires_grad_()
    gtcp = torch.rand(2, 20, 20, 3, generator=g) - 0.5
    nu4 = torch.from_numpy(nu40.astype(np.float32))
    cd1, _ = L.spline_reconstruction_loss_one_sided(nu4, nu4, outp, pts, Cfg)
    lreg, perm = L.control_points_permute_reg_loss(outp, gtcp, 20)
    lap = L.laplacian_loss(outp.reshape(2, 20, 20, 3), perm)
    lclosed, _ = L.control_points_permute_closed_reg_loss(outp, gtcp, 20, 20)
    (0.9 * lreg + 0.1 * (cd1 + lap) + 0.5 * lclosed).backward()
    out.update(tl_pts=pts.numpy(), tl_out=outp.detach().numpy(), tl_gtcp=gtcp.numpy(), tl_cd=cd1.item(),
               tl_reg=lreg.item(), tl_lap=lap.item(), tl_closed=lclosed.item(), tl_gout=outp.grad.numpy())
    # weights_normalize
    wts = torch.randn(7, 300, generator=g).requires_grad_()
    wn = FU.weights_normalize(wts, 0.8)
    ww = torch.randn(wn.shape, generator=g)
    (wn * ww).sum().backward()
    out.update(wn_in=wts.detach().numpy(), wn_out=wn.detach().numpy(), wn_w=ww.numpy(), wn_g=wts.grad.numpy())
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def spline_state(mode, seed):
    M = rl.ref("src.model")
    net = M.DGCNNControlPoints(20, num_points=10, mode=mode)
    shapes = {n: tuple(v.shape) for n, v in net.state_dict().items()}
    sd = common.seeded_state_dict(shapes, seed=seed)
    for i in (1, 2, 3, 4, 5):
        for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    net.load_state_dict(sd)
    return net, shapes


def gen_splinenet():
    out = {}
    for mode in (0, 1):
        net, shapes = spline_state(mode, 30 + mode)
        net.eval()
        g = torch.Generator().manual_seed(mode)
        x = torch.randn(1, 3, 500, generator=g) * 0.4
        w = torch.rand(500, 1, generator=g).requires_grad_()
        o = net(x, w.T)
        c = torch.randn(o.shape, generator=g)
        (o * c).sum().backward()
        out.update({f"m{mode}_x": x.numpy(), f"m{mode}_w": w.detach().numpy(), f"m{mode}_out": o.detach().numpy(),
                    f"m{mode}_c": c.numpy(), f"m{mode}_gw": w.grad.numpy()})
        out[f"m{mode}_keys"] = np.array(sorted(shapes.keys()))
        out[f"m{mode}_shapes"] = np.array([str(shapes[k_]) for k_ in sorted(shapes.keys())])
    # training mode (batch statistics), all parameter gradients (summaries)
    net, shapes = spline_state(0, 33)
    net.train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 420, generator=g) * 0.4
    o = net(x)
    c = torch.randn(o.shape, generator=g)
    (o * c).sum().backward()
    out.update(tr_x=x.numpy(), tr_out=o.detach().numpy(), tr_c=c.numpy())
    for n, p in net.named_parameters():
        out["trgrad:" + n] = grad_summary(p.grad)
    out["tr_rm5"] = net.bn5.running_mean.numpy(); out["tr_rv5"] = net.bn5.running_var.numpy()
    out["tr_rm1"] = net.bn1.running_mean.numpy(); out["tr_rv1"] = net.bn1.running_var.numpy()
    np.savez_compressed(os.path.join(OUT, "splinenet.npz"), **out)


from oracle.make_golden_helpers import e2e_inputs  # noqa: E402


def gen_e2e(no_cylinder=False, fname="e2e.npz"):
    RU = rl.ref("src.residual_utils"); FO = rl.ref("src.fitting_optimization"); PF = rl.ref("src.primitive_forward")
    PR = rl.ref("src.primitives"); MS = rl.ref("src.mean_shift"); L = rl.ref("src.loss")
    N = 1400
    pts, nrm, lab, prim, emb, logp = e2e_inputs(N, 77, no_cylinder)
    ev = RU.Evaluation.__new__(RU.Evaluation)
    ev.res_loss = PR.ResidualLoss()
    fm = FO.FittingModule.__new__(FO.FittingModule)
    fm.fitting = PF.Fit()
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, 30)
    fm.nu = torch.from_numpy(nu.astype(np.float32)); fm.nv = torch.from_numpy(nv.astype(np.float32))
    fm.open_control_decoder, _ = spline_state(0, 41); fm.closed_control_decoder, _ = spline_state(1, 42)
    fm.open_control_decoder.eval(); fm.closed_control_decoder.eval()
    ev.fitter = fm
    ev.ms = MS.MeanShift()
    E = emb.clone().requires_grad_()
    np.random.seed(5)
    captured = {}
    orig_sep = ev.separate_losses

    def sep(distance, gt_points, lamb=1.0):
        captured.update({k: (v[0], float(v[1])) for k, v in distance.items()})
        return orig_sep(distance, gt_points, lamb=lamb)

    ev.separate_losses = sep
    res, extra = ev.fitting_loss(E, torch.from_numpy(pts), torch.from_numpy(nrm), lab, prim.copy(), logp,
                                 quantile=0.015, iterations=10, lamb=0.1)
    res[0].backward()
    params, cluster_ids, weights = extra
    out = dict(N=np.array(N), loss=res[0].detach().numpy(), geo=np.array(res[1] if res[1] is not None else np.nan),
               spl=np.array(res[2] if res[2] is not None else np.nan), s_iou=np.array(res[3]), p_iou=np.array(res[4]),
               cluster_ids=cluster_ids, gradE=E.grad.numpy(), weights=weights.detach().numpy())
    kinds = []
    for k, v in sorted(params.items()):
        if v is None:
            kinds.append(f"{k}:none"); continue
        kinds.append(f"{k}:{v[0]}")
        for i, t in enumerate(v[1:]):
            out[f"par_{k}_{i}"] = t.detach().numpy()
    out["kinds"] = np.array(kinds)
    out["seg_kind"] = np.array([captured[k][0] for k in sorted(captured)])
    out["seg_dist"] = np.array([captured[k][1] for k in sorted(captured)])
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print("e2e loss", res[0].item(), "geo", res[1], "spline", res[2], "siou", res[3], "kinds", kinds)


def gen_cfg1():
    """BASELINE config 1: open-spline fit only.  Random 20x20x3 control grid (seed 0) -> 30x30 surface samples with the
    reference's sample_points_from_control_points_ -> control points recovered by the reference's own gridded solve
    approximation.fit_bezier_surface (float64) and by its Kronecker lstsq (fit_bezier_surface_fit_kronecker)."""
    FU = rl.ref("src.fitting_utils"); L = rl.ref("src.loss"); AP = rl.ref("src.approximation")
    rs = np.random.RandomState(0)
    cp = rs.rand(2, 20, 20, 3)
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, 30)
    S = FU.sample_points_from_control_points_(torch.from_numpy(nu), torch.from_numpy(nv),
                                              torch.from_numpy(cp.reshape(2, 400, 3)), 2).numpy()      # (2,900,3) f64
    rec = np.stack([AP.fit_bezier_surface(S[b].reshape(30, 30, 3), nu, nv) for b in range(2)], 0)
    A_u = np.repeat(nu, 30, axis=0); A_v = np.tile(nv, (30, 1))                                        # per-point basis rows
    rec_k = AP.fit_bezier_surface_fit_kronecker(S[0], A_u, A_v)
    # a noisy, non-interpolating case: least-squares control points of perturbed samples
    Sn = S + 0.01 * rs.randn(*S.shape)
    rec_n = np.stack([AP.fit_bezier_surface(Sn[b].reshape(30, 30, 3), nu, nv) for b in range(2)], 0)
    print("cfg1: max |cp - rec| gridded", np.abs(rec - cp).max(), "kronecker", np.abs(rec_k - cp[0]).max(),
          "cond(nu)", np.linalg.cond(nu))
    np.savez_compressed(os.path.join(OUT, "cfg1.npz"), cp=cp, S=S, rec=rec, rec_k=rec_k, Sn=Sn, rec_n=rec_n, nu=nu, nv=nv)


def gen_guard():
    """the retry loops around mean_shift (more than 49 clusters -> larger quantile): Evaluation.guard_mean_shift
    (residual_utils.py:69-84, x1.2, 10000 samples) and MeanShift.guard_mean_shift (mean_shift.py:81-96, x2, 5000 samples)
    on an embedding with 60 tight clusters, so that the first attempts do exceed 49"""
    from oracle.make_golden_helpers import clustered_embedding
    RU = rl.ref("src.residual_utils"); MS = rl.ref("src.mean_shift")
    X, _ = clustered_embedding(1500, 128, 60, 11, spread=0.05)
    # (the input is regenerated from its seed by the tests: clustered_embedding(1500, 128, 60, 11, spread=0.05))
    out = {"x_checksum": np.array(float(X.double().sum())), "quantile": np.array(0.002), "iterations": np.array(10),
           "seed": np.array(4)}
    ev = RU.Evaluation.__new__(RU.Evaluation); ev.ms = MS.MeanShift()
    calls = {"n": 0}
    orig = ev.ms.mean_shift

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    ev.ms.mean_shift = counting
    np.random.seed(4)
    c, bw, lab = ev.guard_mean_shift(X, 0.002, 10, kernel_type="gaussian")
    out.update(ev_center=c.numpy(), ev_bw=bw.numpy(), ev_labels=lab.numpy(), ev_attempts=np.array(calls["n"]))
    m = MS.MeanShift()
    calls["n"] = 0
    orig2 = m.mean_shift
    m.mean_shift = lambda *a, **k: (calls.__setitem__("n", calls["n"] + 1), orig2(*a, **k))[1]
    np.random.seed(4)
    c, bw, lab = m.guard_mean_shift(X, 0.002, 10, kernel_type="gaussian")
    out.update(ms_center=c.numpy(), ms_bw=bw.numpy(), ms_labels=lab.numpy(), ms_attempts=np.array(calls["n"]))
    print("guard: Evaluation attempts", int(out["ev_attempts"]), "clusters", len(np.unique(out["ev_labels"])),
          "| MeanShift attempts", int(out["ms_attempts"]), "clusters", len(np.unique(out["ms_labels"])))
    np.savez_compressed(os.path.join(OUT, "guard.npz"), **out)


def gen_upsample():
    """up_sample_points_torch (2 rounds), its memory-efficient variant (1 round) and up_sample_points_in_range
    (np.random.seed(3)) of the unmodified reference on uniform random points"""
    FU = rl.ref("src.fitting_utils")
    g = torch.Generator().manual_seed(21)
    out = {}
    for name, n in (("a", 333), ("b", 1000)):
        p = torch.rand(n, 3, generator=g)
        out[name + "_p"] = p.numpy()
        out[name + "_up2"] = FU.up_sample_points_torch(p.clone(), 2).numpy()
        out[name + "_me1"] = FU.up_sample_points_torch_memory_efficient(p.clone(), 1).numpy()
    p = torch.rand(700, 3, generator=g); w = torch.rand(700, 1, generator=g)
    np.random.seed(3)
    rp, rw = FU.up_sample_points_in_range(p.clone(), w.clone(), 1400, 1800)
    out.update(r_p=p.numpy(), r_w=w.numpy(), r_out_p=rp.numpy(), r_out_w=rw.numpy())
    np.savez_compressed(os.path.join(OUT, "upsample.npz"), **out)


def gen_pipeline():
    """Dataset.get_train of the unmodified reference (dataset_segments.py:95-157) on synthetic clouds, constructed
    without its h5 loader: align_canonical with and without normal noise / anisotropic scaling (np.random.seed(11))"""
    DS = rl.ref("src.dataset_segments")
    from tools.synth import ALL_KINDS, synth_cloud
    B, N = 4, 3000
    pts, nrm, lab, prim = synth_cloud(B, N, seed=77, n_patches=6, kinds=ALL_KINDS)
    pts = (pts * np.array([1.7, 0.9, 1.2], np.float32)).astype(np.float32)
    out = {"pts": pts.copy(), "nrm": nrm.copy()}
    for name, kw in (("plain", dict(if_normal_noise=False, anisotropic=False)),
                     ("noise_aniso", dict(if_normal_noise=True, anisotropic=True))):
        d = DS.Dataset.__new__(DS.Dataset)
        d.batch_size = B; d.normals = True; d.primitives = True
        d.train_points = pts.copy(); d.train_labels = lab.copy(); d.train_normals = nrm.copy(); d.train_primitives = prim.copy()
        d.augment_routines = []
        np.random.seed(11)
        p, l, n, pr = next(d.get_train(randomize=False, augment=False, align_canonical=True, **kw))
        out[name + "_p"] = np.asarray(p, np.float32); out[name + "_n"] = np.asarray(n, np.float32)
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), **out)


GENS = {"pipeline": gen_pipeline, "upsample": gen_upsample, "guard": gen_guard, "cfg1": gen_cfg1, "knn": gen_knn, "segnet": gen_segnet, "meanshift": gen_meanshift, "fits": gen_fits, "losses": gen_losses,
        "splinenet": gen_splinenet, "e2e": gen_e2e, "e2e_nocyl": lambda: gen_e2e(True, "e2e_nocyl.npz")}

def gen_kronecker():
    """the pieces of the Kronecker post-fit optimisers that run without geomdl / lapsolver / open3d, from the
    unmodified reference: DrawSurfs.boundary_parameterization / regular_parameterization, uniform_knot_bspline_ knots,
    BSpline.basis_functions per sample (the optimisers' NU / NV), and fit_bezier_surface_fit_kronecker on matched points of
    the optimisers' shape (1600 scattered samples, 10 x 10 control points, degree 2 and 3)."""
    AP = rl.ref("src.approximation"); CU = rl.ref("src.curve_utils")
    ds, bs = CU.DrawSurfs(), AP.BSpline()
    rs = np.random.RandomState(5)
    out = {"bpar20": ds.boundary_parameterization(20), "bpar30": ds.boundary_parameterization(30),
           "rpar": ds.regular_parameterization(30, 30)}
    for deg in (2, 3):
        bpar = out["bpar20"] if deg == 2 else out["bpar30"]
        par = np.concatenate([rs.random_sample((1600 - bpar.shape[0], 2)), bpar], 0)
        _, _, ku, kv = AP.uniform_knot_bspline_(10, 10, deg, deg, 2)
        NU, NV = [], []
        for i in range(par.shape[0]):
            nu, nv = bs.basis_functions(par[i], 10, 10, ku, kv, deg, deg)
            NU.append(nu); NV.append(nv)
        NU = np.concatenate(NU, 1).T; NV = np.concatenate(NV, 1).T
        cp = rs.rand(10, 10, 3)
        pts = np.einsum("ia,ib,abc->ic", NU, NV, cp) + 0.01 * rs.randn(1600, 3)
        rec = AP.fit_bezier_surface_fit_kronecker(pts, NU, NV)
        out.update({f"par{deg}": par, f"NU{deg}": NU, f"NV{deg}": NV, f"pts{deg}": pts, f"rec{deg}": rec,
                    f"ku{deg}": np.array(ku), f"kv{deg}": np.array(kv)})
    # basis rows of the predicted 20 x 20 / 21 x 20 cubic surfaces at a few parameters (the optimisers' first evaluation)
    par = np.concatenate([rs.random_sample((50, 2)), out["bpar20"][:30]], 0)
    for cu in (20, 21):
        _, _, ku, kv = AP.uniform_knot_bspline_(cu, 20, 3, 3, 2)
        rows = [bs.basis_functions(par[i], cu, 20, ku, kv, 3, 3) for i in range(par.shape[0])]
        out[f"old_NU{cu}"] = np.concatenate([r[0] for r in rows], 1).T
        out[f"old_NV{cu}"] = np.concatenate([r[1] for r in rows], 1).T
    out["old_par"] = par
    np.savez_compressed(os.path.join(OUT, "kronecker.npz"), **out)


GENS["kronecker"] = gen_kronecker

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(GENS)
    for n in names:
        GENS[n]()
        print("wrote", n)
