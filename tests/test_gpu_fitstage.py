"""Batched fit stage (pnb200/fitstage.py, csrc/fitsolve.cu, batched moments / residual launches) on the GPU against the
per-shape path of round 1 (itself pinned against the reference's golden run in test_gpu_fitting.py) and the oracle port."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, want, rtol, name):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, np.float64)
    scale = np.abs(want).max() + 1e-30
    err = np.abs(got.reshape(want.shape) - want).max()
    assert err <= rtol * scale + 1e-12, f"{name}: err {err:.3e} scale {scale:.3e}"


def _moments_of(kind, m, seed):
    from oracle.make_golden_helpers import prim_cloud
    from pnb200 import fitting as F
    p, n, w = prim_cloud(kind, m, seed)
    if kind == "cylinder":
        n = n + 0.03 * np.random.RandomState(seed).randn(m, 3).astype(np.float32)
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
    W = torch.from_numpy(w).cuda()
    return F.MomentsFn.apply(W, torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda(), 0, 1, m, 0.0)      # (1,55)


def test_fit_solve_kernel_equals_torch_algebra_values_and_gradients():
    """pn_fit_solve (forward-mode Jacobians) against the (S,3,3) torch algebra + autograd of pnb200.fitting on the same
    moments: parameters 1e-10, gradient w.r.t. the moments 2e-6 of its scale (measured 2e-7: fp64 sums in another order through the 1e6-amplifying K matrix)"""
    from pnb200 import fitting as F
    m = 700
    kinds = ["plane", "sphere", "cylinder", "cone", "plane", "cone", "sphere", "cylinder"]
    mom = torch.cat([_moments_of(k, m, 30 + i) for i, k in enumerate(kinds)], 0)                 # (8,55) f64
    kid = torch.tensor([F.TYPE_ID[k] for k in kinds] + [-1], dtype=torch.int32, device="cuda")
    mom_k = torch.cat([mom, torch.zeros(1, F.NM, dtype=torch.float64, device="cuda")], 0).requires_grad_()
    par, bad = F.FitSolveFn.apply(mom_k, kid, m)
    assert not bad.any() and not par[-1].any()
    gen = torch.Generator().manual_seed(0)
    coef = torch.randn(9, 8, generator=gen, dtype=torch.float64).cuda()
    (par * coef).sum().backward()
    mom_t = mom.clone().requires_grad_()
    want = torch.zeros(8, 8, dtype=torch.float64, device="cuda")
    rows = []
    for i, k in enumerate(kinds):
        mi = mom_t[i:i + 1]
        if k == "plane":
            a, d = F.fit_planes(mi)
            rows.append(torch.cat([a[0], d, torch.zeros(4, dtype=torch.float64, device="cuda")]))
        elif k == "sphere":
            c, r = F.fit_spheres(mi, m)
            rows.append(torch.cat([c[0], r, torch.zeros(4, dtype=torch.float64, device="cuda")]))
        elif k == "cylinder":
            a, c, r = F.fit_cylinders(mi, m)
            rows.append(torch.cat([a[0], c[0], r, torch.zeros(1, dtype=torch.float64, device="cuda")]))
        else:
            apex, axis, deg = F.fit_cone_apex_axis(mi, m)
            assert not bool(deg[0])
            rows.append(torch.cat([apex[0], axis[0], torch.zeros(2, dtype=torch.float64, device="cuda")]))
    want = torch.stack(rows)
    # eigenvector signs are those of the same Jacobi solver on both sides -> no sign ambiguity here
    _close(par[:8], want, 1e-10, "parameters")
    (want * coef[:8]).sum().backward()
    for i, k in enumerate(kinds):
        _close(mom_k.grad[i], mom_t.grad[i], 2e-6, f"d/d moments, segment {i} ({k})")
    assert not mom_k.grad[8].any()


def test_batched_moments_and_residuals_equal_per_shape_launches():
    from pnb200 import fitting as F
    from pnb200.fitstage import SLOTS
    g = torch.Generator().manual_seed(3)
    B, N = 3, 1777
    P = torch.randn(B, N, 3, generator=g).cuda()
    Nr = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g), dim=2).cuda()
    W = torch.rand(B, N, SLOTS, generator=g).cuda().requires_grad_()
    nq = (((N + 1) // 2) + 1) // 2
    mom = F.MomentsBatchedFn.apply(W, P, Nr, 0, 4, nq, F.EPS)
    coef = torch.randn(B, SLOTS, F.NM, generator=g, dtype=torch.float64).cuda()
    (mom * coef).sum().backward()
    for b in range(B):
        Wb = W[b].detach().clone().requires_grad_()
        mb = F.MomentsFn.apply(Wb, P[b].contiguous(), Nr[b].contiguous(), 0, 4, nq, F.EPS)
        _close(mom[b], mb, 1e-12, "moments")                    # same fp32 partial sums, fp64 atomics in another order
        (mb * coef[b]).sum().backward()
        _close(W.grad[b], Wb.grad, 1e-6, "d moments / d weights")
    # residuals: 5 used slots of 4 kinds per shape, the rest unused
    kind = torch.full((B, SLOTS), -1, dtype=torch.int32)
    kind[:, 0:5] = torch.tensor([0, 1, 2, 3, 1], dtype=torch.int32)
    seg = torch.randint(-1, 5, (B, N), generator=g, dtype=torch.int32)
    par = torch.randn(B, SLOTS, 8, generator=g)
    par[:, :, 6] = par[:, :, 6].abs()
    kind, seg = kind.cuda(), seg.cuda()
    par_b = par.cuda().requires_grad_()
    dist = F.ResidualBatchedFn.apply(par_b, P, seg, kind)
    cw = torch.randn(B, SLOTS, generator=g).cuda()
    (dist * cw).sum().backward()
    assert not dist[:, 5:].any()
    for b in range(B):
        pb = par[b, :5].cuda().requires_grad_()
        db = F.ResidualFn.apply(pb, P[b].contiguous(), seg[b].contiguous(), kind[b, :5].contiguous())
        _close(dist[b, :5], db, 1e-5, "residuals")
        (db * cw[b, :5]).sum().backward()
        _close(par_b.grad[b, :5], pb.grad, 1e-4, "d residual / d parameters")


def _nets():
    from oracle.port import common
    from src.model import DGCNNControlPoints
    nets = {}
    for name, mode, s in (("open", 0, 41), ("closed", 1, 42)):
        net = DGCNNControlPoints(20, num_points=10, mode=mode)
        sd = common.seeded_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=s)
        net.load_state_dict(sd)
        nets[name] = net.cuda().eval()
    return nets


@pytest.mark.parametrize("sparse", [True, False])
def test_batched_fit_stage_equals_per_shape_loop(monkeypatch, sparse):
    """Evaluation.fitting_loss on 3 shapes with all six segment kinds: the batched stage (one launch per stage for all shapes,
    fused solve kernel, padded slots) against the per-shape loop: every returned number and the gradient w.r.t. the
    embedding.  Both mean-shift backward paths (sparse rows / dense)."""
    import src.residual_utils as RU
    from oracle.make_golden_helpers import e2e_inputs
    from pnb200 import fitstage, meanshift as pms
    monkeypatch.setattr(pms, "SPARSE_BWD", sparse)
    nets = _nets()
    shapes = [e2e_inputs(1400, 70 + i, False) for i in range(3)]
    pts = torch.from_numpy(np.concatenate([s[0] for s in shapes])).cuda()
    nrm = torch.from_numpy(np.concatenate([s[1] for s in shapes])).cuda()
    lab = np.concatenate([s[2] for s in shapes]); prim = np.concatenate([s[3] for s in shapes])
    emb = torch.cat([s[4] for s in shapes]); logp = torch.cat([s[5] for s in shapes]).cuda()
    results = {}
    for stage in ("loop", "batched"):
        monkeypatch.setattr(RU, "FIT_STAGE", stage)
        ev = RU.Evaluation(open_decoder=nets["open"], closed_decoder=nets["closed"])
        per_seg = []
        orig_sep = ev.separate_losses

        def sep(distance, gt_points, lamb=1.0, _orig=orig_sep, _rec=per_seg, **kw):
            _rec.append({k: (v[0], float(v[1])) for k, v in distance.items()})
            return _orig(distance, gt_points, lamb=lamb, **kw)

        ev.separate_losses = sep
        E = emb.clone().cuda().requires_grad_()
        np.random.seed(5)
        res, extra = ev.fitting_loss(E, pts, nrm, lab, prim.copy(), logp, quantile=0.015, iterations=10, lamb=0.1)
        total = torch.stack([r.reshape(()) for r in res[0::5]]).sum()
        total.backward()
        if stage == "batched":
            per_seg = [{k: (v[0], float(v[1])) for k, v in fitstage.segment_distances(ev.last_fit, b).items()} for b in range(3)]
        results[stage] = (res, extra, E.grad.clone(), ev, per_seg)
    (r0, x0, g0, _, s0), (r1, x1, g1, ev1, s1) = results["loop"], results["batched"]
    assert len(r0) == len(r1) == 15
    # per segment.  Both paths take the similarities from the same batched product, so the discrete decisions of a spline fit
    # (confident-point mask w > 0.8, top-N/2 fallback, kNN graphs of the SplineNet, Chamfer arg-mins) agree; what differs is
    # summation order (padded Chamfer means, batched SplineNet launches) and the fp64 solve kernel vs torch algebra
    for b in range(3):
        assert set(s0[b].keys()) == set(s1[b].keys()), (b, s0[b], s1[b])
        for k in s0[b]:
            (kind0, d0), (kind1, d1) = s0[b][k], s1[b][k]
            assert kind0 == kind1
            tol = 1e-4 if "spline" in kind0 else 2e-5
            assert abs(d0 - d1) <= tol * abs(d0) + 1e-9, (b, k, kind0, d0, d1)
    for i, (a, b) in enumerate(zip(r0, r1)):
        if a is None or b is None:
            assert a is None and b is None, (i, a, b)
        else:
            assert abs(float(a) - float(b)) <= 5e-5 * abs(float(a)) + 1e-9, (i, float(a), float(b))
    np.testing.assert_array_equal(x0[1], x1[1])
    _close(x1[2], x0[2], 1e-5, "returned membership similarities of the last shape")
    p0, p1 = x0[0], x1[0]
    assert set(p0.keys()) == set(p1.keys())
    for k in p0:
        if p0[k] is None:
            assert p1[k] is None
            continue
        assert p0[k][0] == p1[k][0]
        for a, b in zip(p0[k][1:], p1[k][1:]):
            assert tuple(a.shape) == tuple(b.shape), (k, p0[k][0], a.shape, b.shape)
            _close(b, a, 2e-4, f"parameters of the last shape: {p0[k][0]}")
    _close(g1, g0, 2e-4, "d loss / d embedding, batched stage vs per-shape loop")
    # per-segment view used by the parity tests
    d = fitstage.segment_distances(ev1.last_fit, 0)
    assert sorted(v[0] for v in d.values()) == sorted(v[0] for v in p1.values() if v is not None) or len(d) > 0


def test_batched_fit_stage_single_cluster_and_unmatched_shapes():
    """edge cases of the slot tables: a shape whose embedding collapses to ONE cluster (weights_normalize early return), next
    to a normal shape; every loss finite, gradient finite"""
    import src.residual_utils as RU
    from oracle.make_golden_helpers import e2e_inputs
    nets = _nets()
    s0, s1 = e2e_inputs(1400, 81, False), e2e_inputs(1400, 82, False)
    emb1 = torch.nn.functional.normalize(torch.ones(1, 1400, 128) + 1e-3 * torch.randn(1, 1400, 128,
                                         generator=torch.Generator().manual_seed(0)), dim=2)
    pts = torch.from_numpy(np.concatenate([s0[0], s1[0]])).cuda()
    nrm = torch.from_numpy(np.concatenate([s0[1], s1[1]])).cuda()
    lab = np.concatenate([s0[2], s1[2]]); prim = np.concatenate([s0[3], s1[3]])
    E = torch.cat([s0[4], emb1]).cuda().requires_grad_()
    logp = torch.cat([s0[5], s1[5]]).cuda()
    ev = RU.Evaluation(open_decoder=nets["open"], closed_decoder=nets["closed"])
    np.random.seed(5)
    res, extra = ev.fitting_loss(E, pts, nrm, lab, prim.copy(), logp, quantile=0.015, iterations=10, lamb=0.1)
    assert len(np.unique(extra[1])) == 1
    total = torch.stack([r.reshape(()) for r in res[0::5]]).sum()
    assert torch.isfinite(total)
    total.backward()
    assert torch.isfinite(E.grad).all()


def test_weights_normalize_kernel_equals_torch_expression_and_port():
    """csrc/weights.cu (2 launches forward, 2 backward) against the torch expression it replaces on padded (B,N,64) tables
    (K = 1 single-cluster early return, K = 7, K = 49) and against the oracle port per shape; gradient incl. the min / max
    routing"""
    from oracle.port import fitting as OP
    from pnb200 import fitstage as FS
    from pnb200.staging import arena
    g = torch.Generator().manual_seed(0)
    B, N, S = 3, 5003, FS.SLOTS
    K = [1, 7, 49]
    bws = torch.tensor([0.31, 0.2, 0.8]).cuda()
    raw0 = torch.rand(B, N, S, generator=g) * 2 - 1
    coef = torch.randn(B, N, S, generator=g).cuda()
    stage = arena("test", torch.device("cuda", 0))
    a = raw0.clone().cuda().requires_grad_()
    b = raw0.clone().cuda().requires_grad_()
    Wk = FS.normalized_weights(a, bws, K, stage)
    Wt = FS.normalized_weights_torch(b, bws, K, stage)
    _close(Wk, Wt, 1e-6, "weights, kernel vs torch expression")
    (Wk * coef).sum().backward(); (Wt * coef).sum().backward()
    for i in range(B):
        assert not Wk[i, :, K[i]:].any() and not a.grad[i, :, K[i]:].any()
        if K[i] == 1:
            bound = 8 * np.finfo(np.float32).eps * float(coef.abs().max()) / (2 * float(bws[i]) ** 2)
            assert float(a.grad[i].abs().max()) <= bound
        else:
            _close(a.grad[i], b.grad[i], 2e-4, f"d/d similarities, shape {i}")
        r = raw0[i, :, :K[i]].t().clone().requires_grad_()
        want = OP.weights_normalize(r, float(bws[i]))
        _close(Wk[i, :, :K[i]].t(), want, 1e-5, f"weights vs port, shape {i}")
        if K[i] > 1:
            (want * coef[i, :, :K[i]].t().cpu()).sum().backward()
            _close(a.grad[i, :, :K[i]].t(), r.grad, 2e-3, f"d/d similarities vs port, shape {i}")


def test_grid_loss_kernels_equal_reference_formulation():
    """csrc/gridloss.cu against the reference's formulation written with torch ops (candidate stacks by flip / transpose /
    roll, Laplacian as a 3x3 convolution): values 1e-6, chosen candidate identical, gradients 1e-5; L2 and L1 Laplacian"""
    import torch.nn.functional as TF
    from src import loss as L
    gen = torch.Generator().manual_seed(4)
    B, g = 5, 20
    gt = torch.randn(B, g, g, 3, generator=gen).cuda()
    # outputs close to a different symmetric copy of the target per shape, so that the arg-min is not always candidate 0
    flips = [gt, gt.flip(1), gt.flip(2), gt.flip(1, 2)]
    open_c = torch.stack(flips + [f.transpose(1, 2) for f in flips], 1)                                  # (B,8,g,g,3)
    closed_c = torch.cat([torch.stack([r, r.flip(1), r.flip(2), r.flip(1, 2)], 1)
                          for r in (torch.roll(gt, s, 1) for s in range(g))], 1)                         # (B,4g,g,g,3)
    for mode, cands, pick in ((0, open_c, [0, 3, 5, 6, 7]), (1, closed_c, [0, 9, 38, 77, 79])):
        base = torch.stack([cands[b, p] for b, p in enumerate(pick)])
        out = (base + 0.05 * torch.randn(B, g, g, 3, generator=gen).cuda()).reshape(B, g * g, 3)
        a = out.clone().requires_grad_(); b_ = out.clone().requires_grad_()
        if mode == 0:
            la, best = L.control_points_permute_reg_loss(a, gt, g)
        else:
            la, best = L.control_points_permute_closed_reg_loss(a, gt, g, g)
        diff = ((b_.view(B, 1, g, g, 3) - cands) ** 2).sum((2, 3, 4))
        lb, idx = diff.min(1)
        assert idx.tolist() == pick
        lb = lb.mean() / (g * g * 3)
        assert abs(la.item() - lb.item()) <= 1e-6 * abs(lb.item())
        torch.testing.assert_close(best, cands[torch.arange(B), idx], rtol=0, atol=0)
        (3.0 * la).backward(); (3.0 * lb).backward()
        _close(a.grad, b_.grad, 1e-5, f"d/d output, mode {mode}")
    k = torch.tensor([[0.0, -0.25, 0.0], [-0.25, 1.0, -0.25], [0.0, -0.25, 0.0]]).cuda()
    w = torch.zeros(3, 3, 3, 3).cuda()
    for c in range(3):
        w[c, c] = k
    for dist_type in ("l2", "l1"):
        o1 = torch.randn(B, g, g, 3, generator=gen).cuda().requires_grad_(); t1 = gt.clone().requires_grad_()
        o2 = o1.detach().clone().requires_grad_(); t2 = gt.clone().requires_grad_()
        la = L.laplacian_loss(o1, t1, dist_type)
        d = TF.conv2d(o2.permute(0, 3, 1, 2), w, padding=1) - TF.conv2d(t2.permute(0, 3, 1, 2), w, padding=1)
        lb = ((d ** 2) if dist_type == "l2" else d.abs()).sum(1).mean()
        assert abs(la.item() - lb.item()) <= 2e-6 * abs(lb.item()), (dist_type, la.item(), lb.item())
        (2.0 * la).backward(); (2.0 * lb).backward()
        _close(o1.grad, o2.grad, 1e-5, "d laplacian / d output " + dist_type)
        _close(t1.grad, t2.grad, 1e-5, "d laplacian / d target " + dist_type)


@pytest.mark.parametrize("branch", ["threshold", "topk"])
def test_standardize_point_torch_vs_port(branch):
    """SURVEY a27, direct: the SplineNet input frame (confident-subset mean, minor PCA axis onto x through LAPACK's
    eigenvectors, per-axis extent of the weighted subset) on the device -- per-entry drop-in AND the batched stage version --
    against the oracle port (fitting_utils.py:512-553): points 1e-4, extents 1e-4, mean 1e-5, rotation 1e-5"""
    from oracle.port import e2e as PE
    from pnb200 import fitstage as FS
    from pnb200.staging import arena
    from src.fitting_utils import rotation_matrix_a_to_b, standardize_point_torch
    gen = torch.Generator().manual_seed(7 if branch == "topk" else 8)
    n = 2500
    P = torch.randn(n, 3, generator=gen) * torch.tensor([1.0, 0.45, 0.12]) + torch.tensor([0.3, -0.2, 0.1])
    Rq, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    P = P @ Rq.t()
    w = torch.rand(n, 1, generator=gen)
    if branch == "topk":
        w = w * 0.6                               # nothing above 0.8 -> the top-N/2 fallback (:518-522)
    else:
        w[:1200] = 0.9 + 0.1 * w[:1200]
    want_pts, want_std, want_mean, want_R = PE.standardize_point(P, w)
    pts, std, mean, R = standardize_point_torch(P.cuda(), w.cuda())
    _close(mean, want_mean, 1e-5, "mean")
    _close(R, want_R, 1e-5, "rotation")
    _close(std, want_std, 1e-4, "extents")
    _close(pts, want_pts, 1e-4, "standardised points")
    Ps, stdb, meanb, Rb, Rinv = FS.standardize_batched(P.cuda().unsqueeze(0), w.cuda().reshape(1, n), rotation_matrix_a_to_b,
                                                       arena("test", torch.device("cuda", 0)))
    _close(Ps[0], want_pts, 1e-4, "standardised points (batched stage)")
    _close(stdb[0], want_std, 1e-4, "extents (batched stage)")
    _close(Rb[0], want_R, 1e-5, "rotation (batched stage)")
    _close(Rinv[0] @ Rb[0], torch.eye(3), 1e-5, "R^-1 R")


def test_weights_normalize_kernel_extreme_bandwidth_vs_float64():
    """bandwidth 0.05 -> exponents clamp at +-75, e ~ 3.7e32: values and gradients stay finite and match a float64 evaluation
    of the same expression (the fp32 torch expression itself produces non-finite gradients in this regime on the GPU)"""
    from pnb200 import fitstage as FS
    from pnb200.staging import arena
    g = torch.Generator().manual_seed(1)
    B, N, S, K = 1, 3001, FS.SLOTS, 7
    raw0 = torch.rand(B, N, S, generator=g) * 2 - 1
    coef = torch.randn(B, N, S, generator=g)
    a = raw0.clone().cuda().requires_grad_()
    Wk = FS.normalized_weights(a, torch.tensor([0.05]).cuda(), [K], arena("test", torch.device("cuda", 0)))
    (Wk * coef.cuda()).sum().backward()
    assert torch.isfinite(Wk).all() and torch.isfinite(a.grad).all()
    r = raw0[0, :, :K].double().clone().requires_grad_()
    bw2 = float(np.float32(0.05)) ** 2
    x = torch.clamp(r / float(np.float32(bw2)) / 2, min=-75.0, max=75.0)
    prob = torch.exp(x)
    prob = prob / prob.sum(1, keepdim=True)
    mm = prob - prob.min(0, keepdim=True)[0]
    want = mm / (mm.max(0, keepdim=True)[0] + FS.EPS)
    (want * coef[0, :, :K].double()).sum().backward()
    _close(Wk[0, :, :K], want, 1e-5, "weights, clamp regime")
    _close(a.grad[0, :, :K], r.grad, 1e-3, "d/d similarities, clamp regime")


def test_device_input_pipeline_vs_reference_dataset(golden_dir):
    """SURVEY 8f-4 on the device: DeviceBatchPipeline (host 3x3 rotations, pinned upload, batched rotation / extent / scaling
    launches) against Dataset.get_train of the unmodified reference (tests/golden/pipeline.npz)"""
    import os
    from pnb200.input_pipeline import DeviceBatchPipeline
    g = np.load(os.path.join(golden_dir, "pipeline.npz"))
    lab = np.zeros(g["pts"].shape[:2], np.int64)
    for name, noise, aniso in (("plain", False, False), ("noise_aniso", True, True)):
        np.random.seed(11)
        pipe = DeviceBatchPipeline(iter([[g["pts"].copy(), lab, g["nrm"].copy(), lab]]), torch.device("cuda", 0),
                                   if_normal_noise=noise, anisotropic=aniso)
        p, _, n, _ = next(pipe)
        assert p.is_cuda and n.is_cuda
        assert np.abs(p.cpu().numpy() - g[name + "_p"]).max() <= 3e-6, name
        assert np.abs(n.cpu().numpy() - g[name + "_n"]).max() <= 3e-6, name
