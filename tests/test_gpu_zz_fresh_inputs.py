"""Fitting / loss stage on the GPU vs the oracle port (oracle/port/fitting.py, pinned against the reference's golden vectors
in tests/test_oracle_golden.py) on FRESH seeded inputs: sizes and shapes the fixed fixtures do not cover (tiny and ragged
segments, point counts that are not tile multiples, single-cluster weights, full-size Chamfer through properties).

First executed on a B200 in the first GPU call of round 2 (gpurun_out/r02_first/00_gpu_tests.txt: all green except the
K = 1 case of test_weights_normalize_fresh_vs_port, whose expected gradient is identically zero and whose tolerance was
relative to that zero -- fixed below); they are hard tests now.  The file name sorts last on purpose: should one of the
edge sizes (1-point segments, 1 x 1 Chamfer) ever fault a kernel, the sticky CUDA error cannot take the other files down.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, want, rtol, name):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, np.float64)
    scale = np.abs(want).max() + 1e-30
    err = np.abs(got.reshape(want.shape) - want).max()
    assert err <= rtol * scale + 1e-9, f"{name}: err {err:.3e} scale {scale:.3e}"


def _cu(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).cuda()
    return t.requires_grad_() if grad else t


def _cpu(a, grad=False):
    t = torch.from_numpy(np.asarray(a))
    return t.requires_grad_() if grad else t


# ------------------------------------------------------------------------------------------------ primitive fits
@pytest.mark.parametrize("kind,m,seed", [("plane", 61, 11), ("plane", 4999, 12), ("sphere", 61, 13), ("sphere", 4999, 14),
                                         ("cone", 300, 15), ("cone", 4999, 16), ("cylinder", 2500, 17)])
def test_fits_on_fresh_clouds_vs_port(kind, m, seed):
    """Fit.fit_*_torch on new noisy primitive samples (odd sizes) vs the port: parameters 2e-4, d/dweights 2e-3"""
    from oracle.make_golden_helpers import prim_cloud
    from oracle.port import fitting as OP
    from src.primitive_forward import Fit
    p, n, w = prim_cloud(kind, m, seed)
    Wc = _cpu(w, True)
    want = {"plane": lambda: OP.fit_plane(_cpu(p), Wc), "sphere": lambda: OP.fit_sphere(_cpu(p), Wc),
            "cylinder": lambda: OP.fit_cylinder(_cpu(p), _cpu(n), Wc), "cone": lambda: OP.fit_cone(_cpu(p), _cpu(n), Wc)}[kind]()
    Wg = _cu(w, True)
    got = getattr(Fit(), f"fit_{kind}_torch")(_cu(p), _cu(n), Wg)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert tuple(a.shape) == tuple(b.shape), (kind, a.shape, b.shape)
    sign = 1.0
    if kind in ("plane", "cylinder"):          # eigenvector sign convention (LAPACK vs Jacobi), see DESIGN.md §4
        sign = float(np.sign((got[0].detach().cpu().reshape(-1) * want[0].detach().reshape(-1)).sum()))
    g = torch.Generator().manual_seed(seed)
    coefs = [torch.randn(t.shape, generator=g) for t in want]
    if kind == "cylinder":
        # centre / radius carry the reference's along-axis fp32 noise (declared deviation): compare axis and the
        # perpendicular centre component only
        _close(got[0] * sign, want[0], 2e-4, "cylinder axis")
        ax = want[0].detach().double().reshape(3)
        perp = lambda c: c - (c * ax).sum() * ax
        _close(perp(got[1].detach().cpu().double().reshape(3)), perp(want[1].detach().double().reshape(3)), 3e-2,
               "cylinder centre (perpendicular)")
        return
    loss_g = loss_c = 0
    for i, (a, b) in enumerate(zip(got, want)):
        s = sign if (kind == "plane") else 1.0
        _close(a * s, b, 2e-4, f"{kind} out{i}")
        loss_g = loss_g + (a * s * coefs[i].cuda()).sum()
        loss_c = loss_c + (b * coefs[i]).sum()
    loss_g.backward(); loss_c.backward()
    _close(Wg.grad, Wc.grad, 2e-3, f"{kind} d/dweights")


@pytest.mark.parametrize("m", [1, 17, 3001])
def test_residual_distances_on_fresh_points_vs_port(m):
    """ComputePrimitiveDistance.* for all four analytic kinds (ragged point counts down to a single point)"""
    from oracle.port import fitting as OP
    from src.primitives import ComputePrimitiveDistance
    g = torch.Generator().manual_seed(100 + m)
    q = torch.randn(m, 3, generator=g) * 0.5
    unit = lambda v: v / v.norm()
    params = {
        "plane": [unit(torch.randn(3, 1, generator=g)), torch.tensor(0.13)],
        "sphere": [torch.randn(1, 3, generator=g) * 0.2, torch.tensor(0.6)],
        "cylinder": [unit(torch.randn(3, 1, generator=g)), torch.randn(1, 3, generator=g) * 0.2, torch.tensor(0.35)],
        "cone": [torch.randn(1, 3, generator=g) * 0.2, unit(torch.randn(3, 1, generator=g)), torch.tensor(0.5)],
    }
    cp = ComputePrimitiveDistance(reduce=True)
    for kind, ps in params.items():
        pc = [t.clone().requires_grad_() for t in ps]
        pg = [t.clone().cuda().requires_grad_() for t in ps]
        want = OP.DISTANCES[kind](q, pc)
        got = getattr(cp, "distance_from_" + kind)(points=q.cuda(), params=pg, sqrt=False)
        _close(got, want, 1e-4, f"{kind} residual m={m}")
        got.backward(); want.backward()
        for i in range(len(ps)):
            _close(pg[i].grad, pc[i].grad, 1e-3, f"{kind} dpar{i} m={m}")


# ------------------------------------------------------------------------------------------------ Chamfer
@pytest.mark.parametrize("B,Np,M", [(1, 1, 1), (2, 7, 1030), (3, 1024, 1025), (1, 2500, 900)])
def test_chamfer_ragged_sizes_vs_port(B, Np, M):
    from oracle.port import fitting as OP
    from src import utils as U
    g = torch.Generator().manual_seed(B * 1000 + Np)
    pred, gt = torch.randn(B, Np, 3, generator=g), torch.randn(B, M, 3, generator=g)
    pc, gc = pred.clone().requires_grad_(), gt.clone().requires_grad_()
    pg, gg = pred.clone().cuda().requires_grad_(), gt.clone().cuda().requires_grad_()
    want = [OP.chamfer_distance(pc, gc), OP.chamfer_distance(pc, gc, sqrt=True), OP.chamfer_distance_one_side(pc, gc, 0),
            OP.chamfer_distance_one_side(pc, gc, 1), OP.chamfer_distance_single_shape(pc[0], gc[0]),
            OP.chamfer_distance_single_shape(pc[0], gc[0], one_side=True, sqrt=True)]
    got = [U.chamfer_distance(pg, gg), U.chamfer_distance(pg, gg, sqrt=True), U.chamfer_distance_one_side(pg, gg, 0),
           U.chamfer_distance_one_side(pg, gg, 1), U.chamfer_distance_single_shape(pg[0], gg[0]),
           U.chamfer_distance_single_shape(pg[0], gg[0], one_side=True, sqrt=True)]
    for i, (a, b) in enumerate(zip(got, want)):
        assert abs(a.item() - b.item()) <= 1e-4 * abs(b.item()) + 1e-9, (i, a.item(), b.item())
    sum((i + 1) * a for i, a in enumerate(got)).backward()
    sum((i + 1) * b for i, b in enumerate(want)).backward()
    _close(pg.grad, pc.grad, 1e-4, "chamfer d/dpred"); _close(gg.grad, gc.grad, 1e-4, "chamfer d/dgt")
    # per-point vectors (reduce=False) and numpy inputs (utils.py:280-284)
    v_got = U.chamfer_distance_single_shape(pred[0].numpy(), gt[0].numpy(), one_side=True, reduce=False)
    v_want = OP.chamfer_distance_single_shape(pred[0], gt[0], one_side=True, reduce=False)
    assert v_got.shape == (M,)
    _close(v_got, v_want, 1e-4, "per-point one-sided distances")


def test_chamfer_full_size_properties():
    """cfg-3-scale clouds (36 x 2000 x 1600) and a 10^4 x 10^4 pair through size-independent properties: symmetry under
    swapping the arguments, zero on identical clouds, invariance under a permutation of the points, exact value on a
    translated copy of a lattice"""
    from src import utils as U
    g = torch.Generator().manual_seed(5)
    a = torch.randn(36, 1600, 3, generator=g).cuda(); b = torch.randn(36, 2000, 3, generator=g).cuda()
    ab, ba = U.chamfer_distance(a, b).item(), U.chamfer_distance(b, a).item()
    assert abs(ab - ba) <= 1e-6 * abs(ab)
    assert U.chamfer_distance(a, a).item() == 0.0
    perm = torch.randperm(2000, generator=g).cuda()
    assert abs(U.chamfer_distance(a, b[:, perm]).item() - ab) <= 1e-6 * abs(ab)
    assert abs(U.chamfer_distance_one_side(a, b, 0).item() + U.chamfer_distance_one_side(a, b, 1).item() - 2 * ab) <= 1e-5 * ab
    big = torch.randn(1, 10000, 3, generator=g).cuda()
    assert U.chamfer_distance(big, big.flip(1)).item() == 0.0
    # integer lattice shifted by 0.25 along x: every nearest neighbour is the own copy, squared distance 1/16 exactly
    ax = torch.arange(22, dtype=torch.float32)
    lat = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).cuda()      # 10 648 points
    shifted = lat + torch.tensor([0.25, 0.0, 0.0]).cuda()
    assert abs(U.chamfer_distance(lat, shifted).item() - 0.0625) <= 1e-7


# ------------------------------------------------------------------------------------------------ splines
@pytest.mark.parametrize("B,grid", [(1, 30), (5, 40)])
def test_spline_evaluation_and_losses_fresh_vs_port(B, grid):
    from oracle.port import fitting as OP
    from src import loss as L
    from src.fitting_utils import sample_points_from_control_points_
    g = torch.Generator().manual_seed(B + grid)
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, grid)
    onu, onv = OP.uniform_knot_bspline(20, 20, 3, 3, grid)
    np.testing.assert_allclose(nu, onu, rtol=0, atol=1e-13); np.testing.assert_allclose(nv, onv, rtol=0, atol=1e-13)
    assert np.allclose(nu.sum(1), 1.0) and (np.count_nonzero(nu, axis=1) <= 4).all()
    nuf, nvf = torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32))
    cp = torch.rand(B, 400, 3, generator=g) - 0.5
    cc, cg = cp.clone().requires_grad_(), cp.clone().cuda().requires_grad_()
    want = OP.sample_points_from_control_points_(nuf, nvf, cc, B)
    got = sample_points_from_control_points_(nuf, nvf, cg, B)
    _close(got, want, 1e-4, "surface points")
    pts = torch.randn(B, 3, 777, generator=g) * 0.3
    gtcp = torch.rand(B, 20, 20, 3, generator=g) - 0.5

    class Cfg:
        batch_size = B
        grid_size = 20
    w_cd, _ = OP.spline_reconstruction_loss_one_sided(nuf, nvf, cc, pts, B, 20)
    w_cd2, _ = OP.spline_reconstruction_loss(nuf, nvf, cc, pts, B, sqrt=True)
    w_reg, w_perm = OP.control_points_permute_reg_loss(cc, gtcp, 20)
    w_lap = OP.laplacian_loss(cc.reshape(B, 20, 20, 3), w_perm)
    w_closed, _ = OP.control_points_permute_closed_reg_loss(cc, gtcp, 20, 20)
    g_cd, _ = L.spline_reconstruction_loss_one_sided(nuf, nvf, cg, pts.cuda(), Cfg)
    g_cd2, _ = L.spline_reconstruction_loss(nuf, nvf, cg, pts.cuda(), Cfg, sqrt=True)
    g_reg, g_perm = L.control_points_permute_reg_loss(cg, gtcp.cuda(), 20)
    g_lap = L.laplacian_loss(cg.reshape(B, 20, 20, 3), g_perm)
    g_closed, _ = L.control_points_permute_closed_reg_loss(cg, gtcp.cuda(), 20, 20)
    for name, a, b in [("one-sided", g_cd, w_cd), ("two-sided sqrt", g_cd2, w_cd2), ("reg", g_reg, w_reg),
                       ("laplacian", g_lap, w_lap), ("closed reg", g_closed, w_closed)]:
        assert abs(a.item() - b.item()) <= 1e-4 * abs(b.item()), (name, a.item(), b.item())
    _close(g_perm, w_perm, 1e-6, "best-matching permutation of the gt grid")
    (g_cd + 0.5 * g_cd2 + 0.9 * g_reg + 0.1 * g_lap + 0.5 * g_closed).backward()
    (w_cd + 0.5 * w_cd2 + 0.9 * w_reg + 0.1 * w_lap + 0.5 * w_closed).backward()
    _close(cg.grad, cc.grad, 2e-4, "d/dcontrol points")


@pytest.mark.parametrize("K,N", [(1, 333), (7, 1000), (49, 10000)])
def test_weights_normalize_fresh_vs_port(K, N):
    """membership weights incl. the single-cluster early return (fitting_utils.py:318-319) and the 49-cluster maximum"""
    from oracle.port import fitting as OP
    from src.fitting_utils import weights_normalize
    g = torch.Generator().manual_seed(K)
    w = torch.rand(K, N, generator=g) * 2 - 1
    wc, wg = w.clone().requires_grad_(), w.clone().cuda().requires_grad_()
    want, got = OP.weights_normalize(wc, 0.31), weights_normalize(wg, 0.31)
    _close(got, want, 1e-4, "weights")
    coef = torch.randn(K, N, generator=g)
    (got * coef.cuda()).sum().backward(); (want * coef).sum().backward()
    if K == 1:
        # single cluster: prob = e / e == 1 for every point, the exact gradient is 0; what both sides hold is the rounding
        # residue of coef / e - coef * e / e^2 (a few ulp of |coef| / (2 b^2)), so compare against that scale, not against 0
        bound = 4 * np.finfo(np.float32).eps * float(coef.abs().max()) / (2 * 0.31 ** 2)
        assert float(wg.grad.abs().max()) <= bound and float(wc.grad.abs().max()) <= bound
    else:
        _close(wg.grad, wc.grad, 2e-3, "d/dweights")


# ------------------------------------------------------------------------------------------------ end to end
@pytest.mark.parametrize("stage", ["batched", "loop"])
@pytest.mark.parametrize("N,seed", [(1800, 91), (2600, 92)])
def test_fitting_loss_fresh_shape_vs_port(N, seed, stage, monkeypatch):
    """Evaluation.fitting_loss on a NEW synthetic shape (not the golden one) against the complete oracle port
    (oracle/port/e2e.py, pinned on CPU against the reference's own run): identical partition and segment kinds,
    per-segment residuals 1e-3 (cylinder excluded: declared deviation), loss 1e-3 when no cylinder is fitted"""
    from oracle.make_golden_helpers import e2e_inputs
    from oracle.port import common, e2e as pe2e
    from src.model import DGCNNControlPoints
    from src.residual_utils import Evaluation
    import src.residual_utils as RU
    monkeypatch.setattr(RU, "FIT_STAGE", stage)
    pts, nrm, lab, prim, emb, logp = e2e_inputs(N, seed, True)
    nets, mods = {}, {}
    for name, mode, s in (("open", 0, 41), ("closed", 1, 42)):
        net = DGCNNControlPoints(20, num_points=10, mode=mode)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        sd = common.seeded_state_dict(shapes, seed=s)
        for i in (1, 2, 3, 4, 5):
            for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
                if a in sd and b in sd:
                    sd[b] = sd[a]
        net.load_state_dict(sd)
        nets[name], mods[name] = sd, net.cuda().eval()
    np.random.seed(5)
    want, wparams, wdist, wcl = pe2e.fitting_loss(emb[0].clone(), torch.from_numpy(pts[0]), torch.from_numpy(nrm[0]), lab[0],
                                                  prim[0].copy(), nets, 0.015, 10, 0.1)
    ev = Evaluation(open_decoder=mods["open"], closed_decoder=mods["closed"])
    captured = {}
    orig = ev.separate_losses

    def sep(distance, gt_points, lamb=1.0, **kw):
        captured.update({k: (v[0], float(v[1])) for k, v in distance.items()})
        return orig(distance, gt_points, lamb=lamb, **kw)

    ev.separate_losses = sep
    np.random.seed(5)
    res, extra = ev.fitting_loss(emb.cuda(), torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda(), lab, prim.copy(),
                                 logp.cuda(), quantile=0.015, iterations=10, lamb=0.1)

    def canon(l):
        l = np.asarray(l)
        _, first = np.unique(l, return_index=True)
        m = {int(v): i for i, v in enumerate(l[np.sort(first)])}
        return np.array([m[int(v)] for v in l])
    np.testing.assert_array_equal(canon(extra[1]), canon(wcl))
    if stage == "batched":
        from pnb200 import fitstage
        captured.update({k: (v[0], float(v[1])) for k, v in fitstage.segment_distances(ev.last_fit, 0).items()})
    mine = sorted(captured.values())
    ref = sorted((v[0], float(v[1])) for v in wdist.values())
    assert [k for k, _ in mine] == [k for k, _ in ref]
    for (kind, d1), (_, d2) in zip(mine, ref):
        if kind != "cylinder":
            assert abs(d1 - d2) <= 1e-3 * d2, (kind, d1, d2)
    if "cylinder" not in [k for k, _ in ref]:
        assert abs(res[0].item() - want[0].item()) <= 1e-3 * abs(want[0].item())


# ------------------------------------------------------------------------------------------------ config 3 training step
@pytest.mark.parametrize("B,M", [(4, 700), (36, 1000)])
def test_open_spline_training_step_vs_port(B, M):
    """one optimisation step of train_open_splines.py:143-178 on fresh patches (SplineNet in TRAIN mode: batch statistics,
    one-sided spline reconstruction loss + permutation-invariant control-point regression + Laplacian loss, backward)
    against the oracle port: losses 1e-4 relative (BASELINE config 3), parameter gradient norms 5e-3.
    (36, 1000) is the batch size of configs/config_open_splines.yml at a patch size inside its 400..2000 range."""
    from oracle.port import common, e2e as pe2e, fitting as OP
    from src import loss as L
    from src.model import DGCNNControlPoints
    g = torch.Generator().manual_seed(3)
    if B == 36:     # config 3 as SURVEY 8d describes it: smooth random bicubic patches with their control grids
        from tools.synth import open_spline_batch
        p_np, cp_np = open_spline_batch(B, M, seed=3)
        pts, gtcp = torch.from_numpy(p_np), torch.from_numpy(cp_np)
    else:
        pts = torch.randn(B, 3, M, generator=g) * 0.3
        gtcp = torch.rand(B, 20, 20, 3, generator=g) - 0.5
    net = DGCNNControlPoints(20, num_points=10, mode=0)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = common.seeded_state_dict(shapes, seed=7)
    for i in (1, 2, 3, 4, 5):
        for s_ in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            a, b = f"bn{i}.{s_}", f"conv{i}.1.{s_}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    net.load_state_dict(sd)
    net = net.cuda().train()
    nu, nv = L.uniform_knot_bspline(20, 20, 3, 3, 40)
    nuf, nvf = torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32))

    class Cfg:
        batch_size = B
        grid_size = 20
    out = net(pts.cuda())
    cd, _ = L.spline_reconstruction_loss_one_sided(nuf, nvf, out, pts.cuda(), Cfg)
    reg, perm = L.control_points_permute_reg_loss(out, gtcp.cuda(), 20)
    lap = L.laplacian_loss(out.reshape(B, 20, 20, 3), perm)
    (0.9 * reg + 0.1 * (cd + lap)).backward()
    sdp = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    o_r = pe2e.splinenet_fwd(sdp, pts, 10, None, train=True)
    cd_r, _ = OP.spline_reconstruction_loss_one_sided(nuf, nvf, o_r, pts, B, 20)
    reg_r, perm_r = OP.control_points_permute_reg_loss(o_r, gtcp, 20)
    lap_r = OP.laplacian_loss(o_r.reshape(B, 20, 20, 3), perm_r)
    (0.9 * reg_r + 0.1 * (cd_r + lap_r)).backward()
    if B == 36:
        # At this batch size the train-mode network is ill-conditioned in fp32 (BatchNorm over the batch after the global
        # max-pool): the reference's OWN fp32 result differs from a float64 evaluation of the same network by 3e-2 .. 6e-2
        # of the output scale (profiles/r02_parity_bounds.md).  A 1e-4 comparison against the fp32 port is therefore not
        # meaningful; the criterion is that the CUDA path is as close to the float64 result as the fp32 reference is.
        sd64 = {k: (v.detach().double() if v.is_floating_point() else v) for k, v in sd.items()}
        with torch.no_grad():
            o64 = pe2e.splinenet_fwd(sd64, pts.double(), 10, None, train=True)
            cd64, _ = OP.spline_reconstruction_loss_one_sided(nuf.double(), nvf.double(), o64, pts.double(), B, 20)
        scale = o64.abs().max().item()
        ref_err = (o_r.detach().double() - o64).abs().max().item() / scale
        our_err = (out.detach().cpu().double() - o64).abs().max().item() / scale
        print(f"B=36 train mode: fp32 reference vs float64 {ref_err:.2e}, CUDA path vs float64 {our_err:.2e}")
        assert our_err <= 4 * ref_err + 1e-4, (our_err, ref_err)
        assert abs(cd.item() - cd64.item()) <= 4 * abs(cd_r.item() - cd64.item()) + 1e-4 * abs(cd64.item()), \
            (cd.item(), cd_r.item(), cd64.item())
        return
    _close(out, o_r, 3e-4, "control points (train mode)")
    for name, a, b in (("chamfer", cd, cd_r), ("regression", reg, reg_r), ("laplacian", lap, lap_r)):
        assert abs(a.item() - b.item()) <= 1e-4 * abs(b.item()), (name, a.item(), b.item())
    params = dict(net.named_parameters())
    for key in ("conv8.weight", "conv7.weight", "conv5.0.weight", "conv3.0.weight", "conv1.0.weight", "bn5.weight", "bn2.bias"):
        got, want = params[key].grad.norm().item(), sdp[key].grad.norm().item()
        assert abs(got - want) <= 5e-3 * want + 1e-9, (key, got, want)


# ------------------------------------------------------------------------------------------------ inference path
def test_inference_clustering_path_vs_port():
    """generate_predictions.py:131-156: normalised embedding -> Evaluation.guard_mean_shift (50 iterations, no grad) ->
    one-hot weights -> SIOU_matched_segments.  Partition identical to the oracle port's, hence identical IoU metrics."""
    from oracle.make_golden_helpers import clustered_embedding
    from oracle.port import meanshift as oms
    from src.residual_utils import Evaluation
    from src.segment_utils import SIOU_matched_segments, to_one_hot
    N = 2000
    X, gt = clustered_embedding(N, 128, 7, 123, spread=0.25)
    np.random.seed(9)
    _, _, lab_r = oms.guard_mean_shift(X, 0.015, 50, num_samples=10000, growth=1.2)
    ev = Evaluation.__new__(Evaluation)            # only the clustering half is exercised (no SplineNets needed)
    from src.mean_shift import MeanShift
    ev.ms = MeanShift()
    np.random.seed(9)
    with torch.no_grad():
        _, _, cluster_ids = ev.guard_mean_shift(X.cuda(), 0.015, 50, kernel_type="gaussian")

    def canon(l):
        l = np.asarray(l)
        _, first = np.unique(l, return_index=True)
        m = {int(v): i for i, v in enumerate(l[np.sort(first)])}
        return np.array([m[int(v)] for v in l])
    got, want = canon(cluster_ids.cpu().numpy()), canon(lab_r.numpy())
    np.testing.assert_array_equal(got, want)
    prim = (gt.numpy() % 6).astype(np.int64)
    w = to_one_hot(torch.from_numpy(got).cuda(), int(got.max()) + 1)
    s_iou, p_iou, _, _ = SIOU_matched_segments(gt.numpy(), got, prim, prim, w)
    s_iou_r, p_iou_r, _, _ = SIOU_matched_segments(gt.numpy(), want, prim, prim, w)
    assert abs(s_iou - s_iou_r) < 1e-12 and abs(p_iou - p_iou_r) < 1e-12 and 0.0 <= s_iou <= 1.0


@pytest.mark.parametrize("B,N,K", [(2, 1000, 5), (3, 4999, 49)])
def test_sparse_row_backward_equals_dense_backward(B, N, K):
    """gradient w.r.t. X of a loss that sees K rows of the last iterate: the sparse-row path (pn_ms_rows_bwd) must equal
    the dense tcgen05 backward (which spends N^2 work on rows that contribute zero) and the oracle's closed form"""
    from oracle.port import meanshift as oms
    from pnb200 import meanshift as pms
    d, its = 128, 4
    g = torch.Generator().manual_seed(B * N)
    X0 = torch.nn.functional.normalize(torch.randn(B, N, d, generator=g), dim=2)
    bws = torch.tensor([0.3, 0.5, 0.8])[:B]
    ids = [torch.randperm(N, generator=g)[:K].sort()[0] for _ in range(B)]
    w = [torch.randn(K, d, generator=g) for _ in range(B)]
    # dense
    Xd = X0.clone().cuda().requires_grad_()
    Y = pms.mean_shift_iters(Xd, bws.cuda(), its)
    sum((Y[b][ids[b].cuda()] * w[b].cuda()).sum() for b in range(B)).backward()
    # sparse
    Xs = X0.clone().cuda().requires_grad_()
    Yk, state = pms.mean_shift_iters_keep(Xs, bws.cuda(), its)
    centers = pms.centers_sparse(Xs, state, [i.cuda() for i in ids])
    sum((centers[b] * w[b].cuda()).sum() for b in range(B)).backward()
    assert torch.equal(Yk, Y.detach())
    _close(Xs.grad, Xd.grad, 1e-4, "sparse vs dense d loss / d X")
    # oracle closed form for shape 0
    Ys = [t[0].cpu() for t in state[1]]
    gX = torch.zeros(N, d); gg = w[0].clone()
    for t in range(its, 0, -1):
        gg, gx = oms.sparse_rows_backward(gg, Ys[t][ids[0]], Ys[t - 1][ids[0]], state[2][t - 1][0].cpu()[ids[0]],
                                          state[3][t - 1][0].cpu()[ids[0]], X0[0], float(bws[0]))
        gX += gx
    gX[ids[0]] += gg
    _close(Xs.grad[0], gX, 1e-3, "sparse path vs oracle closed form")


def test_guard_mean_shift_retry_loops_vs_reference(golden_dir):
    """the retry loops (more than 49 clusters -> larger quantile) of Evaluation.guard_mean_shift and
    MeanShift.guard_mean_shift against the unmodified reference's run (tests/golden/guard.npz: three attempts each)"""
    import os
    from oracle.make_golden_helpers import clustered_embedding
    from src.mean_shift import MeanShift
    from src.residual_utils import Evaluation
    g = np.load(os.path.join(golden_dir, "guard.npz"))
    X, _ = clustered_embedding(1500, 128, 60, 11, spread=0.05)

    def canon(l):
        l = np.asarray(l)
        _, first = np.unique(l, return_index=True)
        m = {int(v): i for i, v in enumerate(l[np.sort(first)])}
        return np.array([m[int(v)] for v in l])
    ev = Evaluation.__new__(Evaluation)
    ev.ms = MeanShift()
    for prefix, fn in (("ev", ev.guard_mean_shift), ("ms", MeanShift().guard_mean_shift)):
        np.random.seed(int(g["seed"]))
        with torch.no_grad():
            c, bw, lab = fn(X.cuda(), float(g["quantile"]), int(g["iterations"]), kernel_type="gaussian")
        assert abs(float(bw) - float(g[prefix + "_bw"])) <= 1e-5 * float(g[prefix + "_bw"]), prefix
        np.testing.assert_array_equal(canon(lab.cpu().numpy()), canon(g[prefix + "_labels"]))
