"""CPU tests (no GPU) of the host-side rewrites of the fitting pipeline: each must reproduce the formulation it
replaced — the reference's torch one-hot IoU cost (src/segment_utils.py:356-374 + fitting_utils.py:362-376), its per-pair
boolean IoU loop (segment_utils.py:66-112), scipy.stats.mode (residual_utils.py:187), np.random draw order of
EmbeddingLoss.triplet_loss (segment_loss.py:60-96) — bit for bit; and the 3x3 solvers of csrc/small3.cuh (compiled
here with g++ from the very header the device kernels include) must agree with LAPACK."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ matching / IoU
def _one_hot(t, K=50):
    o = torch.zeros(t.shape[0], K)
    return o.scatter_(1, torch.from_numpy(t).long().unsqueeze(1), 1)


@pytest.mark.parametrize("trial", range(6))
def test_iou_cost_host_equals_one_hot_torch_formulation(trial):
    from src.segment_utils import iou_cost_host, relaxed_iou_fast
    rs = np.random.RandomState(trial)
    N = 10000
    tgt = rs.randint(0, rs.randint(1, 50), N)
    pred = (tgt + rs.randint(0, 2, N)) % 50 if trial % 2 else rs.randint(0, rs.randint(1, 50), N)
    ref = 1.0 - relaxed_iou_fast(_one_hot(pred).unsqueeze(0), _one_hot(tgt).unsqueeze(0)).numpy()[0]
    mine = iou_cost_host(pred, tgt)
    assert mine.dtype == np.float32 and np.array_equal(ref, mine)
    with pytest.raises(ValueError):
        iou_cost_host(np.array([0, 50]), np.array([0, 1]))


def _mean_iou_reference_style(matching, predicted_labels, labels, pred_prim, gt_prim):
    ious, prim_ious, pairs = [], [], []
    for b in range(labels.shape[0]):
        iou_b, prim_b, pairs = [], [], []
        for r, c in zip(*matching[b]):
            pi, gi = predicted_labels[b] == r, labels[b] == c
            if gi.sum() == 0 or pi.sum() == 0 or gi.sum() < 100:
                continue
            iou_b.append(np.logical_and(pi, gi).sum() / (np.logical_or(pi, gi).sum() + 1e-8))
            g_t, p_t = gt_prim[b][gi][0], pred_prim[b][r]
            prim_b.append(g_t == p_t)
            pairs.append([g_t, p_t])
        ious.append(np.mean(iou_b))
        prim_ious.append(np.mean(prim_b))
    return np.mean(ious), np.mean(prim_ious), pairs


@pytest.mark.parametrize("trial", range(6))
def test_segment_iou_metrics_equal_boolean_mask_loop(trial):
    from src.segment_utils import (SIOU_matched_segments, iou_cost_host, mean_IOU_primitive_segment, solve_dense)
    rs = np.random.RandomState(100 + trial)
    N, kg, kp = 10000, rs.randint(2, 20), rs.randint(1, 10)
    labels = rs.randint(0, kg, N)
    clusters = labels % kp if trial % 2 else rs.randint(0, kp, N)
    prims = (labels * 3) % 10                                  # one primitive type per gt segment
    seg_types = rs.randint(0, 10, kp)
    matching = [list(solve_dense(iou_cost_host(clusters, labels)))]
    a = _mean_iou_reference_style(matching, clusters[None], labels[None], seg_types[None], prims[None])
    b = mean_IOU_primitive_segment(matching, clusters[None], labels[None], seg_types[None], prims[None])
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    # the full metric with / without a precomputed matching, incl. the in-place type merge (0,6,7 -> 9; 8 -> 2)
    p1, p2 = prims.copy(), prims.copy()
    m1 = SIOU_matched_segments(labels, clusters, None, p1, None, prim_pred_seg=seg_types)
    m2 = SIOU_matched_segments(labels, clusters, None, p2, None, prim_pred_seg=seg_types, matching=matching[0])
    assert m1[0] == m2[0] and m1[1] == m2[1]
    q = prims.copy()
    for s_, d_ in ((0, 9), (6, 9), (7, 9), (8, 2)):
        q[q == s_] = d_
    assert np.array_equal(q, p1) and np.array_equal(q, p2)


def test_bincount_mode_equals_scipy_mode():
    rs = np.random.RandomState(3)
    for _ in range(50):
        y = rs.randint(0, 10, rs.randint(1, 400))
        assert int(stats.mode(y)[0]) == int(np.bincount(y).argmax())


def test_segment_types_device_formulation_on_cpu_tensors():
    from src.segment_utils import segment_types_device
    rs = np.random.RandomState(4)
    N = 5000
    pp = rs.randint(0, 10, N)
    W = torch.randn(7, N)
    merged = pp.copy()
    for s_, d_ in ((0, 9), (6, 9), (7, 9), (8, 2)):
        merged[merged == s_] = d_
    hot = torch.zeros(N, 10).scatter_(1, torch.from_numpy(merged).long().unsqueeze(1), 1)
    assert torch.equal(segment_types_device(torch.from_numpy(pp), W), torch.max(hot.t() @ W.t(), 0)[1])


# ------------------------------------------------------------------------------------------------ triplet sampling
def _triplet_sample_reference_style(labels, N, rng=np.random, max_segments=5):
    """segment_loss.py:60-96 as written: rng.choice(list(np.where(np.isin(p, l))[0]), S), rng.choice(L, 1)[0]"""
    samples, pairs_all = [], []
    for i in range(labels.shape[0]):
        p = labels[i]
        uniq = np.unique(p)
        S = min(N // uniq.shape[0] + 1, 30)
        samples.append({l: rng.choice(list(np.where(np.isin(p, l))[0]), S, replace=True) for l in uniq})
    for i in range(labels.shape[0]):
        keys = sorted(samples[i].keys())
        L = len(keys)
        if L == 1:
            continue
        for _ in range(min(max_segments * max_segments, L * L)):
            k1 = rng.choice(L, 1)[0]
            k2 = rng.choice(L, 1)[0]
            if k1 == k2:
                continue
            pairs_all.append((samples[i][keys[k1]] + i * N, samples[i][keys[k2]] + i * N))
    return pairs_all


def test_triplet_sampling_consumes_identical_np_random_draws():
    from pnb200.losses import triplet_sample
    rs = np.random.RandomState(5)
    labels = rs.randint(0, 9, size=(6, 4000))
    labels[3] = 2                                              # a shape with a single segment
    np.random.seed(11)
    ref = _triplet_sample_reference_style(labels, 4000)
    after_ref = np.random.rand()
    np.random.seed(11)
    groups = triplet_sample(labels, 4000)
    after_new = np.random.rand()
    assert after_ref == after_new                              # same number of draws consumed
    A = np.concatenate([g[1] for g in groups])
    Nn = np.concatenate([g[2] for g in groups])
    assert np.array_equal(A, np.stack([r[0] for r in ref])) and np.array_equal(Nn, np.stack([r[1] for r in ref]))


# ------------------------------------------------------------------------------------------------ 3x3 solvers
@pytest.fixture(scope="module")
def small3(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("small3") / "libsmall3.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "parsenet-codebase_b200", "csrc"),
                           "-o", out, os.path.join(ROOT, "tests", "c", "small3_host.cpp")])
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def test_small3_eigh_matches_lapack(small3):
    rs = np.random.RandomState(0)
    S = 500
    A = rs.randn(S, 40, 3) * rs.rand(S, 1, 3) * np.array([1, 1e-3, 1])[None, None]
    G = np.ascontiguousarray(np.einsum("smi,smj->sij", A, A))
    G[0] = np.eye(3); G[1] = np.diag([1, 1, 0.0]); G[2] = 0
    v = rs.randn(3); G[3] = np.outer(v, v)                     # rank 1
    w, V = np.zeros((S, 3)), np.zeros((S, 3, 3))
    small3.small3_eigh_host(_p(G), S, _p(w), _p(V))
    wr = np.linalg.eigvalsh(G)
    scale = np.maximum(np.abs(wr).max(1, keepdims=True), 1e-300)
    assert (np.abs(w - wr) / scale).max() < 1e-13
    rec = np.einsum("sik,sk,sjk->sij", V, w, V)
    assert (np.abs(rec - G).max((1, 2)) / scale[:, 0]).max() < 1e-13
    assert np.abs(np.einsum("sik,sjk->sij", V, V) - np.eye(3)).max() < 1e-13


def test_small3_lstsq_matches_reference_rank_rule(small3):
    """lambda of the regularised branch (1e-6 * 10^j, first j making AtA + lambda I full rank) and the solution"""
    EPS = float(np.finfo(np.float32).eps)
    rs = np.random.RandomState(1)
    S, rows = 400, 50
    A = rs.randn(S, rows, 3) * rs.rand(S, 1, 3) * np.array([1, 1e-4, 1])[None, None]
    A[:20, :, 2] = A[:20, :, 0]                                # exactly rank deficient
    G = np.ascontiguousarray(np.einsum("smi,smj->sij", A, A))
    Y = np.ascontiguousarray(rs.randn(S, 3))
    x, mi, lam = np.zeros((S, 3)), np.zeros((S, 3, 3)), np.zeros(S)
    small3.small3_lstsq_host(_p(G), _p(Y), S, rows, ctypes.c_double(EPS), _p(x), _p(mi), _p(lam))
    ev = np.linalg.eigvalsh(G)[:, ::-1]
    s = np.sqrt(np.clip(ev, 0, None))
    bad = s[:, 2] <= s[:, 0] * max(rows, 3) * EPS
    lam_ref = np.zeros(S)
    for i in np.nonzero(bad)[0]:
        cur = 1e-6
        for _ in range(7):
            if ev[i, 2] + cur > (ev[i, 0] + cur) * 3 * EPS:
                break
            cur *= 10
        lam_ref[i] = cur
    assert bad.sum() >= 20 and np.array_equal(lam, lam_ref)
    xr = np.linalg.solve(G + lam_ref[:, None, None] * np.eye(3), Y[:, :, None])[:, :, 0]
    assert (np.abs(x - xr).max(1) / np.abs(xr).max(1)).max() < 1e-6


def test_gram_svd_backward_is_finite_on_rank_deficient_gram_matrices():
    """regression: exact-zero singular values (null directions, e.g. the normals of a planar segment) used to give
    inf * 0 = NaN in the custom SVD backward; it surfaced as NaN weights after ~10 steps of the 2-GPU bench"""
    from pnb200.fitting import GramSVDFn, floor_singular_values

    class Ctx:
        pass

    torch.manual_seed(0)
    for evals in ([4.0, -1e-20, -3e-21], [0.0, 0.0, 0.0], [2.0, 2.0, 0.0], [2.0, 1.0, 1e-18], [3.0, 2.0, 1.0]):
        ev = torch.tensor([evals], dtype=torch.float64)
        sv = floor_singular_values(torch.sqrt(torch.clamp(ev, min=0.0)))
        assert (sv > 0).all()
        ctx = Ctx()
        ctx.saved_tensors = (torch.linalg.qr(torch.randn(1, 3, 3, dtype=torch.float64))[0], sv)
        g = GramSVDFn.backward(ctx, torch.randn(1, 3, 3, dtype=torch.float64), None)
        assert torch.isfinite(g).all(), evals
    # well-separated spectrum: unchanged by the floor, and the gradient matches the closed form of the reference
    sv = torch.tensor([[3.0, 2.0, 1.0]], dtype=torch.float64)
    assert torch.equal(floor_singular_values(sv), sv)


def test_ms_bwd_lite_rounding_model_stays_within_gradient_tolerance():
    """float64 model of PN_MS_BWD_LITE (csrc/meanshift_tc_bwd.cu): the gradient-side operand of the second tile product
    (gS for dY, [gS^T | K^T] for dX) rounded to tf32 (round to nearest), the streamed operand exact.  The error of the
    gradients of one mean-shift iteration stays below the 1e-3 gradient tolerance; rounding BOTH operands (a single
    tf32 MMA) would not keep a 2x margin, which is why that variant does not exist."""
    import torch

    def rna(x):
        i = x.float().contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1fff).view(torch.float32).double()

    def trunc(x):
        return (x.float().contiguous().view(torch.int32) & ~0x1fff).view(torch.float32).double()

    g0 = torch.Generator().manual_seed(1)
    N = 300
    X = torch.nn.functional.normalize(torch.randn(N, 128, generator=g0), dim=1).double()
    Y = torch.nn.functional.normalize(X + 0.05 * torch.randn(N, 128, generator=g0).double(), dim=1)
    g = torch.randn(N, 128, generator=g0).double()
    worst_lite, worst_single = 0.0, 0.0
    for bw in (0.8, 0.5, 1.2):
        c = 1 / bw ** 2
        K = torch.exp(torch.clamp((Y @ X.t() - 1.0) * c, -75, 75))
        den = K.sum(1)
        M = (K @ X) / den[:, None]
        nr = M.norm(dim=1)
        Yn = M / nr[:, None]
        gu = (g - Yn * (g * Yn).sum(1, keepdim=True)) / nr[:, None]
        Gn = gu / den[:, None]
        gd = -((gu * Yn).sum(1) * nr) / den
        gS = (Gn @ X.t() + gd[:, None]) * K * c
        exact = (gS @ X, gS.t() @ Y + K.t() @ Gn)
        lite = (rna(gS) @ X, rna(gS).t() @ Y + rna(K).t() @ Gn)
        single = (rna(gS) @ trunc(X), rna(gS).t() @ trunc(Y) + rna(K).t() @ trunc(Gn))
        for e, l, s in zip(exact, lite, single):
            worst_lite = max(worst_lite, ((e - l).abs().max() / e.abs().max()).item())
            worst_single = max(worst_single, ((e - s).abs().max() / e.abs().max()).item())
    assert 1e-6 < worst_lite < 5e-4, worst_lite
    assert worst_single > worst_lite


@pytest.mark.parametrize("sqrt,reduce", [(True, True), (False, False), (True, False)])
def test_eval_time_residual_variants_match_the_oracle_port(sqrt, reduce):
    """ComputePrimitiveDistance with sqrt=True / reduce=False (evaluation-time flags, reference src/primitives.py:100-195):
    plain torch closed forms, so they can be checked against the oracle port right here on CPU tensors"""
    from oracle.port import fitting as OP
    from src.primitives import ComputePrimitiveDistance
    g = torch.Generator().manual_seed(3)
    q = torch.randn(257, 3, generator=g) * 0.5
    unit = lambda v: v / v.norm()
    params = {"plane": [unit(torch.randn(3, 1, generator=g)), torch.tensor(0.13)],
              "sphere": [torch.randn(1, 3, generator=g) * 0.2, torch.tensor(0.6)],
              "cylinder": [unit(torch.randn(3, 1, generator=g)), torch.randn(1, 3, generator=g) * 0.2, torch.tensor(0.35)],
              "cone": [torch.randn(1, 3, generator=g) * 0.2, unit(torch.randn(3, 1, generator=g)), torch.tensor(0.5)]}
    cp = ComputePrimitiveDistance(reduce=reduce)
    for kind, ps in params.items():
        got = getattr(cp, "distance_from_" + kind)(points=q, params=ps, sqrt=sqrt)
        want = OP.DISTANCES[kind](q, ps, sqrt=sqrt, reduce=reduce)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-9), kind


def test_input_pipeline_vs_reference_dataset(golden_dir):
    """host rotations (the reference's own 3x3 arithmetic) + the batched rotate / extent / scale expressions of
    pnb200.input_pipeline (run on CPU tensors here, on the device in tests/test_gpu_fitstage.py) against Dataset.get_train of
    the unmodified reference (tests/golden/pipeline.npz): aligned points and rotated normals to 2e-6"""
    import os
    from pnb200 import input_pipeline as IP
    g = np.load(os.path.join(golden_dir, "pipeline.npz"))
    for name, noise, aniso in (("plain", False, False), ("noise_aniso", True, True)):
        pts, nrm = g["pts"].copy(), g["nrm"].copy()
        np.random.seed(11)
        if noise:
            pts = pts + nrm * IP.normal_noise(pts.shape[1])
        R = IP.host_rotations(pts)
        p, n = IP.preprocess_on_device(torch.from_numpy(pts), torch.from_numpy(nrm), torch.from_numpy(R), anisotropic=aniso)
        assert np.abs(p.numpy() - g[name + "_p"]).max() <= 2e-6, name
        assert np.abs(n.numpy() - g[name + "_n"]).max() <= 2e-6, name


# ------------------------------------------------------------------------------------------------ SURVEY 8f-1 (host parts)
def test_kronecker_optimiser_host_parts_vs_reference(golden_dir):
    """tests/golden/kronecker.npz (unmodified reference: DrawSurfs parameterisations, BSpline.basis_functions per sample,
    fit_bezier_surface_fit_kronecker): the oracle port reproduces all of it, and so do the product's vectorised
    parameterisations / per-sample basis rows (local-support recurrence instead of one Cox-de Boor triangle per entry)"""
    import os
    from scipy.interpolate import BSpline
    from oracle.port import optimize as PO
    from src import approximation as AP
    from src import primitive_forward as PF
    g = np.load(os.path.join(golden_dir, "kronecker.npz"))
    for grid, key in ((20, "bpar20"), (30, "bpar30")):
        assert np.array_equal(PO.boundary_parameterization(grid), g[key])
        assert np.array_equal(PF._boundary_parameterization(grid), g[key])
    assert np.array_equal(PO.regular_parameterization(30, 30), g["rpar"])
    assert np.abs(PF._regular_parameterization(30, 30) - g["rpar"]).max() == 0
    for deg in (2, 3):
        par = g[f"par{deg}"]
        nu_o, nv_o = PO.basis_rows(par, 10, 10, deg, deg)
        nu_p, nv_p = AP.basis_rows(par, 10, 10, deg, deg)
        for got in (nu_o, nu_p):
            assert np.abs(got - g[f"NU{deg}"]).max() < 1e-14
        for got in (nv_o, nv_p):
            assert np.abs(got - g[f"NV{deg}"]).max() < 1e-14
        _, _, ku, kv = AP.uniform_knot_bspline_(10, 10, deg, deg, 2)
        assert np.array_equal(np.array(ku), g[f"ku{deg}"]) and np.array_equal(np.array(kv), g[f"kv{deg}"])
        # independent check of the rows: scipy's design matrix over the same knots (away from the right end point)
        inner = par[:, 0] < 1
        dm = BSpline.design_matrix(par[inner, 0], np.array(ku), deg).toarray()
        assert np.abs(dm - nu_p[inner]).max() < 1e-14
        rec = PO.fit_bezier_surface_fit_kronecker(g[f"pts{deg}"], g[f"NU{deg}"], g[f"NV{deg}"])
        assert np.abs(rec - g[f"rec{deg}"]).max() < 1e-10
    for cu in (20, 21):
        nu_p, nv_p = AP.basis_rows(g["old_par"], cu, 20, 3, 3)
        assert np.abs(nu_p - g[f"old_NU{cu}"]).max() < 1e-14 and np.abs(nv_p - g[f"old_NV{cu}"]).max() < 1e-14
    # the port's surface evaluation is the tensor-product formula: check it against scipy's BSpline in both directions
    rs = np.random.RandomState(0)
    cp = rs.rand(20, 20, 3)
    par = rs.random_sample((40, 2))
    ku = np.array(AP.uniform_knot_bspline_(20, 20, 3, 3, 2)[2])
    want = np.stack([BSpline(ku, BSpline(ku, cp.transpose(1, 0, 2), 3)(v), 3)(u) for u, v in par])
    assert np.abs(PO.evaluate_list(cp, par, 3, 3) - want).max() < 1e-13


def test_hungarian_host_twin_matches_scipy(tmp_path):
    """csrc/assign.cuh (the Kuhn-Munkres template the device kernel instantiates with 32 lanes) run with one lane on the host
    against scipy.optimize.linear_sum_assignment: same optimal cost on random, heavily tied and IoU-like matrices; the same
    assignment where the optimum is unique"""
    so = str(tmp_path / "assign.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "parsenet-codebase_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "assign_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 17, 50, 64):
        for rep in range(12):
            if rep % 3 == 0:
                c = rng.random((n, n)).astype(np.float32)
            elif rep % 3 == 1:
                c = rng.integers(0, 4, (n, n)).astype(np.float32)
            else:
                c = np.ones((n, n), np.float32); k = min(n, 8); c[:k, :k] = 1 - rng.random((k, k)).astype(np.float32)
            out = np.zeros(n, np.int32)
            lib.hungarian_host(c.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
            assert sorted(out.tolist()) == list(range(n))
            r, cc = linear_sum_assignment(c)
            want = c[r, cc].astype(np.float64).sum()
            assert abs(c[np.arange(n), out].astype(np.float64).sum() - want) < 1e-9 * max(1, abs(want))
            if rep % 3 == 0:
                assert np.array_equal(out, cc)
