"""Segmentation-network parity on the GPU: CUDA path (through the C-ABI) vs the oracle port and the golden vectors.
Tolerance: 1e-4 relative (north_star, fp32) on embedding / log-prob / losses; gradients 2e-3 of the tensor scale."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-4, atol=None, name=""):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, np.float64)
    scale = np.abs(want).max() + 1e-30
    if atol is None:
        atol = rtol * scale
    err = np.abs(got - want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert (err <= atol + rtol * np.abs(want)).all(), f"{name}: max err {err.max():.3e} scale {scale:.3e}"


# ------------------------------------------------------------------------------------------------ linear kernels
@pytest.mark.parametrize("B,Np,K,Nout,G", [(2, 300, 256, 512, 8), (3, 130, 6, 128, 0), (1, 1000, 512, 10, 0),
                                           (2, 257, 64, 256, 4)])
def test_linear_fwd_bwd(B, Np, K, Nout, G):
    from pnb200 import ops
    g = torch.Generator().manual_seed(K + Nout)
    A = torch.randn(B, Np, K, generator=g)
    W = torch.randn(Nout, K, generator=g) / K ** 0.5
    bias = torch.randn(Nout, generator=g)
    sb = torch.randn(B, Nout, generator=g)
    sc = torch.randn(B, K, generator=g); sh = torch.randn(B, K, generator=g) * 0.3
    Ad, Wd = A.cuda(), W.cuda()
    nin = ops.Norm(None, sc.cuda(), sh.cuda(), ops.ACT_RELU, 1, 1.0, True, None, None)
    Y, st = ops.linear_fwd(Ad, Wd, bias=bias.cuda(), sbias=sb.cuda(), in_norm=nin, stats_groups=G)
    Aact = F.relu(A * sc[:, None] + sh[:, None])
    Yw = Aact @ W.T + bias + sb[:, None]
    _close(Y, Yw, name="Y")
    if G:
        yg = Yw.view(B, Np, G, Nout // G).double()
        _close(st[:, :, 0], yg.sum((1, 3)), rtol=1e-5, name="sum")
        _close(st[:, :, 1], (yg ** 2).sum((1, 3)), rtol=1e-5, name="sumsq")
    dY = torch.randn(B, Np, Nout, generator=g)
    dW, db, dsb = ops.linear_bwd_weight(dY.cuda(), Ad, in_norm=nin, want_sbias=True)
    _close(dW, torch.einsum("bno,bnk->ok", dY, Aact), rtol=2e-4, name="dW")
    _close(db, dY.sum((0, 1)), rtol=2e-4, name="db")
    _close(dsb, dY.sum(1), rtol=2e-4, name="dsb")
    dZ, _ = ops.linear_bwd_data(dY.cuda(), Wd)
    _close(dZ, dY @ W, name="dZ")
    # accumulate + strided destination
    buf = torch.randn(B, Np, K + 64, generator=g).cuda()
    want = buf[:, :, 32:32 + K].cpu() + dY @ W
    ops.linear_bwd_data(dY.cuda(), Wd, dZ=buf[:, :, 32:32 + K], accumulate=True)
    _close(buf[:, :, 32:32 + K], want, name="dZ accumulate")


def test_linear_norm_chain_backward_matches_autograd():
    """linear -> GroupNorm -> ReLU -> linear, gradients of all leaves vs torch autograd (CPU)."""
    from pnb200 import ops
    B, Np, K, C, Nout, G = 2, 333, 64, 256, 32, 4
    g = torch.Generator().manual_seed(5)
    A = torch.randn(B, Np, K, generator=g)
    W1 = (torch.randn(C, K, generator=g) / 8).requires_grad_()
    b1 = torch.randn(C, generator=g).requires_grad_()
    ga = (torch.randn(C, generator=g) * 0.5 + 0.2).requires_grad_()
    be = (torch.randn(C, generator=g) * 0.1).requires_grad_()
    W2 = (torch.randn(Nout, C, generator=g) / 16).requires_grad_()
    Ar = A.clone().requires_grad_()
    Y1 = Ar @ W1.T + b1
    H = F.relu(F.group_norm(Y1.permute(0, 2, 1), G, ga, be)).permute(0, 2, 1)
    Y2 = H @ W2.T
    dY2 = torch.randn(B, Np, Nout, generator=g)
    (Y2 * dY2).sum().backward()
    # CUDA
    Ad = A.cuda()
    y1, st = ops.linear_fwd(Ad, W1.detach().cuda(), bias=b1.detach().cuda(), stats_groups=G)
    n1 = ops.norm_finalize(st, ga.detach().cuda(), be.detach().cuda(), B, C, (C // G) * Np, ops.ACT_RELU)
    y2, _ = ops.linear_fwd(y1, W2.detach().cuda(), in_norm=n1)
    _close(y1, Y1, name="y1"); _close(y2, Y2, name="y2")
    d2 = dY2.cuda()
    dW2, _, _ = ops.linear_bwd_weight(d2, y1, in_norm=n1, want_bias=False)
    dZ, gs = ops.linear_bwd_data(d2, W2.detach().cuda(), fin_A=y1, fin_norm=n1)
    dga, dbe = ops.norm_bwd_apply(dZ, y1, n1, gs)
    dW1, db1, _ = ops.linear_bwd_weight(dZ, Ad)
    dA, _ = ops.linear_bwd_data(dZ, W1.detach().cuda())
    _close(dW2, W2.grad, rtol=3e-4, name="dW2")
    _close(dga, ga.grad, rtol=3e-4, name="dgamma")
    _close(dbe, be.grad, rtol=3e-4, name="dbeta")
    _close(dW1, W1.grad, rtol=3e-4, name="dW1")
    _close(db1, b1.grad, rtol=3e-4, name="db1")
    _close(dA, Ar.grad, rtol=3e-4, name="dA")


# ------------------------------------------------------------------------------------------------ one edge-conv layer
@pytest.mark.parametrize("C,Cout,k", [(6, 64, 20), (64, 128, 16)])
def test_edge_conv_layer_forward_backward(C, Cout, k):
    from oracle.port import segnet as port
    from pnb200 import ops
    from pnb200.segnet import _wcat, _wcat_grad
    B, N = 2, 400
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, N, generator=g) * 0.5
    W = (torch.randn(Cout, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5).requires_grad_()
    ga = (torch.randn(Cout, generator=g) * 0.5 + 0.2).requires_grad_()
    be = (torch.randn(Cout, generator=g) * 0.1).requires_grad_()
    xr = x.clone().requires_grad_()
    idx = port.knn_idx(x, k, 0)
    out = port.edge_conv_gn(xr, idx, W, ga, be, 2)                        # (B,Cout,N)
    gout = torch.randn(B, N, Cout, generator=g)
    (out.permute(0, 2, 1) * gout).sum().backward()
    # CUDA
    xd = x.permute(0, 2, 1).contiguous().cuda()
    idxd = ops.knn_graph(xd, k, 0)
    assert (idxd.cpu().long() == idx).all()
    wc = _wcat(W.detach().reshape(Cout, 2 * C).cuda())
    PQ, _ = ops.linear_fwd(xd, wc)
    esel, jsel, esum, st = ops.edge_gather_fwd(PQ, idxd, ga.detach().cuda(), 2)
    nrm = ops.norm_finalize(st, ga.detach().cuda(), be.detach().cuda(), B, Cout, (Cout // 2) * N * k, ops.ACT_LRELU)
    o = torch.empty(B, N, Cout, device="cuda")
    ops.edge_apply(esel, nrm, o)
    _close(o, out.permute(0, 2, 1), name="edge out")
    dPQ, dga, dbe = ops.edge_bwd(gout.cuda(), PQ, idxd, esel, jsel, esum, nrm)
    dwc, _, _ = ops.linear_bwd_weight(dPQ, xd, want_bias=False)
    dW = _wcat_grad(dwc, Cout)
    dx, _ = ops.linear_bwd_data(dPQ, wc)
    _close(dga, ga.grad, rtol=5e-4, name="dgamma")
    _close(dbe, be.grad, rtol=5e-4, name="dbeta")
    _close(dW, W.grad.reshape(Cout, 2 * C), rtol=5e-4, name="dW")
    _close(dx, xr.grad.permute(0, 2, 1), rtol=5e-4, name="dx")


# ------------------------------------------------------------------------------------------------ whole network
def _build(golden_dir, k):
    from oracle.port import common
    from src.PointNet import PrimitivesEmbeddingDGCNGn
    from src.segment_loss import EmbeddingLoss
    loss = EmbeddingLoss(margin=1.0)
    m = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10,
                                  loss_function=loss.triplet_loss, mode=5, num_channels=6, nn_nb=k)
    return m, common


def test_segnet_state_dict_keys_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "segnet.npz"))
    m, _ = _build(golden_dir, 20)
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(g["state_keys"])
    for kname, s in zip(g["state_keys"], g["state_shapes"]):
        assert tuple(sd[kname].shape) == eval(s), kname


def test_segnet_vs_golden_and_port(golden_dir):
    from oracle.port import segnet as port
    from src.segment_loss import primitive_loss
    g = np.load(os.path.join(golden_dir, "segnet.npz"))
    B, N, k, wseed, rseed = [int(v) for v in g["meta"]]
    m, common = _build(golden_dir, k)
    shapes = {n: tuple(v.shape) for n, v in m.state_dict().items()}
    sd = common.seeded_state_dict(shapes, seed=wseed)
    for i in (1, 2, 3):
        for s in ("weight", "bias"):
            sd[f"encoder.conv{i}.1.{s}"] = sd[f"encoder.bn{i}.{s}"]
    m.load_state_dict(sd)
    m.cuda()
    x = torch.from_numpy(g["points"]).cuda()
    idxs = [torch.from_numpy(g[f"idx{i}"].astype(np.int64)).cuda() for i in (1, 2, 3)]
    np.random.seed(rseed)
    emb, lp, el = m(x, torch.from_numpy(g["labels"]).cuda(), True, idx_override=idxs)
    _close(emb, g["embedding"], name="embedding vs reference")
    _close(lp, g["logprob"], name="logprob vs reference")
    _close(el, g["embed_loss"], name="embed_loss vs reference")
    nll = primitive_loss(lp, torch.from_numpy(g["prims"]).cuda())
    _close(nll, g["nll"], name="nll vs reference")
    (el.sum() + nll).backward()
    checked = 0
    for key in g.files:
        if not key.startswith("grad:") or (key.startswith("grad:encoder.conv") and ".1." in key):
            continue
        p = dict(m.named_parameters())[key[5:]]
        t = p.grad.detach().cpu().reshape(-1).double()
        got = np.array([t.sum().item(), t.norm().item()] + t[:14].tolist())
        want = g[key]
        assert abs(got[1] - want[1]) <= 2e-3 * want[1] + 1e-7, (key, got[1], want[1])
        assert np.abs(got[2:] - want[2:]).max() <= 3e-3 * (np.abs(want[2:]).max() + want[1] / np.sqrt(t.numel())), key
        checked += 1
    assert checked >= 30
    # free-running graph: kNN on our own features
    with torch.no_grad():
        emb2, lp2, _ = m(x, torch.from_numpy(g["labels"]).cuda(), False)
    err = (emb2.cpu().numpy() - g["embedding"])
    assert (np.abs(err) < 1e-3 * (np.abs(g["embedding"]) + 1e-2)).mean() > 0.98


def test_segnet_fresh_inputs_full_gradients_vs_port():
    """fresh seeded inputs (not the golden ones), every parameter gradient compared elementwise with the port."""
    from oracle.port import common, segnet as port
    from src.PointNet import PrimitivesEmbeddingDGCNGn
    B, N, k = 2, 500, 24
    pts, nrm, lab, prim = common.synth_cloud(B, N, seed=21)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous()
    m = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10,
                                  loss_function=None, mode=5, num_channels=6, nn_nb=k)
    shapes = {n: tuple(v.shape) for n, v in m.state_dict().items()}
    sd = common.seeded_state_dict(shapes, seed=3)
    for i in (1, 2, 3):
        for s in ("weight", "bias"):
            sd[f"encoder.conv{i}.1.{s}"] = sd[f"encoder.bn{i}.{s}"]
    m.load_state_dict(sd); m.cuda()
    sdr = {n: v.clone().requires_grad_(v.is_floating_point()) for n, v in sd.items()}
    emb_r, lp_r, idxs, x4_r, xf_r = port.segnet_fwd(sdr, x, k, 5)
    g = torch.Generator().manual_seed(9)
    ge = torch.randn(emb_r.shape, generator=g); gl = torch.randn(lp_r.shape, generator=g)
    ((emb_r * ge).sum() + (lp_r * gl).sum()).backward()
    emb, lp, _ = m(x.cuda(), None, False, idx_override=[i.cuda() for i in idxs])
    _close(emb, emb_r, name="embedding"); _close(lp, lp_r, name="logprob")
    ((emb * ge.cuda()).sum() + (lp * gl.cuda()).sum()).backward()
    for n, p in m.named_parameters():
        if n.startswith("encoder.bn4") or n.startswith("encoder.bn5"):
            continue
        ref_name = n
        gr = sdr[ref_name].grad
        if n.startswith("encoder.bn") and n[len("encoder.bn")] in "123":
            alias = sdr["encoder.conv%s.1.%s" % (n[len("encoder.bn")], n.split(".")[-1])].grad
            gr = gr if alias is None else (alias if gr is None else gr + alias)
        assert gr is not None and p.grad is not None, n
        _close(p.grad, gr, rtol=1e-3, name="grad " + n)
