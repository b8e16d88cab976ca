"""SplineNet (DGCNNControlPoints) forward/backward schedule on the edge-conv / linear kernels.

Reference: DGCNNControlPoints.forward src/model.py:140-179 (4 edge-convs with BatchNorm2d + LeakyReLU, k nearest
neighbours in feature space, concat -> Conv1d+BN+LeakyReLU -> (x weights) -> max over points -> 3 FC -> tanh).
  training=True : batch statistics (BatchNorm train mode), gradients for every parameter (train_open_splines.py)
  training=False: running statistics (frozen net inside the end-to-end fit, residual_utils.py:62-66); the only
                  gradient that exists there is w.r.t. the per-point membership `weights` (points are detached,
                  fitting_optimization.py:138) and that is what backward returns.
"""
import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU
from .segnet import _wcat, _wcat_grad

BN_EPS = 1e-5


def _eval_norm(gamma, beta, rm, rv, B, act):
    rstd = torch.rsqrt(rv + BN_EPS)
    scale = gamma * rstd
    shift = beta - rm * scale
    C = gamma.shape[0]
    mr = torch.stack([rm, rstd], 1).view(1, C, 2).contiguous()
    return ops.Norm(mr, scale.view(1, C).expand(B, C).contiguous(), shift.view(1, C).expand(B, C).contiguous(), act,
                    C, 1.0, False, gamma, beta)


class SplineNetFn(torch.autograd.Function):
    """x0 (B,N,3), weights (B,N)|None -> control points (B, ncp*ncp*3) after tanh.
    params: [c1w,g1,b1, c2w,g2,b2, c3w,g3,b3, c4w,g4,b4, c5w,g5,b5, c6w,c6b,g6,b6, c7w,c7b,g7,b7, c8w,c8b]
    running: [(rm, rv)] x 7 (bn1..bn7), used when training=False; batch statistics are returned through
    `stats_out` (list filled with (mean, biased var, count) per BN) when training=True."""

    @staticmethod
    def forward(ctx, x0, weights, training, k, widths, running, stats_out, *params):
        P = [p.detach() for p in params]
        x0 = x0.detach().contiguous()
        B, N, _ = x0.shape
        dev = x0.device
        Ctot = sum(widths)
        xcat = torch.empty((B, N, Ctot), dtype=torch.float32, device=dev)
        layers = []
        cur = x0
        lo = 0
        for li in range(4):
            w, g, bt = P[3 * li], P[3 * li + 1], P[3 * li + 2]
            Cout = widths[li]
            wc = _wcat(w.reshape(Cout, -1))
            idx = ops.knn_graph(cur, k, 0)
            PQ, _ = ops.linear_fwd(cur, wc)
            if training:
                esel, jsel, esum, st = ops.edge_gather_fwd(PQ, idx, g, Cout, per_shape=False)
                norm = ops.norm_finalize(st, g, bt, B, Cout, B * N * k, ACT_LRELU, per_shape=False, eps=BN_EPS)
                if stats_out is not None:
                    stats_out.append((norm.mean_rstd[0, :, 0].clone(), st[0].clone(), float(B * N * k)))
            else:
                norm = _eval_norm(g, bt, running[li][0], running[li][1], B, ACT_LRELU)
                esel, jsel, esum, _ = ops.edge_gather_fwd(PQ, idx, norm.scale[0].contiguous(), Cout, per_shape=False,
                                                          want_stats=False)
            out = xcat[:, :, lo:lo + Cout]
            ops.edge_apply(esel, norm, out)
            layers.append(dict(inp=cur, wc=wc, idx=idx, PQ=PQ, esel=esel, jsel=jsel, esum=esum, norm=norm, lo=lo,
                               hi=lo + Cout))
            cur = out
            lo += Cout
        c5w, g5, b5 = P[12], P[13], P[14]
        W5 = c5w.reshape(1024, Ctot)
        if training:
            Y5, st5 = ops.linear_fwd(xcat, W5, stats_groups=1024, per_shape=False)
            n5 = ops.norm_finalize(st5, g5, b5, B, 1024, B * N, ACT_LRELU, per_shape=False, eps=BN_EPS)
            if stats_out is not None:
                stats_out.append((n5.mean_rstd[0, :, 0].clone(), st5[0].clone(), float(B * N)))
        else:
            Y5, _ = ops.linear_fwd(xcat, W5)
            n5 = _eval_norm(g5, b5, running[4][0], running[4][1], B, ACT_LRELU)
        wts = None if weights is None else weights.detach().reshape(B, N).contiguous().float()
        pooled, arg = ops.colmax_norm(Y5, n5, wts)                       # (B,1024)
        # ---- FC tail on (1, B, 1024): BatchNorm1d statistics run over the batch
        c6w, c6b, g6, b6, c7w, c7b, g7, b7, c8w, c8b = P[15:25]
        W6, W7, W8 = c6w.reshape(1024, 1024), c7w.reshape(1024, 1024), c8w.reshape(c8w.shape[0], 1024)
        pv = pooled.view(1, B, 1024)
        if training:
            Y6, s6 = ops.linear_fwd(pv, W6, bias=c6b, stats_groups=1024, per_shape=False)
            n6 = ops.norm_finalize(s6, g6, b6, 1, 1024, B, ACT_RELU, per_shape=False, eps=BN_EPS)
            Y7, s7 = ops.linear_fwd(Y6, W7, bias=c7b, in_norm=n6, stats_groups=1024, per_shape=False)
            n7 = ops.norm_finalize(s7, g7, b7, 1, 1024, B, ACT_RELU, per_shape=False, eps=BN_EPS)
            if stats_out is not None:
                stats_out.append((n6.mean_rstd[0, :, 0].clone(), s6[0].clone(), float(B)))
                stats_out.append((n7.mean_rstd[0, :, 0].clone(), s7[0].clone(), float(B)))
        else:
            Y6, _ = ops.linear_fwd(pv, W6, bias=c6b)
            n6 = _eval_norm(g6, b6, running[5][0], running[5][1], 1, ACT_RELU)
            Y7, _ = ops.linear_fwd(Y6, W7, bias=c7b, in_norm=n6)
            n7 = _eval_norm(g7, b7, running[6][0], running[6][1], 1, ACT_RELU)
        Y8, _ = ops.linear_fwd(Y7, W8, bias=c8b, in_norm=n7)
        out = torch.tanh(Y8.view(B, -1))
        ctx.t = dict(layers=layers, xcat=xcat, W5=W5, Y5=Y5, n5=n5, wts=wts, pooled=pooled, arg=arg, W6=W6, W7=W7, W8=W8,
                     Y6=Y6, Y7=Y7, n6=n6, n7=n7, out=out, training=training, k=k)
        ctx.shapes = [p.shape for p in params]
        ctx.has_w = weights is not None
        ctx.wshape = None if weights is None else weights.shape
        return out.view_as(out)             # (fresh view: see segnet.EncoderFn.forward)

    @staticmethod
    def backward(ctx, gout):
        t = ctx.t
        B, N, Ctot = t["xcat"].shape
        dev = gout.device
        S = ctx.shapes
        training = t["training"]
        grads = [None] * 25
        dY8 = (gout.contiguous() * (1 - t["out"] * t["out"])).view(1, B, -1).contiguous()
        Y5, n5, arg, wts = t["Y5"], t["n5"], t["arg"], t["wts"]
        if training:
            dW8, db8, _ = ops.linear_bwd_weight(dY8, t["Y7"], in_norm=t["n7"])
            grads[23], grads[24] = dW8.view(S[23]), db8
            dZ7, gs7 = ops.linear_bwd_data(dY8, t["W8"], fin_A=t["Y7"], fin_norm=t["n7"])
            grads[21], grads[22] = ops.norm_bwd_apply(dZ7, t["Y7"], t["n7"], gs7)
            dW7, db7, _ = ops.linear_bwd_weight(dZ7, t["Y6"], in_norm=t["n6"])
            grads[19], grads[20] = dW7.view(S[19]), db7
            dZ6, gs6 = ops.linear_bwd_data(dZ7, t["W7"], fin_A=t["Y6"], fin_norm=t["n6"])
            grads[17], grads[18] = ops.norm_bwd_apply(dZ6, t["Y6"], t["n6"], gs6)
            dW6, db6, _ = ops.linear_bwd_weight(dZ6, t["pooled"].view(1, B, 1024))
            grads[15], grads[16] = dW6.view(S[15]), db6
            dpool, _ = ops.linear_bwd_data(dZ6, t["W6"])
        else:
            # frozen affine norms: mask by the activation, scale by gamma*rstd
            dZ7, _ = ops.linear_bwd_data(dY8, t["W8"], fin_A=t["Y7"], fin_norm=None,
                                         fin_act=ACT_NONE)
            pre7 = t["Y7"] * t["n7"].scale.view(1, 1, -1) + t["n7"].shift.view(1, 1, -1)
            dZ7 = dZ7 * (pre7 > 0).float() * t["n7"].scale.view(1, 1, -1)
            dZ6, _ = ops.linear_bwd_data(dZ7.contiguous(), t["W7"])
            pre6 = t["Y6"] * t["n6"].scale.view(1, 1, -1) + t["n6"].shift.view(1, 1, -1)
            dZ6 = (dZ6 * (pre6 > 0).float() * t["n6"].scale.view(1, 1, -1)).contiguous()
            dpool, _ = ops.linear_bwd_data(dZ6, t["W6"])
        dpool = dpool.view(B, 1024)
        # ---- max over points (+ optional per-point weights) + LeakyReLU + BN5
        y_at = torch.gather(Y5, 1, arg.long().unsqueeze(1)).squeeze(1)               # (B,1024) pre-norm at arg
        pre = y_at * n5.scale + n5.shift
        actv = torch.where(pre > 0, pre, 0.2 * pre)
        gweights = None
        if ctx.has_w:
            w_at = torch.gather(wts, 1, arg.long())                                      # (B,1024)
            gw = torch.zeros((B, N), dtype=torch.float32, device=dev)
            gw.scatter_add_(1, arg.long(), dpool * actv)
            gweights = gw.view(ctx.wshape)
            gact = dpool * w_at
        else:
            gact = dpool
        if not training:
            return (None, gweights, None, None, None, None, None) + tuple(grads)
        gt = (gact * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, 0.2))).contiguous()
        mr = n5.mean_rstd                                                               # (1,1024,2)
        xh = (y_at - mr[0, :, 0].view(1, -1)) * mr[0, :, 1].view(1, -1)
        gg = gt * n5.gamma.view(1, -1)
        gsum = torch.stack([gg.sum(0), (gg * xh).sum(0)], 1).view(1, 1024, 2).double().contiguous()
        grads[13] = (gt * xh).sum(0)
        grads[14] = gt.sum(0)
        dY5 = ops.colmax_bwd_fill(Y5, gt, arg, n5, gsum)
        dW5, _, _ = ops.linear_bwd_weight(dY5, t["xcat"], want_bias=False)
        grads[12] = dW5.view(S[12])
        dxc, _ = ops.linear_bwd_data(dY5, t["W5"])
        del dY5
        layers = t["layers"]
        for li in (3, 2, 1, 0):
            L = layers[li]
            Cout = L["hi"] - L["lo"]
            g = dxc[:, :, L["lo"]:L["hi"]]
            dPQ, dg, db = ops.edge_bwd(g, L["PQ"], L["idx"], L["esel"], L["jsel"], L["esum"], L["norm"])
            dwc, _, _ = ops.linear_bwd_weight(dPQ, L["inp"], want_bias=False)
            grads[3 * li] = _wcat_grad(dwc, Cout).reshape(S[3 * li])
            grads[3 * li + 1], grads[3 * li + 2] = dg, db
            if li > 0:
                Pv = layers[li - 1]
                ops.linear_bwd_data(dPQ, L["wc"], dZ=dxc[:, :, Pv["lo"]:Pv["hi"]], accumulate=True)
        dx0 = None
        if ctx.needs_input_grad[0]:
            dx0, _ = ops.linear_bwd_data(dPQ, layers[0]["wc"])
        return (dx0, gweights, None, None, None, None, None) + tuple(grads)
