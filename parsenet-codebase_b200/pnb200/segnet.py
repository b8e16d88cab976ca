"""Segmentation network (DGCNN encoder + per-point head) as two autograd Functions over the sm_100a kernels.

Forward/backward schedules for DGCNNEncoderGn.forward (src/PointNet.py:172-220) and
PrimitivesEmbeddingDGCNGn.forward (src/PointNet.py:265-284).  Activations are point-major (B,N,C); only PRE-norm
tensors are kept (the consumer GEMM applies the producer's GroupNorm + ReLU while loading its operand).
"""
import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU


def _wcat(w2d):
    """Conv2d weight (Cout, 2C) acting on [x_j - x_i ; x_i]  ->  [W1 ; W2 - W1]  (2Cout, C) acting on x_j / x_i"""
    C = w2d.shape[1] // 2
    return torch.cat([w2d[:, :C], w2d[:, C:] - w2d[:, :C]], 0).contiguous()


def _wcat_grad(dwcat, Cout):
    dP, dQ = dwcat[:Cout], dwcat[Cout:]
    return torch.cat([dP - dQ, dQ], 1)


class EncoderFn(torch.autograd.Function):
    """x0 (B,N,Cin) -> x4 (B,1024), xf (B,N,256).  params: [w1,g1,b1, w2,g2,b2, w3,g3,b3, wm,bm,gm,btm]"""

    @staticmethod
    def forward(ctx, x0, k, mode, idx_override, *params):
        (w1, g1, b1, w2, g2, b2, w3, g3, b3, wm, bm, gm, btm) = [p.detach() for p in params]
        x0 = x0.detach().contiguous()
        B, N, Cin = x0.shape
        dev = x0.device
        xf = torch.empty((B, N, 256), dtype=torch.float32, device=dev)
        layers = []
        cur = x0
        specs = [(w1, g1, b1, 0, 64), (w2, g2, b2, 64, 128), (w3, g3, b3, 128, 256)]
        for li, (w, g, bt, lo, hi) in enumerate(specs):
            Cout = hi - lo
            w2d = w.reshape(Cout, -1)
            wc = _wcat(w2d)
            if idx_override is not None:
                idx = idx_override[li].to(torch.int32).contiguous()
            else:
                idx = ops.knn_graph(cur, k, 1 if (mode == 5 and li == 0) else 0)
            PQ, _ = ops.linear_fwd(cur, wc)
            esel, jsel, esum, stats = ops.edge_gather_fwd(PQ, idx, g, 2)
            norm = ops.norm_finalize(stats, g, bt, B, Cout, (Cout // 2) * N * idx.shape[2], ACT_LRELU)
            out = xf[:, :, lo:hi]
            ops.edge_apply(esel, norm, out)
            layers.append(dict(inp=cur, wc=wc, idx=idx, PQ=PQ, esel=esel, jsel=jsel, esum=esum, norm=norm, lo=lo,
                               hi=hi))
            cur = out
        wm2 = wm.reshape(1024, 256)
        Y1, st = ops.linear_fwd(xf, wm2, bias=bm, stats_groups=8)
        nm = ops.norm_finalize(st, gm, btm, B, 1024, 128 * N, ACT_RELU)
        x4, arg = ops.colmax_norm(Y1, nm)
        ext = torch.gather(Y1, 1, arg.long().unsqueeze(1)).squeeze(1)      # pre-norm value at the arg-max row
        ctx.layers, ctx.xf, ctx.Y1, ctx.nm, ctx.arg, ctx.ext, ctx.wm2 = layers, xf, Y1, nm, arg, ext, wm2
        ctx.shapes = [p.shape for p in params]
        ctx.idx_list = [l["idx"] for l in layers]
        # outputs are returned as fresh views: an output object stored in ctx would close a reference cycle
        # (ctx -> tensor -> grad_fn -> ctx) that only the cyclic GC frees, i.e. GBs of activations per step would pile up
        return x4, xf.view_as(xf)

    @staticmethod
    def backward(ctx, g4, gxf):
        layers, xf, Y1, nm = ctx.layers, ctx.xf, ctx.Y1, ctx.nm
        B, N, _ = xf.shape
        dev = xf.device
        dxf = gxf.contiguous().clone() if gxf is not None else torch.zeros_like(xf)
        grads = [None] * 13
        if g4 is not None:
            g4 = g4.contiguous()
            # ---- mlp1 + GroupNorm(8) + ReLU + max over N   (tiny (B,1024) glue in torch, dense part in kernels)
            pre = nm.scale * ctx.ext + nm.shift
            gt = g4 * (pre > 0).to(g4.dtype)
            mr = nm.mean_rstd                                  # (B,8,2)
            mean_c = mr[:, :, 0].repeat_interleave(128, 1)
            rstd_c = mr[:, :, 1].repeat_interleave(128, 1)
            xh = (ctx.ext - mean_c) * rstd_c
            gg = gt * nm.gamma[None]
            gsum = torch.stack([gg.view(B, 8, 128).sum(2), (gg * xh).view(B, 8, 128).sum(2)], 2).double().contiguous()
            grads[11] = (gt * xh).sum(0)
            grads[12] = gt.sum(0)
            dY1 = ops.colmax_bwd_fill(Y1, gt.contiguous(), ctx.arg, nm, gsum)
            dWm, dbm, _ = ops.linear_bwd_weight(dY1, xf)
            grads[9] = dWm.view(ctx.shapes[9])
            grads[10] = dbm
            ops.linear_bwd_data(dY1, ctx.wm2, dZ=dxf, accumulate=True)
            del dY1
        for li in (2, 1, 0):
            L = layers[li]
            Cout = L["hi"] - L["lo"]
            g = dxf[:, :, L["lo"]:L["hi"]]
            dPQ, dg, db = ops.edge_bwd(g, L["PQ"], L["idx"], L["esel"], L["jsel"], L["esum"], L["norm"])
            dwc, _, _ = ops.linear_bwd_weight(dPQ, L["inp"], want_bias=False)
            grads[3 * li] = _wcat_grad(dwc, Cout).reshape(ctx.shapes[3 * li])
            grads[3 * li + 1] = dg
            grads[3 * li + 2] = db
            if li > 0:
                P = layers[li - 1]
                ops.linear_bwd_data(dPQ, L["wc"], dZ=dxf[:, :, P["lo"]:P["hi"]], accumulate=True)
        dx0 = None
        if ctx.needs_input_grad[0]:
            dx0, _ = ops.linear_bwd_data(dPQ, layers[0]["wc"])
        return (dx0, None, None, None) + tuple(grads)


class HeadFn(torch.autograd.Function):
    """(x4 (B,1024), xf (B,N,256)) -> embedding (B,N,E), logp (B,P,N)
    params: [c1w,c1b,n1g,n1b, c2w,c2b,n2g,n2b, s1w,s1b,nsg,nsb, s2w,s2b, p1w,p1b,npg,npb, p2w,p2b]"""

    @staticmethod
    def forward(ctx, x4, xf, *params):
        P = [p.detach() for p in params]
        (c1w, c1b, n1g, n1b, c2w, c2b, n2g, n2b, s1w, s1b, nsg, nsb, s2w, s2b, p1w, p1b, npg, npb, p2w, p2b) = P
        x4 = x4.detach().contiguous()
        xf = xf.detach()
        B, N, _ = xf.shape
        W1 = c1w.reshape(512, 1280)
        Wg, Wl = W1[:, :1024], W1[:, 1024:]
        sb, _ = ops.linear_fwd(x4.view(1, B, 1024), Wg)
        sb = sb.view(B, 512)
        Y1, st1 = ops.linear_fwd(xf, Wl, bias=c1b, sbias=sb, stats_groups=8)
        f1 = ops.norm_finalize(st1, n1g, n1b, B, 512, 64 * N, ACT_RELU)
        W2 = c2w.reshape(256, 512)
        Y2, st2 = ops.linear_fwd(Y1, W2, bias=c2b, in_norm=f1, stats_groups=4)
        f2 = ops.norm_finalize(st2, n2g, n2b, B, 256, 64 * N, ACT_RELU)
        Ws1 = s1w.reshape(256, 256)
        Ye, ste = ops.linear_fwd(Y2, Ws1, bias=s1b, in_norm=f2, stats_groups=4)
        fe = ops.norm_finalize(ste, nsg, nsb, B, 256, 64 * N, ACT_RELU)
        Ws2 = s2w.reshape(s2w.shape[0], 256)
        emb, _ = ops.linear_fwd(Ye, Ws2, bias=s2b, in_norm=fe)
        Wp1 = p1w.reshape(256, 256)
        Yp, stp = ops.linear_fwd(Y2, Wp1, bias=p1b, in_norm=f2, stats_groups=4)
        fp = ops.norm_finalize(stp, npg, npb, B, 256, 64 * N, ACT_RELU)
        Wp2 = p2w.reshape(p2w.shape[0], 256)
        logits, _ = ops.linear_fwd(Yp, Wp2, bias=p2b, in_norm=fp)
        logp = ops.logsoftmax_fwd(logits)
        ctx.t = dict(x4=x4, xf=xf, W1=W1, Wg=Wg, Wl=Wl, W2=W2, Ws1=Ws1, Ws2=Ws2, Wp1=Wp1, Wp2=Wp2, Y1=Y1, Y2=Y2,
                     Ye=Ye, Yp=Yp, f1=f1, f2=f2, fe=fe, fp=fp, logp=logp)
        ctx.shapes = [p.shape for p in params]
        return emb, logp.view_as(logp)      # (fresh view: see EncoderFn.forward)

    @staticmethod
    def backward(ctx, gemb, glogp):
        t = ctx.t
        xf, x4 = t["xf"], t["x4"]
        B, N, _ = xf.shape
        dev = xf.device
        S = ctx.shapes
        grads = [None] * 20
        dZ2 = None
        # ---- primitive-type branch
        if glogp is not None:
            dlogits = ops.logsoftmax_bwd(t["logp"], glogp.contiguous())
            dW, db, _ = ops.linear_bwd_weight(dlogits, t["Yp"], in_norm=t["fp"])
            grads[18], grads[19] = dW.view(S[18]), db
            dZp, gs = ops.linear_bwd_data(dlogits, t["Wp2"], fin_A=t["Yp"], fin_norm=t["fp"])
            grads[16], grads[17] = ops.norm_bwd_apply(dZp, t["Yp"], t["fp"], gs)
            dW, db, _ = ops.linear_bwd_weight(dZp, t["Y2"], in_norm=t["f2"])
            grads[14], grads[15] = dW.view(S[14]), db
            dZ2, _ = ops.linear_bwd_data(dZp, t["Wp1"])
        # ---- embedding branch
        if gemb is not None:
            gemb = gemb.contiguous()
            dW, db, _ = ops.linear_bwd_weight(gemb, t["Ye"], in_norm=t["fe"])
            grads[12], grads[13] = dW.view(S[12]), db
            dZe, gs = ops.linear_bwd_data(gemb, t["Ws2"], fin_A=t["Ye"], fin_norm=t["fe"])
            grads[10], grads[11] = ops.norm_bwd_apply(dZe, t["Ye"], t["fe"], gs)
            dW, db, _ = ops.linear_bwd_weight(dZe, t["Y2"], in_norm=t["f2"])
            grads[8], grads[9] = dW.view(S[8]), db
            dZ2, gs2 = ops.linear_bwd_data(dZe, t["Ws1"], dZ=dZ2, accumulate=dZ2 is not None, fin_A=t["Y2"],
                                           fin_norm=t["f2"])
        else:
            # finalize on a zero contribution (keeps one code path): dZ2 += 0 @ Ws1 with mask + sums
            zero = torch.zeros((B, N, 256), dtype=torch.float32, device=dev)
            dZ2, gs2 = ops.linear_bwd_data(zero, t["Ws1"], dZ=dZ2, accumulate=True, fin_A=t["Y2"], fin_norm=t["f2"])
        grads[6], grads[7] = ops.norm_bwd_apply(dZ2, t["Y2"], t["f2"], gs2)
        dW, db, _ = ops.linear_bwd_weight(dZ2, t["Y1"], in_norm=t["f1"])
        grads[4], grads[5] = dW.view(S[4]), db
        dZ1, gs1 = ops.linear_bwd_data(dZ2, t["W2"], fin_A=t["Y1"], fin_norm=t["f1"])
        grads[2], grads[3] = ops.norm_bwd_apply(dZ1, t["Y1"], t["f1"], gs1)
        # conv1: local part (xf) + hoisted global part (x4)
        dW1 = torch.zeros((512, 1280), dtype=torch.float32, device=dev)
        _, db1, dsb = ops.linear_bwd_weight(dZ1, xf, dW=dW1[:, 1024:], want_bias=True, want_sbias=True)
        ops.linear_bwd_weight(dsb.view(1, B, 512), x4.view(1, B, 1024), dW=dW1[:, :1024], want_bias=False)
        grads[0], grads[1] = dW1.view(S[0]), db1
        dxf, _ = ops.linear_bwd_data(dZ1, t["Wl"])
        dx4, _ = ops.linear_bwd_data(dsb.view(1, B, 512), t["Wg"])
        return (dx4.view(B, 1024), dxf) + tuple(grads)
