"""Segmentation losses on top of the kernels: triplet embedding loss and NLL over primitive types.

Reference: EmbeddingLoss.triplet_loss src/segment_loss.py:31-124, primitive_loss :151.  The random sampling stays on
the host with numpy so that the draw order (and hence the loss value for a given np.random state) is the
reference's; only index lists go to the device.
"""
import numpy as np
import torch

from . import ops
from .cabi import lib, check, call
from .ops import _ptr, _stream
from .staging import arena


def triplet_sample(labels, N, rng=np.random, max_segments=5):
    """Consumes host RNG draws exactly like segment_loss.py:60-96.  Returns per-S groups of
    (anchor rows, negative rows, weight) with rows indexing the flattened (B*N) embedding, and only_one count."""
    B = labels.shape[0]
    samples, S_of = [], []
    for i in range(B):
        p = labels[i]
        # rows of every label, ascending, via one stable sort (== np.where(np.isin(p, l))[0] per label)
        uniq, counts = np.unique(p, return_counts=True)
        order = np.argsort(p, kind="stable")
        starts = np.concatenate([[0], np.cumsum(counts)])
        S = min(N // uniq.shape[0] + 1, 30)
        d = {}
        for li, l in enumerate(uniq):
            # rng.choice(rows, S, replace=True) == rows[rng.randint(0, len(rows), S)] (same legacy draws)
            d[l] = order[starts[li]:starts[li + 1]][rng.randint(0, counts[li], size=S)]
        samples.append(d)
        S_of.append(S)
    groups = {}
    only_one = 0
    per_shape_norm = []
    for i in range(B):
        keys = sorted(samples[i].keys())
        L = len(keys)
        if L == 1:
            only_one += 1
            per_shape_norm.append(0)
            continue
        norm = 0
        pairs = []
        for _ in range(min(max_segments * max_segments, L * L)):
            k1 = rng.randint(0, L, size=1)[0]          # == rng.choice(L, 1)[0], same draws
            k2 = rng.randint(0, L, size=1)[0]
            if k1 == k2:
                continue
            norm += 1
            pairs.append((samples[i][keys[k1]] + i * N, samples[i][keys[k2]] + i * N))
        per_shape_norm.append(norm)
        g = groups.setdefault(S_of[i], dict(a=[], n=[], shape=[]))
        for a, n in pairs:
            g["a"].append(a); g["n"].append(n); g["shape"].append(i)
    denom = B - only_one + 1e-8
    out = []
    for S, g in groups.items():
        if not g["a"]:
            continue
        w = np.array([1.0 / ((per_shape_norm[i] + 1e-8) * denom) for i in g["shape"]], np.float32)
        out.append((S, np.stack(g["a"]).astype(np.int32), np.stack(g["n"]).astype(np.int32), w))
    return out


class L2NormFn(torch.autograd.Function):
    """F.normalize(x, p=2, dim=-1) for (..., D) tensors (rows contiguous) on the pn_l2norm_* kernels"""

    @staticmethod
    def forward(ctx, x):
        shp = x.shape
        x2 = x.detach().reshape(-1, shp[-1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        y, norms = ops.l2norm_fwd(x2)
        ctx.saved = (y, norms, shp)
        return y.view(shp)

    @staticmethod
    def backward(ctx, g):
        y, norms, shp = ctx.saved
        g2 = g.reshape(-1, shp[-1])
        if g2.stride(1) != 1:
            g2 = g2.contiguous()
        return ops.l2norm_bwd(y, g2, norms).view(shp)


def l2_normalize(x):
    return L2NormFn.apply(x)


class TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb_bnd, groups, margin):
        x = emb_bnd.detach()
        B, N, D = x.shape
        x2 = x.reshape(B * N, D)
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        E, norms = ops.l2norm_fwd(x2)
        dev = x.device
        total = torch.zeros((1,), dtype=torch.float32, device=dev)
        saved = []
        stage = arena("triplet", dev)
        stage.reset()
        for S, a, n, w in groups:
            T = a.shape[0]
            a_d, n_d, w_d = stage.upload(a, dev), stage.upload(n, dev), stage.upload(w, dev)
            pl = torch.empty((T,), dtype=torch.float32, device=dev)
            ps = torch.empty((T,), dtype=torch.float32, device=dev)
            call("pn_triplet_fwd", _ptr(E), D, D, _ptr(a_d), _ptr(n_d), T, S, float(margin), _ptr(pl), _ptr(ps),
                                     _stream())
            total = total + (pl * w_d).sum()
            saved.append((S, a_d, n_d, w_d, ps))
        ctx.saved = (E, norms, saved, margin, (B, N, D))
        return total

    @staticmethod
    def backward(ctx, g):
        E, norms, saved, margin, (B, N, D) = ctx.saved
        dE = torch.zeros_like(E)
        for S, a_d, n_d, w_d, ps in saved:
            pw = (w_d * g.reshape(())).contiguous()
            call("pn_triplet_bwd", _ptr(E), D, D, _ptr(a_d), _ptr(n_d), a_d.shape[0], S, float(margin), _ptr(ps),
                                     _ptr(pw), _ptr(dE), D, _stream())
        dx = ops.l2norm_bwd(E, dE, norms)
        return dx.view(B, N, D), None, None


class NllFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp_bpn, target_bn):
        lp = logp_bpn.detach().contiguous()
        tg = target_bn.detach().to(torch.int64).contiguous()
        B, P, N = lp.shape
        loss = torch.zeros((1,), dtype=torch.float32, device=lp.device)
        call("pn_nll_fwd", _ptr(lp), _ptr(tg), B, N, P, _ptr(loss), _stream())
        ctx.tg, ctx.shape = tg, (B, P, N)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        B, P, N = ctx.shape
        g = g.reshape(1).contiguous().float()
        dlp = torch.zeros((B, P, N), dtype=torch.float32, device=g.device)
        call("pn_nll_bwd", _ptr(ctx.tg), _ptr(g), B, N, P, _ptr(dlp), _stream())
        return dlp, None


# ------------------------------------------------------------------------------------------------ control-grid losses
class GridPermLossFn(torch.autograd.Function):
    """min over the symmetric re-orderings of the target control grid of the squared error (csrc/gridloss.cu).
    out, gt (B,g,g,3); mode 0 = the 8 dihedral candidates (open surfaces, loss.py:76-97), mode 1 = g cyclic shifts along u
    x 4 flips (closed in u, loss.py:100-124) -> (mean_b min_p / (g*g*3), best-matching candidate of gt (B,g,g,3))."""

    @staticmethod
    def forward(ctx, out, gt, mode):
        o = out.detach().contiguous().float()
        t = gt.detach().contiguous().float()
        B, g = o.shape[0], o.shape[1]
        P = 8 if mode == 0 else 4 * g
        dev = o.device
        diff = torch.empty((B, P), dtype=torch.float32, device=dev)
        loss_b = torch.empty((B,), dtype=torch.float32, device=dev)
        pick = torch.empty((B,), dtype=torch.int32, device=dev)
        best = torch.empty_like(t)
        call("pn_grid_perm_fwd", _ptr(o), _ptr(t), B, g, int(mode), _ptr(diff), _ptr(loss_b), _ptr(pick), _ptr(best), _stream())
        ctx.saved = (o, best, 1.0 / (B * g * g * 3))
        ctx.mark_non_differentiable(best)
        return loss_b.mean() / (g * g * 3), best

    @staticmethod
    def backward(ctx, gl, _gb):
        o, best, inv = ctx.saved
        dout = torch.empty_like(o)
        gs = gl.reshape(1).float().contiguous()
        call("pn_grid_perm_bwd", _ptr(o), _ptr(best), o.numel(), _ptr(gs), float(inv), _ptr(dout), _stream())
        return dout, None, None


class GridLaplacianLossFn(torch.autograd.Function):
    """mean over cells of the channel-summed squared (or absolute) difference of the zero-padded 4-neighbour Laplacians of two
    (B,g,g,3) grids (loss.py:213-239) as a 5-point stencil kernel instead of two cuDNN convolutions."""

    @staticmethod
    def forward(ctx, out, gt, l1):
        o = out.detach().contiguous().float()
        t = gt.detach().contiguous().float()
        B, g = o.shape[0], o.shape[1]
        l = torch.empty_like(o)
        part = torch.empty((B,), dtype=torch.float32, device=o.device)
        call("pn_grid_laplacian_fwd", _ptr(o), _ptr(t), B, g, int(l1), _ptr(l), _ptr(part), _stream())
        ctx.saved = (l, B, g, int(l1), 1.0 / (B * g * g))
        return part.sum() / (B * g * g)

    @staticmethod
    def backward(ctx, gl):
        l, B, g, l1, inv = ctx.saved
        dout = torch.empty_like(l)
        dgt = torch.empty_like(l) if ctx.needs_input_grad[1] else None
        gs = gl.reshape(1).float().contiguous()
        call("pn_grid_laplacian_bwd", _ptr(l), B, g, l1, _ptr(gs), float(inv), _ptr(dout), _ptr(dgt), _stream())
        return dout, dgt, None
