"""Primitive fitting from weighted moments + batched residuals, Chamfer and B-spline evaluation wrappers.

The O(points) streaming work runs in csrc/{fit,primitives,chamfer,spline}.cu; the 3x3 algebra below runs on
(S,3,3) float64 tensors (S = segments of one type in a shape) and reproduces, including the reference's custom SVD
gradient, Fit.fit_{plane,sphere,cylinder,cone}_torch (reference src/primitive_forward.py:708-843),
LeastSquares.lstsq (src/fitting_utils.py:36-65) and CustomSVD (src/fitting_utils.py:385-455).
"""
import numpy as np
import torch

from .cabi import call
from .ops import _need_cuda, _ptr, _stream

EPS = float(np.finfo(np.float32).eps)
NM = 55
# moment layout (power of w, monomial):  see csrc/fit.cu eval_phi
M1_1, M1_P, M1_PP, M1_N = 0, slice(1, 4), slice(4, 10), slice(10, 13)
M2_1, M2_P, M2_PP, M2_N, M2_NN, M2_NPN = 13, slice(14, 17), slice(17, 23), slice(23, 26), slice(26, 32), slice(32, 35)
M3_PP, M3_T, M0_N, M0_1 = slice(35, 41), slice(41, 51), slice(51, 54), 54

_SYM6 = torch.tensor([[0, 1, 2], [1, 3, 4], [2, 4, 5]])
_T10 = {(0, 0, 0): 0, (0, 0, 1): 1, (0, 0, 2): 2, (0, 1, 1): 3, (0, 1, 2): 4, (0, 2, 2): 5, (1, 1, 1): 6, (1, 1, 2): 7,
        (1, 2, 2): 8, (2, 2, 2): 9}
_T27 = torch.tensor([[[_T10[tuple(sorted((i, j, k)))] for k in range(3)] for j in range(3)] for i in range(3)])


_DEV_CONST = {}


def _dev_const(name, t, device):
    """index tables live on the device once (a .to(device) per call is a blocking pageable H2D copy)"""
    key = (name, str(device))
    c = _DEV_CONST.get(key)
    if c is None:
        c = _DEV_CONST[key] = t.to(device)
    return c


def sym6(v):
    """(S,6) [xx,xy,xz,yy,yz,zz] -> (S,3,3)"""
    return v[:, _dev_const("sym6", _SYM6, v.device)]


def sym10(v):
    """(S,10) unique entries of a symmetric 3-tensor -> (S,3,3,3)"""
    return v[:, _dev_const("t27", _T27, v.device)]


def eigh3(G):
    """(S,3,3) float64 symmetric -> eigenvalues ascending (S,3), eigenvectors in columns (S,3,3); never synchronises
    (csrc/small3.cu; torch.linalg.eigh on CUDA blocks the host on the cusolver info read-back)"""
    G = G.detach().to(torch.float64).contiguous()
    S = G.shape[0]
    w = torch.empty((S, 3), dtype=torch.float64, device=G.device)
    V = torch.empty((S, 3, 3), dtype=torch.float64, device=G.device)
    if S:
        call("pn_sym3_eigh", _ptr(G), S, _ptr(w), _ptr(V), _stream())
    return w, V


# ------------------------------------------------------------------------------------------------ moments
class MomentsFn(torch.autograd.Function):
    """W (N,K) membership weights -> (K, 55) float64 moments of the point set {start + i*step, i < m} with
    w = W[n,s] + EPS (fit_one_shape_torch adds EPS, primitive_forward.py:942)."""

    @staticmethod
    def forward(ctx, W, P, Nr, start, step, m, eps=EPS):
        _need_cuda(W, P)
        W = W.detach()
        assert (W.shape[1] == 1 or W.stride(1) == 1) and P.is_contiguous() and (Nr is None or Nr.is_contiguous())
        N, K = W.shape
        mom = torch.zeros((K, NM), dtype=torch.float64, device=W.device)
        call("pn_fit_moments_fwd", _ptr(P), _ptr(Nr), _ptr(W), W.stride(0), K, start, step, m, float(eps), _ptr(mom),
             _stream())
        ctx.saved = (W, P, Nr, start, step, m, float(eps))
        return mom

    @staticmethod
    def backward(ctx, gmom):
        W, P, Nr, start, step, m, eps = ctx.saved
        N, K = W.shape
        gW = torch.zeros((N, K), dtype=torch.float32, device=W.device)
        g32 = gmom.to(torch.float32).contiguous()
        call("pn_fit_moments_bwd", _ptr(P), _ptr(Nr), _ptr(W), W.stride(0), K, start, step, m, eps, _ptr(g32),
             _ptr(gW), K, _stream())
        return gW, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------ custom SVD of a Gram
def floor_singular_values(sv):
    """(S,3) descending singular values.  A null direction of the Gram matrix comes out of the eigen-solver as a tiny
    NEGATIVE eigenvalue about half of the time -> singular value exactly 0 -> 1/(s_i + s_i) = inf on the diagonal of K
    -> inf * 0 = NaN in the backward (seen as NaN gradients after ~10 optimizer steps, spread to every rank by the
    gradient all-reduce).  An fp32 SVD of the (m,3) matrix, which is what the reference differentiates, returns
    ~eps32 * s_max there; use the same floor."""
    return torch.maximum(sv, (EPS * sv[:, :1]).clamp(min=1e-30))


class GramSVDFn(torch.autograd.Function):
    """G = A^T A (S,3,3) -> V (S,3,3) with columns ordered by DEcreasing singular value of A, sv (S,3).
    Backward is the reference's custom rule (fitting_utils.py:385-417): grad_A = 2 U S sym(K^T o V^T gV) V^T, i.e.
    grad_G = V sym(K^T o V^T gV) V^T, K_ij = 1/((s_i - s_j)(s_i + s_j)) with |s_i - s_j| floored at 1e-6."""

    @staticmethod
    def forward(ctx, G):
        evals, evecs = eigh3(G)                                # ascending
        V = torch.flip(evecs, dims=[2])
        sv = floor_singular_values(torch.sqrt(torch.clamp(torch.flip(evals, dims=[1]), min=0.0)))
        ctx.save_for_backward(V, sv)
        return V, sv

    @staticmethod
    def backward(ctx, gV, gS):
        V, sv = ctx.saved_tensors
        s_col = sv.unsqueeze(2)             # s_i along rows
        s_row = sv.unsqueeze(1)
        diff = s_col - s_row
        plus = s_col + s_row
        # (sign(0) := +1: two null directions floored to the same value must not give 1/0)
        sgn = torch.where(diff >= 0, torch.ones_like(diff), -torch.ones_like(diff))
        kneg = sgn * torch.clamp(diff.abs(), min=1e-6)
        eye = torch.eye(3, dtype=sv.dtype, device=sv.device)
        kneg = kneg * (1 - eye) + 1e-6 * eye
        K = (1.0 / kneg) * (1.0 / torch.clamp(plus, min=1e-8)) * (1 - eye)     # (|K| <= 1e14, the reference's fp32 range)
        inner = K.transpose(1, 2) * (V.transpose(1, 2) @ gV)
        inner = (inner + inner.transpose(1, 2)) / 2.0
        return V @ inner @ V.transpose(1, 2)


class Lstsq3Fn(torch.autograd.Function):
    """x = (AtA + lambda I)^-1 AtY with the reference's rank rule for lambda (csrc/small3.cuh lstsq3), (S,3,3),(S,3)
    float64 -> (S,3).  Backward is the one of a linear solve with lambda held fixed (what autograd does for the
    reference's torch.inverse path): gAtY = M^-1 g, gAtA = -gAtY x^T."""

    @staticmethod
    def forward(ctx, AtA, AtY, rows):
        A = AtA.detach().to(torch.float64).contiguous()
        Y = AtY.detach().to(torch.float64).contiguous()
        S = A.shape[0]
        x = torch.empty((S, 3), dtype=torch.float64, device=A.device)
        minv = torch.empty((S, 3, 3), dtype=torch.float64, device=A.device)
        lam = torch.empty((S,), dtype=torch.float64, device=A.device)
        if S:
            call("pn_lstsq3", _ptr(A), _ptr(Y), S, int(rows), float(EPS), _ptr(x), _ptr(minv), _ptr(lam), _stream())
        ctx.save_for_backward(x, minv)
        return x

    @staticmethod
    def backward(ctx, g):
        x, minv = ctx.saved_tensors
        gy = (minv @ g.unsqueeze(2)).squeeze(2)
        return -gy.unsqueeze(2) * x.unsqueeze(1), gy, None


def solve_normal(AtA, AtY, rows):
    """LeastSquares.lstsq in normal-equation form: x = argmin |A x - Y|.  Rank-deficient systems follow the
    reference's regularised branch (lambda = 1e-6 * 10^j, first j making AtA + lambda I full rank)."""
    return Lstsq3Fn.apply(AtA, AtY, rows)


# ------------------------------------------------------------------------------------------------ fits from moments
def plane_from(w1, s1p, w2, s2p, s2pp):
    """-> a (S,3) unit normal (smallest right singular vector of w*(p-c)), d (S,) = sum w (a.p)/sum w = a.c"""
    sw = w1 + EPS
    c = s1p / sw.unsqueeze(1)
    G = s2pp - c.unsqueeze(2) * s2p.unsqueeze(1) - s2p.unsqueeze(2) * c.unsqueeze(1) \
        + w2.view(-1, 1, 1) * c.unsqueeze(2) * c.unsqueeze(1)
    V, _ = GramSVDFn.apply(G)
    a = V[:, :, -1]
    d = (a * s1p).sum(1) / sw
    return a, d


def sphere_from(w1, s1p, q1, w2, s2p, s2pp, q3, r3, rows):
    """linearised weighted sphere fit (primitive_forward.py:746-769) -> centre (S,3), radius (S,)"""
    sw = w1 + EPS
    c = s1p / sw.unsqueeze(1)
    nu = q1 / sw
    AtA = 4.0 * (w2.view(-1, 1, 1) * c.unsqueeze(2) * c.unsqueeze(1) - c.unsqueeze(2) * s2p.unsqueeze(1)
                 - s2p.unsqueeze(2) * c.unsqueeze(1) + s2pp)
    AtY = 2.0 * (c * q3.unsqueeze(1) - r3 - (nu * w2).unsqueeze(1) * c + nu.unsqueeze(1) * s2p)
    center = -solve_normal(AtA, AtY, rows)
    r2 = (q1 - 2.0 * (center * s1p).sum(1) + (center * center).sum(1) * w1) / sw
    radius = torch.sqrt(torch.clamp(torch.clamp(r2, min=1e-3), min=1e-5))
    return center, radius


def fit_planes(mom):
    return plane_from(mom[:, M1_1], mom[:, M1_P], mom[:, M2_1], mom[:, M2_P], sym6(mom[:, M2_PP]))


def fit_spheres(mom, rows):
    pp1, pp3, T = sym6(mom[:, M1_PP]), sym6(mom[:, M3_PP]), sym10(mom[:, M3_T])
    q1 = pp1.diagonal(dim1=1, dim2=2).sum(1)
    q3 = pp3.diagonal(dim1=1, dim2=2).sum(1)
    r3 = torch.einsum("siik->sk", T)
    return sphere_from(mom[:, M1_1], mom[:, M1_P], q1, mom[:, M2_1], mom[:, M2_P], sym6(mom[:, M2_PP]), q3, r3, rows)


def fit_cylinders(mom, rows):
    """axis = smallest right singular vector of w*normals; circle fit of the points projected along it
    (primitive_forward.py:784-806) -> a (S,3), centre (S,3), radius (S,)"""
    V, _ = GramSVDFn.apply(sym6(mom[:, M2_NN]))
    a = V[:, :, -1]
    a = a / (a.norm(dim=1, keepdim=True) + EPS)
    eye = torch.eye(3, dtype=mom.dtype, device=mom.device)
    Pm = eye - a.unsqueeze(2) * a.unsqueeze(1)
    M2 = Pm.transpose(1, 2) @ Pm
    pp1, pp2, pp3, T = sym6(mom[:, M1_PP]), sym6(mom[:, M2_PP]), sym6(mom[:, M3_PP]), sym10(mom[:, M3_T])
    s1p = torch.einsum("sij,sj->si", Pm, mom[:, M1_P])
    s2p = torch.einsum("sij,sj->si", Pm, mom[:, M2_P])
    s2pp = Pm @ pp2 @ Pm.transpose(1, 2)
    q1 = (M2 * pp1).sum((1, 2))
    q3 = (M2 * pp3).sum((1, 2))
    r3 = torch.einsum("skl,sij,sijl->sk", Pm, M2, T)
    center, radius = sphere_from(mom[:, M1_1], s1p, q1, mom[:, M2_1], s2p, s2pp, q3, r3, rows)
    return a, center, radius


def fit_cone_apex_axis(mom, rows):
    """apex from the normal-equation solve, axis from a plane fit of the normals, sign so that sum(n.a) <= 0
    (primitive_forward.py:808-831).  Returns apex (S,3), axis (S,3), degenerate (S,) bool (cond > 1e5)."""
    nn = sym6(mom[:, M2_NN])
    with torch.no_grad():
        ev = torch.flip(eigh3(nn)[0], dims=[1])
        s = torch.sqrt(torch.clamp(ev, min=0.0))
        degenerate = (s[:, 0] / s[:, 2]) > 1e5
    apex = solve_normal(nn, mom[:, M2_NPN], rows)
    a, _ = plane_from(mom[:, M1_1], mom[:, M1_N], mom[:, M2_1], mom[:, M2_N], nn)
    flip = ((mom[:, M0_N] * a).sum(1) > 0).to(a.dtype).unsqueeze(1)
    a = a * (1 - 2 * flip)
    return apex, a, degenerate


def cone_theta(points, weights, apex, axis):
    """weighted mean opening angle (primitive_forward.py:833-842); points (m,3), weights (m,1), apex/axis (3,)"""
    diff = torch.nn.functional.normalize(points - apex.view(1, 3), p=2, dim=1)
    v = torch.clamp((diff @ axis.view(3, 1)).abs(), max=0.999)
    theta = (weights * torch.acos(v)).sum() / (weights.sum() + EPS)
    return torch.clamp(theta, min=1e-3, max=3.142 / 2 - 1e-3)


# ------------------------------------------------------------------------------------------------ residuals
TYPE_ID = {"plane": 0, "sphere": 1, "cylinder": 2, "cone": 3}


class ResidualFn(torch.autograd.Function):
    """mean squared point-to-primitive distance for every segment in one launch.
    par (S,8) fp32 parameter rows, types (S,) int32, seg (N,) int32 (segment of every point or -1)."""

    @staticmethod
    def forward(ctx, par, points, seg, types):
        _need_cuda(par, points)
        par = par.detach().contiguous()
        S = par.shape[0]
        dev = par.device
        sumf = torch.zeros((S,), dtype=torch.float32, device=dev)
        jac = torch.zeros((S, 8), dtype=torch.float32, device=dev)
        cnt = torch.zeros((S,), dtype=torch.float32, device=dev)
        call("pn_residual_fwd", _ptr(points), _ptr(seg), points.shape[0], _ptr(types), _ptr(par), S, _ptr(sumf),
             _ptr(jac), _ptr(cnt), _stream())
        cnt = torch.clamp(cnt, min=1.0)
        ctx.save_for_backward(jac, cnt)
        return sumf / cnt

    @staticmethod
    def backward(ctx, g):
        jac, cnt = ctx.saved_tensors
        return jac * (g / cnt).unsqueeze(1), None, None, None


# ------------------------------------------------------------------------------------------------ chamfer
class NearestFn(torch.autograd.Function):
    """A (B,Na,3), Bs (B,Nb,3) -> squared distance of every A point to its nearest Bs point (B,Na)"""

    @staticmethod
    def forward(ctx, A, Bs):
        _need_cuda(A, Bs)
        A = A.detach().contiguous().float()
        Bs = Bs.detach().contiguous().float()
        B, Na, _ = A.shape
        Nb = Bs.shape[1]
        mind = torch.empty((B, Na), dtype=torch.float32, device=A.device)
        arg = torch.empty((B, Na), dtype=torch.int32, device=A.device)
        call("pn_chamfer_nn_fwd", _ptr(A), Na, _ptr(Bs), Nb, B, _ptr(mind), _ptr(arg), _stream())
        ctx.saved = (A, Bs, arg)
        return mind

    @staticmethod
    def backward(ctx, g):
        A, Bs, arg = ctx.saved
        B, Na, _ = A.shape
        dA = torch.zeros_like(A) if ctx.needs_input_grad[0] else None
        dB = torch.zeros_like(Bs) if ctx.needs_input_grad[1] else None
        g = g.contiguous()
        call("pn_chamfer_nn_bwd", _ptr(A), Na, _ptr(Bs), Bs.shape[1], B, _ptr(arg), _ptr(g), _ptr(dA), _ptr(dB),
             _stream())
        return dA, dB


def nearest_sqdist(A, Bs):
    return NearestFn.apply(A, Bs)


def nearest_index(A, Bs):
    """A (B,Na,3), Bs (B,Nb,3) -> index of the nearest Bs point of every A point (B,Na) int32 (the argmin the Chamfer kernel
    keeps for its backward)"""
    _need_cuda(A, Bs)
    A = A.detach().contiguous().float()
    Bs = Bs.detach().contiguous().float()
    B, Na, _ = A.shape
    mind = torch.empty((B, Na), dtype=torch.float32, device=A.device)
    arg = torch.empty((B, Na), dtype=torch.int32, device=A.device)
    call("pn_chamfer_nn_fwd", _ptr(A), Na, _ptr(Bs), Bs.shape[1], B, _ptr(mind), _ptr(arg), _stream())
    return arg


# ------------------------------------------------------------------------------------------------ spline evaluation
class SplineEvalFn(torch.autograd.Function):
    """P (B,cu,cv,3) -> (B, gu*gv, 3) = Nu P Nv^T per coordinate.  float32, or float64 when P is float64"""

    @staticmethod
    def forward(ctx, P, Nu, Nv):
        _need_cuda(P, Nu, Nv)
        f64 = P.dtype == torch.float64
        dt = torch.float64 if f64 else torch.float32
        P = P.detach().contiguous().to(dt)
        Nu = Nu.detach().contiguous().to(dt)
        Nv = Nv.detach().contiguous().to(dt)
        B, cu, cv, _ = P.shape
        gu, gv = Nu.shape[0], Nv.shape[0]
        out = torch.empty((B, gu * gv, 3), dtype=dt, device=P.device)
        call("pn_spline_eval_fwd_f64" if f64 else "pn_spline_eval_fwd", _ptr(Nu), _ptr(Nv), _ptr(P), B, gu, gv, cu, cv,
             _ptr(out), _stream())
        ctx.saved = (Nu, Nv, (B, cu, cv, gu, gv), f64)
        return out

    @staticmethod
    def backward(ctx, g):
        Nu, Nv, (B, cu, cv, gu, gv), f64 = ctx.saved
        g = g.contiguous().to(Nu.dtype)
        dP = torch.empty((B, cu, cv, 3), dtype=Nu.dtype, device=g.device)
        call("pn_spline_eval_bwd_f64" if f64 else "pn_spline_eval_bwd", _ptr(Nu), _ptr(Nv), _ptr(g), B, gu, gv, cu, cv,
             _ptr(dP), _stream())
        return dP, None, None


def spline_eval(P_bijc, Nu, Nv):
    return SplineEvalFn.apply(P_bijc, Nu, Nv)


# ------------------------------------------------------------------------------------------------ control-point solve
_PINV = {}


def pinv_basis(N_gc, device):
    """(g, c) basis matrix -> its left pseudo-inverse (N^T N)^-1 N^T  (c, g), float64 on the host once, cached as a
    float64 device constant (the reference forms exactly this product, src/approximation.py:319-323)"""
    N64 = np.asarray(N_gc.detach().cpu().numpy() if isinstance(N_gc, torch.Tensor) else N_gc, dtype=np.float64)
    key = (N64.shape, hash(N64.tobytes()), str(device))
    t = _PINV.get(key)
    if t is None:
        pinv = np.linalg.inv(N64.T @ N64) @ N64.T
        t = _PINV[key] = torch.from_numpy(pinv).to(device)
    return t


def fit_control_points_grid(S_bguv3, nu, nv):
    """Least-squares control grid of gridded surface samples: P = Nu^+ S (Nv^+)^T per coordinate.
    S (B, gu, gv, 3) cuda -> (B, cu, cv, 3), same dtype as S.  Replaces approximation.fit_bezier_surface (reference
    src/approximation.py:308-334, numpy float64, one shape at a time): the solve is the SAME tensor-product kernel as
    the surface evaluation, run with the pseudo-inverse basis matrices, and is differentiable w.r.t. the samples.
    The arithmetic is float64 like the reference's: ||Nu^+||_1 ||Nv^+||_1 ~ 8e4 for the 30 -> 20 cubic basis, so an
    fp32 solve (even fp32 rounding of the samples alone) sits at 3e-4, above the 1e-4 parity bar."""
    _need_cuda(S_bguv3)
    B, gu, gv, _ = S_bguv3.shape
    pu, pv = pinv_basis(nu, S_bguv3.device), pinv_basis(nv, S_bguv3.device)
    assert pu.shape[1] == gu and pv.shape[1] == gv, "basis matrices do not match the sample grid"
    out = SplineEvalFn.apply(S_bguv3.double(), pu, pv).view(B, pu.shape[0], pv.shape[0], 3)
    return out.to(S_bguv3.dtype)


def kron_fit(P_sm3, U_smn, V_smm):
    """least-squares control points of scattered surface samples with per-sample basis rows (csrc/kronfit.cu):
    P (S,M,3), U (S,M,n), V (S,M,m) cuda -> (ctrl (S,n,m,3) float64, flag (S,) int32; flag 1 = rank-deficient, ctrl[s] unset).
    Replaces fit_bezier_surface_fit_kronecker (reference src/approximation.py:338-364) for a batch of surfaces."""
    _need_cuda(P_sm3, U_smn, V_smm)
    P, U, V = (t.detach().to(torch.float64).contiguous() for t in (P_sm3, U_smn, V_smm))
    S, M, n = U.shape
    m = V.shape[2]
    assert P.shape == (S, M, 3) and V.shape[:2] == (S, M)
    ctrl = torch.zeros((S, n, m, 3), dtype=torch.float64, device=P.device)
    flag = torch.empty((S,), dtype=torch.int32, device=P.device)
    call("pn_kron_fit", _ptr(U), _ptr(V), _ptr(P), S, M, n, m, _ptr(ctrl), _ptr(flag), _stream())
    return ctrl, flag


def kron_eval(C_snm3, U_smn, V_smm):
    """surface points at scattered parameters: out[s,i] = sum_ab U[s,i,a] V[s,i,b] C[s,a,b]  (float64; C (n,m,3) is shared by
    all surfaces).  Replaces geomdl's evaluate_list in the post-fit optimisers (reference src/primitive_forward.py:186,258)."""
    _need_cuda(C_snm3, U_smn, V_smm)
    C, U, V = (t.detach().to(torch.float64).contiguous() for t in (C_snm3, U_smn, V_smm))
    S, M, n = U.shape
    m = V.shape[2]
    shared = C.dim() == 3
    assert C.shape[-3:] == (n, m, 3) and (shared or C.shape[0] == S)
    out = torch.empty((S, M, 3), dtype=torch.float64, device=U.device)
    call("pn_kron_eval", _ptr(U), _ptr(V), _ptr(C), 0 if shared else n * m * 3, S, M, n, m, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------ batched fit stage
# Slot-indexed versions of the three fit kernels for Evaluation.fitting_loss over a whole batch of shapes (fitstage.py):
# every table is (B, S, .) with S = 64 weight columns ("slots") per shape; unused slots carry kind -1.
KIND_NONE = -1


class MomentsBatchedFn(torch.autograd.Function):
    """W (B,N,S) membership weights -> (B,S,55) float64 moments of the point set {start + i*step, i < m} of every shape,
    w = W + eps, in ONE launch (csrc/fit.cu, gridDim.y = shape)"""

    @staticmethod
    def forward(ctx, W, P, Nr, start, step, m, eps):
        _need_cuda(W, P)
        W = W.detach()
        assert W.is_contiguous() and P.is_contiguous() and (Nr is None or Nr.is_contiguous())
        B, N, S = W.shape
        mom = torch.zeros((B, S, NM), dtype=torch.float64, device=W.device)
        call("pn_fit_moments_fwd_batched", _ptr(P), _ptr(Nr), _ptr(W), S, B, N, S, start, step, m, float(eps), _ptr(mom),
             _stream())
        ctx.saved = (W, P, Nr, start, step, m, float(eps))
        return mom

    @staticmethod
    def backward(ctx, gmom):
        W, P, Nr, start, step, m, eps = ctx.saved
        B, N, S = W.shape
        gW = torch.zeros((B, N, S), dtype=torch.float32, device=W.device)
        g32 = gmom.to(torch.float32).contiguous()
        call("pn_fit_moments_bwd_batched", _ptr(P), _ptr(Nr), _ptr(W), S, B, N, S, start, step, m, eps, _ptr(g32),
             _ptr(gW), S, _stream())
        return gW, None, None, None, None, None, None


class FitSolveFn(torch.autograd.Function):
    """mom (S,55) float64, kind (S,) int32 -> par (S,8) float64 parameter rows (residual-kernel layout, cone angle slot
    left 0) and bad (S,) float32 degenerate-cone flags, ONE launch for every segment of every shape (csrc/fitsolve.cu).
    The kernel also returns d par / d mom (forward mode), so the backward is one small batched product."""

    @staticmethod
    def forward(ctx, mom, kind, rows):
        _need_cuda(mom, kind)
        mom = mom.detach().contiguous()
        assert mom.dtype == torch.float64 and kind.dtype == torch.int32 and kind.is_contiguous()
        S = mom.shape[0]
        par = torch.empty((S, 8), dtype=torch.float64, device=mom.device)
        jac = torch.empty((S, 8, NM), dtype=torch.float64, device=mom.device)
        bad = torch.empty((S,), dtype=torch.float32, device=mom.device)
        call("pn_fit_solve", _ptr(mom), _ptr(kind), S, int(rows), _ptr(par), _ptr(jac), _ptr(bad), _stream())
        ctx.save_for_backward(jac)
        ctx.mark_non_differentiable(bad)
        return par, bad

    @staticmethod
    def backward(ctx, gpar, _gbad):
        (jac,) = ctx.saved_tensors
        return torch.bmm(gpar.unsqueeze(1), jac).squeeze(1), None, None


class ResidualBatchedFn(torch.autograd.Function):
    """mean squared point-to-primitive distance of every segment slot of every shape in ONE launch.
    par (B,S,8) fp32, points (B,N,3), seg (B,N) int32 slot of every point or -1, kind (B,S) int32 -> (B,S)"""

    @staticmethod
    def forward(ctx, par, points, seg, kind):
        _need_cuda(par, points)
        par = par.detach().contiguous()
        B, S, _ = par.shape
        dev = par.device
        assert points.is_contiguous() and seg.is_contiguous() and kind.is_contiguous()
        sumf_c = torch.zeros((B, S), dtype=torch.float32, device=dev)
        cnt_c = torch.zeros((B, S), dtype=torch.float32, device=dev)
        jac_c = torch.zeros((B, S, 8), dtype=torch.float32, device=dev)
        call("pn_residual_fwd_batched", _ptr(points), _ptr(seg), B, points.shape[1], _ptr(kind), _ptr(par), S,
             _ptr(sumf_c), _ptr(jac_c), _ptr(cnt_c), _stream())
        cnt_c = torch.clamp(cnt_c, min=1.0)
        ctx.save_for_backward(jac_c, cnt_c)
        return sumf_c / cnt_c

    @staticmethod
    def backward(ctx, g):
        jac, cnt = ctx.saved_tensors
        return jac * (g / cnt).unsqueeze(2), None, None, None


class WeightsNormalizeFn(torch.autograd.Function):
    """weights_normalize (fitting_utils.py:306-325) for every shape of a batch: raw (B,N,64) centre . point similarities,
    bw2 (B,) squared bandwidths, K (B,) int32 clusters per shape -> membership weights (B,N,64); 2 launches forward,
    2 backward (csrc/weights.cu).  Padded slots (>= K[b]) come out as exact zeros and receive no gradient."""

    @staticmethod
    def forward(ctx, raw, bw2, K):
        _need_cuda(raw, bw2, K)
        raw = raw.detach().contiguous()
        B, N, S = raw.shape
        assert raw.dtype == torch.float32 and K.dtype == torch.int32 and bw2.dtype == torch.float32
        out = torch.empty_like(raw)
        keys = torch.empty((2, B, S), dtype=torch.int64, device=raw.device)
        keys[0].fill_(-1)                   # all-ones = +inf for the running minima
        keys[1].zero_()
        call("pn_weights_normalize_fwd", _ptr(raw), _ptr(bw2), _ptr(K), B, N, S, _ptr(out), _ptr(keys), _stream())
        ctx.saved = (raw, bw2, K, keys)
        return out

    @staticmethod
    def backward(ctx, g):
        raw, bw2, K, keys = ctx.saved
        B, N, S = raw.shape
        g = g.contiguous()
        red = torch.zeros((B, S, 2), dtype=torch.float64, device=raw.device)
        graw = torch.empty_like(raw)
        call("pn_weights_normalize_bwd", _ptr(raw), _ptr(g), _ptr(bw2), _ptr(K), B, N, S, _ptr(keys), _ptr(red), _ptr(graw),
             _stream())
        return graw, None, None
