"""Drop-in for the control-point solves of reference src/approximation.py: fit_bezier_surface (:308-334, gridded
samples) and fit_bezier_surface_fit_kronecker (:338-364, scattered samples with per-point basis rows).  BASELINE
config 1 ("open-spline fit only ... control-point residual check") and the core of SURVEY 8f row 1.
The geomdl / bernstein helpers of that file are visualisation / dataset-side and out of scope."""
import numpy as np
import torch

from pnb200.fitting import fit_control_points_grid


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("parsenet_b200 has no CPU fallback: approximation.* needs a CUDA device")
    return torch.device("cuda", torch.cuda.current_device())


def fit_bezier_surface(points, basis_u, basis_v):
    """points (g, g, 3) gridded samples, basis_u (g, n+1), basis_v (g, m+1) -> control points (n+1, m+1, 3).
    numpy in / numpy out like the reference; torch cuda tensors (optionally batched (B, g, g, 3)) stay on the device
    and keep their autograd graph."""
    if isinstance(points, torch.Tensor):
        P = points if points.dim() == 4 else points.unsqueeze(0)
        out = fit_control_points_grid(P, basis_u, basis_v)
        return out if points.dim() == 4 else out[0]
    S = torch.from_numpy(np.asarray(points, dtype=np.float64)).to(_dev()).unsqueeze(0)
    return fit_control_points_grid(S, basis_u, basis_v)[0].cpu().numpy()


def fit_bezier_surface_fit_kronecker(points, basis_u, basis_v):
    """points (N, 3) scattered samples with per-point basis rows basis_u (N, n+1), basis_v (N, m+1) -> control points
    (n+1, m+1, 3): least squares on the Kronecker rows A_i = u_i (x) v_i.  The (n+1)(m+1)-square normal equations are
    assembled, Cholesky-factored and solved inside one kernel (csrc/kronfit.cu, float64, one CTA per surface; batched
    inputs (S, N, .) are one launch).  A rank-deficient sampling (fewer basis functions excited than control points) is
    flagged by the kernel; numpy's lstsq returns the minimum-norm solution there, which the pseudo-inverse of the
    (symmetric) normal matrix reproduces."""
    from pnb200.fitting import kron_fit
    dev = _dev()
    as_t = lambda a: a.to(dev).double() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).to(dev).double()
    P, U, V = as_t(points), as_t(basis_u), as_t(basis_v)
    batched = P.dim() == 3
    if not batched:
        P, U, V = P.unsqueeze(0), U.unsqueeze(0), V.unsqueeze(0)
    def pinv_solve(s):
        A = (U[s].unsqueeze(2) * V[s].unsqueeze(1)).reshape(U.shape[1], -1)
        return (torch.linalg.pinv(A.t() @ A, hermitian=True) @ (A.t() @ P[s])).reshape(U.shape[2], V.shape[2], 3)

    if U.shape[2] * V.shape[2] > 128:            # larger grids than the optimisers use (the kernel keeps G in shared memory)
        ctrl = torch.stack([pinv_solve(s) for s in range(P.shape[0])], 0)
    else:
        ctrl, flag = kron_fit(P, U, V)
        for s in torch.nonzero(flag).flatten().tolist():     # (one read-back; this row is outside the training step)
            ctrl[s] = pinv_solve(s)
    if not batched:
        ctrl = ctrl[0]
    return ctrl if isinstance(points, torch.Tensor) else ctrl.cpu().numpy()


def uniform_knot_bspline_(control_points_u, control_points_v, degree_u, degree_v, grid_size=30):
    """reference src/approximation.py:494-514: basis matrices on the regular grid plus the two clamped uniform knot vectors"""
    from src.loss import uniform_knot_bspline
    nu, nv = uniform_knot_bspline(control_points_u, control_points_v, degree_u, degree_v, grid_size)
    ku = [0.0] * degree_u + np.arange(0, 1.01, 1 / (control_points_u - degree_u)).tolist() + [1.0] * degree_u
    kv = [0.0] * degree_v + np.arange(0, 1.01, 1 / (control_points_v - degree_v)).tolist() + [1.0] * degree_v
    return nu, nv, ku, kv


def basis_rows(params, n_ctrl_u, n_ctrl_v, degree_u, degree_v):
    """per-sample basis rows of a clamped uniform B-spline surface: params (M, 2) in [0, 1]^2 -> NU (M, cu), NV (M, cv)
    float64 -- what the reference builds one sample at a time with BSpline.basis_functions (src/approximation.py:55-70,
    src/primitive_forward.py:202-208), evaluated for all M samples at once with the local-support recurrence
    (The NURBS Book A2.2): only the degree + 1 non-zero functions of a span are computed."""
    return _basis_rows_1d(params[:, 0], n_ctrl_u, degree_u), _basis_rows_1d(params[:, 1], n_ctrl_v, degree_v)


def _basis_rows_1d(u, n_ctrl, p):
    u = np.asarray(u, np.float64)
    U = np.array([0.0] * p + np.arange(0, 1.01, 1 / (n_ctrl - p)).tolist() + [1.0] * p)
    M = u.shape[0]
    # knot span of every sample (u == last knot belongs to the last non-empty span, like A2.4's special case)
    span = np.clip(np.searchsorted(U, u, side="right") - 1, p, n_ctrl - 1)
    N = np.zeros((M, p + 1))
    N[:, 0] = 1.0
    left = np.zeros((M, p + 1)); right = np.zeros((M, p + 1))
    for j in range(1, p + 1):
        left[:, j] = u - U[span + 1 - j]
        right[:, j] = U[span + j] - u
        saved = np.zeros(M)
        for r in range(j):
            den = right[:, r + 1] + left[:, j - r]
            t = np.divide(N[:, r], den, out=np.zeros(M), where=den != 0)
            N[:, r] = saved + right[:, r + 1] * t
            saved = left[:, j - r] * t
        N[:, j] = saved
    out = np.zeros((M, n_ctrl))
    cols = span[:, None] - p + np.arange(p + 1)[None, :]
    np.put_along_axis(out, cols, N, axis=1)
    return out


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
