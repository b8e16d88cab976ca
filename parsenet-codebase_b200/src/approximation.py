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
    assembled and solved on the device in float64 (cuBLAS / cuSOLVER through torch: this row is outside the measured
    hot path; a batched in-kernel Cholesky is the 8f follow-up)."""
    dev = _dev()
    as_t = lambda a: a.to(dev).double() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).to(dev).double()
    P, U, V = as_t(points), as_t(basis_u), as_t(basis_v)
    A = (U.unsqueeze(2) * V.unsqueeze(1)).reshape(U.shape[0], -1)
    G = A.t() @ A
    # the Gram matrix is singular when fewer basis functions than (n+1)(m+1) are excited; lstsq's minimum-norm solution
    # is what numpy returns in the reference, so solve through the pseudo-inverse of the (symmetric) normal matrix
    ctrl = torch.linalg.pinv(G, hermitian=True) @ (A.t() @ P)
    ctrl = ctrl.reshape(U.shape[1], V.shape[1], 3)
    return ctrl if isinstance(points, torch.Tensor) else ctrl.cpu().numpy()


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
