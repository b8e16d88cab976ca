"""Drop-in for reference src/loss.py: B-spline basis (uniform_knot_bspline :190, basis_function_one :242), spline
reconstruction losses (:142-187), permutation-invariant control-point regression (:76-124), Laplacian loss (:213)."""
import numpy as np
import torch
from pnb200.fitting import spline_eval
from pnb200.losses import GridLaplacianLossFn, GridPermLossFn
from src.utils import chamfer_distance, chamfer_distance_one_side


def basis_function_one(degree, knot_vector, span, knot):
    """N_{span,degree}(knot) by the Cox-de Boor triangle (The NURBS Book, A2.4)."""
    U, p, i, u = knot_vector, degree, span, knot
    m = len(U) - 1
    if (i == 0 and u == U[0]) or (i == m - p - 1 and u == U[m]):
        return 1.0
    if u < U[i] or u >= U[i + p + 1]:
        return 0.0
    tri = [1.0 if U[i + j] <= u < U[i + j + 1] else 0.0 for j in range(p + 1)] + [0.0] * i
    for k in range(1, p + 1):
        saved = 0.0 if tri[0] == 0.0 else ((u - U[i]) * tri[0]) / (U[i + k] - U[i])
        for j in range(p - k + 1):
            left, right = U[i + j + 1], U[i + j + k + 1]
            if tri[j + 1] == 0.0:
                tri[j], saved = saved, 0.0
            else:
                t = tri[j + 1] / (right - left)
                tri[j], saved = saved + (right - u) * t, (u - left) * t
    return tri[0]


def _basis_matrix(n_ctrl, degree, params):
    knots = [0.0] * degree + np.arange(0, 1.01, 1 / (n_ctrl - degree)).tolist() + [1.0] * degree
    return np.array([[basis_function_one(degree, knots, j, u) for j in range(n_ctrl)] for u in params])


def uniform_knot_bspline(control_points_u, control_points_v, degree_u, degree_v, grid_size=30):
    """clamped uniform basis matrices sampled at u = 0, 1/g, ..., (g-1)/g  -> nu (g,cu), nv (g,cv) float64"""
    u = np.arange(0., 1, 1 / grid_size)
    return _basis_matrix(control_points_u, degree_u, u), _basis_matrix(control_points_v, degree_v, u)


def _eval_grid(nu, nv, output, batch_size, cu, cv):
    P = output.reshape(batch_size, cu, cv, 3)
    return spline_eval(P, nu.to(P.device), nv.to(P.device))


def spline_reconstruction_loss_one_sided(nu, nv, output, points, config, side=1):
    """one-sided Chamfer between the evaluated surface (gu*gv points) and the input points (B,3,M)"""
    rec = _eval_grid(nu, nv, output, config.batch_size, config.grid_size, config.grid_size)
    return chamfer_distance_one_side(rec, points.permute(0, 2, 1), side), rec


def spline_reconstruction_loss(nu, nv, output, points, config, sqrt=False):
    rec = _eval_grid(nu, nv, output, config.batch_size, nu.shape[1], nv.shape[1])
    return chamfer_distance(rec, points.permute(0, 2, 1), sqrt=sqrt), rec


def all_permutations(array):
    """the 8 dihedral re-orderings of a (B,g,g,3) control grid -> (B,8,g,g,3)"""
    flips = [array, torch.flip(array, (1,)), torch.flip(array, (2,)), torch.flip(array, (1, 2))]
    return torch.stack(flips + [torch.transpose(f, 2, 1) for f in flips], 1)


def all_permutations_half(array):
    return torch.stack([array, torch.flip(array, (1,)), torch.flip(array, (2,)), torch.flip(array, (1, 2))], 1)


def roll(x, shift, dim=-1, fill_pad=None):
    return x if shift == 0 else torch.roll(x, shifts=shift, dims=dim)


def _best_permutation(output, candidates, norm):
    diff = ((output.unsqueeze(1) - candidates) ** 2).sum((2, 3, 4))
    loss, index = diff.min(1)
    return loss.mean() / norm, candidates[torch.arange(output.shape[0], device=output.device), index]


def control_points_permute_reg_loss(output, control_points, grid_size):
    """min over the 8 grid symmetries of the squared error; also returns the best-matching permuted target.
    One kernel evaluates the 8 candidates as index maps (csrc/gridloss.cu); nothing is stacked."""
    out = output.view(output.shape[0], grid_size, grid_size, 3)
    return GridPermLossFn.apply(out, control_points, 0)


def control_points_permute_closed_reg_loss(output, control_points, grid_size_x, grid_size_y):
    """closed in u: min over (cyclic shifts along u) x (4 flips)"""
    out = output.view(output.shape[0], grid_size_x, grid_size_y, 3)
    if grid_size_x == grid_size_y:
        return GridPermLossFn.apply(out, control_points, 1)
    # (rectangular grids: no caller of the path; the reference's formulation as torch expressions)
    cands = torch.cat([all_permutations_half(roll(control_points, i, 1)) for i in range(grid_size_y)], 1)
    return _best_permutation(out, cands, grid_size_x * grid_size_y * 3)


def control_points_loss(output, control_points, grid_size):
    out = output.view(output.shape[0], grid_size, grid_size, 3)
    return ((out - control_points) ** 2).sum((1, 2, 3)).mean() / (grid_size * grid_size * 3)


def laplacian_loss(output, gt, dist_type="l2"):
    """difference of the 4-neighbour Laplacians (zero padding) of two (B,g,g,3) grids: 5-point stencil kernel
    (csrc/gridloss.cu) instead of the reference's two 3x3 convolutions"""
    return GridLaplacianLossFn.apply(output, gt, 0 if dist_type == "l2" else 1)


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
