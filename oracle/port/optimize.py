"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's Kronecker post-fit optimisers.

  fit_bezier_surface_fit_kronecker    src/approximation.py:338-364
  optimize_open_spline_kronecker      src/primitive_forward.py:229-296   (deform=False)
  optimize_close_spline_kronecker     src/primitive_forward.py:153-226   (deform=False)
  DrawSurfs.boundary_parameterization / regular_parameterization   src/curve_utils.py:200-221

PARITY UNPINNED for the two optimisers: the reference evaluates its surfaces with geomdl (`create_geomdl_surface(...).evaluate_list`,
src/approximation.py:72-88) and solves the assignment with lapsolver; neither package (geomdl 5.x, lapsolver 1.x, unpinned in the
reference's README) exists in this image and there is no network, so the reference functions cannot be run here to dump golden
vectors.  What is restated: geomdl's evaluate_list is the textbook tensor-product B-spline surface S(u,v) = sum_ab N_a(u) N_b(v) P_ab
over the knot vectors the reference hands it (its own `uniform_knot_bspline_`), evaluated below with the reference's own
`basis_function_one` recursion (pinned: tests/golden/losses.npz basis matrices, and against scipy.interpolate.BSpline in
tests/test_cpu_host_logic.py); lapsolver.solve_dense returns the optimal assignment, as scipy.optimize.linear_sum_assignment does.
`fit_bezier_surface_fit_kronecker` IS pinned (pure numpy in the reference: tests/golden/kronecker.npz from the unmodified function).
"""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from . import fitting as F


def boundary_parameterization(grid_u):
    """curve_utils.py:211-221"""
    u = np.arange(grid_u)
    rows = [np.stack([np.zeros(grid_u), u], 1),
            np.stack([np.arange(1, grid_u), np.zeros(grid_u - 1)], 1),
            np.stack([np.arange(1, grid_u), np.ones(grid_u - 1) * (grid_u - 1)], 1),
            np.stack([np.ones(grid_u - 2) * (grid_u - 1), np.arange(1, grid_u - 1)], 1)]
    return np.concatenate(rows, 0) / (grid_u - 1)


def regular_parameterization(grid_u, grid_v):
    """curve_utils.py:200-209"""
    xv, yv = np.meshgrid(np.linspace(0, 1, grid_u), np.linspace(0, 1, grid_v))
    return np.concatenate([xv.transpose().reshape(-1, 1), yv.transpose().reshape(-1, 1)], 1)


def basis_rows(parameters, n_u, n_v, deg_u, deg_v):
    """primitive_forward.py:202-208 / approximation.py:55-70: one (nu, nv) pair per sample, stacked to (M, n_u), (M, n_v)"""
    ku, kv = F._clamped_knots(n_u, deg_u), F._clamped_knots(n_v, deg_v)
    NU = np.array([[F.basis_value(deg_u, ku, j, float(p[0])) for j in range(n_u)] for p in parameters])
    NV = np.array([[F.basis_value(deg_v, kv, j, float(p[1])) for j in range(n_v)] for p in parameters])
    return NU, NV


def evaluate_list(control_points, parameters, deg_u, deg_v):
    """what geomdl's Surface.evaluate_list computes for the surface of approximation.py:72-88: control_points (cu, cv, 3)"""
    cu, cv, _ = control_points.shape
    NU, NV = basis_rows(parameters, cu, cv, deg_u, deg_v)
    return np.einsum("ia,ib,abc->ic", NU, NV, control_points)


def fit_bezier_surface_fit_kronecker(points, basis_u, basis_v):
    """approximation.py:338-364: A_i = u_i^T v_i flattened, one lstsq per coordinate"""
    N = basis_u.shape[0]
    A = np.stack([np.matmul(basis_u[i:i + 1].T, basis_v[i:i + 1]) for i in range(N)], 0).reshape(N, -1)
    n1 = basis_v.shape[1]
    return np.stack([np.linalg.lstsq(A, points[:, i], rcond=None)[0].reshape(n1, n1) for i in range(3)], 2)


def _optimize(input_points, control_points, deg_old, boundary, n_input, subsample, new_cp_size, new_degree, rng):
    bpar = boundary_parameterization(boundary)
    parameters = np.concatenate([rng.random((1600 - bpar.shape[0], 2)), bpar], 0)
    points = evaluate_list(control_points, parameters, deg_old, deg_old)
    inp = F.up_sample_points_in_range(torch.from_numpy(input_points), n_input[0], n_input[1], rng=rng)
    if subsample:
        L = rng.choice(np.arange(inp.shape[0]), 1600, replace=False)
        inp = inp[L]
    inp = inp.numpy()
    dist = np.linalg.norm(np.expand_dims(points, 1) - np.expand_dims(inp, 0), axis=2)
    _, cids = linear_sum_assignment(dist)
    matched = inp[cids]
    NU, NV = basis_rows(parameters, new_cp_size, new_cp_size, new_degree, new_degree)
    new_cp = fit_bezier_surface_fit_kronecker(matched, NU, NV)
    return evaluate_list(new_cp, regular_parameterization(30, 30), new_degree, new_degree).astype(np.float32), new_cp


def optimize_open_spline_kronecker(input_points, control_points, new_cp_size=10, new_degree=2, rng=np.random):
    """primitive_forward.py:229-296, deform=False: input_points (n,3) float32, control_points (400,3) -> (900,3)"""
    return _optimize(input_points, control_points.reshape(20, 20, 3), 3, 20, (1600, 2000), True, new_cp_size, new_degree, rng)


def optimize_close_spline_kronecker(input_points, control_points, new_cp_size=10, new_degree=3, rng=np.random):
    """primitive_forward.py:153-226, deform=False: control_points (21,20,3) -> (930,3) (first grid row appended again)"""
    pts, cp = _optimize(input_points, control_points.reshape(21, 20, 3), 3, 30, (2000, 2100), False, new_cp_size, new_degree, rng)
    pts = pts.reshape(30, 30, 3)
    return np.concatenate([pts, pts[0:1]], 0).reshape(930, 3), cp
