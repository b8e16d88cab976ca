"""Drop-in for reference src/fitting_optimization.py FittingModule (:117-242): owns Fit, the two frozen SplineNets and
the 30x20 basis matrices; per-primitive forward_pass_* record `fitting.parameters[ids]` exactly like the reference.
Arap (:32, open3d) is out of scope."""
import numpy as np
import torch

from src.loss import uniform_knot_bspline
from src.primitive_forward import (Fit, forward_closed_splines, forward_pass_open_spline,
                                   initialize_closed_spline_model, initialize_open_spline_model)


class FittingModule:
    def __init__(self, closed_splinenet_path, open_splinenet_path, open_decoder=None, closed_decoder=None):
        self.fitting = Fit()
        self.closed_splinenet_path, self.open_splinenet_path = closed_splinenet_path, open_splinenet_path
        nu, nv = uniform_knot_bspline(20, 20, 3, 3, 30)
        self.nu, self.nv = torch.from_numpy(nu.astype(np.float32)), torch.from_numpy(nv.astype(np.float32))
        self.open_control_decoder = open_decoder if open_decoder is not None else \
            initialize_open_spline_model(open_splinenet_path, 0)
        self.closed_control_decoder = closed_decoder if closed_decoder is not None else \
            initialize_closed_spline_model(closed_splinenet_path, 1)

    def _basis_on(self, device):
        """the 30x20 basis matrices move to the device once (a per-call .to(device) is a blocking pageable copy)"""
        if self.nu.device != device:
            self.nu, self.nv = self.nu.to(device), self.nv.to(device)

    def forward_pass_open_spline(self, points, ids, weights, if_optimize=False):
        self._basis_on(points.device)
        rec = forward_pass_open_spline(points.detach().unsqueeze(0), self.open_control_decoder, self.nu, self.nv,
                                       if_optimize=if_optimize, weights=weights)[1]
        self.fitting.parameters[ids] = ["open-spline", rec]
        return rec

    def forward_pass_closed_spline(self, points, ids, weights, if_optimize=False):
        self._basis_on(points.device)
        rec = forward_closed_splines(points.detach().unsqueeze(0), self.closed_control_decoder, self.nu, self.nv,
                                     if_optimize=if_optimize, weights=weights)[2]
        self.fitting.parameters[ids] = ["closed-spline", rec]
        return rec

    def forward_pass_plane(self, points, normals, weights, ids, sample_points=False):
        axis, distance = self.fitting.fit_plane_torch(points, normals, weights, ids=ids)
        self.fitting.parameters[ids] = ["plane", axis.reshape(3, 1), distance]
        if sample_points:
            raise NotImplementedError("surface sampling is eval-only (outside the hot path)")

    def forward_pass_cone(self, points, normals, weights, ids, sample_points=False):
        apex, axis, theta = self.fitting.fit_cone_torch(points, normals, weights=weights, ids=ids)
        self.fitting.parameters[ids] = ["cone", apex.reshape(1, 3), axis.reshape(3, 1), theta]
        if sample_points:
            raise NotImplementedError("surface sampling is eval-only (outside the hot path)")

    def forward_pass_cylinder(self, points, normals, weights, ids, sample_points=False):
        a, center, radius = self.fitting.fit_cylinder_torch(points, normals, weights, ids=ids)
        self.fitting.parameters[ids] = ["cylinder", a, center, radius]
        if sample_points:
            raise NotImplementedError("surface sampling is eval-only (outside the hot path)")

    def forward_pass_sphere(self, points, normals, weights, ids, sample_points=False):
        center, radius = self.fitting.fit_sphere_torch(points, normals, weights, ids=ids)
        self.fitting.parameters[ids] = ["sphere", center, radius]
        if sample_points:
            raise NotImplementedError("surface sampling is eval-only (outside the hot path)")


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
