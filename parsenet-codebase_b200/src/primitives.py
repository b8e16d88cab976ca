"""Drop-in for reference src/primitives.py: ResidualLoss (:18) / ComputePrimitiveDistance (:47).  All analytic
primitives of a shape go through ONE pn_residual_fwd launch (value + parameter Jacobian); splines use the Chamfer
kernels.  SaveParameters (:209, result serialisation) is out of scope."""
import numpy as np
import torch

from pnb200.fitting import TYPE_ID, ResidualFn
from src.guard import guard_sqrt
from src.utils import chamfer_distance_single_shape

EPS = np.finfo(np.float32).eps


_TYPE_CONST = {}


def _type_const(kind, dev):
    """0-d int32 device constant holding the kernel's type id (uploading a fresh id list per call would block)"""
    key = (kind, str(dev))
    t = _TYPE_CONST.get(key)
    if t is None:
        t = _TYPE_CONST[key] = torch.tensor(TYPE_ID[kind], dtype=torch.int32, device=dev)
    return t


def _pack(kind, params):
    """-> (8,) parameter row in the kernel's layout"""
    z = params[0].new_zeros
    if kind == "plane":
        a, d = params
        return torch.cat([a.reshape(3), d.reshape(1), z(4)])
    if kind == "sphere":
        c, r = params
        return torch.cat([c.reshape(3), r.reshape(1), z(4)])
    if kind == "cylinder":
        a, c, r = params
        return torch.cat([a.reshape(3), c.reshape(3), r.reshape(1), z(1)])
    apex, a, th = params
    return torch.cat([apex.reshape(3), a.reshape(3), th.reshape(1), z(1)])


def _closed_form(kind, points, params, sqrt, reduce):
    """squared point-to-primitive distances of reference src/primitives.py:100-195 (per point; optional guard_sqrt / mean)"""
    if kind == "plane":
        a, d = params
        dist = ((points @ a.reshape(3, 1) - d) ** 2).sum(1)
    elif kind == "sphere":
        c, r = params
        dist = (torch.norm(points - c.reshape(1, 3), p=2, dim=1) - r) ** 2
    elif kind == "cylinder":
        axis, c, r = params
        v = points - c.reshape(1, 3)
        radial2 = torch.clamp((v * v).sum(1) - (v @ axis.reshape(3, 1))[:, 0] ** 2, min=1e-5)
        dist = (torch.sqrt(radial2) - r) ** 2
    else:
        apex, axis, theta = params
        v = points - apex.reshape(1, 3) + 1e-8
        mod = torch.norm(v, dim=1, p=2)
        cosang = torch.clamp((v @ axis.reshape(3, 1))[:, 0] / (mod + 1e-7), min=-0.999, max=0.999)
        off = torch.clamp((torch.acos(cosang) - theta).abs(), max=3.142 / 2.0)
        dist = (mod * torch.sin(off)) ** 2
    if sqrt:
        dist = guard_sqrt(dist)
    return dist.mean() if reduce else dist


class ComputePrimitiveDistance:
    def __init__(self, reduce=True, one_side=False):
        self.reduce, self.one_side = reduce, one_side

    def _analytic(self, kind, points, params, sqrt):
        if sqrt or not self.reduce:
            # evaluation-time variants (per-point vectors, guarded square roots; reference :100-195 with sqrt / reduce
            # flags): off the training hot path, closed forms as plain torch expressions on the caller's device
            return _closed_form(kind, points, params, sqrt, self.reduce)
        par = _pack(kind, [p.float() for p in params]).unsqueeze(0)
        seg = torch.zeros((points.shape[0],), dtype=torch.int32, device=points.device)
        typ = torch.full((1,), TYPE_ID[kind], dtype=torch.int32, device=points.device)
        return ResidualFn.apply(par, points.contiguous().float(), seg, typ)[0]

    def distance_from_plane(self, points, params, sqrt=False):
        return self._analytic("plane", points, params, sqrt)

    def distance_from_sphere(self, points, params, sqrt=False):
        return self._analytic("sphere", points, params, sqrt)

    def distance_from_cylinder(self, points, params, sqrt=False):
        return self._analytic("cylinder", points, params, sqrt)

    def distance_from_cone(self, points, params, sqrt=False):
        return self._analytic("cone", points, params, sqrt)

    def distance_from_torus(self, points, params, sqrt=False):
        axis, center, R, r = params
        axis = axis.reshape(3, 1) / torch.norm(axis, p=2)
        v = points - center.reshape(1, 3)
        z = v @ axis
        x = guard_sqrt((v ** 2).sum(1, keepdim=True) - z ** 2)
        d = torch.min((guard_sqrt((x - R) ** 2 + z ** 2) - r) ** 2, (guard_sqrt((x + R) ** 2 + z ** 2) - r) ** 2).squeeze()
        d = guard_sqrt(d) if sqrt else d
        return d.mean() if self.reduce else d

    def distance_from_bspline(self, points, params, sqrt=False):
        return chamfer_distance_single_shape(params[0][0], points, one_side=self.one_side, sqrt=sqrt,
                                             reduce=self.reduce)


class ResidualLoss:
    def __init__(self, reduce=True, one_side=False):
        self.cp = ComputePrimitiveDistance(reduce, one_side=one_side)
        self.routines = {"torus": self.cp.distance_from_torus, "sphere": self.cp.distance_from_sphere,
                         "cylinder": self.cp.distance_from_cylinder, "cone": self.cp.distance_from_cone,
                         "plane": self.cp.distance_from_plane, "closed-spline": self.cp.distance_from_bspline,
                         "open-spline": self.cp.distance_from_bspline}

    def residual_loss(self, Points, parameters, sqrt=False):
        """Points: {key: (m,3) gt points}, parameters: {key: [kind, params...] | None} -> {key: [kind, distance]}.
        All analytic segments are batched into one kernel launch."""
        out = {}
        keys = [k for k, v in parameters.items() if v is not None]
        analytic = [k for k in keys if parameters[k][0] in TYPE_ID]
        if analytic and not sqrt and self.cp.reduce:
            dev = Points[analytic[0]].device
            par = torch.stack([_pack(parameters[k][0], [p.float() for p in parameters[k][1:]]) for k in analytic], 0)
            pts = torch.cat([Points[k] for k in analytic], 0).contiguous().float()
            seg = torch.cat([torch.full((Points[k].shape[0],), i, dtype=torch.int32, device=dev)
                             for i, k in enumerate(analytic)])
            typ = torch.stack([_type_const(parameters[k][0], dev) for k in analytic])
            dist = ResidualFn.apply(par, pts, seg, typ)
            for i, k in enumerate(analytic):
                out[k] = [parameters[k][0], dist[i]]
        for k in keys:
            if k not in out:
                v = parameters[k]
                out[k] = [v[0], self.routines[v[0]](points=Points[k], params=v[1:], sqrt=sqrt)]
        return {k: out[k] for k in keys}


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
