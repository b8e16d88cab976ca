"""Drop-in for the hot-path part of reference src/primitive_forward.py: Fit (:418, fit_*_torch :708-843),
fit_one_shape_torch (:925), forward_pass_open_spline (:34), forward_closed_splines (:347) and the SplineNet loaders
(:88, :400), and the Kronecker post-fit optimisers optimize_open_spline_kronecker (:229) / optimize_close_spline_kronecker
(:153) without the ARAP pre-deformation.  ARAP (open3d), the geomdl `approximate_surface` optimisers (:105-150,
:299-344) and Fit.sample_* are out of scope (`deform=True`, `if_optimize=True`, `sample_points=True`, `eval=True` raise
NotImplementedError)."""
import numpy as np
import torch

from pnb200 import fitting as _f
from src.fitting_utils import LeastSquares, customsvd, sample_points_from_control_points_, standardize_points_torch
from src.guard import guard_sqrt

EPS = float(np.finfo(np.float32).eps)
CLOSED_IDS, OPEN_IDS = (0, 6, 7, 9), (2, 8)
from pnb200.fitstage import STATS          # host counters of fitted segments, shared with the batched path (bench.py reports them)


# ------------------------------------------------------------------------------------------------ SplineNet passes
def _unstandardize(x, scale, R, mean):
    """(n,3) standardised -> original frame: scale per axis, inverse rotation, shift"""
    t = x * scale.reshape(1, 3)
    Rinv = getattr(R, "_pn_inv", None)
    if Rinv is None:
        Rinv = torch.inverse(R)
    return (Rinv @ t.t()).t() + mean


# ------------------------------------------------------------------------------------------------ post-fit optimisation
def _boundary_parameterization(grid_u):
    """the 4 (grid_u - 1) boundary nodes of the regular grid_u x grid_u parameter grid, in the reference's order
    (curve_utils.py:211-221: u = 0 edge, v = 0 edge, v = 1 edge, u = 1 edge without its corners)"""
    g, last = np.arange(grid_u, dtype=np.float64), float(grid_u - 1)
    edges = [np.stack([np.zeros(grid_u), g], 1),
             np.stack([g[1:], np.zeros(grid_u - 1)], 1),
             np.stack([g[1:], np.full(grid_u - 1, last)], 1),
             np.stack([np.full(grid_u - 2, last), g[1:-1]], 1)]
    return np.concatenate(edges, 0) / last


def _regular_parameterization(grid_u, grid_v):
    """(grid_u * grid_v, 2) nodes of linspace(0,1) x linspace(0,1), u-major (curve_utils.py:200-209)"""
    u, v = np.meshgrid(np.linspace(0, 1, grid_u), np.linspace(0, 1, grid_v), indexing="ij")
    return np.stack([u.reshape(-1), v.reshape(-1)], 1)


def _optimize_spline_kronecker(input_points_, control_points, cp_u, cp_v, boundary, n_input, new_cp_size, new_degree,
                               subsample, matcher):
    """shared body of the two Kronecker optimisers (reference src/primitive_forward.py:153-296, deform=False):
    1600 surface samples of the predicted spline (boundary nodes + random parameters) are matched one-to-one with 1600
    (up-sampled) input points, a new_cp_size^2 control grid of degree new_degree is fitted to the matched points at the same
    parameters, and the new surface is sampled on the regular 30 x 30 grid.  Device work: up-sampling (kNN kernel), the two
    surface evaluations and the 100 x 100 normal-equation solve (csrc/kronfit.cu), the distance matrix.  The assignment is
    the host's (scipy linear_sum_assignment instead of lapsolver: same optimum) unless matcher == "nearest" (each surface
    sample takes its nearest input point: the Chamfer argmin kernel, no read-back)."""
    from scipy.optimize import linear_sum_assignment
    from src.approximation import basis_rows
    from src.fitting_utils import up_sample_points_torch_in_range
    from src.utils import chamfer_argmin
    dev = input_points_.device
    bpar = _boundary_parameterization(boundary)
    parameters = np.concatenate([np.random.random((1600 - bpar.shape[0], 2)), bpar], 0)
    NU, NV = basis_rows(parameters, cp_u, cp_v, 3, 3)
    NUd, NVd = torch.from_numpy(NU).to(dev).unsqueeze(0), torch.from_numpy(NV).to(dev).unsqueeze(0)
    cp = control_points[0].detach().reshape(cp_u, cp_v, 3)
    points = _f.kron_eval(cp, NUd, NVd)[0]                                    # (1600, 3) float64, on the predicted surface
    inp = up_sample_points_torch_in_range(input_points_[0].detach(), n_input[0], n_input[1])
    if subsample:
        L = np.random.choice(np.arange(inp.shape[0]), 1600, replace=False)
        inp = inp[torch.from_numpy(L).to(dev)]
    if matcher == "hungarian":
        dist = torch.cdist(points, inp.double()).cpu().numpy()
        _, cids = linear_sum_assignment(dist)
        matched = inp[torch.from_numpy(cids).to(dev)]
    elif matcher == "nearest":
        matched = inp[chamfer_argmin(points.float().unsqueeze(0), inp.float().unsqueeze(0))[0].long()]
    else:
        raise ValueError(f"unknown matcher {matcher!r}")
    NU2, NV2 = basis_rows(parameters, new_cp_size, new_cp_size, new_degree, new_degree)
    NU2d, NV2d = torch.from_numpy(NU2).to(dev).unsqueeze(0), torch.from_numpy(NV2).to(dev).unsqueeze(0)
    new_cp, flag = _f.kron_fit(matched.double().unsqueeze(0), NU2d, NV2d)
    if bool(flag[0]):
        from src.approximation import fit_bezier_surface_fit_kronecker
        new_cp = fit_bezier_surface_fit_kronecker(matched.double(), NU2d[0], NV2d[0]).unsqueeze(0)
    RU, RV = basis_rows(_regular_parameterization(30, 30), new_cp_size, new_cp_size, new_degree, new_degree)
    out = _f.kron_eval(new_cp[0], torch.from_numpy(RU).to(dev).unsqueeze(0), torch.from_numpy(RV).to(dev).unsqueeze(0))
    return out[0].float(), new_cp[0]


def optimize_open_spline_kronecker(reconstructed_points, input_points_, control_points, new_cp_size=10, new_degree=2,
                                   deform=False, matcher="hungarian"):
    """reference src/primitive_forward.py:229-296.  control_points (1, 400, 3) of the predicted 20 x 20 cubic surface, input
    points (1, n, 3) -> (1, 900, 3) samples of the re-fitted surface.  Consumes np.random like the reference (the random
    parameters, then the 1600-point subsample)."""
    if deform:
        raise NotImplementedError("ARAP pre-deformation (open3d) is outside the hot path: call with deform=False")
    pts, _ = _optimize_spline_kronecker(input_points_, control_points, 20, 20, 20, (1600, 2000), new_cp_size, new_degree,
                                        True, matcher)
    return pts.unsqueeze(0)


def optimize_close_spline_kronecker(reconstructed_points, input_points_, control_points, new_cp_size=10, new_degree=3,
                                    deform=True, matcher="hungarian"):
    """reference src/primitive_forward.py:153-226.  control_points (1, 21, 20, 3) (the closed grid with its first row
    repeated), input points (1, n, 3) -> (1, 930, 3): the 30 x 30 samples of the re-fitted surface plus the first row again.
    The reference's default deform=True needs ARAP (open3d, out of scope): pass deform=False."""
    if deform:
        raise NotImplementedError("ARAP pre-deformation (open3d) is outside the hot path: call with deform=False")
    pts, _ = _optimize_spline_kronecker(input_points_, control_points, 21, 20, 30, (2000, 2100), new_cp_size, new_degree,
                                        False, matcher)
    pts = pts.reshape(30, 30, 3)
    return torch.cat([pts, pts[0:1]], 0).reshape(1, 930, 3)


def forward_pass_open_spline(input_points_, control_decoder, nu, nv, viz=False, weights=None, if_optimize=True):
    """points (1,n,3) + membership weights (n,1) -> reconstructed surface samples (1, g*g, 3) in the input frame"""
    if if_optimize or viz:
        raise NotImplementedError("post-fit optimisation / visualisation are outside the hot path")
    dev = input_points_.device
    nu, nv = nu.to(dev), nv.to(dev)
    with torch.no_grad():
        pts, scales, means, RS = standardize_points_torch(input_points_, weights)
    B = pts.shape[0]
    output = control_decoder(pts.permute(0, 2, 1), weights.t())
    rec = sample_points_from_control_points_(nu, nv, output, B)
    rec = torch.stack([_unstandardize(rec[b], scales[b], RS[b], means[b]) for b in range(B)], 0)
    return rec, rec


def forward_closed_splines(input_points_, control_decoder, nu, nv, viz=False, weights=None, if_optimize=True):
    """closed (in u) spline: as above, then the first grid row is appended again -> (1, 930, 3)"""
    if (if_optimize and input_points_.shape[1] > 200) or viz:
        raise NotImplementedError("post-fit optimisation / visualisation are outside the hot path")
    dev = input_points_.device
    nu, nv = nu.to(dev), nv.to(dev)
    with torch.no_grad():
        pts, scales, means, RS = standardize_points_torch(input_points_, weights)
    B = pts.shape[0]
    output = control_decoder(pts.permute(0, 2, 1), weights.t())
    rec = sample_points_from_control_points_(nu, nv, output, B)
    g = nu.shape[0]
    closed = []
    for b in range(B):
        t = _unstandardize(rec[b], scales[b], RS[b], means[b]).reshape(g, g, 3)
        closed.append(torch.cat([t, t[0:1]], 0))
    rec = torch.stack(closed, 0).reshape(1, (g + 1) * g, 3)
    return rec, None, rec


def _load_splinenet(modelname, mode):
    from src.model import DGCNNControlPoints
    net = DGCNNControlPoints(20, num_points=10, mode=mode)
    state = torch.load(modelname, map_location="cpu")
    net.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in state.items()})
    # The reference parks both SplineNets on cuda:1 whenever a second GPU exists (primitive_forward.py:100, :411) because
    # its training script keeps the fit stage on `alt_gpu`.  Under one-process-per-GPU data parallelism that would pile
    # every rank's copy onto GPU 1 and hand cuda:<rank> activations to cuda:1 weights, so the default here is the
    # process's current device; Evaluation.fitting_loss moves the decoders to the device of its inputs if they differ.
    # PN_SPLINENET_DEVICE=reference restores the reference placement.
    import os
    if os.environ.get("PN_SPLINENET_DEVICE", "") == "reference":
        net.cuda(1 if torch.cuda.device_count() > 1 else 0)
    else:
        net.cuda(torch.cuda.current_device())
    net.eval()
    return net


def initialize_open_spline_model(modelname, mode):
    return _load_splinenet(modelname, mode)


def initialize_closed_spline_model(modelname, mode):
    return _load_splinenet(modelname, mode)


from src._fallthrough import ReferenceMethods as _ReferenceMethods  # noqa: E402


# ------------------------------------------------------------------------------------------------ primitive fits
class Fit(_ReferenceMethods):
    """Weighted fits of plane / sphere / cylinder / cone to (points, normals, weights); same outputs as the
    reference.  Each call computes the weighted moments with pn_fit_moments_* and solves the 3x3 problems with the
    reference's gradient conventions (pnb200.fitting).  The numpy fits and the surface samplers (`fit_*_numpy`,
    `sample_*`, outside the hot path) resolve to the reference's own methods when PARSENET_REFERENCE_SRC is set."""
    _reference_module = "primitive_forward"

    def __init__(self):
        self.lstsq = LeastSquares().lstsq
        self.parameters = {}

    @staticmethod
    def _moments(points, normals, weights):
        W = weights.reshape(-1, 1).contiguous().float()
        nr = None if normals is None else normals.contiguous().float()
        return _f.MomentsFn.apply(W, points.contiguous().float(), nr, 0, 1, points.shape[0], 0.0)

    def fit_plane_torch(self, points, normals, weights, ids=0, show_warning=False):
        a, d = _f.fit_planes(self._moments(points, None, weights))
        return a.float().reshape(1, 3), d.float().reshape(())

    def fit_sphere_torch(self, points, normals, weights, ids=0, show_warning=False):
        c, r = _f.fit_spheres(self._moments(points, None, weights), points.shape[0])
        return c.float().reshape(1, 3), r.float().reshape(())

    def fit_cylinder_torch(self, points, normals, weights, ids=0, show_warning=False):
        a, c, r = _f.fit_cylinders(self._moments(points, normals, weights), points.shape[0])
        return a.float().reshape(3, 1), c.float().reshape(1, 3), r.float().reshape(())

    def fit_cone_torch(self, points, normals, weights, ids=0, show_warning=False):
        mom = self._moments(points, normals, weights)
        apex, axis, degenerate = _f.fit_cone_apex_axis(mom, points.shape[0])
        if bool(degenerate[0]):
            dev = points.device
            return (torch.zeros((1, 3), device=dev), torch.tensor([[1.0, 0.0, 0.0]], device=dev),
                    torch.zeros(1, device=dev))
        apex, axis = apex.float(), axis.float()
        theta = _f.cone_theta(points, weights.reshape(-1, 1), apex[0], axis[0])
        return apex.reshape(3, 1), axis.reshape(1, 3), theta


def fit_one_shape_torch(data, fitter, weights, bw, eval=False, sample_points=False, if_optimize=False,
                        if_visualize=False):
    """Training-mode fit of every matched segment of one shape (reference :925-1047).
    data: list of [points (N,3), normals (N,3), primitive id, gt points (m,3), None, (weight column, label index)]
    — in training mode every entry carries the SAME full point set and the segment is defined by its weight column.
    All analytic primitives share ONE moment pass over the decimated points; splines go through the SplineNets."""
    if eval or sample_points or if_optimize or if_visualize:
        raise NotImplementedError("fit_one_shape_torch: only the training path (eval=False) is on the hot path")
    fitter.fitting.parameters = {}
    gt_points, recon = {}, []
    if not data:
        return gt_points, recon
    points, normals = data[0][0].contiguous(), data[0][1].contiguous()
    N = points.shape[0]
    n_half = (N + 1) // 2            # points[0::2]
    n_quarter = (n_half + 1) // 2    # ...[0::2] again for analytic primitives
    W = weights if (weights.shape[1] == 1 or weights.stride(1) == 1) else weights.contiguous()
    mom = None
    spline_count = 0
    plan = []
    for d in data:
        prim = int(np.asarray(d[2]).reshape(-1)[0])
        col, label_index = d[5]
        if prim in CLOSED_IDS + OPEN_IDS:
            spline_count += 1
            if spline_count > 4:
                plan.append((None, d, col, label_index))
                continue
            m = n_half
        else:
            m = n_quarter
        if m < 20 or (prim in CLOSED_IDS + OPEN_IDS and m < 100):
            plan.append((None, d, col, label_index))
            continue
        plan.append((prim, d, col, label_index))
    analytic = [(prim, col) for prim, _, col, _ in plan if prim in (1, 3, 4, 5)]
    fits = {}
    if analytic:
        mom = _f.MomentsFn.apply(W, points, normals, 0, 4, n_quarter, EPS)
        for kind_id, fn in ((1, "plane"), (5, "sphere"), (4, "cylinder"), (3, "cone")):
            cols = [c for p, c in analytic if p == kind_id]
            if not cols:
                continue
            sub = torch.stack([mom[c] for c in cols], 0)         # (no index upload: that would block the host)
            if fn == "plane":
                a, dd = _f.fit_planes(sub)
                for i, c in enumerate(cols):
                    fits[c] = ["plane", a[i].float().reshape(3, 1), dd[i].float()]
            elif fn == "sphere":
                ce, r = _f.fit_spheres(sub, n_quarter)
                for i, c in enumerate(cols):
                    fits[c] = ["sphere", ce[i].float().reshape(1, 3), r[i].float()]
            elif fn == "cylinder":
                a, ce, r = _f.fit_cylinders(sub, n_quarter)
                for i, c in enumerate(cols):
                    fits[c] = ["cylinder", a[i].float().reshape(3, 1), ce[i].float().reshape(1, 3), r[i].float()]
            else:
                apex, axis, deg = _f.fit_cone_apex_axis(sub, n_quarter)
                pq = points[0::4]
                dev = points.device
                x_axis = _f._dev_const("x_axis", torch.tensor([1.0, 0.0, 0.0]), dev)
                for i, c in enumerate(cols):
                    wq = (W[0::4, c:c + 1] + EPS)
                    th = _f.cone_theta(pq, wq, apex[i].float(), axis[i].float())
                    # ill-conditioned normals (cond > 1e5): constant zero apex / x axis / zero angle (reference
                    # :818-823), selected on the device so that no flag has to be read back
                    bad = deg[i]
                    ap = torch.where(bad, torch.zeros_like(apex[i]), apex[i]).float()
                    ax = torch.where(bad, x_axis.to(axis.dtype), axis[i]).float()
                    th = torch.where(bad, torch.zeros_like(th), th)
                    fits[c] = ["cone", ap.reshape(1, 3), ax.reshape(3, 1), th]
    for prim, d, col, label_index in plan:
        if prim is None:
            recon.append(None)
            gt_points[label_index] = None
            fitter.fitting.parameters[label_index] = None
            continue
        if prim in CLOSED_IDS + OPEN_IDS:
            pts_h = points[0::2]
            w_h = W[0::2, col:col + 1] + EPS
            if prim in CLOSED_IDS:
                rec = fitter.forward_pass_closed_spline(pts_h, weights=w_h, ids=label_index, if_optimize=False)
                STATS["closed_spline_fits"] += 1
            else:
                rec = fitter.forward_pass_open_spline(pts_h, weights=w_h, ids=label_index, if_optimize=False)
                STATS["open_spline_fits"] += 1
            recon.append(rec)
        else:
            fitter.fitting.parameters[label_index] = fits[col]
            STATS["analytic_fits"] += 1
            recon.append(None)
        gt_points[label_index] = d[3]
    return gt_points, recon



from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
