"""Drop-in for the hot-path part of reference src/primitive_forward.py: Fit (:418, fit_*_torch :708-843),
fit_one_shape_torch (:925), forward_pass_open_spline (:34), forward_closed_splines (:347) and the SplineNet loaders
(:88, :400).  The geomdl / ARAP / Hungarian post-fit optimisers (:105-344) and Fit.sample_* are out of scope
(`if_optimize=True`, `sample_points=True`, `eval=True` raise NotImplementedError)."""
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
