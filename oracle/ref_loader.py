"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (read-only at /root/reference).

Only usable in the build container (the GPU box has no /root/reference).  It is used by
`oracle/make_golden.py` to dump golden vectors into `tests/golden/`, which is what pins the
CPU restatement in `oracle/port/`.  Nothing in the product package imports this.

What it does (no edits to the reference, only process-local monkeypatches):
  * inserts MagicMock stubs for absent viz/IO wheels (open3d, lapsolver, lap, geomdl, matplotlib,
    h5py, configobj, trimesh, transforms3d, tensorboard_logger, ipdb);
  * `lapsolver.solve_dense` -> scipy.optimize.linear_sum_assignment;
  * `torch.matrix_rank`, `torch.eig`, `torch.svd(some=)`, `torch.qr` shims for torch>=2;
  * CPU mode: `Tensor.cuda`/`Module.cuda` -> identity, `torch.device('cuda')` -> cpu,
    `Tensor.get_device` -> 'cpu'-compatible, `torch.cuda.FloatTensor` -> torch.FloatTensor.
"""
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

REF_ROOT = "/root/reference"
_LOADED = False


def _stub(name, **attrs):
    m = MagicMock(name=name)
    m.__name__ = name
    m.__all__ = []
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(cpu=True):
    global _LOADED
    if _LOADED:
        return
    _LOADED = True
    for name in ["open3d", "open3d.utility", "open3d.geometry", "open3d.visualization",
                 "lap", "geomdl", "geomdl.fitting", "geomdl.BSpline", "geomdl.utilities",
                 "geomdl.tessellate", "geomdl.visualization", "geomdl.visualization.VisMPL",
                 "geomdl.exchange", "geomdl.operations", "geomdl.NURBS", "geomdl.helpers",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "h5py", "configobj", "trimesh", "transforms3d",
                 "transforms3d.affines", "transforms3d.euler", "tensorboard_logger", "ipdb",
                 "skimage", "skimage.measure"]:
        _stub(name)
    sys.modules["open3d"].__all__ = ["utility", "geometry", "visualization"]
    sys.modules["open3d"].utility = sys.modules["open3d.utility"]
    sys.modules["open3d"].geometry = sys.modules["open3d.geometry"]
    sys.modules["open3d"].visualization = sys.modules["open3d.visualization"]

    from scipy.optimize import linear_sum_assignment

    def solve_dense(cost):
        r, c = linear_sum_assignment(np.asarray(cost))
        return r, c

    _stub("lapsolver", solve_dense=solve_dense)

    # ---- removed torch APIs (reference pins torch==1.2.0, environment.yml:20) ----
    torch.matrix_rank = torch.linalg.matrix_rank

    def _eig(a, eigenvectors=False):
        w, v = torch.linalg.eig(a)
        vals = torch.stack([w.real, w.imag], 1)
        return vals, v.real

    torch.eig = _eig

    def _svd(a, some=True, compute_uv=True):
        u, s, vh = torch.linalg.svd(a, full_matrices=not some)
        return u, s, vh.transpose(-2, -1).conj()

    torch.svd = _svd

    def _qr(a, some=True):
        return torch.linalg.qr(a, mode="reduced" if some else "complete")

    torch.qr = _qr

    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _orig_device = torch.device

        class _DevMeta(type):
            def __instancecheck__(cls, inst):
                return isinstance(inst, _orig_device)

        class _Dev(metaclass=_DevMeta):
            def __new__(cls, *a, **k):
                if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                    return _orig_device("cpu")
                return _orig_device(*a, **k)

        torch.device = _Dev
        torch.Tensor.get_device = lambda self: "cpu"
        torch.get_device = lambda t: "cpu"
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.empty_cache = lambda: None
        _orig_eye = torch.eye

        def _eye(*a, **k):
            k.pop("device", None) if k.get("device", None) == "cpu" else None
            return _orig_eye(*a, **k)

        torch.eye = _eye

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def ref(module):
    """import a reference module, e.g. ref('src.PointNet')"""
    install()
    import importlib
    return importlib.import_module(module)
