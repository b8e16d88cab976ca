"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares, the ctypes table matches
the header, host-side logic (triplet sampling order, batch sharding, basis matrices, matching) behaves like the
reference, and the data-parallel gradient averaging works across 2 gloo processes."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "parsenet_b200.h")).read()
    return sorted(set(re.findall(r"\b(pn_\w+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    from pnb200 import cabi
    syms = _header_symbols()
    assert len(syms) >= 30
    for name in syms:
        assert hasattr(cabi.lib, name), f"{name} declared in include/parsenet_b200.h but not exported"
    assert cabi.lib.pn_abi_version() == 1


def test_ctypes_table_matches_header_arity():
    from pnb200 import cabi
    txt = open(os.path.join(ROOT, "include", "parsenet_b200.h")).read()
    for name, argtypes in cabi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*)\);", txt, re.S)
        assert m, f"{name} bound in cabi.py but missing from the header"
        n_args = len([a for a in m.group(1).split(",") if a.strip()])
        assert n_args == len(argtypes), (name, n_args, len(argtypes))


def test_ops_refuse_cpu_tensors():
    from pnb200 import cabi, ops
    with pytest.raises(cabi.PnError):
        ops.knn_graph(torch.zeros(1, 100, 3), 10)


def test_triplet_sampling_consumes_rng_like_reference_port():
    """same np.random draws in the same order as the oracle port's restatement of segment_loss.py:60-96"""
    from oracle.port import segnet as port
    from pnb200.losses import triplet_sample
    rng = np.random.RandomState(3)
    labels = rng.randint(0, 4, (2, 300))
    np.random.seed(11)
    groups = triplet_sample(labels, 300)
    after_ours = np.random.rand()
    np.random.seed(11)
    port.triplet_loss(torch.randn(2, 8, 300), labels, 1.0)
    after_port = np.random.rand()
    assert after_ours == after_port
    assert sum(g[1].shape[0] for g in groups) > 0 and all(g[1].max() < 2 * 300 for g in groups)


def test_shard_batch_covers_everything():
    from pnb200.parallel import shard_batch
    for n, w in [(16, 8), (16, 3), (5, 8)]:
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_basis_matrices_partition_of_unity():
    from src.loss import uniform_knot_bspline
    nu, nv = uniform_knot_bspline(20, 20, 3, 3, 30)
    assert nu.shape == (30, 20) and np.allclose(nu.sum(1), 1.0) and (nu != 0).sum(1).max() <= 4 and (nu >= 0).all()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "losses.npz"))
    np.testing.assert_allclose(nu, golden["nu"], atol=1e-15)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "parsenet-codebase_b200"))
import torch, torch.distributed as dist
from pnb200.parallel import allreduce_mean_grads, shard_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
lin = torch.nn.Linear(6, 3)
data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
lo, hi = shard_batch(8, rank, world)
(lin(data[lo:hi]).pow(2).sum() / 8 * world).backward()      # per-rank mean of its shard, scaled to the global mean
allreduce_mean_grads(list(lin.parameters()), world)
ref = torch.nn.Linear(6, 3); ref.load_state_dict(lin.state_dict())
(ref(data).pow(2).sum() / 8).backward()
ok = all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(lin.parameters(), ref.parameters()))
print("RANK", rank, "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
'''


def test_two_process_gloo_gradient_average(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", str(script), ROOT]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("OK") == 2


# ------------------------------------------------------------------------------------------ drop-in completeness
_FALLTHROUGH_SCRIPT = r"""
import os, sys
from unittest.mock import MagicMock
ROOT, PKG, REF = sys.argv[1:4]
sys.path.insert(0, PKG)
# the deployment environment of the reference has these wheels; this container does not
for name in ["open3d", "open3d.utility", "open3d.geometry", "open3d.visualization", "lap", "lapsolver", "geomdl",
             "geomdl.fitting", "geomdl.BSpline", "geomdl.utilities", "geomdl.tessellate", "geomdl.visualization",
             "geomdl.visualization.VisMPL", "geomdl.exchange", "geomdl.operations", "geomdl.NURBS", "geomdl.helpers",
             "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "h5py",
             "configobj", "trimesh", "transforms3d", "transforms3d.affines", "transforms3d.euler",
             "tensorboard_logger", "ipdb", "skimage", "skimage.measure"]:
    m = MagicMock(name=name); m.__name__ = name; m.__all__ = []; m.__path__ = []
    sys.modules[name] = m
o3d = sys.modules["open3d"]; o3d.__all__ = ["utility", "geometry", "visualization"]      # `from open3d import *`
o3d.utility, o3d.geometry, o3d.visualization = (sys.modules["open3d." + n] for n in o3d.__all__)
os.environ["PARSENET_REFERENCE_SRC"] = os.path.join(REF, "src")
import src
assert os.path.realpath(src.__path__[0]).startswith(os.path.realpath(PKG)), src.__path__
# 1. modules the drop-in does not provide come from the reference's files
from src.dataset import generator_iter
from src.dataset_segments import Dataset
import src.dataset, src.augment_utils
assert os.path.realpath(src.dataset.__file__).startswith(os.path.realpath(REF)), src.dataset.__file__
# 2. hot-path names are the drop-in's, whatever the variable says
from src.PointNet import PrimitivesEmbeddingDGCNGn, DGCNNEncoderGn
from src.mean_shift import MeanShift
from src.residual_utils import Evaluation
from src.utils import chamfer_distance, grad_norm
from src.segment_utils import to_one_hot, SIOU_matched_segments
import src.PointNet, src.utils, src.mean_shift, src.segment_utils, src.residual_utils
for mod in (src.PointNet, src.utils, src.mean_shift, src.segment_utils, src.residual_utils):
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(PKG)), mod.__file__
for obj in (PrimitivesEmbeddingDGCNGn, DGCNNEncoderGn, MeanShift, Evaluation, chamfer_distance, to_one_hot,
            SIOU_matched_segments):
    assert obj.__module__.startswith("src."), (obj, obj.__module__)            # not the private reference copy
# 3. names outside the hot path fall through to the reference module of the same name
from src.utils import visualize_uv_maps, visualize_fitted_surface, fit_surface_sample_points
from src.segment_utils import cluster
from src.primitives import SaveParameters
from src.fitting_utils import up_sample_points_torch_in_range, remove_outliers
assert visualize_uv_maps.__module__ == "_parsenet_reference_src.utils", visualize_uv_maps.__module__
assert cluster.__module__ == "_parsenet_reference_src.segment_utils"
# every import statement of the reference's training / inference scripts that targets `src.` now resolves
import ast
for script in ("train_parsenet.py", "train_parsenet_e2e.py", "train_open_splines.py", "train_closed_control_points.py",
               "generate_predictions.py"):
    path = os.path.join(REF, script)
    if not os.path.exists(path):
        continue
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("src."):
            mod = __import__(node.module, fromlist=["x"])
            for a in node.names:
                assert hasattr(mod, a.name), (script, node.module, a.name)
# 3b. methods of a drop-in class outside the hot path bind to the drop-in instance (SURVEY 8b: Fit numpy variants)
import numpy as np
from src.primitive_forward import Fit
f = Fit()
assert type(f).__module__ == "src.primitive_forward"
rng = np.random.RandomState(0)
n = rng.randn(500, 3); n /= np.linalg.norm(n, axis=1, keepdims=True)
c, r = f.fit_sphere_numpy(0.5 * n + 0.1, n, np.ones((500, 1)))
assert np.allclose(c, 0.1, atol=1e-6) and abs(r - 0.5) < 1e-6, (c, r)
assert f.fit_plane_torch.__func__.__module__ == "src.primitive_forward"          # hot-path methods stay ours
try:
    f.no_such_method
    raise SystemExit("expected AttributeError")
except AttributeError:
    pass
# 4. a name nobody defines is still an AttributeError
try:
    src.utils.no_such_function
    raise SystemExit("expected AttributeError")
except AttributeError:
    pass
print("FALLTHROUGH-OK")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference checkout (build container only)")
def test_drop_in_package_resolves_every_src_import_of_the_reference_scripts():
    """`parsenet-codebase_b200/` first on sys.path + PARSENET_REFERENCE_SRC: hot-path names are ours, everything else the
    reference's training / inference scripts import from `src.*` still resolves (drop-in for the callers)"""
    import subprocess
    r = subprocess.run([sys.executable, "-c", _FALLTHROUGH_SCRIPT, ROOT, os.path.join(ROOT, "parsenet-codebase_b200"),
                        "/root/reference"], capture_output=True, text=True, timeout=300)
    assert "FALLTHROUGH-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_missing_name_without_reference_is_attribute_error():
    import subprocess
    code = ("import sys, os; os.environ.pop('PARSENET_REFERENCE_SRC', None); sys.path.insert(0, sys.argv[1]); import src.utils\n"
            "try:\n    src.utils.visualize_uv_maps\nexcept AttributeError as e:\n    print('OK', e)\n")
    r = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "parsenet-codebase_b200")], capture_output=True,
                       text=True, timeout=300)
    assert r.stdout.startswith("OK") and "PARSENET_REFERENCE_SRC" in r.stdout, r.stdout + r.stderr[-2000:]


def test_product_package_never_imports_the_oracle_or_falls_back_to_cpu():
    """the oracle is test infrastructure: no module of the product package may import it, and `bench.py`'s product arm
    (everything outside cpu_baseline / run_reference) may not either"""
    import ast
    pkg = os.path.join(ROOT, "parsenet-codebase_b200")
    offenders = []
    for sub in ("src", "pnb200"):
        for f in sorted(os.listdir(os.path.join(pkg, sub))):
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(pkg, sub, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module]
                offenders += [(sub, f, n) for n in names if n == "oracle" or n.startswith("oracle.")]
    assert not offenders, offenders
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn) if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("oracle")]
        assert not uses or fn.name in ("cpu_baseline", "run_reference", "_port_weights", "_port_one_shape",
                                       "gpu_eager_baseline"), f"bench.py::{fn.name} imports the oracle"


@pytest.mark.parametrize("use_tma", [False, True])
def test_meanshift_host_wiring_matches_abi_arity(monkeypatch, use_tma):
    """dry run of the mean-shift autograd schedule on CPU tensors with a recording stand-in for the C-ABI call: every
    call site (default path and the experimental TMA path) passes exactly the arguments include/parsenet_b200.h declares,
    pointers where pointers are expected, and the forward / backward sequences have the expected shape"""
    import ctypes
    from pnb200 import cabi, meanshift as pms
    calls = []

    def fake_call(name, *args):
        sig = cabi.SIGNATURES[name]
        assert len(args) == len(sig), f"{name}: {len(args)} arguments, header declares {len(sig)}"
        for i, (a, t) in enumerate(zip(args, sig)):
            if t is ctypes.c_void_p:
                assert a is None or isinstance(a, int), f"{name} arg {i}: expected a pointer"
            elif t in (ctypes.c_int, ctypes.c_longlong):
                assert isinstance(a, int) and not isinstance(a, bool), f"{name} arg {i}: expected an int, got {type(a)}"
        calls.append(name)

    monkeypatch.setattr(pms, "call", fake_call)
    monkeypatch.setattr(pms, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(pms, "_stream", lambda: 0)
    monkeypatch.setattr(pms, "_ptr", lambda t: None if t is None else t.data_ptr())     # (CPU tensors in a dry run)
    monkeypatch.setattr(pms, "USE_TMA", use_tma)
    B, N, d, its = 2, 70, 128, 3
    X = torch.nn.functional.normalize(torch.randn(B, N, d), dim=2).requires_grad_()
    Y = pms.mean_shift_iters(X, torch.tensor([0.3, 0.5]), its)
    assert Y.shape == (B, N, d)
    Y.sum().backward()
    assert X.grad is not None and X.grad.shape == X.shape
    if use_tma:
        assert calls == ["pn_ms_prepare_operands"] + ["pn_ms_iter_fwd_tma"] * its + ["pn_ms_prepare_operands"] + \
            ["pn_ms_iter_bwd_tma"] * its
    else:
        assert calls == ["pn_ms_iter_fwd_tc"] * its + ["pn_ms_iter_bwd_tc"] * its


def test_sparse_row_backward_host_wiring_matches_abi_arity(monkeypatch):
    """dry run (recording stand-in for the C-ABI) of the experimental sparse-row backward schedule: forward iterations
    without an autograd node, centres gathered from the last iterate, backward = one pn_ms_rows_bwd per iteration"""
    import ctypes
    from pnb200 import cabi, meanshift as pms
    calls = []

    def fake_call(name, *args):
        sig = cabi.SIGNATURES[name]
        assert len(args) == len(sig), f"{name}: {len(args)} arguments, header declares {len(sig)}"
        for i, (a, t) in enumerate(zip(args, sig)):
            if t is ctypes.c_void_p:
                assert a is None or isinstance(a, int), f"{name} arg {i}: expected a pointer"
            elif t in (ctypes.c_int, ctypes.c_longlong):
                assert isinstance(a, int) and not isinstance(a, bool), f"{name} arg {i}: expected an int, got {type(a)}"
        calls.append(name)

    monkeypatch.setattr(pms, "call", fake_call)
    monkeypatch.setattr(pms, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(pms, "_stream", lambda: 0)
    monkeypatch.setattr(pms, "_ptr", lambda t: None if t is None else t.data_ptr())     # (CPU tensors in a dry run)
    B, N, d, its = 2, 300, 128, 3
    X = torch.nn.functional.normalize(torch.randn(B, N, d), dim=2).requires_grad_()
    Y, state = pms.mean_shift_iters_keep(X, torch.tensor([0.3, 0.5]), its)
    assert Y.shape == (B, N, d) and not Y.requires_grad
    ids = [torch.tensor([5, 17, 200]), torch.tensor([0, 299])]
    centers = pms.centers_sparse(X, state, ids)
    assert [tuple(c.shape) for c in centers] == [(3, d), (2, d)]
    assert torch.equal(centers[0], Y[0][ids[0]]) and torch.equal(centers[1], Y[1][ids[1]])
    (centers[0].sum() + centers[1].sum()).backward()
    assert X.grad is not None and X.grad.shape == X.shape
    assert calls == ["pn_ms_prepare_operands"] + ["pn_ms_iter_fwd_tma"] * its + ["pn_ms_rows_bwd"] * its


def test_sparse_row_backward_host_schedule_is_numerically_right_with_reference_kernels(monkeypatch):
    """the Python half of the experimental sparse-row backward (forward without an autograd node, gather of the centre
    rows, per-iteration backward over the compact row set, padding slots, final scatter into the rows of X) driven on the
    CPU by stand-in kernels that implement the C-ABI contracts with the oracle's formulas, writing through the raw
    pointers they are given: the gradient must equal dense autograd through the port's iteration"""
    import ctypes
    from oracle.port import meanshift as oms
    from pnb200 import meanshift as pms

    def arr(ptr, *shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr)).reshape(shape)

    def fake_call(name, *a):
        if name == "pn_ms_iter_fwd_tc":
            Yp, Xp, B, N, d, cp, Ynp, dnp, unp, _ = a
            Y, X, c = arr(Yp, B, N, d), arr(Xp, B, N, d), arr(cp, B)
            Yn, dn, un = arr(Ynp, B, N, d), arr(dnp, B, N), arr(unp, B, N)
            for b in range(B):
                y, x = torch.from_numpy(Y[b].copy()), torch.from_numpy(X[b].copy())
                K = torch.exp(torch.clamp((y @ x.t() - 1.0) * float(c[b]), -75.0, 75.0))
                den = K.sum(1)
                u = y + ((K @ x) / den[:, None] - y)
                nr = u.norm(dim=1)
                Yn[b], dn[b], un[b] = (u / nr[:, None]).numpy(), den.numpy(), nr.numpy()
        elif name == "pn_ms_rows_bwd":
            gp, ynp, ypp, dnp, unp, Xp, B, R, N, d, cp, _, _, _, gyp, gxp, _ = a
            g, yn, yp = arr(gp, B, R, d), arr(ynp, B, R, d), arr(ypp, B, R, d)
            dn, un, X, c = arr(dnp, B, R), arr(unp, B, R), arr(Xp, B, N, d), arr(cp, B)
            gy, gx = arr(gyp, B, R, d), arr(gxp, B, N, d)
            for b in range(B):
                t = [torch.from_numpy(v[b].copy()) for v in (g, yn, yp, dn, un, X)]
                gyb, gxb = oms.sparse_rows_backward(*t, float(c[b]) ** -0.5)
                gy[b] = gyb.numpy()
                gx[b] += gxb.numpy()
        else:
            raise AssertionError(name)

    monkeypatch.setattr(pms, "call", fake_call)
    monkeypatch.setattr(pms, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(pms, "_stream", lambda: 0)
    monkeypatch.setattr(pms, "_ptr", lambda t: None if t is None else t.data_ptr())     # (CPU tensors in a dry run)
    monkeypatch.setattr(pms, "USE_TMA", False)         # (the stand-in implements the loader-warp entry point's contract)
    gen = torch.Generator().manual_seed(0)
    B, N, d, its = 2, 90, 128, 3
    X0 = torch.nn.functional.normalize(torch.randn(B, N, d, generator=gen), dim=2)
    bws = torch.tensor([0.4, 0.7])
    ids = [torch.tensor([1, 40, 41, 89]), torch.tensor([0, 7])]
    w = [torch.randn(len(i), d, generator=gen) for i in ids]
    X = X0.clone().requires_grad_()
    Yfin, state = pms.mean_shift_iters_keep(X, bws, its)
    centers = pms.centers_sparse(X, state, ids)
    sum((c * wi).sum() for c, wi in zip(centers, w)).backward()
    for b in range(B):                                  # dense autograd through the port's iteration
        xr = X0[b].clone().requires_grad_()
        yr = oms.mean_shift_iters(xr, bws[b], its)
        assert ((yr.detach() - Yfin[b]).abs().max() / yr.abs().max()).item() < 1e-5
        (yr[ids[b]] * w[b]).sum().backward()
        err = ((X.grad[b] - xr.grad).abs().max() / xr.grad.abs().max()).item()
        assert err < 1e-4, (b, err)


def test_launch_device_bookkeeping_rejects_mixed_devices_and_cpu_pointers():
    """every pointer of one launch must live on ONE cuda device; the launch then runs on that device's stream (ADVICE r1:
    the reference's scripts keep the fit stage on cuda:alt_gpu while the current device is cuda:0)"""
    from pnb200 import cabi
    cabi.close_device_set()
    cabi.note_device(torch.device("cuda", 1)); cabi.note_device(torch.device("cuda", 1))
    assert cabi.close_device_set() == 1
    assert cabi.close_device_set() is None                      # nothing noted since the last launch
    cabi.note_device(torch.device("cuda", 0)); cabi.note_device(torch.device("cuda", 1))
    with pytest.raises(cabi.PnError, match="different devices"):
        cabi.close_device_set()
    assert cabi.close_device_set() is None                      # the failed launch does not leak into the next one
    with pytest.raises(cabi.PnError, match="no CPU fallback"):
        cabi.note_device(torch.device("cpu"))
