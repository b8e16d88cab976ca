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
