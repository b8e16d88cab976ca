"""TEST INFRASTRUCTURE ONLY — input generators shared by make_golden.py and the tests (no reference import)."""
import numpy as np
import torch

from oracle.port import common


def clustered_embedding(N, d, n_clusters, seed, spread=0.2):
    """unit-norm embedding rows around `n_clusters` random unit centres"""
    g = torch.Generator().manual_seed(seed)
    cent = torch.nn.functional.normalize(torch.randn(n_clusters, d, generator=g), dim=1)
    lab = torch.randint(0, n_clusters, (N,), generator=g)
    x = cent[lab] + spread * torch.randn(N, d, generator=g) / d ** 0.5
    return torch.nn.functional.normalize(x, dim=1), lab


def e2e_inputs(N, seed, no_cylinder=False):
    """one synthetic shape (6 patches; two of them re-typed as open / closed spline) + an embedding clustered by the
    gt segment with some noise"""
    pts, nrm, lab, prim = common.synth_cloud(1, N, seed=seed, n_patches=6)
    prim = prim.copy()
    prim[lab == 4] = 2      # open b-spline
    prim[lab == 5] = 9      # closed b-spline
    if no_cylinder:
        prim[prim == 4] = 5     # fit a sphere to the cylinder patch instead (see tests: reference cylinder noise)
    g = torch.Generator().manual_seed(seed)
    cent = torch.nn.functional.normalize(torch.randn(6, 128, generator=g), dim=1)
    emb = cent[torch.from_numpy(lab[0])] + 0.25 * torch.randn(N, 128, generator=g) / 128 ** 0.5
    logp = torch.log_softmax(torch.randn(1, 10, N, generator=g), 1)
    return pts, nrm, lab, prim, emb.unsqueeze(0), logp


def prim_cloud(kind, m, seed):
    """noisy samples of one analytic primitive + outward normals + random positive weights"""
    rng = np.random.RandomState(seed)
    u, v = rng.rand(m), rng.rand(m)
    if kind == "plane":
        p = np.stack([u - 0.5, v - 0.5, 0 * u], 1); n = np.tile([0, 0, 1.0], (m, 1))
    elif kind == "sphere":
        th, ph = 2 * np.pi * u, np.arccos(1 - 1.4 * v)
        n = np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], 1); p = 0.7 * n
    elif kind == "cylinder":
        th = 2 * np.pi * u * 0.8
        n = np.stack([np.cos(th), np.sin(th), 0 * u], 1); p = 0.4 * n + np.stack([0 * u, 0 * u, 1.5 * (v - 0.5)], 1)
    else:
        a = np.pi / 5; th = 2 * np.pi * u * 0.9; h = 0.3 + 0.7 * v
        p = np.stack([h * np.tan(a) * np.cos(th), h * np.tan(a) * np.sin(th), h], 1)
        n = np.stack([np.cos(a) * np.cos(th), np.cos(a) * np.sin(th), -np.sin(a) * np.ones(m)], 1)
    R, _ = np.linalg.qr(rng.randn(3, 3))
    p = p @ R.T + rng.randn(3) * 0.3 + rng.randn(m, 3) * 0.004
    n = n @ R.T
    w = 0.15 + 0.85 * rng.rand(m, 1)
    return p.astype(np.float32), n.astype(np.float32), w.astype(np.float32)
