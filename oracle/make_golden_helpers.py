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
