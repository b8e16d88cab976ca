"""TEST INFRASTRUCTURE ONLY — input generators shared by make_golden.py and the tests (no reference import)."""
import torch


def clustered_embedding(N, d, n_clusters, seed, spread=0.2):
    """unit-norm embedding rows around `n_clusters` random unit centres"""
    g = torch.Generator().manual_seed(seed)
    cent = torch.nn.functional.normalize(torch.randn(n_clusters, d, generator=g), dim=1)
    lab = torch.randint(0, n_clusters, (N,), generator=g)
    x = cent[lab] + spread * torch.randn(N, d, generator=g) / d ** 0.5
    return torch.nn.functional.normalize(x, dim=1), lab
