"""TEST INFRASTRUCTURE ONLY — shared helpers for the oracle port and the golden generator."""
import numpy as np
import torch


def seeded_state_dict(shapes, seed, scale=None):
    """Deterministic weights keyed by parameter name (sorted): used by the golden generator (loaded into the
    unmodified reference model) and by the tests (loaded into the port and the CUDA model)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        t = torch.randn(shp, generator=g)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        if name.endswith("running_var"):
            out[name] = t.abs() * 0.5 + 0.5
            continue
        if name.endswith("running_mean"):
            out[name] = t * 0.1
            continue
        if len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            out[name] = t * (1.0 / np.sqrt(fan_in))
        elif name.endswith("weight"):
            out[name] = t * 0.5 + 0.2          # 1-D weights are norm gammas: mixed signs on purpose
        else:
            out[name] = t * 0.1
    return out


def synth_cloud(B, N, seed, n_patches=5):
    """Seeded multi-patch point cloud with unit normals, labels and primitive ids (SURVEY.md §8d input 2,
    reduced): planes / spheres / cylinders / cones, jitter sigma=0.002, centred, unit max extent."""
    rng = np.random.RandomState(seed)
    pts = np.zeros((B, N, 3), np.float32); nrm = np.zeros((B, N, 3), np.float32)
    lab = np.zeros((B, N), np.int64); prim = np.zeros((B, N), np.int64)
    for b in range(B):
        sizes = rng.multinomial(N - 40 * n_patches, np.ones(n_patches) / n_patches) + 40
        o = 0
        for s_i, m in enumerate(sizes):
            kind = [1, 5, 4, 3][s_i % 4]
            u = rng.rand(m); v = rng.rand(m)
            c = rng.randn(3) * 0.4
            R, _ = np.linalg.qr(rng.randn(3, 3))
            if kind == 1:      # plane
                p = np.stack([u - 0.5, v - 0.5, np.zeros(m)], 1); n = np.tile([0, 0, 1.0], (m, 1))
            elif kind == 5:    # sphere
                th = 2 * np.pi * u; ph = np.arccos(1 - 1.2 * v); r = 0.3 + 0.2 * rng.rand()
                n = np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], 1); p = r * n
            elif kind == 4:    # cylinder
                th = 2 * np.pi * u * 0.7; r = 0.2 + 0.2 * rng.rand()
                n = np.stack([np.cos(th), np.sin(th), np.zeros(m)], 1); p = r * n + np.stack([0 * u, 0 * u, v - 0.5], 1)
            else:              # cone, half angle a
                a = 0.3 + 0.4 * rng.rand(); th = 2 * np.pi * u * 0.8; h = 0.2 + 0.6 * v
                p = np.stack([h * np.tan(a) * np.cos(th), h * np.tan(a) * np.sin(th), h], 1)
                n = np.stack([np.cos(a) * np.cos(th), np.cos(a) * np.sin(th), -np.sin(a) * np.ones(m)], 1)
            p = p @ R.T + c + rng.randn(m, 3) * 0.002
            n = n @ R.T
            pts[b, o:o + m] = p; nrm[b, o:o + m] = n; lab[b, o:o + m] = s_i; prim[b, o:o + m] = kind
            o += m
        perm = rng.permutation(N)
        pts[b] = pts[b, perm]; nrm[b] = nrm[b, perm]; lab[b] = lab[b, perm]; prim[b] = prim[b, perm]
        pts[b] -= pts[b].mean(0, keepdims=True)
        pts[b] /= np.max(pts[b].max(0) - pts[b].min(0))
    return pts, nrm, lab, prim
