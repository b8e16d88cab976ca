"""Seeded synthetic ParSeNet-style point clouds (SURVEY.md 8d input 2): every shape is a union of analytic patches
(plane / sphere / cylinder / cone), open bicubic B-spline patches over smooth random 20x20 control grids and closed
(periodic in u) swept surfaces, with unit normals, per-point segment labels and primitive-type ids, jittered
(sigma = 0.002), permuted, centred and scaled to unit max extent like the reference's dataset_segments.py:128-144.

Plain numpy, no dependency on the product package or on the oracle: bench.py, __graft_entry__.smoke(), the tests and the
golden generator all draw their inputs from here.  Primitive ids follow the reference's convention
(primitive_forward.py:925-1047): 1 plane, 3 cone, 4 cylinder, 5 sphere, {2, 8} open spline, {0, 6, 7, 9} closed spline.
"""
import numpy as np

ANALYTIC_KINDS = (1, 5, 4, 3)
ALL_KINDS = (1, 5, 4, 3, 2, 9)


def _clamped_uniform_knots(n_ctrl, degree):
    inner = np.linspace(0.0, 1.0, n_ctrl - degree + 1)
    return np.concatenate([np.zeros(degree), inner, np.ones(degree)])


def bspline_basis(n_ctrl, degree, t):
    """(len(t), n_ctrl) clamped uniform B-spline basis values by the Cox-de Boor recursion, vectorised over t in [0,1)"""
    t = np.asarray(t, np.float64)
    kn = _clamped_uniform_knots(n_ctrl, degree)
    m = len(kn) - 1
    N = np.zeros((t.shape[0], m), np.float64)
    for i in range(m):
        N[:, i] = (t >= kn[i]) & (t < kn[i + 1])
    for p in range(1, degree + 1):
        Nn = np.zeros((t.shape[0], m - p), np.float64)
        for i in range(m - p):
            d1, d2 = kn[i + p] - kn[i], kn[i + p + 1] - kn[i + 1]
            if d1 > 0:
                Nn[:, i] += (t - kn[i]) / d1 * N[:, i]
            if d2 > 0:
                Nn[:, i] += (kn[i + p + 1] - t) / d2 * N[:, i + 1]
        N = Nn
    return N


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)


def _normals_fd(f, u, v, h=1e-4):
    """unit normals of the parametric surface f(u, v) -> (m,3) by central differences"""
    du = f(np.clip(u + h, 0, 1 - 1e-9), v) - f(np.clip(u - h, 0, 1 - 1e-9), v)
    dv = f(u, np.clip(v + h, 0, 1 - 1e-9)) - f(u, np.clip(v - h, 0, 1 - 1e-9))
    return _unit(np.cross(du, dv))


def open_spline_patch(rng, u, v):
    """smooth random bicubic patch: 20x20 control grid = regular xy lattice + a low-frequency height field"""
    g = np.linspace(-0.5, 0.5, 20)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    a = rng.uniform(0.05, 0.15, 3); f = rng.uniform(0.5, 2.0, (3, 2)); ph = rng.uniform(0, 2 * np.pi, 3)
    gz = sum(a[k] * np.sin(2 * np.pi * (f[k, 0] * gx + f[k, 1] * gy) + ph[k]) for k in range(3))
    cp = np.stack([gx, gy, gz], 2)                                                  # (20,20,3)

    def surf(uu, vv):
        Bu, Bv = bspline_basis(20, 3, uu), bspline_basis(20, 3, vv)
        return np.einsum("mi,ijc,mj->mc", Bu, cp, Bv)
    return surf(u, v), _normals_fd(surf, u, v)


def closed_spline_patch(rng, u, v):
    """surface periodic in u: a cylinder whose radius is modulated around and along the axis"""
    r0 = rng.uniform(0.2, 0.35); a = rng.uniform(0.05, 0.2, 2); ph = rng.uniform(0, 2 * np.pi, 2)

    def surf(uu, vv):
        th, z = 2 * np.pi * uu, vv - 0.5
        r = r0 * (1 + a[0] * np.sin(2 * th + ph[0]) + a[1] * np.sin(3 * th + ph[1]) * np.cos(np.pi * z))
        return np.stack([r * np.cos(th), r * np.sin(th), z], 1)
    du = surf((u + 1e-4) % 1.0, v) - surf((u - 1e-4) % 1.0, v)
    dv = surf(u, np.clip(v + 1e-4, 0, 1)) - surf(u, np.clip(v - 1e-4, 0, 1))
    return surf(u, v), _unit(np.cross(du, dv))


def synth_cloud(B, N, seed, n_patches=5, kinds=ANALYTIC_KINDS):
    """-> points (B,N,3) f32, unit normals (B,N,3) f32, labels (B,N) i64 (patch id), primitives (B,N) i64 (type id).
    Patch s of a shape has type kinds[s % len(kinds)].  The default kinds (analytic only) keeps the random stream of
    the generator the committed golden vectors were made with."""
    rng = np.random.RandomState(seed)
    pts = np.zeros((B, N, 3), np.float32); nrm = np.zeros((B, N, 3), np.float32)
    lab = np.zeros((B, N), np.int64); prim = np.zeros((B, N), np.int64)
    for b in range(B):
        sizes = rng.multinomial(N - 40 * n_patches, np.ones(n_patches) / n_patches) + 40
        o = 0
        for s_i, m in enumerate(sizes):
            kind = kinds[s_i % len(kinds)]
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
            elif kind == 3:    # cone, half angle a
                a = 0.3 + 0.4 * rng.rand(); th = 2 * np.pi * u * 0.8; h = 0.2 + 0.6 * v
                p = np.stack([h * np.tan(a) * np.cos(th), h * np.tan(a) * np.sin(th), h], 1)
                n = np.stack([np.cos(a) * np.cos(th), np.cos(a) * np.sin(th), -np.sin(a) * np.ones(m)], 1)
            elif kind in (2, 8):
                p, n = open_spline_patch(rng, u * (1 - 1e-9), v * (1 - 1e-9))
            elif kind in (0, 6, 7, 9):
                p, n = closed_spline_patch(rng, u, v)
            else:
                raise ValueError(f"unknown primitive id {kind}")
            p = p @ R.T + c + rng.randn(m, 3) * 0.002
            n = n @ R.T
            pts[b, o:o + m] = p; nrm[b, o:o + m] = n; lab[b, o:o + m] = s_i; prim[b, o:o + m] = kind
            o += m
        perm = rng.permutation(N)
        pts[b] = pts[b, perm]; nrm[b] = nrm[b, perm]; lab[b] = lab[b, perm]; prim[b] = prim[b, perm]
        pts[b] -= pts[b].mean(0, keepdims=True)
        pts[b] /= np.max(pts[b].max(0) - pts[b].min(0))
    return pts, nrm, lab, prim


def open_spline_batch(B, M, seed):
    """BASELINE config 3 input: B smooth random bicubic 20x20 control grids, M surface points each at random
    (u, v), centred and scaled to unit max extent together with their control grid.
    -> points (B,3,M) f32 (channel-major like train_open_splines.py feeds them), control grids (B,20,20,3) f32"""
    rng = np.random.RandomState(seed)
    pts = np.zeros((B, 3, M), np.float32)
    cps = np.zeros((B, 20, 20, 3), np.float32)
    g = np.linspace(-0.5, 0.5, 20)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    for b in range(B):
        a = rng.uniform(0.05, 0.15, 3); f = rng.uniform(0.5, 2.0, (3, 2)); ph = rng.uniform(0, 2 * np.pi, 3)
        gz = sum(a[k] * np.sin(2 * np.pi * (f[k, 0] * gx + f[k, 1] * gy) + ph[k]) for k in range(3))
        cp = np.stack([gx, gy, gz], 2)
        R, _ = np.linalg.qr(rng.randn(3, 3))
        cp = cp @ R.T
        u, v = rng.rand(M) * (1 - 1e-9), rng.rand(M) * (1 - 1e-9)
        p = np.einsum("mi,ijc,mj->mc", bspline_basis(20, 3, u), cp, bspline_basis(20, 3, v))
        c = p.mean(0, keepdims=True)
        s = np.max(p.max(0) - p.min(0))
        pts[b] = ((p - c) / s).T
        cps[b] = (cp - c.reshape(1, 1, 3)) / s
    return pts, cps
