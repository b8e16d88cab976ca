"""Per-batch input preprocessing of the training loop on the device.

Reference: Dataset.get_train src/dataset_segments.py:95-157 as the training scripts configure it (train_parsenet_e2e.py:108-114:
randomize, augment=False, align_canonical=True, anisotropic=False, optional normal noise): per shape the minor principal axis of
the cloud is rotated onto x (PCA of X^T X through numpy's LAPACK eig + rotation_matrix_a_to_b :276-298), points AND normals are
rotated, and the points are scaled by their largest (or per-axis) extent.  The reference does all of it in numpy on the loader
thread: 16 x (10 000 x 3) rotations + extent reductions per step, which bounds the loop once the model runs at ~100 shapes/s.

Split here: the 3x3 work stays on the host with the reference's own arithmetic (float32 X^T X, LAPACK eig: the eigenvector SIGN
decides the rotation, so it has to be the reference's), the O(N) work -- noise, rotations, extents, scaling -- runs on the device
on the batch that is being uploaded anyway.  Nothing synchronises: the host never reads anything back.
"""
import numpy as np
import torch

EPS = float(np.finfo(np.float32).eps)


def rotation_matrix_a_to_b(A, B):
    """dataset_segments.py:276-298 (numpy): rotation with R A = B"""
    cos, sin = np.dot(A, B), np.linalg.norm(np.cross(B, A))
    v = B - np.dot(A, B) * A
    v = v / (np.linalg.norm(v) + EPS)
    w = np.cross(B, A)
    w = w / (np.linalg.norm(w) + EPS)
    Fm = np.stack([A, v, w], 1)
    G = np.array([[cos, -sin, 0], [sin, cos, 0], [0, 0, 1]])
    try:
        return Fm @ G @ np.linalg.inv(Fm)
    except np.linalg.LinAlgError:
        return np.eye(3, dtype=np.float32)


def host_rotations(points_np):
    """(B,N,3) float32 host clouds -> (B,3,3) float32 rotations taking each cloud's minor principal axis onto x
    (pca_numpy :300-302 + rotation_matrix_a_to_b; 3x3 problems, ~50 us per shape)"""
    B = points_np.shape[0]
    R = np.empty((B, 3, 3), np.float32)
    for j in range(B):
        X = points_np[j]
        S, U = np.linalg.eig(X.T @ X)
        R[j] = rotation_matrix_a_to_b(U[:, np.argmin(S)], np.array([1, 0, 0]))
    return R


def normal_noise(n_points, rng=np.random):
    """the reference's draw for if_normal_noise (:119-123): one clipped gaussian per point, shared by the batch -> (1,N,1)"""
    return np.clip(rng.randn(1, n_points, 1) * 0.01, a_min=-0.01, a_max=0.01).astype(np.float32)


def preprocess_on_device(points, normals, R, anisotropic=False):
    """points, normals (B,N,3) device tensors (noise already applied to the points), R (B,3,3) device -> aligned + scaled
    points, rotated normals.  Three batched launches for the rotations / extents, no host round trip."""
    Rt = R.transpose(1, 2)
    points = torch.bmm(points, Rt)
    normals = torch.bmm(normals, Rt) if normals is not None else None
    std = points.amax(1) - points.amin(1)                                   # (B,3)
    if anisotropic:
        points = points / (std.unsqueeze(1) + EPS)
    else:
        points = points / (std.amax(1).view(-1, 1, 1) + EPS)
    return points, normals


class DeviceBatchPipeline:
    """wraps a host iterator of raw batches [points (B,N,3), labels, normals, primitives] (what Dataset.get_train yields with
    align_canonical=False) and yields the aligned batch as device tensors, uploads from pinned staging buffers"""

    def __init__(self, host_iter, device, if_normal_noise=False, anisotropic=False, rng=np.random):
        self.it, self.dev, self.noise, self.aniso, self.rng = host_iter, device, if_normal_noise, anisotropic, rng
        self._pin = {}

    def _upload(self, name, arr):
        arr = np.ascontiguousarray(arr)
        buf = self._pin.get(name)
        if buf is None or buf.shape != arr.shape or buf.dtype != torch.from_numpy(arr).dtype:
            buf = self._pin[name] = torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype).pin_memory()
        buf.numpy()[...] = arr
        return buf.to(self.dev, non_blocking=True)

    def __iter__(self):
        return self

    def __next__(self):
        points, labels, normals, primitives = next(self.it)
        points = np.asarray(points, np.float32)
        if self.noise and normals is not None:
            points = points + normals * normal_noise(points.shape[1], self.rng)
        R = host_rotations(points)
        p_d, n_d = self._upload("p", points), (self._upload("n", np.asarray(normals, np.float32)) if normals is not None else None)
        p_d, n_d = preprocess_on_device(p_d, n_d, self._upload("R", R), self.aniso)
        return p_d, labels, n_d, primitives
