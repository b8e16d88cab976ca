"""TEST INFRASTRUCTURE ONLY — ctypes wrapper over oracle/c/knn_oracle.c (see that file's header)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _get():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.pn_oracle_knn.restype = ctypes.c_int
        _lib.pn_oracle_knn_row.restype = ctypes.c_int
    return _lib


def knn(x_bnc, k, metric=0, return_dist=False):
    """x_bnc: (B,N,C) float32 point-major.  Returns idx (B,N,k) int64 sorted best-first
    (reference: src/PointNet.py:9-26 metric 0, :29-69 metric 1)."""
    x = np.ascontiguousarray(x_bnc, dtype=np.float32)
    B, N, C = x.shape
    idx = np.empty((B, N, k), dtype=np.int32)
    dist = np.empty((B, N, k), dtype=np.float32) if return_dist else None
    rc = _get().pn_oracle_knn(x.ctypes.data_as(ctypes.c_void_p), B, N, C, C, k, metric,
                              idx.ctypes.data_as(ctypes.c_void_p),
                              dist.ctypes.data_as(ctypes.c_void_p) if return_dist else None)
    if rc != 0:
        raise RuntimeError(f"pn_oracle_knn failed rc={rc}")
    return (idx.astype(np.int64), dist) if return_dist else idx.astype(np.int64)


def knn_row(x_nc, i, metric=0):
    """all N fp32 distance values of query i (oracle arithmetic)"""
    x = np.ascontiguousarray(x_nc, dtype=np.float32)
    N, C = x.shape
    row = np.empty(N, dtype=np.float32)
    _get().pn_oracle_knn_row(x.ctypes.data_as(ctypes.c_void_p), N, C, C, metric, i,
                             row.ctypes.data_as(ctypes.c_void_p))
    return row
