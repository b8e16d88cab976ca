"""Clamped exp / sqrt used throughout the fitting code (semantics of reference src/guard.py:7-14).

These two are elementwise glue on tiny tensors; the hot kernels re-implement the same clamps internally
(mean-shift kernel: exponent clamp +-75; residual kernels: sqrt floor)."""
import torch

EXP_CLAMP = 75.0
SQRT_FLOOR = 1e-5


def guard_exp(x, max_value=EXP_CLAMP, min_value=-EXP_CLAMP):
    return torch.exp(x.clamp(min=min_value, max=max_value))


def guard_sqrt(x, minimum=SQRT_FLOOR):
    return torch.sqrt(x.clamp(min=minimum))


from src._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__)     # non-hot-path names: reference module of the same name (opt-in, see _fallthrough.py)
