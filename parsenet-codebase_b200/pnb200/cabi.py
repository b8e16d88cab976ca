"""ctypes binding of lib/libparsenet_b200.so (declared in include/parsenet_b200.h).

Fails loudly when the library is missing — there is deliberately no fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libparsenet_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_ll = ctypes.c_longlong

# name -> argtypes (all return int status unless listed in _SPECIAL)
c_d = ctypes.c_double
SIGNATURES = {
    "pn_knn": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p],
    "pn_knn_tma": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p],
    # linear.cu
    "pn_linear_fwd": [c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_i, c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "pn_linear_fwd_tc": [c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_i, c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "pn_linear_bwd_data": [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_p, c_ll, c_p, c_p, c_i, c_p, c_p, c_p,
                           c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "pn_linear_bwd_data_tc": [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_p, c_ll, c_p, c_p, c_i, c_p, c_p, c_p,
                              c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "pn_linear_bwd_weight": [c_p, c_ll, c_p, c_ll, c_p, c_p, c_i, c_p, c_ll, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "pn_linear_bwd_weight_tc": [c_p, c_ll, c_p, c_ll, c_p, c_p, c_i, c_p, c_ll, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "pn_norm_finalize": [c_p, c_p, c_p, c_i, c_i, c_i, c_d, c_f, c_p, c_p, c_p, c_p],
    "pn_norm_bwd_apply": [c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_p, c_p, c_p],
    # edgeconv.cu
    "pn_edge_gather_fwd": [c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "pn_edge_apply": [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p],
    "pn_edge_bwd_prep": [c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_knn_csr_transpose": [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_edge_bwd": [c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_i,
                    c_p, c_ll, c_p],
    # pointwise.cu
    "pn_colmax_norm": [c_p, c_ll, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p],  # (.., act, wts, out, arg, stream)
    "pn_colmax_bwd_fill": [c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_i, c_p, c_ll, c_p],
    "pn_logsoftmax_fwd": [c_p, c_ll, c_i, c_i, c_i, c_p, c_p],
    "pn_logsoftmax_bwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_ll, c_p],
    "pn_nll_fwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "pn_nll_bwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "pn_l2norm_fwd": [c_p, c_ll, c_ll, c_i, c_f, c_p, c_ll, c_p, c_p],
    "pn_l2norm_bwd": [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_p, c_ll, c_i, c_p],
    "pn_triplet_fwd": [c_p, c_ll, c_i, c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p],
    "pn_triplet_bwd": [c_p, c_ll, c_i, c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p, c_ll, c_p],
    # meanshift.cu
    "pn_ms_iter_fwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_ms_iter_fwd_tc": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_ms_iter_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    "pn_ms_iter_bwd_tc": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    "pn_debug_set_progress": [c_p],
    "pn_ms_rows_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "pn_ms_bwd_prep_tc": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "pn_ms_bwd_cols_tc": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_p],
    # meanshift_tma.cu (default path; PN_MS_TMA=0: loader-warp kernels)
    "pn_ms_prepare_operands": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "pn_ms_iter_fwd_tma": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_ms_iter_bwd_tma": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    "pn_ms_kth_dist": [c_p, c_p, c_i, c_i, c_ll, c_i, c_i, c_p, c_p],
    "pn_ms_kth_dist_tc": [c_p, c_p, c_i, c_i, c_ll, c_i, c_i, c_p, c_p],
    "pn_knn_lowdim": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p],
    "pn_knn_flagged": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p, c_p],
    "pn_knn_tc": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p],
    "pn_knn_tma_flagged": [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p, c_p],
    "pn_iou_cost": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "pn_hungarian": [c_p, c_i, c_i, c_p, c_p],
    "pn_kron_fit": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "pn_kron_eval": [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_p, c_p],
    "pn_ms_kth_dist_tc_flagged": [c_p, c_p, c_i, c_i, c_ll, c_i, c_i, c_p, c_p, c_p],
    "pn_ms_kth_dist_tma": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p],
    "pn_ms_argsel": [c_i, c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "pn_ms_argsel_tc": [c_i, c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "pn_ms_argsel_tma": [c_i, c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    # chamfer.cu / spline.cu / fit.cu / primitives.cu
    "pn_chamfer_nn_fwd": [c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p],
    "pn_chamfer_nn_bwd": [c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_spline_eval_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "pn_spline_eval_bwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "pn_spline_eval_fwd_f64": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "pn_spline_eval_bwd_f64": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "pn_fit_moments_fwd": [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_f, c_p, c_p],
    "pn_fit_moments_bwd": [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_ll, c_p],
    "pn_residual_fwd": [c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p],
    "pn_fit_moments_fwd_batched": [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p],
    "pn_fit_moments_bwd_batched": [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_ll, c_p],
    "pn_fit_solve": [c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p],
    "pn_grid_perm_fwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "pn_grid_perm_bwd": [c_p, c_p, c_ll, c_p, c_f, c_p, c_p],
    "pn_grid_laplacian_fwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "pn_grid_laplacian_bwd": [c_p, c_i, c_i, c_i, c_p, c_f, c_p, c_p, c_p],
    "pn_weights_normalize_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "pn_weights_normalize_bwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "pn_residual_fwd_batched": [c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p],
    # small3.cu
    "pn_sym3_eigh": [c_p, c_i, c_p, c_p, c_p],
    "pn_lstsq3": [c_p, c_p, c_i, c_i, c_d, c_p, c_p, c_p, c_p],
}
_SPECIAL = {
    "pn_last_error": (ctypes.c_char_p, []),
    "pn_launch_count": (ctypes.c_ulonglong, []),
    "pn_reset_launch_count": (None, []),
    "pn_abi_version": (c_i, []),
    "pn_knn_tma_supported": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i]),
    "pn_knn_tc_supported": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i]),
    "pn_ms_argsel_tma_supported": (c_i, [c_p, c_ll, c_i, c_i]),
    "pn_knn_lowdim_supported": (c_i, [c_i, c_i, c_i, c_i]),
    "pn_linear_fwd_tc_supported": (c_i, [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_i]),
    "pn_linear_bwd_weight_tc_supported": (c_i, [c_p, c_ll, c_p, c_ll, c_i, c_i, c_i]),
    "pn_linear_bwd_data_tc_supported": (c_i, [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_i]),
}


class PnError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python parsenet-codebase_b200/build.py` "
            "(or __graft_entry__.build()). parsenet_b200 has no CPU/eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.argtypes = argtypes
        fn.restype = c_i
    for name, (res, argtypes) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = res
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        raise PnError(f"{what} failed (rc={rc}): {lib.pn_last_error().decode()}")


def launch_count():
    return int(lib.pn_launch_count())


def reset_launch_count():
    lib.pn_reset_launch_count()


# ---- device of a launch.  Every pointer argument goes through ops._ptr (which calls note_device) and the stream
# argument through ops._stream (which calls close_device_set): one launch = one device, and `call` makes that device
# current for the duration of the launch.  Kernels launched under another current device than the one owning their
# pointers / stream fail with an invalid-resource-handle error (or, with peer access, race with the torch ops queued on
# the owning device's stream).
import threading

_tls = threading.local()


def note_device(device):
    if device.type != "cuda":
        raise PnError("parsenet_b200 ops need CUDA tensors (there is no CPU fallback)")
    seen = getattr(_tls, "seen", None)
    if seen is None:
        _tls.seen = device.index
    elif seen != device.index:
        _tls.mixed = (seen, device.index)


def close_device_set():
    """-> device index shared by the pointers noted since the previous launch (None = no tensor argument)"""
    dev = getattr(_tls, "seen", None)
    mixed = getattr(_tls, "mixed", None)
    _tls.seen = None
    _tls.mixed = None
    if mixed is not None:
        raise PnError("pointer arguments of one launch live on different devices: cuda:%d and cuda:%d" % mixed)
    _tls.launch_dev = dev
    return dev


def _take_launch_device():
    dev = getattr(_tls, "launch_dev", None)
    _tls.launch_dev = None
    _tls.seen = None            # (a launch without a stream argument must not leak its pointers into the next one)
    _tls.mixed = None
    return dev


# ---- optional per-entry-point device timing (bench.py uses it for the roofline of the dominant kernel)
TIMED = {}          # name -> list of (start_event, end_event); register a name to start collecting
EVENT_POOL = []     # optional pre-created timing events (bench.py fills it before its timed loop: creating an event is a driver
                    # call that can block behind other driver activity; recording one is not)


def call(name, *args):
    fn = getattr(lib, name)
    dev = _take_launch_device()
    if dev is not None:
        import torch
        if torch.cuda.current_device() != dev:
            with torch.cuda.device(dev):
                return _call_on_current(name, fn, args, dev)
    return _call_on_current(name, fn, args, dev)


def _call_on_current(name, fn, args, dev):
    ev = TIMED.get(name)
    if ev is not None:
        import torch
        st = torch.cuda.current_stream(dev)
        a = EVENT_POOL.pop() if EVENT_POOL else torch.cuda.Event(enable_timing=True)
        b = EVENT_POOL.pop() if EVENT_POOL else torch.cuda.Event(enable_timing=True)
        a.record(st)
        rc = fn(*args)
        b.record(st)
        ev.append((a, b))
    else:
        rc = fn(*args)
    if rc != 0:
        raise PnError(f"{name} failed (rc={rc}): {lib.pn_last_error().decode()}")
