"""parsenet_b200 host layer: ctypes binding of the C-ABI library (`cabi`) and torch autograd ops (`ops`).

PyTorch is plumbing here (device memory, streams, autograd graph, torch.distributed); every hot operation is a
hand-written sm_100a kernel behind include/parsenet_b200.h.  There is no CPU / eager fallback: importing `ops`
without the built library, or calling an op on a non-CUDA tensor, raises.
"""
from . import cabi  # noqa: F401
