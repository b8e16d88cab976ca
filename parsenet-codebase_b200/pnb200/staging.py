"""Pinned host staging for the small index / weight arrays the host logic produces every step.

A host->device copy from PAGEABLE memory synchronises the stream before it starts (CUDA API sync behaviour), i.e.
`torch.from_numpy(a).to(device)` stalls the host until every queued kernel has finished.  Uploads on the hot path go
through a pinned arena instead: the numpy array is memcpy'd into pinned memory and copied with non_blocking=True.
The arena is recycled with `reset()`, which waits (normally a no-op) for the last upload issued from it.
"""
import numpy as np
import torch

_TORCH_OF = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
             np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool}


class PinnedArena:
    def __init__(self, nbytes=16 << 20):
        self.cap = int(nbytes)
        self.buf = None
        self.off = 0
        self.last = None

    def reset(self):
        if self.last is not None:
            self.last.synchronize()
            self.last = None
        self.off = 0

    def upload(self, arr, device):
        """numpy array -> device tensor of the same dtype/shape, without a stream synchronisation"""
        arr = np.ascontiguousarray(arr)
        tdt = _TORCH_OF.get(arr.dtype)
        n = arr.nbytes
        if tdt is None or n == 0 or self.off + n > self.cap:
            return torch.from_numpy(arr).to(device)            # correct, but blocks
        if self.buf is None:
            self.buf = torch.empty(self.cap, dtype=torch.uint8).pin_memory()
        view = self.buf[self.off:self.off + n].view(tdt).view(arr.shape)
        view.numpy()[...] = arr
        self.off += (n + 63) & ~63
        out = view.to(device, non_blocking=True)
        # (the copy runs on the current stream of the DESTINATION device; record the recycle event there, not on the
        # stream of torch.cuda.current_device())
        self.last = torch.cuda.Event()
        self.last.record(torch.cuda.current_stream(out.device))
        return out


_ARENAS = {}


def arena(name, device=None):
    key = (name, str(device))
    a = _ARENAS.get(key)
    if a is None:
        a = _ARENAS[key] = PinnedArena()
    return a
