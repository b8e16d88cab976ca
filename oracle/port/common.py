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


# the synthetic input generator is neither oracle nor product: it lives in tools/synth.py (numpy only) and is re-exported
# here for the golden generator / tests that historically imported it from this module
from tools.synth import synth_cloud  # noqa: E402,F401
