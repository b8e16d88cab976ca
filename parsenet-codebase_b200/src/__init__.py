"""Drop-in module surface for the ParSeNet hot path (same module / class / function names as the reference's
`src` package), backed by the sm_100a kernels in ../csrc through ../pnb200.

Put `parsenet-codebase_b200/` on sys.path and the reference's training scripts' imports
(`from src.PointNet import PrimitivesEmbeddingDGCNGn`, `from src.mean_shift import MeanShift`, ...) resolve here.
"""

from ._fallthrough import extend_package_path as _extend

_extend(__path__)        # opt-in (PARSENET_REFERENCE_SRC): modules outside the hot path resolve to the reference's files
