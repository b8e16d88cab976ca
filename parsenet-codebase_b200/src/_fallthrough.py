"""Fall-through to the reference's own `src` package for everything that is NOT on the hot path.

The drop-in package only re-implements the hot-path modules and names.  The reference's training /
inference scripts also import data loaders, augmentation, visualisation and evaluation helpers from `src.*`
(`from src.dataset import generator_iter`, `from src.utils import visualize_uv_maps`, ...).  When the environment
variable PARSENET_REFERENCE_SRC points at the reference's `src/` directory

  * modules the drop-in does not provide (`src.dataset`, `src.dataset_segments`, `src.augment_utils`, ...) are found
    there (`src.__path__` is extended, drop-in directory first), and
  * names the drop-in modules do not define (`src.utils.visualize_uv_maps`, `src.segment_utils.cluster`, ...) are taken
    from the reference module of the same name, loaded privately from its file (module-level `__getattr__`, PEP 562).

Hot-path names are never looked up in the reference: a name the drop-in defines always wins, so there is no way to
fall back to the reference's torch implementation of a kernel by accident.  Without the variable, missing names raise
AttributeError / ImportError as usual.
"""
import importlib.util
import os
import sys

_ENV = "PARSENET_REFERENCE_SRC"
_PRIVATE = "_parsenet_reference_src"
_loading = set()


def reference_dir():
    d = os.environ.get(_ENV)
    return d if d and os.path.isdir(d) else None


def extend_package_path(path_list):
    d = reference_dir()
    if d and d not in path_list:
        path_list.append(d)


def _load_reference_module(short):
    """the reference's src/<short>.py as a private module (its own `from src.x import y` resolve to the drop-in first)"""
    d = reference_dir()
    if d is None:
        return None
    name = _PRIVATE + "." + short
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(d, short + ".py")
    if not os.path.isfile(path) or short in _loading:
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    _loading.add(short)
    try:
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    finally:
        _loading.discard(short)
    return mod


def module_getattr(module_name):
    """-> a module-level __getattr__ for the drop-in module `module_name` (e.g. 'src.utils')"""
    short = module_name.rsplit(".", 1)[-1]

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        ref = _load_reference_module(short)
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
        hint = "" if reference_dir() else f" (set {_ENV} to the reference's src/ directory for non-hot-path names)"
        raise AttributeError(f"module {module_name!r} has no attribute {name!r}{hint}")

    return __getattr__


class ReferenceMethods:
    """Mixin for drop-in classes: methods the drop-in class does not define (e.g. `Fit.fit_*_numpy`, `Fit.sample_*`,
    host-side numpy helpers outside the hot path) are taken from the reference class of the same name in the
    reference module of the same name and bound to the drop-in instance.  Opt-in like everything here
    (PARSENET_REFERENCE_SRC); methods the drop-in defines always win."""

    _reference_module = None        # short module name, e.g. "primitive_forward"

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(self)
        ref = _load_reference_module(cls._reference_module) if cls._reference_module else None
        ref_cls = getattr(ref, cls.__name__, None) if ref is not None else None
        attr = getattr(ref_cls, name, None) if ref_cls is not None else None
        if attr is None:
            raise AttributeError(f"{cls.__name__!r} object has no attribute {name!r}")
        return attr.__get__(self, cls) if hasattr(attr, "__get__") else attr
