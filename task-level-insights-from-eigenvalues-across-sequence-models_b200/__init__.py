"""eigb200 -- B200-native (sm_100a) eigenvalue analysis for sequence models.

Drop-in for the hot path of the reference's `analysis/eval_eig.py`: the `get_eig_*` extractors, `threshold_analysis`,
the layer recurrences they drive and `eval_eig` itself, executed by hand-written CUDA kernels in `libeigb200.so`
(C ABI: include/eigb200.h).  No CPU fallback: importing is cheap, calling without the built library or without a CUDA
device raises `Eigb200Error`.
"""
from ._lib import Eigb200Error, LIB_PATH, load as load_library  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Eigb200Error", "LIB_PATH", "load_library", "ops"]


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import eigb200` stays light
    if name in ("ops", "extractors", "analysis", "layers", "dist", "ssm"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
