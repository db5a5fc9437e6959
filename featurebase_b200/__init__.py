"""featurebase_b200 — B200-native roaring-bitmap query executor for FeatureBase's hot path.

csrc/      CUDA kernels (sm_100a) + C++ host runtime behind the C ABI in include/fbgpu.h  -> libfbgpu.so
lib.py     ctypes binding (no CPU fallback)
executor.py / pql.py / roaring_io.py   host-side mirror of the reference's executor interface for this path
datagen.py synthetic fragments (tests, bench)
"""
import importlib

_LAZY = {"lib": None, "Context": "lib", "FbgpuError": "lib"}


def __getattr__(name):
    # lazy: importing the package (e.g. for datagen in the CPU reference arm of bench.py) must not touch the CUDA binding
    if name in _LAZY:
        mod = importlib.import_module(".lib", __name__)
        return mod if _LAZY[name] is None else getattr(mod, name)
    raise AttributeError(name)
