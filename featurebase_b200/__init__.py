"""featurebase_b200 — B200-native roaring-bitmap query executor for FeatureBase's hot path.

csrc/      CUDA kernels (sm_100a) + C++ host runtime behind the C ABI in include/fbgpu.h  -> libfbgpu.so
lib.py     ctypes binding (no CPU fallback)
executor.py / pql.py / roaring_io.py   host-side mirror of the reference's executor interface for this path
datagen.py synthetic fragments (tests, bench)
"""
from . import lib  # noqa: F401
from .lib import Context, FbgpuError  # noqa: F401
