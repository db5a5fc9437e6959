"""Builds tuning variants of libfbgpu.so (different eval-kernel launch shapes) for one-shot A/B runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from featurebase_b200 import build as B
VARIANTS = {
    "b7": ["FBGPU_EVAL_MIN_BLOCKS=7"],
    "t512": ["FBGPU_EVAL_THREADS=512", "FBGPU_EVAL_MIN_BLOCKS=4"],
}
if __name__ == "__main__":
    for name, defs in VARIANTS.items():
        out = B.build_fbgpu(force=True, defines=defs, out_name=f"libfbgpu_{name}.so")
        print(out)
