"""Builds tuning variants of libfbgpu.so (different eval-kernel launch shapes) for one-shot A/B runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from featurebase_b200 import build as B
VARIANTS = {
    "v1_b6": ["FBGPU_BATCH_IMPL=1", "FBGPU_EVAL_MIN_BLOCKS=6"],
    "v1_b7": ["FBGPU_BATCH_IMPL=1", "FBGPU_EVAL_MIN_BLOCKS=7"],
    "v1_b8": ["FBGPU_BATCH_IMPL=1", "FBGPU_EVAL_MIN_BLOCKS=8"],
}
if __name__ == "__main__":
    for name, defs in VARIANTS.items():
        out = B.build_fbgpu(force=True, defines=defs, out_name=f"libfbgpu_{name}.so")
        print(out)
