"""Builds tuning variants of libfbgpu.so (different eval-kernel launch shapes) for one-shot A/B runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from featurebase_b200 import build as B
VARIANTS = {   # name -> -D defines; edit freely, outputs featurebase_b200/libfbgpu_<name>.so (select with FBGPU_LIB=...)
    "eval_b7": ["FBGPU_EVAL_MIN_BLOCKS=7"],
    "wp2": ["FBGPU_WP_SLICES=2"],
}
if __name__ == "__main__":
    for name, defs in VARIANTS.items():
        out = B.build_fbgpu(force=True, defines=defs, out_name=f"libfbgpu_{name}.so")
        print(out)
