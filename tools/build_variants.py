"""Builds tuning variants of libfbgpu.so (different eval-kernel launch shapes) for one-shot A/B runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from featurebase_b200 import build as B
VARIANTS = {   # name -> -D defines; edit freely, outputs featurebase_b200/libfbgpu_<name>.so (select with FBGPU_LIB=...)
    "wp_unroll3": ["FBGPU_WP_UNROLL3"],          # experimental fixed-register op loop of the word-parallel kernel (csrc/wp_machine.h)
    "pair_unscatter": ["FBGPU_PAIR_UNSCATTER"],  # pair_count_kernel array x array: clear the a-side bits after the probe instead of wiping 8 KiB per pair
}
only = sys.argv[1:]
if only:
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in only}
if __name__ == "__main__":
    for name, defs in VARIANTS.items():
        out = B.build_fbgpu(force=True, defines=defs, out_name=f"libfbgpu_{name}.so")
        print(out)
