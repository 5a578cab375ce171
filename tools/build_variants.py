"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

VARIANTS = {
    "base": [],
    "fastslab": ["RT_FAST_SLAB=1"],
    "postpone": ["RT_POSTPONE_LEAF=1"],
    "fastslab_postpone": ["RT_FAST_SLAB=1", "RT_POSTPONE_LEAF=1"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
