"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

VARIANTS = {
    "base": [],
    "mb6": ["RT_TRACE_MIN_BLOCKS=6"],
    "mb5": ["RT_TRACE_MIN_BLOCKS=5"],
    "thr8": ["RT_REFILL_THRESHOLD=8"],
    "thr20": ["RT_REFILL_THRESHOLD=20"],
    "thr26": ["RT_REFILL_THRESHOLD=26"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
