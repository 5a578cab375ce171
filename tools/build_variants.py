"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

VARIANTS = {
    "base": [],
    "pm1": ["RT_PRIM_MIN=1"],
    "pm4": ["RT_PRIM_MIN=4"],
    "pm12": ["RT_PRIM_MIN=12"],
    "pm16": ["RT_PRIM_MIN=16"],
    "pm12_thr20": ["RT_PRIM_MIN=12", "RT_REFILL_THRESHOLD=20"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
