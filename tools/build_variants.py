"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

VARIANTS = {
    "base": [],
    "thr12": ["RT_REFILL_THRESHOLD=12"],
    "thr16": ["RT_REFILL_THRESHOLD=16"],
    "cap0": ["RT_NODE_CAP=0"],
    "cap4": ["RT_NODE_CAP=4"],
    "cap4_thr12": ["RT_NODE_CAP=4", "RT_REFILL_THRESHOLD=12"],
    "cap8_thr12": ["RT_NODE_CAP=8", "RT_REFILL_THRESHOLD=12"],
    "cap12_thr12": ["RT_NODE_CAP=12", "RT_REFILL_THRESHOLD=12"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
