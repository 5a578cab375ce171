"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

def _split(t0, t1):
    return ["RT_SPLIT_SHADE=1", f"RT_SPLIT_THRESHOLD_S0={t0}", f"RT_SPLIT_THRESHOLD_S1={t1}"]


VARIANTS = {
    "base": [],
    "split_16_16": _split(16, 16),
    "split_20_20": _split(20, 20),
    "split_24_24": _split(24, 24),
    "split_28_28": _split(28, 28),
    "split_24_16": _split(24, 16),
    "split_28_20": _split(28, 20),
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
