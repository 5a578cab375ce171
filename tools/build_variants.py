"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

def _split(t0, t1):
    return ["RT_SPLIT_SHADE=1", f"RT_SPLIT_THRESHOLD_S0={t0}", f"RT_SPLIT_THRESHOLD_S1={t1}"]


VARIANTS = {
    "base": [],
    "tile128": ["RT_TREE_TILE=128"],
    "thr8": ["RT_REFILL_THRESHOLD=8"], "thr10": ["RT_REFILL_THRESHOLD=10"], "thr14": ["RT_REFILL_THRESHOLD=14"], "thr16": ["RT_REFILL_THRESHOLD=16"],
    "cap0": ["RT_NODE_CAP=0"], "cap2": ["RT_NODE_CAP=2"], "cap6": ["RT_NODE_CAP=6"], "cap8": ["RT_NODE_CAP=8"],
    "blk7": ["RT_TRACE_MIN_BLOCKS=7"], "blk6": ["RT_TRACE_MIN_BLOCKS=6"], "blk10": ["RT_TRACE_MIN_BLOCKS=10"],
    "fastslab": ["RT_FAST_SLAB=1"],
    "leaf1": ["RT_BLAS_LEAF_MAX=1"],
    "leaf2": ["RT_BLAS_LEAF_MAX=2"],
    "leaf3": ["RT_BLAS_LEAF_MAX=3"],
    "leaf6": ["RT_BLAS_LEAF_MAX=6"],
    "leaf8": ["RT_BLAS_LEAF_MAX=8"],
    "setup1": ["RT_SETUP_BATCH=1"],
    "setup2": ["RT_SETUP_BATCH=2"],
    "setup8": ["RT_SETUP_BATCH=8"],
    "bounce_unordered": ["RT_BOUNCE_ORDERED=0", "RT_FUSED_STAGES=0"],
    "two_launches": ["RT_FUSED_STAGES=0"],
    "always_fused": ["RT_FUSED_MAX_PIXELS=4000000000u"],
    "match": ["RT_SORT_USE_MATCH=1"],
    "ballot_5": ["RT_SORT_MIN_CTAS=5"],
    "ballot_6": ["RT_SORT_MIN_CTAS=6"],
    "lb1": ["RT_LOOKBACK_WINDOW=1"],
    "lb4": ["RT_LOOKBACK_WINDOW=4"],
    "lb16": ["RT_LOOKBACK_WINDOW=16"],
    "lb8_6": ["RT_LOOKBACK_WINDOW=8", "RT_SORT_MIN_CTAS=6"],
    "sort8_5": ["RT_SORT_ITEMS=8", "RT_SORT_MIN_CTAS=5"],
    "sort8_6": ["RT_SORT_ITEMS=8", "RT_SORT_MIN_CTAS=6"],
    "sort12_5": ["RT_SORT_ITEMS=12", "RT_SORT_MIN_CTAS=5"],
    "sort12_6": ["RT_SORT_ITEMS=12", "RT_SORT_MIN_CTAS=6"],
    "sort16_4": ["RT_SORT_ITEMS=16", "RT_SORT_MIN_CTAS=4"],
    "sort16_3": ["RT_SORT_ITEMS=16", "RT_SORT_MIN_CTAS=3"],
    "tile256": ["RT_TREE_TILE=256"],
    "tile512": ["RT_TREE_TILE=512"],
    "tile64": ["RT_TREE_TILE=64"],
    "tile256_4": ["RT_TREE_TILE=256", "RT_TREE_MIN_CTAS=4"],
    "tile512_2": ["RT_TREE_TILE=512", "RT_TREE_MIN_CTAS=2"],
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
