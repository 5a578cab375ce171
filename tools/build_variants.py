"""Builds tuning variants of librtcore into build-up-phase_b200/build/ (compared on the GPU by tools/gpu_variants.sh).
Every entry is a list of -D flags; the measured outcome of each family is recorded in profiles/README.md."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import build as b

VARIANTS = {
    "base": [],
    # trace
    "two_launches": ["RT_FUSED_STAGES=0"],                       # never fuse the two stages (r01u)
    "always_fused": ["RT_FUSED_MAX_PIXELS=4000000000u"],         # always fuse
    "bounce_unordered": ["RT_BOUNCE_ORDERED=0", "RT_FUSED_STAGES=0"],   # append queue instead of the tile-ordered index (r01p)
    "thr8": ["RT_REFILL_THRESHOLD=8"], "thr10": ["RT_REFILL_THRESHOLD=10"], "thr14": ["RT_REFILL_THRESHOLD=14"], "thr16": ["RT_REFILL_THRESHOLD=16"],
    "cap0": ["RT_NODE_CAP=0"], "cap2": ["RT_NODE_CAP=2"], "cap4": ["RT_NODE_CAP=4"], "cap8": ["RT_NODE_CAP=8"],
    "blk7": ["RT_TRACE_MIN_BLOCKS=7"], "blk6": ["RT_TRACE_MIN_BLOCKS=6"], "blk10": ["RT_TRACE_MIN_BLOCKS=10"],
    "ieee_slab": ["RT_FAST_SLAB=0"],
    "leaf1": ["RT_BLAS_LEAF_MAX=1"], "leaf3": ["RT_BLAS_LEAF_MAX=3"], "leaf4": ["RT_BLAS_LEAF_MAX=4"], "leaf8": ["RT_BLAS_LEAF_MAX=8"],   # (r02a)
    # ray numbering / fetch counters (r2_f) and the shared-memory short stack
    "reg0": ["RT_REGIONS=0"], "reg1": ["RT_REGIONS=1"], "reg2": ["RT_REGIONS=2"],
    "nosmemtlas": ["RT_SMEM_TLAS=0"], "smemtlas": ["RT_SMEM_TLAS=1"], "ldg256": ["RT_LDG256=1"], "tos": ["RT_STACK_TOS=1"], "branchless": ["RT_BRANCHLESS_NODE=1"], "minmax": ["RT_SLAB_MINMAX=1"], "no_oconst": ["RT_PRIMARY_ORIGIN_CONST=0"], "oconst_blk10": ["RT_TRACE_MIN_BLOCKS=10"], "blk9": ["RT_TRACE_MIN_BLOCKS=9"], "minmax_blk10": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=10"],
    "minmax_blk9_thr10": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=9", "RT_REFILL_THRESHOLD=10"], "minmax_blk9_thr14": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=9", "RT_REFILL_THRESHOLD=14"],
    "minmax_blk9_cap8": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=9", "RT_NODE_CAP=8"], "minmax_blk9_s1blk8": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=9", "RT_TRACE_MIN_BLOCKS_S1=8"],
    "minmax_blk8_s1blk9": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=8", "RT_TRACE_MIN_BLOCKS_S1=9"], "minmax_blk9": ["RT_SLAB_MINMAX=1", "RT_TRACE_MIN_BLOCKS=9"], "minmax_cap8": ["RT_SLAB_MINMAX=1", "RT_NODE_CAP=8"], "branchless_cap8": ["RT_BRANCHLESS_NODE=1", "RT_NODE_CAP=8"], "branchless_thr10": ["RT_BRANCHLESS_NODE=1", "RT_REFILL_THRESHOLD=10"], "s1blk10": ["RT_TRACE_MIN_BLOCKS_S1=10"], "s1blk7": ["RT_TRACE_MIN_BLOCKS_S1=7"], "s1blk6": ["RT_TRACE_MIN_BLOCKS_S1=6"],
    "tos_s1blk7": ["RT_STACK_TOS=1", "RT_TRACE_MIN_BLOCKS_S1=7"], "ldg256_blk7": ["RT_LDG256=1", "RT_TRACE_MIN_BLOCKS=7"], "smemtlas_reg2": ["RT_REGIONS=2"], "smemtlas_stack8": ["RT_SMEM_STACK=8"],
    "smemtlas_cap8": ["RT_NODE_CAP=8"], "smemtlas_thr16": ["RT_REFILL_THRESHOLD=16"],
    "reg2_8x8": ["RT_REGION_TW=8", "RT_REGION_TH=8"], "reg2_32x16": ["RT_REGION_TW=32", "RT_REGION_TH=16"], "reg2_8x16": ["RT_REGION_TW=8", "RT_REGION_TH=16"],
    "reg1_32x32": ["RT_REGIONS=1", "RT_REGION_TW=32", "RT_REGION_TH=32"],
    "smem8": ["RT_SMEM_STACK=8"], "smem16": ["RT_SMEM_STACK=16"], "reg0_smem8": ["RT_REGIONS=0", "RT_SMEM_STACK=8"],
    "capevery2": ["RT_CAP_EVERY=2"],
    # build
    "inline_border": ["RT_TREE_INLINE_BORDER=1"], "inline_border_t256": ["RT_TREE_INLINE_BORDER=1", "RT_TREE_TILE=256"],
    "tile64": ["RT_TREE_TILE=64"], "tile256": ["RT_TREE_TILE=256"], "tile512": ["RT_TREE_TILE=512"],
    "sort8_6": ["RT_SORT_ITEMS=8", "RT_SORT_MIN_CTAS=6"], "sort12_6": ["RT_SORT_ITEMS=12", "RT_SORT_MIN_CTAS=6"],
    "ballot": ["RT_SORT_USE_BALLOT=1"],
    "lb1": ["RT_LOOKBACK_WINDOW=1"], "lb16": ["RT_LOOKBACK_WINDOW=16"],
    "setup2": ["RT_SETUP_BATCH=2"], "setup4": ["RT_SETUP_BATCH=4"],
    # r2_x: regrouped tile climb (RT_TREE_REGROUP merges per thread, then the survivors finish in the first warp(s)); trace cache hints
    "rg0": ["RT_TREE_REGROUP=0"], "rg1": ["RT_TREE_REGROUP=1"], "rg2": ["RT_TREE_REGROUP=2"], "rg4": ["RT_TREE_REGROUP=4"], "rg6": ["RT_TREE_REGROUP=6"],
    "tile256_rg3": ["RT_TREE_TILE=256"], "tile256_rg5": ["RT_TREE_TILE=256", "RT_TREE_REGROUP=5"], "tile512_rg3": ["RT_TREE_TILE=512"],
    "tile64_rg3": ["RT_TREE_TILE=64"],
    "tri_noalloc": ["RT_TRI_NOALLOC=1"], "tri_evict_first": ["RT_TRI_NOALLOC=2"], "carve0": ["RT_L1_CARVEOUT=0"], "carve0_tri_noalloc": ["RT_L1_CARVEOUT=0", "RT_TRI_NOALLOC=1"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or ["base"]
    for n in names:
        print(n, b.build_variant(n, VARIANTS[n]))
