"""Python models of the hierarchy algorithms (CPU-only). The reference is Karras 2012 top-down (what the oracle restates in C++);
against it: (1) the bottom-up construction the CUDA tree pass uses (a finished subtree [l, r] merges right iff delta(r) > delta(l-1);
nodes numbered so that the children of split g are slots g and g + 1), in an arbitrary merge order; (2) the thread-coarsened variant
planned for the next round: every "thread" first merges a run of K consecutive leaves with a stack, only the leftover
pieces enter the arrival protocol. Both must reproduce Karras' tree exactly: same slots, ranges and splits, duplicates included."""
import random

import numpy as np
import pytest


def _delta_fn(keys):
    n = len(keys)

    def clz64(v):
        return 64 - int(v).bit_length()

    def clz32(v):
        return 32 - int(v).bit_length()

    def delta(i, j):
        if j < 0 or j >= n:
            return -1
        if keys[i] == keys[j]:
            return 64 + clz32(i ^ j)
        return clz64(int(keys[i]) ^ int(keys[j]))
    return delta


def karras_top_down(keys):
    """{slot i: (first, last, gamma)} for the n - 1 internal nodes (Karras 2012, Fig. 4, index-augmented keys)."""
    n, delta, out = len(keys), _delta_fn(keys), {}
    for i in range(n - 1):
        d = 1 if (i == 0 or delta(i, i + 1) - delta(i, i - 1) >= 0) else -1
        dmin = delta(i, i - d)
        lmax = 2
        while delta(i, i + lmax * d) > dmin:
            lmax *= 2
        l, t = 0, lmax // 2
        while t >= 1:
            if delta(i, i + (l + t) * d) > dmin:
                l += t
            t //= 2
        j = i + l * d
        dnode = delta(i, j)
        s, t = 0, l
        while True:
            t = (t + 1) // 2
            if delta(i, i + (s + t) * d) > dnode:
                s += t
            if t <= 1:
                break
        gamma = i + s * d + min(d, 0)
        out[i] = (min(i, j), max(i, j), gamma)
    return out


class BottomUp:
    """The arrival protocol: subtrees meet at their split; the second arriver creates the node and climbs on."""

    def __init__(self, keys):
        self.n, self.delta, self.nodes, self.waiting = len(keys), _delta_fn(keys), {}, {}

    def adj(self, i):                       # delta between sorted primitives i and i + 1 (-1 outside)
        return self.delta(i, i + 1) if 0 <= i < self.n - 1 else -1

    def merges_right(self, l, r):
        return self.adj(r) > self.adj(l - 1)

    def slot_of(self, l, r):                # Karras numbering: a left child sits at its right end, a right child at its left end
        if l == 0 and r == self.n - 1:
            return 0
        return r if self.merges_right(l, r) else l

    def make_node(self, l, g, r):
        slot = self.slot_of(l, r)
        assert slot not in self.nodes, f"slot {slot} written twice"
        self.nodes[slot] = (l, r, g)

    def climb(self, l, r):
        """A finished subtree [l, r] arrives; returns when it waits for its sibling or has completed the root."""
        while not (l == 0 and r == self.n - 1):
            right = self.merges_right(l, r)
            g = r if right else l - 1
            if g not in self.waiting:
                self.waiting[g] = (l, r)
                return
            sl, sr = self.waiting.pop(g)
            l, r = (l, sr) if right else (sl, r)
            assert (sl == g + 1) if right else (sr == g)
            self.make_node(l, g, r)


def bottom_up(keys, order):
    b = BottomUp(keys)
    for leaf in order:
        b.climb(leaf, leaf)
    assert not b.waiting
    return b.nodes


def coarsened(keys, K, order):
    """Every thread owns K consecutive leaves: stack-merge inside the run (siblings merge when they select EACH OTHER), then the
    leftover pieces go through the arrival protocol in any order."""
    b = BottomUp(keys)
    n = len(keys)
    pieces = []
    for c0 in range(0, n, K):
        stack = []
        for leaf in range(c0, min(n, c0 + K)):
            cur = (leaf, leaf)
            while stack:
                tl, tr = stack[-1]
                g = tr
                assert cur[0] == g + 1
                if (tl == 0 and tr == n - 1) or not (b.merges_right(tl, tr) and not b.merges_right(cur[0], cur[1])):
                    break
                stack.pop()
                cur = (tl, cur[1])
                b.make_node(cur[0], g, cur[1])
            stack.append(cur)
        pieces.extend(stack)
    for idx in order(len(pieces)):
        b.climb(*pieces[idx])
    assert not b.waiting
    return b.nodes


def _keys(kind, n, rng):
    if kind == "random":
        k = np.sort(rng.integers(0, 1 << 30, size=n, dtype=np.uint64))
    elif kind == "dups":
        k = np.sort(rng.integers(0, max(2, n // 8), size=n, dtype=np.uint64) << np.uint64(7))
    elif kind == "alleq":
        k = np.full(n, 12345, dtype=np.uint64)
    elif n < 5:
        k = np.sort(rng.integers(0, 1 << 30, size=n, dtype=np.uint64))
    else:                                   # two segments (BLAS id as key prefix), one of them tiny
        a = np.sort(rng.integers(0, 1 << 30, size=n - 3, dtype=np.uint64))
        b_ = (np.uint64(1) << np.uint64(30)) | np.sort(rng.integers(0, 1 << 30, size=3, dtype=np.uint64))
        k = np.concatenate([a, b_])
    return k


@pytest.mark.parametrize("kind", ["random", "dups", "alleq", "segments"])
@pytest.mark.parametrize("n", [2, 3, 17, 200])
def test_bottom_up_equals_karras(kind, n):
    rng = np.random.default_rng(n * 7 + len(kind))
    keys = _keys(kind, n, rng)
    ref = karras_top_down(keys)
    for seed in range(3):
        order = list(range(n))
        random.Random(seed).shuffle(order)
        assert bottom_up(keys, order) == ref
    assert bottom_up(keys, range(n)) == ref and bottom_up(keys, range(n - 1, -1, -1)) == ref


@pytest.mark.parametrize("kind", ["random", "dups", "alleq", "segments"])
@pytest.mark.parametrize("K", [2, 4, 8])
def test_coarsened_equals_karras(kind, K):
    rng = np.random.default_rng(K * 11 + len(kind))
    for n in (2, 5, 64, 333):
        keys = _keys(kind, n, rng)
        ref = karras_top_down(keys)
        for seed in range(3):
            def order(m, seed=seed):
                o = list(range(m))
                random.Random(seed).shuffle(o)
                return o
            assert coarsened(keys, K, order) == ref, (kind, K, n, seed)


def test_oracle_lbvh_is_karras_tree(oracle):
    """The oracle's C++ LBVH (what the GPU build is compared with bit for bit) against the Python statement of the paper above: walking
    the exported tree from the root, every live node sits in Karras' slot and splits its range exactly where the paper says; leaves hold
    at most two primitives and together cover every sorted primitive once."""
    from build_up_phase_b200 import scenes
    for scene in (scenes.tess_scene(nx=23, ny=17, width=32, height=32, bounces=0), scenes.duplicate_key_scene(700)):
        o = oracle.OracleScene(scene)
        info, nodes, tris, keys, prims = o.blas_export(0)
        o.close()
        n = info.triangle_count
        ref = karras_top_down(keys)
        covered = np.zeros(n, dtype=np.int32)

        def span(ref_, slot_range):
            r = int(np.int32(ref_))
            if r < 0:
                first, count = (~r) >> 3, ((~r) & 7) + 1
                assert count <= 2
                covered[first:first + count] += 1
                return first, first + count - 1
            return ref[r][0], ref[r][1]

        stack = [int(info.root_ref)]
        live = 0
        while stack:
            slot = stack.pop()
            first, last, gamma = ref[slot]
            lref, rref = nodes[slot][6], nodes[slot][14]
            assert span(lref, None) == (first, gamma), (slot, "left child range")
            assert span(rref, None) == (gamma + 1, last), (slot, "right child range")
            for r in (lref, rref):
                if int(np.int32(r)) >= 0:
                    assert int(np.int32(r)) in (gamma, gamma + 1)          # Karras numbering: children of split gamma
                    stack.append(int(np.int32(r)))
            live += 1
        assert np.all(covered == 1) and live >= n // 4


def live_rule(keys, i, leaf_max=2):
    """The product's compaction rule (lbvh_build.cu slot_is_live), from the sorted keys alone: slot i is written by the build iff Karras' node i
    covers more than leaf_max leaves, i.e. iff delta(i, i + leaf_max * d) > delta(i, i - d) with d the node's direction."""
    n, delta = len(keys), _delta_fn(keys)
    if n <= leaf_max or i > n - 2:
        return False
    dr, dl = delta(i, i + 1), delta(i, i - 1)
    d = 1 if dr > dl else -1
    dmin = dl if d > 0 else dr
    return delta(i, i + leaf_max * d) > dmin


@pytest.mark.parametrize("kind", ["random", "dups", "alleq", "segments"])
@pytest.mark.parametrize("leaf_max", [1, 2, 4])
def test_compaction_live_rule_equals_karras_range_sizes(kind, leaf_max):
    """rt_compact_blas decides which Karras slots hold a node WITHOUT reading the (uninitialised) dead slots: from the sorted records.
    Against the paper's own ranges: node i is written iff its range holds more than leaf_max leaves."""
    rng = np.random.default_rng(leaf_max * 13 + len(kind))
    for n in (2, 3, 4, 9, 64, 333):
        keys = _keys(kind, n, rng) if kind != "segments" else np.sort(rng.integers(0, 1 << 30, size=n, dtype=np.uint64))   # the rule works per segment
        ref = karras_top_down(keys)
        for i in range(n - 1):
            first, last, _ = ref[i]
            assert live_rule(keys, i, leaf_max) == (last - first + 1 > leaf_max), (kind, leaf_max, n, i)


def test_compaction_live_rule_equals_reachable_nodes_of_the_oracle_tree(oracle):
    """... and against the oracle's finished LBVH: the slots the rule calls live are exactly the nodes reachable from the root."""
    from build_up_phase_b200 import scenes
    for scene in (scenes.tess_scene(nx=23, ny=17, width=32, height=32, bounces=0), scenes.duplicate_key_scene(700)):
        o = oracle.OracleScene(scene)
        info, nodes, tris, keys, prims = o.blas_export(0)
        o.close()
        reach, stack = set(), [int(info.root_ref)]
        while stack:
            s = stack.pop()
            reach.add(s)
            for ref_ in (int(np.int32(nodes[s][6])), int(np.int32(nodes[s][14]))):
                if 0 <= ref_ < 0x7FFFFFF0:
                    stack.append(ref_)
        rule = {i for i in range(info.triangle_count - 1) if live_rule(keys, i, 2)}
        assert rule == reach and 0.3 * info.triangle_count < len(rule) < 0.8 * info.triangle_count
