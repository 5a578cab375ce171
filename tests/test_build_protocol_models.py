"""Python models (CPU-only) of what the round-2 tree pass and per-BLAS sort do differently from the textbook versions, each against an
independent statement:
  * the TILED tree pass of lbvh_build.cu: splits inside a tile are resolved in the tile, a thread climbs for at most R merges and the
    survivors are regrouped (bound: at most TILE / (R + 1) of them), what is left becomes border jobs that carry the deltas at both ends
    of their range; the border phase never looks at a key again (the union of two siblings inherits the outer delta of each) - the result
    must be Karras' tree (tests/test_tree_algorithms_model.py::karras_top_down);
  * the border ARRIVAL protocol: one add on a per-split counter (0 -> first child: deposit, then +1 again = deposit complete; 2 -> the
    deposit is complete: merge; 1 -> look again later), under random interleavings of the jobs' steps: every split is merged exactly once,
    nobody reads an incomplete deposit, everything terminates;
  * the LIGHT segment sort of seg_sort.cuh (512 threads x 22 records counted in two halves of 11, 4-bit counters, packed 16-bit scan):
    the very arithmetic of the kernel against numpy's stable argsort."""
import random

import numpy as np
import pytest

from test_tree_algorithms_model import _delta_fn, _keys, karras_top_down


# ------------------------------------------------------------------------------------------------------------------------------------
# tiled tree pass with regroup + border jobs that carry their deltas
# ------------------------------------------------------------------------------------------------------------------------------------
class Tiled:
    def __init__(self, keys, tile, regroup, seg_of=None):
        self.keys, self.n, self.tile, self.R = keys, len(keys), tile, regroup
        d = _delta_fn(keys)
        self.adj = lambda i: d(i, i + 1) if 0 <= i < self.n - 1 else -1
        self.seg_of = seg_of or (lambda leaf: (0, self.n))
        self.nodes, self.roots, self.max_survivors = {}, [], 0

    def slot_of(self, l, r, dl, dr, seg_count):
        return r if (r - l + 1 != seg_count and dr > dl) else l

    def run_tile(self, L0, order_seed):
        """Returns the border jobs of the tile: (l, r, dl, dr, seg_first, seg_count)."""
        T = min(self.tile, self.n - L0)
        sdelta = [self.adj(L0 - 1 + k) for k in range(T + 1)]              # s_delta[k] = delta(L0 - 1 + k)
        if L0 + T == self.n:
            sdelta[T] = -1
        deposits, jobs, carried = {}, [], []                               # slot -> (side, lt, rt)

        def climb(lt, rt, seg, max_merges):
            merges = 0
            while True:
                if rt - lt + 1 == seg[1]:
                    self.roots.append((L0 + lt, L0 + rt))
                    return None
                right = sdelta[rt + 1] > sdelta[lt]
                slot = rt if right else lt - 1
                if slot < 0 or slot >= self.tile - 1:                     # the kernel's unsigned `slot >= TILE - 1`: next split outside the tile
                    jobs.append((L0 + lt, L0 + rt, sdelta[lt], sdelta[rt + 1], seg[0], seg[1]))
                    carried.append((lt, rt))
                    return None
                side = 0 if right else 1
                if slot not in deposits:
                    deposits[slot] = (side, lt, rt)
                    return None
                oside, olt, ort = deposits.pop(slot)
                assert oside != side
                lt, rt = (lt, ort) if right else (olt, rt)
                g = L0 + slot
                ns = self.slot_of(L0 + lt, L0 + rt, sdelta[lt], sdelta[rt + 1], seg[1])
                assert ns not in self.nodes
                self.nodes[ns] = (L0 + lt, L0 + rt, g)
                merges += 1
                if max_merges and merges == max_merges:
                    return (lt, rt, seg)

        order = list(range(T))
        random.Random(order_seed).shuffle(order)
        survivors = []
        for t in order:                                                    # phase A: every leaf thread, at most R merges
            s = climb(t, t, self.seg_of(L0 + t), self.R)
            if s:
                survivors.append(s)
        if self.R:
            assert len(survivors) <= self.tile // (self.R + 1)             # the bound the regroup buffer is sized by
            for (lt, rt, seg) in survivors:
                assert rt - lt + 1 > self.R                                # a survivor holds more than R leaves
        self.max_survivors = max(self.max_survivors, len(survivors))
        random.Random(order_seed + 1).shuffle(survivors)
        for (lt, rt, seg) in survivors:                                    # phase B: the regrouped survivors climb to the end
            assert climb(lt, rt, seg, 0) is None
        assert len(carried) <= 2                                           # the kernel's s_carry holds two records
        # orphans: a deposit whose sibling straddles the tile border
        for slot, (side, lt, rt) in deposits.items():
            l, r = L0 + lt, L0 + rt
            seg = self.seg_of(L0 + slot)                                   # the kernel takes the segment of leaf `slot` for both sides
            assert seg == self.seg_of(l) == self.seg_of(r)
            jobs.append((l, r, sdelta[lt], sdelta[rt + 1], seg[0], seg[1]))
        return jobs

    def run_border(self, jobs, seed):
        """Border phase without keys: deposits carry (far end, dl, dr)."""
        rng = random.Random(seed)
        jobs = list(jobs)
        rng.shuffle(jobs)
        dep = {}
        for (l, r, dl, dr, sf, sc) in jobs:
            while True:
                if r - l + 1 == sc:
                    self.roots.append((l, r))
                    break
                right = dr > dl
                g = r if right else l - 1
                side = 0 if right else 1
                if g not in dep:
                    dep[g] = (side, l, r, dl, dr)
                    break
                oside, ol, orr, odl, odr = dep.pop(g)
                assert oside != side
                if right:
                    assert ol == g + 1
                    r, dr = orr, odr
                else:
                    assert orr == g
                    l, dl = ol, odl
                ns = self.slot_of(l, r, dl, dr, sc)
                assert ns not in self.nodes
                self.nodes[ns] = (l, r, g)
        assert not dep


def _tiled_tree(keys, tile, regroup, seed, seg_of=None):
    t = Tiled(keys, tile, regroup, seg_of)
    jobs = []
    tiles = list(range(0, len(keys), tile))
    random.Random(seed).shuffle(tiles)
    for L0 in tiles:
        jobs += t.run_tile(L0, seed * 31 + L0)
    t.run_border(jobs, seed + 5)
    return t


@pytest.mark.parametrize("kind", ["random", "dups", "alleq"])
@pytest.mark.parametrize("tile,regroup", [(8, 0), (8, 1), (16, 2), (32, 3), (64, 2)])
def test_tiled_regrouped_tree_equals_karras(kind, tile, regroup):
    rng = np.random.default_rng(tile * 13 + regroup + len(kind))
    for n in (2, 7, tile, tile + 1, 5 * tile + 3, 333):
        keys = _keys(kind, n, rng)
        ref = karras_top_down(keys)
        for seed in range(3):
            t = _tiled_tree(keys, tile, regroup, seed)
            assert t.nodes == ref, (kind, tile, regroup, n, seed)
            assert t.roots == [(0, n - 1)]


def test_tiled_tree_with_segments():
    """A batch: the BLAS id is the key prefix, every segment is one subtree; splits between segments are never nodes."""
    rng = np.random.default_rng(5)
    sizes = [1, 2, 37, 3, 120, 64, 5]
    keys, bounds, first = [], [], 0
    for b, s in enumerate(sizes):
        k = np.sort(rng.integers(0, 1 << 30, size=s, dtype=np.uint64)) | (np.uint64(b) << np.uint64(30))
        keys.append(k)
        bounds.append((first, s))
        first += s
    keys = np.concatenate(keys)
    owner = np.repeat(np.arange(len(sizes)), sizes)
    seg_of = lambda leaf: bounds[int(owner[leaf])]
    for tile, regroup in ((16, 2), (32, 0), (64, 3)):
        for seed in range(3):
            t = _tiled_tree(keys, tile, regroup, seed, seg_of)
            assert sorted(t.roots) == sorted((f, f + s - 1) for f, s in bounds)
            expect = {}
            for (f, s) in bounds:                                       # every segment on its own: Karras' tree, slots relative to the segment
                if s > 1:
                    for slot, (l, r, g) in karras_top_down(keys[f:f + s]).items():
                        expect[f + slot] = (f + l, f + r, f + g)
            assert t.nodes == expect, (tile, regroup, seed)


# ------------------------------------------------------------------------------------------------------------------------------------
# arrival protocol of the border kernel under random interleavings
# ------------------------------------------------------------------------------------------------------------------------------------
def test_border_arrival_protocol_interleavings():
    """Each split gets exactly two arrivals. An arrival is an atomic add (+1) returning the old value: 0 -> first: (later) store the deposit,
    (later) add 1 again; 1 -> wait: re-read the counter on a later turn until it is 3; 2 -> merge. Steps of different jobs interleave at random."""
    for seed in range(200):
        rng = random.Random(seed)
        n_splits = rng.randint(1, 12)
        counter = [0] * n_splits
        deposit_done = [False] * n_splits
        merged = [0] * n_splits
        # a job = a list of pending micro-steps on one split
        jobs = []
        for s in range(n_splits):
            jobs.append({"split": s, "state": "arrive"})
            jobs.append({"split": s, "state": "arrive"})
        steps = 0
        while jobs:
            steps += 1
            assert steps < 10000, "protocol does not terminate"
            j = rng.choice(jobs)
            s = j["split"]
            if j["state"] == "arrive":
                seen = counter[s]
                counter[s] += 1
                if seen == 0:
                    j["state"] = "store"
                elif seen == 1:
                    j["state"] = "wait"
                else:
                    assert seen == 2 and deposit_done[s]
                    merged[s] += 1
                    jobs.remove(j)
            elif j["state"] == "store":
                deposit_done[s] = True
                j["state"] = "release"
            elif j["state"] == "release":
                counter[s] += 1
                jobs.remove(j)
            elif j["state"] == "wait":
                if counter[s] == 3:
                    assert deposit_done[s]
                    merged[s] += 1
                    jobs.remove(j)
        assert merged == [1] * n_splits and counter == [3] * n_splits


# ------------------------------------------------------------------------------------------------------------------------------------
# the light segment sort, with the kernel's own packed arithmetic
# ------------------------------------------------------------------------------------------------------------------------------------
T, ITEMS, HALF, W = 512, 22, 11, 16
CAP = T * ITEMS


def _expand4(cnt, q):
    f = (cnt >> (16 * q)) & 0xFFFF
    return (f & 15) | (((f >> 4) & 15) << 16) | (((f >> 8) & 15) << 32) | (((f >> 12) & 15) << 48)


def _light_pass(s_m, s_id, shift):
    m = s_m.reshape(T, ITEMS)
    dig = ((m >> np.uint64(shift)) & np.uint64(15)).astype(np.int64)
    cnt_a, cnt_b = [0] * T, [0] * T
    rank = np.zeros((T, ITEMS), dtype=np.int64)
    for t in range(T):
        ca = cb = 0
        for i in range(HALF):
            d4 = 4 * int(dig[t, i])
            rank[t, i] = (ca >> d4) & 15
            ca += 1 << d4
        for i in range(HALF):
            d4 = 4 * int(dig[t, HALF + i])
            rank[t, HALF + i] = ((ca >> d4) & 15) + ((cb >> d4) & 15)      # behind all of the first half's records of that digit
            cb += 1 << d4
        cnt_a[t], cnt_b[t] = ca, cb
    ex = [[0] * 4 for _ in range(T)]
    wsum = [[0] * W for _ in range(4)]
    tot = [0] * 4
    for q in range(4):
        for w in range(W):
            inc = 0
            for lane in range(32):
                t = w * 32 + lane
                ex[t][q] = inc
                inc += _expand4(cnt_a[t], q) + _expand4(cnt_b[t], q)
            wsum[q][w] = inc
        inc = 0
        for w in range(W):
            own, wsum[q][w] = wsum[q][w], inc
            inc += own
        tot[q] = inc
    new_m, new_id = np.empty_like(s_m), np.empty_like(s_id)
    seen = np.zeros(CAP, dtype=bool)
    idv = s_id.reshape(T, ITEMS)
    for t in range(T):
        off, dbase = [0] * 16, 0
        for q in range(4):
            e = ex[t][q] + wsum[q][t >> 5]
            for k in range(4):
                off[4 * q + k] = (dbase + ((e >> (16 * k)) & 0xFFFF)) & 0xFFFF
                dbase += (tot[q] >> (16 * k)) & 0xFFFF
        for i in range(ITEMS):
            pos = off[int(dig[t, i])] + int(rank[t, i])
            assert not seen[pos]
            seen[pos] = True
            new_m[pos], new_id[pos] = m[t, i], idv[t, i]
    assert seen.all()
    return new_m, new_id


@pytest.mark.parametrize("n,kind", [(300, "random"), (9800, "random"), (CAP, "few"), (5000, "few")])
def test_light_segment_sort_arithmetic(n, kind):
    rng = np.random.default_rng(n)
    s_m = np.full(CAP, 0xFFFFFFFF, dtype=np.uint64)
    s_m[:n] = rng.integers(0, 1 << 30, size=n, dtype=np.uint64) if kind == "random" else rng.integers(0, 37, size=n, dtype=np.uint64) * np.uint64(0x1234567) % np.uint64(1 << 30)
    s_id = np.arange(CAP, dtype=np.uint64)
    order = np.argsort(s_m[:n], kind="stable")
    want_m, want_id = s_m[:n][order], order.astype(np.uint64)
    for shift in range(0, 32, 4):
        s_m, s_id = _light_pass(s_m, s_id, shift)
    assert np.array_equal(s_m[:n], want_m) and np.array_equal(s_id[:n], want_id)
    assert (s_m[n:] == 0xFFFFFFFF).all()                                   # the padding sorts last and stays last


def test_light_sort_row_pitch_division():
    """seg2_m_at: p + p / 22 through a multiply-high (conflict-free row pitch of 23 words per thread)."""
    p = np.arange(0, 1 << 16, dtype=np.uint64)
    assert np.array_equal((p * np.uint64(195225787)) >> np.uint64(32), p // np.uint64(22))
