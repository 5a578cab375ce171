// radix_sort.cu — "onesweep" LSD radix sort of (u64 Morton key, u32 primitive id) pairs for sm_100a.
//
// Two record formats: (u64 key, u32 value) pairs, or PACKED 64-bit words `key << val_bits | value` when
// key_bits + val_bits <= 64 (every configuration of BASELINE.json): then only 8 B are read and 8 B written per
// element and pass instead of 12 + 12.
// One up-front histogram kernel reads the keys once and produces the global digit histogram of
// EVERY 8-bit pass; each pass is then a single kernel: a tile (3072 pairs) is ranked in
// shared memory with warp-level match_any multi-split, its per-digit counts are published to a
// (tile x 256) state array, and the exclusive prefix over preceding tiles is obtained by
// decoupled look-back (status|value packed in one 32-bit word, so no fence is needed between
// them). Tile ids are handed out by an atomic counter so a tile only ever waits on tiles that
// have already started (forward progress without co-residency assumptions). Keys and values are
// reordered through shared memory so global stores are coalesced runs. HBM traffic per pass is
// one read + one write of each pair (24 B/pair), the bound this kernel is measured against.
//
// A spin watchdog turns a (theoretically impossible) look-back stall into an error flag instead
// of a hung GPU.
#include "rt_internal.h"
#include "seg_sort.cuh"

namespace rt {

namespace {

constexpr int RADIX = 256;
constexpr int SORT_THREADS = 256;            // == RADIX: thread d owns digit d in the scan / look-back
constexpr int SORT_WARPS = SORT_THREADS / 32;
#ifndef RT_SORT_ITEMS
#define RT_SORT_ITEMS 12
#endif
#ifndef RT_SORT_MIN_CTAS
#define RT_SORT_MIN_CTAS 4
#endif
constexpr int SORT_ITEMS = RT_SORT_ITEMS;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 3072 pairs
constexpr int MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u << 30, FLAG_PREFIX = 2u << 30, VALUE_MASK = (1u << 30) - 1u;
constexpr uint32_t SPIN_LIMIT = 1u << 24;
#ifndef RT_LOOKBACK_WINDOW
#define RT_LOOKBACK_WINDOW 4
#endif
constexpr int LOOKBACK_WINDOW = RT_LOOKBACK_WINDOW;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Peer mask of the lanes holding the same 8-bit digit: MATCH.ANY, or 8 ballots. MATCH.ANY costs more the more DISTINCT digits the 32 lanes
// hold (measured on B200, profiles/README.md r01l / r2_zd: on a surface mesh in grid order, few distinct digits per warp, it beats the ballots,
// 0.57 vs 0.63 ms for the five passes of inst10m; on the uniformly random keys of the triangle soup, ~30 distinct digits per warp, it loses,
// 0.85 vs 0.74 ms), the ballots cost the same whatever the data. Every warp therefore ranks its FIRST item with ballots, counts the distinct
// digits it saw and picks the form for its other items (warp-uniform branch). RT_SORT_USE_BALLOT forces the ballots, RT_SORT_BALLOT_DISTINCT=33
// forces MATCH.ANY.
#ifndef RT_SORT_BALLOT_DISTINCT
#define RT_SORT_BALLOT_DISTINCT 24
#endif
__device__ __forceinline__ uint32_t match_digit_ballot(uint32_t d, bool valid) {
    uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t v = __ballot_sync(0xffffffffu, bit);
        m &= bit ? v : ~v;
    }
    return m;
}
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool valid, bool use_ballot) {
#if defined(RT_SORT_USE_BALLOT)
    (void)use_ballot;
    return match_digit_ballot(d, valid);
#else
    if (use_ballot) return match_digit_ballot(d, valid);
    return __match_any_sync(0xffffffffu, valid ? (d & 255u) : 0xFFFFFFFFu);
#endif
}

// hist[p][d] += number of keys whose p-th 8-bit digit (counted from bit base_shift) is d
__global__ void __launch_bounds__(512) k_sort_hist(const uint64_t* __restrict__ keys, uint32_t n, int passes, int base_shift, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[MAX_PASSES * RADIX];
    for (int j = threadIdx.x; j < passes * RADIX; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = keys[i] >> base_shift;
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * RADIX + (uint32_t)((k >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < passes * RADIX; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

// in-place exclusive scan of each pass's 256-bin histogram; one block of 256 threads per pass
__global__ void __launch_bounds__(RADIX) k_sort_scan_hist(uint32_t* hist) {
    __shared__ uint32_t wtot[SORT_WARPS];
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    uint32_t* h = hist + blockIdx.x * RADIX;
    uint32_t v = h[d], inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += wtot[w];
    h[d] = base + inc - v;
}

// HAS_VALS = false: the primitive id rides in the low bits of the 64-bit word (below `shift`), nothing else is moved
template <bool HAS_VALS>
__global__ void __launch_bounds__(SORT_THREADS, RT_SORT_MIN_CTAS) k_onesweep_pass(
    const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
    const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
    uint32_t n, int shift, const uint32_t* __restrict__ digit_base,
    uint32_t* tile_state, uint32_t* tile_counter, int* error_flag) {
    __shared__ uint32_t s_cnt[SORT_WARPS * RADIX];   // per-warp digit counts -> per-warp exclusive offsets
    __shared__ uint64_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[HAS_VALS ? SORT_TILE : 1];
    __shared__ uint32_t s_loff[RADIX];               // start of digit d inside the tile-local order
    __shared__ uint32_t s_goff[RADIX];               // global position of local slot p with digit d = s_goff[d] + p
    __shared__ uint32_t s_wtot[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int j = tid; j < SORT_WARPS * RADIX; j += SORT_THREADS) s_cnt[j] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * (uint32_t)SORT_TILE;
    const uint32_t tile_n = min((uint32_t)SORT_TILE, n - base);
    const uint32_t wbase = warp * (32 * SORT_ITEMS) + lane;     // tile-local index of item 0 of this lane

    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        uint32_t li = wbase + i * 32;
        key[i] = li < tile_n ? keys_in[base + li] : ~0ull;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    // Ranking in three software-pipelined sweeps (no dependent chain through shared memory between items):
    // (1) the peer mask of every item (lanes holding the same digit), (2) one shared-memory atomic per distinct digit
    // and item by the lowest peer lane — a warp's atomics are issued in program order, so item i is counted before
    // item i + 1 and the ranking is stable —, (3) broadcast of the old counter value to the peers.
    constexpr int RANK_BATCH = SORT_ITEMS % 6 == 0 ? 6 : 4;     // items in flight per sweep (bounds the live registers)
    static_assert(SORT_ITEMS % RANK_BATCH == 0, "SORT_ITEMS");
    bool use_ballot = true;                                     // item 0: ballots; they also tell how many distinct digits the warp holds
#pragma unroll
    for (int i0 = 0; i0 < SORT_ITEMS; i0 += RANK_BATCH) {
        uint32_t prev[RANK_BATCH];
#pragma unroll
        for (int k = 0; k < RANK_BATCH; ++k) {
            const int i = i0 + k;
            const bool valid = (wbase + i * 32) < tile_n;
            rank[i] = match_digit((uint32_t)(key[i] >> shift), valid, use_ballot);
            if (i == 0) {
                const uint32_t m0 = rank[0];
                use_ballot = __popc(__ballot_sync(0xffffffffu, valid && lane == __ffs(m0) - 1)) >= RT_SORT_BALLOT_DISTINCT;
            }
        }
#pragma unroll
        for (int k = 0; k < RANK_BATCH; ++k) {
            const int i = i0 + k;
            const bool valid = (wbase + i * 32) < tile_n;
            const uint32_t m = rank[i];
            prev[k] = 0;
            if (valid && lane == __ffs(m) - 1) prev[k] = atomicAdd(&s_cnt[warp * RADIX + (uint32_t)((key[i] >> shift) & 255u)], (uint32_t)__popc(m));
        }
#pragma unroll
        for (int k = 0; k < RANK_BATCH; ++k) {
            const int i = i0 + k;
            const uint32_t m = rank[i];
            rank[i] = __shfl_sync(0xffffffffu, prev[k], __ffs(m) - 1) + __popc(m & lt_mask);
        }
    }
    __syncthreads();

    // thread d: exclusive scan over warps for digit d, tile count, publish, look back
    {
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) { uint32_t c = s_cnt[w * RADIX + d]; s_cnt[w * RADIX + d] = run; run += c; }
        uint32_t* my_state = tile_state + (size_t)tile * RADIX + d;
        if (tile == 0) st_volatile_u32(my_state, FLAG_PREFIX | run);
        else st_volatile_u32(my_state, FLAG_AGG | run);

        // exclusive scan of the tile counts over digits -> local offsets
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; ++w) wb += s_wtot[w];
        const uint32_t loff = wb + inc - run;

        // Decoupled look-back, LOOKBACK_WINDOW predecessor tiles per step: the window's loads are independent (one L2
        // round trip per step instead of one per tile); they are consumed in order up to the first tile that has not
        // published yet, or the first inclusive prefix. Tile -1 is a virtual inclusive prefix of 0.
        uint32_t excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            uint32_t spins = 0;
            for (;;) {
                uint32_t sv[LOOKBACK_WINDOW];
#pragma unroll
                for (int k = 0; k < LOOKBACK_WINDOW; ++k)
                    sv[k] = t - k >= 0 ? ld_volatile_u32(tile_state + (size_t)(t - k) * RADIX + d) : FLAG_PREFIX;
                bool done = false;
                int adv = 0;
#pragma unroll
                for (int k = 0; k < LOOKBACK_WINDOW; ++k) {
                    const uint32_t flag = sv[k] >> 30;
                    if (!done && adv == k && flag != 0u) { excl += sv[k] & VALUE_MASK; adv = k + 1; done = flag == 2u; }
                }
                if (done) break;
                t -= adv;
                if (adv == 0) {
                    if (++spins > SPIN_LIMIT) { atomicExch(error_flag, 1); break; }
                    __nanosleep(20);
                }
            }
            st_volatile_u32(my_state, FLAG_PREFIX | ((excl + run) & VALUE_MASK));
        }
        s_loff[d] = loff;
        s_goff[d] = digit_base[d] + excl - loff;
    }
    __syncthreads();

    // scatter keys into tile-local sorted order
    uint32_t lpos[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const bool valid = (wbase + i * 32) < tile_n;
        if (valid) {
            const uint32_t d = (uint32_t)((key[i] >> shift) & 255u);
            lpos[i] = s_loff[d] + s_cnt[warp * RADIX + d] + rank[i];
            s_keys[lpos[i]] = key[i];
        } else lpos[i] = 0xFFFFFFFFu;
    }
    if (HAS_VALS) {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            uint32_t li = wbase + i * 32;
            if (li < tile_n) s_vals[lpos[i]] = vals_in[base + li];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t p = i * SORT_THREADS + tid;
        if (p < tile_n) {
            const uint64_t k = s_keys[p];
            const uint32_t d = (uint32_t)((k >> shift) & 255u);
            const uint32_t g = s_goff[d] + p;
            keys_out[g] = k;
            if (HAS_VALS) vals_out[g] = s_vals[p];
        }
    }
}

// ---- segmented sort (seg_sort.cuh): the stand-alone kernel ---------------------------------------------------------
__global__ void __launch_bounds__(SEG_THREADS, 1) k_seg_sort(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                            const BlasRecord* __restrict__ recs, int shift0, int key_bits, uint32_t single_n) {
    extern __shared__ __align__(16) unsigned char seg_smem[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(seg_smem);
    const int tid = threadIdx.x;
    const uint32_t first = recs ? recs[blockIdx.x].first : 0u, n = recs ? recs[blockIdx.x].tri_count : single_n;   // recs == nullptr: the whole input is one segment
    if (n == 0) return;
    // the segment: one TMA bulk copy (cp.async.bulk + mbarrier); padding (~0) sorts last and stays last
    const int items = n <= 1u * SEG_THREADS ? 1 : n <= 3u * SEG_THREADS ? 3 : n <= 5u * SEG_THREADS ? 5 : SEG_ITEMS;   // records per thread (block-uniform)
    for (uint32_t i = n + tid; i < (uint32_t)items * SEG_THREADS; i += SEG_THREADS) s_keys[i] = ~0ull;
    seg_load_bulk(s_keys, in + first, n, reinterpret_cast<uint64_t*>(seg_smem + SEG_SMEM_MBAR_OFFSET));
    __syncthreads();
    if (items == 1) seg_sort_passes<1>(seg_smem, shift0, key_bits, n);
    else if (items == 3) seg_sort_passes<3>(seg_smem, shift0, key_bits, n);
    else if (items == 5) seg_sort_passes<5>(seg_smem, shift0, key_bits, n);
    else seg_sort_passes<SEG_ITEMS>(seg_smem, shift0, key_bits, n);
    seg_store_bulk(out + first, s_keys, n);
}

}  // namespace

SortPlan sort_plan(uint32_t n, int key_bits) {
    SortPlan p;
    p.n = n;
    p.passes = (key_bits + 7) / 8;
    if (p.passes < 1) p.passes = 1;
    if (p.passes > MAX_PASSES) p.passes = MAX_PASSES;
    p.tiles = (n + SORT_TILE - 1) / SORT_TILE;
    p.scratch_bytes = (size_t)MAX_PASSES * RADIX * 4 + 64 + (size_t)p.passes * p.tiles * RADIX * 4;
    return p;
}

int sort_pairs(const SortPlan& plan, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
               void* scratch, int* device_error_flag, cudaStream_t stream, bool* result_in_b) {
    const bool has_vals = vals_a != nullptr;
    const int base_shift = has_vals ? 0 : plan.packed_val_bits;
    *result_in_b = false;
    if (plan.n == 0) return 0;
    if ((plan.seg_records || plan.seg_single) && !has_vals) {    // segmented path: one kernel, result in keys_b
        // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set it once per device
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            if (cudaFuncSetAttribute(k_seg_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEG_SMEM_BYTES) != cudaSuccess) return -1;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        k_seg_sort<<<plan.seg_single ? 1u : plan.n_segments, SEG_THREADS, SEG_SMEM_BYTES, stream>>>(keys_a, keys_b, plan.seg_single ? nullptr : plan.seg_records, base_shift,
                                                                                              plan.seg_key_bits, plan.n);
        *result_in_b = true;
        if (cudaGetLastError() != cudaSuccess) return -1;
        return 1;
    }
    uint32_t* hist = (uint32_t*)scratch;
    uint32_t* counters = hist + MAX_PASSES * RADIX;
    uint32_t* states = counters + 16;
    if (cudaMemsetAsync(scratch, 0, plan.scratch_bytes, stream) != cudaSuccess) return -1;
    int launches = 0;
    int hist_blocks = (int)((plan.n + 512 * 16 - 1) / (512 * 16));
    if (hist_blocks > 148 * 4) hist_blocks = 148 * 4;
    if (hist_blocks < 1) hist_blocks = 1;
    k_sort_hist<<<hist_blocks, 512, 0, stream>>>(keys_a, plan.n, plan.passes, base_shift, hist);
    k_sort_scan_hist<<<plan.passes, RADIX, 0, stream>>>(hist);
    launches += 2;
    uint64_t* kin = keys_a; uint64_t* kout = keys_b;
    uint32_t* vin = vals_a; uint32_t* vout = vals_b;
    for (int p = 0; p < plan.passes; ++p) {
        if (has_vals)
            k_onesweep_pass<true><<<plan.tiles, SORT_THREADS, 0, stream>>>(kin, kout, vin, vout, plan.n, 8 * p, hist + p * RADIX,
                                                                           states + (size_t)p * plan.tiles * RADIX, counters + p,
                                                                           device_error_flag);
        else
            k_onesweep_pass<false><<<plan.tiles, SORT_THREADS, 0, stream>>>(kin, kout, nullptr, nullptr, plan.n, base_shift + 8 * p, hist + p * RADIX,
                                                                            states + (size_t)p * plan.tiles * RADIX, counters + p,
                                                                            device_error_flag);
        ++launches;
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *result_in_b = (plan.passes & 1) != 0;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace rt
