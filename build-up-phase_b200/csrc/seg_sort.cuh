// seg_sort.cuh — one CTA sorts one whole segment (a BLAS of a batch) in shared memory; shared by the stand-alone
// k_seg_sort (radix_sort.cu) and the fused setup + Morton + sort kernel of the batched build (lbvh_build.cu).
#pragma once
#include "rt_internal.h"

namespace rt {

// The build input is already grouped by BLAS, so sorting every segment by its Morton bits gives exactly what the global
// sort by (BLAS id, Morton) gives — with one global read and one global write of the records instead of one per 8-bit
// pass, no look-back chain between tiles and no scattered global stores.
// LSD passes of FOUR bits with a match-free ranking: a thread owns SEG_ITEMS consecutive records (blocked order), counts
// its 16 digits in one 64-bit register of 4-bit fields (which also yields each record's rank among the thread's own
// records), and a block-wide exclusive scan of the per-thread counts (4 registers of 4 x 16-bit fields: warp shuffles,
// then one warp over the 32 warp totals) gives every (thread, digit) its start in the sorted order. Stable by
// construction. (The first version ranked 8-bit digits with MATCH.ANY like the onesweep pass: 0.33 ms for 10 M records,
// bound by the rate of MATCH.ANY itself, ~60 cycles per warp instruction and SM; profiles/README.md r01x/r01y.)
constexpr int SEG_THREADS = 1024, SEG_WARPS = SEG_THREADS / 32, SEG_ITEMS = 11;     // 11: odd stride -> blocked shared-memory access without bank pile-ups
static_assert(SEG_THREADS * SEG_ITEMS == (int)SEG_SORT_CAPACITY, "segment capacity");
constexpr size_t SEG_SMEM_MBAR_OFFSET = sizeof(uint64_t) * SEG_SORT_CAPACITY + sizeof(uint16_t) * 16 * SEG_THREADS + sizeof(uint64_t) * (4 * SEG_WARPS + 4);   // 8-byte aligned
constexpr size_t SEG_SMEM_BYTES = SEG_SMEM_MBAR_OFFSET + 64;

__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int o) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, o), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    return ((uint64_t)hi << 32) | lo;
}
// digits 4q .. 4q+3 of a register of 4-bit fields -> 16-bit fields
__device__ __forceinline__ uint64_t expand4(uint64_t cnt, int q) {
    const uint32_t f = (uint32_t)(cnt >> (16 * q)) & 0xFFFFu;
    return (uint64_t)(f & 15u) | ((uint64_t)((f >> 4) & 15u) << 16) | ((uint64_t)((f >> 8) & 15u) << 32) | ((uint64_t)((f >> 12) & 15u) << 48);
}

// Sorts the SEG_SORT_CAPACITY records in s_keys (padding = ~0 sorts last) by bits [shift0, shift0 + key_bits). All threads of
// the 1024-thread CTA call it after a __syncthreads(); the records are sorted and visible to all threads on return.
//
// n_real <= SEG_RANK_SORT_MAX (the sample scene's 4 triangles, small TLASes): a RANK sort instead of the 8 counting passes over all 11,264
// padded slots - thread i counts the records that sort before its own (key bits, then position: stable, hence the very same order the
// passes produce) with broadcast shared-memory reads, ~8 instructions per comparison. Measured on B200 (rt_update_tlas, 1024 instances,
// profiles/README.md r2_j): rank sort 0.123 ms vs 0.074 ms with the passes - n^2 comparisons on ONE SM lose beyond a few hundred records.
constexpr uint32_t SEG_RANK_SORT_MAX = 256;
// ITEMS: records per thread (odd: conflict-free blocked access); the passes cover the first SEG_THREADS * ITEMS slots, which must hold every
// real record and be padded. The stand-alone kernel picks the smallest that fits (a 1024-instance TLAS sorts 1 record per thread, not 11).
template <int ITEMS = SEG_ITEMS>
__device__ __forceinline__ void seg_sort_passes(unsigned char* seg_smem, int shift0, int key_bits, uint32_t n_real = SEG_SORT_CAPACITY) {
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(seg_smem);                         // the segment, sorted by the passes so far
    if (n_real <= SEG_RANK_SORT_MAX) {
        const uint32_t i = threadIdx.x;
        const uint64_t mask = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
        uint64_t mine = 0, mk = 0;
        uint32_t rank = 0;
        if (i < n_real) {
            mine = s_keys[i]; mk = (mine >> shift0) & mask;
            for (uint32_t j = 0; j < n_real; ++j) {
                const uint64_t kj = (s_keys[j] >> shift0) & mask;
                rank += (kj < mk || (kj == mk && j < i)) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (i < n_real) s_keys[rank] = mine;
        __syncthreads();
        return;
    }
    uint64_t* s_wsum = s_keys + SEG_SORT_CAPACITY;                                     // [4][SEG_WARPS] packed warp totals -> exclusive warp bases
    uint64_t* s_tot = s_wsum + 4 * SEG_WARPS;                                          // [4] packed digit totals
    uint16_t* s_off = reinterpret_cast<uint16_t*>(s_tot + 4);                          // [16][SEG_THREADS] start of (digit, thread)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t key[ITEMS];
    for (int shift = shift0; shift < shift0 + key_bits; shift += 4) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = s_keys[tid * ITEMS + i];
        uint64_t cnt = 0, rk = 0;                                                      // 16 x 4-bit digit counts; 4-bit rank of record i among this thread's records of its digit
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int d4 = 4 * ((int)(key[i] >> shift) & 15);
            rk |= ((cnt >> d4) & 15ull) << (4 * i);
            cnt += 1ull << d4;
        }
        uint64_t ex[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint64_t own = expand4(cnt, q);
            uint64_t inc = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint64_t t = shfl_up_u64(inc, o); if (lane >= o) inc += t; }
            if (lane == 31) s_wsum[q * SEG_WARPS + warp] = inc;
            ex[q] = inc - own;                                                         // exclusive within the warp
        }
        __syncthreads();                                                               // also: every thread has read its records
        if (warp < 4) {                                                                // warp q: exclusive scan of the 32 warp totals of digits 4q .. 4q+3
            const uint64_t own = s_wsum[warp * SEG_WARPS + lane];
            uint64_t inc = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint64_t t = shfl_up_u64(inc, o); if (lane >= o) inc += t; }
            s_wsum[warp * SEG_WARPS + lane] = inc - own;
            if (lane == 31) s_tot[warp] = inc;
        }
        __syncthreads();
        {
            uint32_t dbase = 0;                                                        // start of the digit in the sorted order
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint64_t tot = s_tot[q], e = ex[q] + s_wsum[q * SEG_WARPS + warp];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s_off[(4 * q + k) * SEG_THREADS + tid] = (uint16_t)(dbase + (uint32_t)((e >> (16 * k)) & 0xFFFFu));
                    dbase += (uint32_t)((tot >> (16 * k)) & 0xFFFFu);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int d = (int)(key[i] >> shift) & 15;
            s_keys[(uint32_t)s_off[d * SEG_THREADS + tid] + (uint32_t)((rk >> (4 * i)) & 15ull)] = key[i];
        }
        __syncthreads();
    }
}


// ---- the LIGHT segment sort: 32-bit Morton key + 16-bit local index, 512 threads x 22 records ------------------------------------------------
// Same LSD passes of four bits, same stable order, but the records are {u32 Morton, u16 position in the segment} (6 B instead of 8) and a CTA
// is 512 threads with 22 records each: 85 KB of shared memory and <= 64 registers, so TWO CTAs share an SM and one BLAS's gather phase runs
// beside another BLAS's sort passes (one 1024-thread CTA per SM exposes every memory phase between its barriers: profiles/README.md r2_zg).
// A thread counts its 22 records in two halves of 11 (4-bit fields hold at most 15).
constexpr int SEG2_THREADS = 512, SEG2_WARPS = SEG2_THREADS / 32, SEG2_ITEMS = 22, SEG2_HALF = 11;
static_assert(SEG2_THREADS * SEG2_ITEMS == (int)SEG_SORT_CAPACITY && SEG2_ITEMS == 2 * SEG2_HALF, "segment capacity");
// s_m is stored with a row pitch of 23 words per thread (record p lives at word p + p / 22): a thread's 22 consecutive records then start at an odd
// stride from its neighbour's and the blocked loads are bank-conflict free (a pitch of 22 is a two-way conflict on every load)
constexpr uint32_t SEG2_M_WORDS = SEG2_THREADS * (SEG2_ITEMS + 1);
__device__ __forceinline__ uint32_t seg2_m_at(uint32_t p) { return p + __umulhi(p, 195225787u);  }   // p / 22 for p < 2^16: ceil(2^32 / 22) = 195225787
constexpr size_t SEG2_OFF_M = 0, SEG2_OFF_ID = SEG2_OFF_M + sizeof(uint32_t) * SEG2_M_WORDS, SEG2_OFF_WSUM = SEG2_OFF_ID + sizeof(uint16_t) * SEG_SORT_CAPACITY,
                 SEG2_OFF_TOT = SEG2_OFF_WSUM + sizeof(uint64_t) * 4 * SEG2_WARPS, SEG2_OFF_OFF = SEG2_OFF_TOT + sizeof(uint64_t) * 4,
                 SEG2_SMEM_BYTES = SEG2_OFF_OFF + sizeof(uint16_t) * 16 * SEG2_THREADS;
static_assert(SEG2_OFF_ID % 8 == 0 && SEG2_OFF_WSUM % 8 == 0 && SEG2_OFF_OFF % 8 == 0, "alignment");

// Sorts the SEG_SORT_CAPACITY records {s_m[i], s_id[i]} (padding: s_m = ~0 sorts last and stays last) by the low 30 bits of s_m, stable.
// All 512 threads call it after a __syncthreads(); sorted and visible to all threads on return.
__device__ __forceinline__ void seg2_sort_passes(unsigned char* smem, uint32_t n_real) {
    uint32_t* s_m = reinterpret_cast<uint32_t*>(smem + SEG2_OFF_M);
    uint16_t* s_id = reinterpret_cast<uint16_t*>(smem + SEG2_OFF_ID);
    if (n_real <= SEG_RANK_SORT_MAX) {                                                 // tiny segments: rank sort (see seg_sort_passes)
        const uint32_t i = threadIdx.x;
        uint32_t mine = 0, rank = 0;
        uint16_t mid = 0;
        if (i < n_real) {
            mine = s_m[seg2_m_at(i)]; mid = s_id[i];
            for (uint32_t j = 0; j < n_real; ++j) {
                const uint32_t mj = s_m[seg2_m_at(j)];
                rank += (mj < mine || (mj == mine && j < i)) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (i < n_real) { s_m[seg2_m_at(rank)] = mine; s_id[rank] = mid; }
        __syncthreads();
        return;
    }
    uint64_t* s_wsum = reinterpret_cast<uint64_t*>(smem + SEG2_OFF_WSUM);             // [4][SEG2_WARPS] packed warp totals -> exclusive warp bases
    uint64_t* s_tot = reinterpret_cast<uint64_t*>(smem + SEG2_OFF_TOT);               // [4] packed digit totals
    uint16_t* s_off = reinterpret_cast<uint16_t*>(smem + SEG2_OFF_OFF);               // [16][SEG2_THREADS] start of (digit, thread)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t m[SEG2_ITEMS], idp[SEG2_HALF];                                             // idp[k] = ids of records 2k (low half) and 2k + 1 (high half)
    for (int shift = 0; shift < 32; shift += 4) {
#pragma unroll
        for (int i = 0; i < SEG2_ITEMS; ++i) m[i] = s_m[tid * (SEG2_ITEMS + 1) + i];
#pragma unroll
        for (int k = 0; k < SEG2_HALF; ++k) idp[k] = reinterpret_cast<const uint32_t*>(s_id)[tid * SEG2_HALF + k];
        uint64_t cnt_a = 0, cnt_b = 0, rk_a = 0, rk_b = 0;                             // per half: 16 x 4-bit digit counts; 4-bit rank of record i within its half
#pragma unroll
        for (int i = 0; i < SEG2_HALF; ++i) {
            const int d4 = 4 * (int)((m[i] >> shift) & 15u);
            rk_a |= ((cnt_a >> d4) & 15ull) << (4 * i);
            cnt_a += 1ull << d4;
        }
#pragma unroll
        for (int i = 0; i < SEG2_HALF; ++i) {
            const int d4 = 4 * (int)((m[SEG2_HALF + i] >> shift) & 15u);
            rk_b |= ((cnt_b >> d4) & 15ull) << (4 * i);
            cnt_b += 1ull << d4;
        }
        uint64_t ex[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint64_t own = expand4(cnt_a, q) + expand4(cnt_b, q);                // <= 22 per 16-bit field
            uint64_t inc = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint64_t t = shfl_up_u64(inc, o); if (lane >= o) inc += t; }
            if (lane == 31) s_wsum[q * SEG2_WARPS + warp] = inc;
            ex[q] = inc - own;                                                         // exclusive within the warp
        }
        __syncthreads();                                                               // also: every thread has read its records
        if (warp < 4) {                                                                // warp q: exclusive scan of the 16 warp totals of digits 4q .. 4q+3
            const uint64_t own = lane < SEG2_WARPS ? s_wsum[warp * SEG2_WARPS + lane] : 0ull;
            uint64_t inc = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint64_t t = shfl_up_u64(inc, o); if (lane >= o) inc += t; }
            if (lane < SEG2_WARPS) s_wsum[warp * SEG2_WARPS + lane] = inc - own;
            if (lane == 31) s_tot[warp] = inc;
        }
        __syncthreads();
        {
            uint32_t dbase = 0;                                                        // start of the digit in the sorted order
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint64_t tot = s_tot[q], e = ex[q] + s_wsum[q * SEG2_WARPS + warp];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s_off[(4 * q + k) * SEG2_THREADS + tid] = (uint16_t)(dbase + (uint32_t)((e >> (16 * k)) & 0xFFFFu));
                    dbase += (uint32_t)((tot >> (16 * k)) & 0xFFFFu);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < SEG2_ITEMS; ++i) {
            const int d = (int)((m[i] >> shift) & 15u);
            // rank among this thread's records of digit d: first half by its own rank, second half behind all of the first half's
            const uint32_t r = i < SEG2_HALF ? (uint32_t)((rk_a >> (4 * i)) & 15ull)
                                             : (uint32_t)((cnt_a >> (4 * d)) & 15ull) + (uint32_t)((rk_b >> (4 * (i - SEG2_HALF))) & 15ull);
            const uint32_t pos = (uint32_t)s_off[d * SEG2_THREADS + tid] + r;
            s_m[seg2_m_at(pos)] = m[i];
            s_id[pos] = (uint16_t)((i & 1) ? (idp[i >> 1] >> 16) : (idp[i >> 1] & 0xFFFFu));
        }
        __syncthreads();
    }
}

// ---- TMA bulk copies of a whole sorted segment (sm_100a: cp.async.bulk, SASS UBLKCP) ----------------------------------------------------
// The sorted records of a segment are ONE contiguous run of n x 8 bytes in shared and in global memory, so one elected thread moves them
// with a single bulk copy instead of every thread looping over 8-byte stores. Needs 16-byte alignment on both sides: the even part of an
// evenly placed segment goes through the bulk copy, an odd head / tail record through plain stores.
__device__ __forceinline__ void seg_store_bulk(uint64_t* __restrict__ gdst, const uint64_t* s_src, uint32_t n) {
    const bool aligned = (reinterpret_cast<uintptr_t>(gdst) & 15u) == 0;
    const uint32_t n_bulk = aligned ? (n & ~1u) : 0u;
    if (n_bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // this thread's generic-proxy writes to the segment -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(s_src)), "r"(n_bulk * 8u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    for (uint32_t i = n_bulk + threadIdx.x; i < n; i += blockDim.x) gdst[i] = s_src[i];
    if (n_bulk && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // the copy has left shared memory AND landed before the CTA retires
}
// global -> shared: one bulk copy completing on an mbarrier (8 bytes at `bar`, 8-byte aligned shared memory); every thread waits on phase 0
__device__ __forceinline__ void seg_load_bulk(uint64_t* s_dst, const uint64_t* __restrict__ gsrc, uint32_t n, uint64_t* bar) {
    const bool aligned = (reinterpret_cast<uintptr_t>(gsrc) & 15u) == 0;
    const uint32_t n_bulk = aligned ? (n & ~1u) : 0u;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    if (n_bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n_bulk * 8u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(s_dst)), "l"(gsrc), "r"(n_bulk * 8u), "r"(b) : "memory");
        }
    }
    for (uint32_t i = n_bulk + threadIdx.x; i < n; i += blockDim.x) s_dst[i] = gsrc[i];
    if (n_bulk) {
        __syncthreads();                                                         // the barrier is initialised
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
    }
}

}  // namespace rt
