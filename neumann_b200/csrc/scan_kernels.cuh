// scan_kernels.cuh — sm_100a kernels of the SIMILAR brute-force scan.
//
// What the kernels compute is the reference's per-row score followed by "sort all, keep k":
//   simd::dot_product / sum_of_squares  tensor_store/src/hnsw.rs:168-222  (8-lane f32x8 tree)
//   cosine_similarity                   vector_engine/src/lib.rs:2257-2266
//   euclidean_distance (scalar fold)    vector_engine/src/lib.rs:2249-2253
//   compute_score                       vector_engine/src/lib.rs:2231-2246
//   sort_by(desc) + truncate(k)         vector_engine/src/lib.rs:2027-2034
// The arithmetic reproduces the reference's summation ORDER so scores are bit-identical:
// every multiply and add is a separate IEEE-754 round-to-nearest op (__fmul_rn/__fadd_rn are
// never contracted into FMA), lane j of the f32x8 accumulator owns elements i = j (mod 8) in
// ascending i, lanes are folded 0..7 left to right from 0.0, the dim%8 tail is added after.
//
// Data movement: the corpus is row-major f32 [rows, pitch] in HBM.  A persistent CTA per SM
// walks row blocks of 256 rows; one producer thread streams [256 rows x 32 floats] boxes
// (128 B per row, SWIZZLE_128B) with TMA into a ring of shared-memory stages guarded by
// full/empty mbarriers; 256 consumer threads own one row each and keep that row's lane
// accumulators in registers across the dim/32 boxes of the block.  With the 128 B swizzle the
// eight 16-byte units of a row are XOR-permuted by (row & 7), so the LDS.128 of a quarter
// warp (8 consecutive rows, same logical unit) hits 8 distinct bank groups: conflict-free.
//
// Selection: scores become 64-bit keys (order-preserving score bits | inverted row), so
// "better" is plain unsigned >.  Each CTA keeps a candidate buffer with a running k-th-best
// threshold, pruned by an in-smem bitonic sort when full; CTAs publish their k best and the
// last CTA to finish merges all lists and writes the result (no second launch).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_types.hpp"

namespace nm {


// ---------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------
// [63:32] order-preserving score (NaN -> 0 = worst, -0.0 folded onto +0.0)
// [31:1]  0x7fffffff - local_row  (lower row wins ties)
// [0]     1 iff the score bits were -0.0 (restored on decode; never decides an order because
//         rows are unique)
__host__ __device__ __forceinline__ uint32_t score_to_ord(uint32_t u) {
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0u;
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t ord_to_score_bits(uint32_t o) {
    if (o == 0u) return 0x7fc00000u;  // canonical NaN
    return (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
}
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t score_bits, uint32_t local_row) {
    uint32_t negzero = (score_bits == 0x80000000u) ? 1u : 0u;
    return ((uint64_t)score_to_ord(score_bits) << 32) |
           ((uint64_t)(0x7fffffffu - local_row) << 1) | negzero;
}
__host__ __device__ __forceinline__ uint32_t key_local_row(uint64_t key) {
    return 0x7fffffffu - (uint32_t)((key >> 1) & 0x7fffffffu);
}
__host__ __device__ __forceinline__ uint32_t key_score_bits(uint64_t key) {
    if (key & 1ull) return 0x80000000u;
    return ord_to_score_bits((uint32_t)(key >> 32));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// PTX helpers (mbarrier, TMA, named barriers)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
// 2D tiled TMA load global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int32_t x,
                                            int32_t y, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Named barrier 1 = the 256 consumer threads (the producer warp never joins it).
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kRowsPerBlock) : "memory");
}
// Barrier + population count of `pred` over the 256 consumer threads; uniform result.
__device__ __forceinline__ uint32_t consumer_sync_popc(bool pred) {
    uint32_t total;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.popc.u32 %0, 1, %2, p;\n\t"
        "}\n"
        : "=r"(total)
        : "r"((uint32_t)pred), "n"(kRowsPerBlock)
        : "memory");
    return total;
}

// Programmatic dependent launch (PDL): consecutive scans on one stream overlap — the next
// query's CTAs start streaming on SMs this query has already left while its last CTA is still
// merging / exchanging.
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grids() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------
// candidate buffer (shared memory, consumer threads only)
// ---------------------------------------------------------------------------------------
struct TopKState {
    uint64_t *buf;       // kCandCap keys
    uint32_t *cnt_smem;  // append cursor
    uint64_t *thr_smem;  // k-th best key after the last prune (0 = none yet)
    uint32_t count;      // uniform copy of *cnt_smem
    uint32_t k;
    uint32_t cap;        // entries of buf in use: small k prunes a smaller buffer (cheaper sort)
};

// Bitonic sort, descending, of buf[0..n) (n = power of two) by the 256 consumer threads.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t *buf, uint32_t n, uint32_t t) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = t; i < (n >> 1); i += kRowsPerBlock) {
                uint32_t lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = buf[lo], b = buf[hi];
                bool swap = desc ? (a < b) : (a > b);
                if (swap) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
            consumer_sync();
        }
    }
}

// Bitonic sort, descending, of buf[0..64) by ONE warp in registers (two keys per lane): no
// block barriers, ~20x cheaper than the block-wide network for the small buffers that the
// eager threshold refresh of the pre-filter scan produces.
__device__ __forceinline__ void warp_sort_desc_64(uint64_t *buf, uint32_t lane) {
    uint64_t a = buf[lane], b = buf[lane + 32u];  // element indices: lane, lane + 32
    // network over 64 elements; element e lives in (e < 32 ? a : b) of lane e & 31
#pragma unroll
    for (uint32_t size = 2; size <= 64u; size <<= 1) {
#pragma unroll
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32u) {
                // partners are (lane, lane+32): both in this lane; size == 64 -> descending
                if (a < b) {
                    uint64_t tmp = a;
                    a = b;
                    b = tmp;
                }
            } else {
                const uint64_t pa = __shfl_xor_sync(0xffffffffu, a, stride);
                const uint64_t pb = __shfl_xor_sync(0xffffffffu, b, stride);
                const bool upper = (lane & stride) != 0;  // this lane holds the higher index
                // direction of the block of `size` elements containing the element
                const bool desc_a = ((lane & size) == 0);            // element index = lane
                const bool desc_b = (((lane + 32u) & size) == 0);    // element index = lane + 32
                // lower index keeps the larger key when descending
                const bool take_max_a = (desc_a != upper);
                const bool take_max_b = (desc_b != upper);
                a = take_max_a ? (a > pa ? a : pa) : (a < pa ? a : pa);
                b = take_max_b ? (b > pb ? b : pb) : (b < pb ? b : pb);
            }
        }
    }
    buf[lane] = a;
    buf[lane + 32u] = b;
}

// Sort the buffer, keep the k best, refresh the threshold.  Called by all consumer threads
// with uniform state; contains barriers.
__device__ __forceinline__ void topk_prune(TopKState &st, uint32_t t) {
    uint32_t n = 2;
    while (n < st.count) n <<= 1;
    if (n <= 64u) {
        n = 64u;
        for (uint32_t i = st.count + t; i < n; i += kRowsPerBlock) st.buf[i] = 0ull;
        consumer_sync();
        if (t < 32u) warp_sort_desc_64(st.buf, t);
    } else {
        for (uint32_t i = st.count + t; i < n; i += kRowsPerBlock) st.buf[i] = 0ull;
        consumer_sync();
        bitonic_sort_desc(st.buf, n, t);
    }
    consumer_sync();
    uint32_t kept = st.count < st.k ? st.count : st.k;
    if (t == 0) {
        *st.cnt_smem = kept;
        *st.thr_smem = (st.count >= st.k) ? st.buf[st.k - 1] : 0ull;
    }
    st.count = kept;
    consumer_sync();
}

// Offer one key per consumer thread (key == 0 -> nothing to offer).
__device__ __forceinline__ void topk_offer(TopKState &st, uint64_t key, uint32_t t) {
    bool cand = key > *st.thr_smem;
    uint32_t total = consumer_sync_popc(cand);
    if (st.count + total > st.cap) {
        topk_prune(st, t);
        cand = key > *st.thr_smem;
        total = consumer_sync_popc(cand);
    }
    uint32_t ballot = __ballot_sync(0xffffffffu, cand);
    if (ballot) {
        uint32_t lane = t & 31;
        uint32_t leader = __ffs(ballot) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(st.cnt_smem, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (cand) st.buf[base + __popc(ballot & ((1u << lane) - 1u))] = key;
    }
    st.count += total;
}


// ---------------------------------------------------------------------------------------
// final merge of the per-CTA lists (last CTA only)
// ---------------------------------------------------------------------------------------
struct MergeScratch {
    uint32_t *hist;  // 256 bins
    uint32_t *sc;    // [0..1] prefix (lo,hi) [2] need [3] exact flag [4] cursor [8..16) warp totals
};

// Select the k best of `total` published keys (0 = empty slot) into st.buf[0..count), sorted
// descending.  Small inputs are sorted directly; larger ones go through an MSB-first radix
// select (8-bit digits, early exit as soon as a digit bin is wholly selected), so the cost is
// a few passes over the L2-resident lists instead of hundreds of buffer prunes.  Loads are
// issued kMergeTile at a time per thread so the passes are bandwidth-, not latency-bound.
constexpr int kMergeTile = 8;
__device__ __forceinline__ void merge_published(TopKState &st, uint32_t t, const uint64_t *all,
                                                uint32_t total, uint32_t k,
                                                const MergeScratch ms) {
    const uint32_t lane = t & 31u, warp = t >> 5;
    if (total <= 512u) {
        for (uint32_t i = t; i < total; i += kRowsPerBlock) st.buf[i] = __ldcg(all + i);
        st.count = total;
        consumer_sync();
        topk_prune(st, t);  // zeros sort to the end
        const uint32_t kept = st.count;
        uint32_t valid = 0;
        for (uint32_t base = 0; base < kept; base += kRowsPerBlock) {
            const uint32_t i = base + t;
            valid += consumer_sync_popc(i < kept && st.buf[i] != 0ull);
        }
        st.count = valid;
        return;
    }
    if (t < 8) ms.sc[t] = 0u;
    uint64_t prefix = 0ull;
    uint32_t need = 0, kk = 0;
    int shift = 56;
    for (;; shift -= 8) {
        ms.hist[t] = 0u;
        consumer_sync();
        for (uint32_t base = 0; base < total; base += kRowsPerBlock * kMergeTile) {
            uint64_t keys[kMergeTile];
#pragma unroll
            for (int j = 0; j < kMergeTile; ++j) {
                const uint32_t i = base + j * kRowsPerBlock + t;
                keys[j] = (i < total) ? __ldcg(all + i) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < kMergeTile; ++j) {
                const uint64_t key = keys[j];
                const bool in = key != 0ull && (shift == 56 || (key >> (shift + 8)) == prefix);
                const uint32_t bin = in ? (uint32_t)(key >> shift) & 255u : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                if (in && lane == (uint32_t)(__ffs(peers) - 1))
                    atomicAdd(&ms.hist[bin], (uint32_t)__popc(peers));
            }
        }
        consumer_sync();
        // thread t owns bin t; suffix sums by warp scan: s = bins t..(end of this warp)
        const uint32_t mine = ms.hist[t];
        uint32_t sfx = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, sfx, off);
            if (lane + off < 32u) sfx += v;
        }
        if (lane == 0) ms.sc[8 + warp] = sfx;  // warp totals
        consumer_sync();
        uint32_t above = sfx - mine;
        for (uint32_t w = warp + 1; w < (uint32_t)kConsumerWarps; ++w) above += ms.sc[8 + w];
        if (shift == 56) {
            uint32_t valid = 0;
            for (uint32_t w = 0; w < (uint32_t)kConsumerWarps; ++w) valid += ms.sc[8 + w];
            kk = min(k, valid);
            need = kk;
        }
        if (kk == 0) {
            st.count = 0;
            consumer_sync();
            return;
        }
        if (above < need && need <= above + mine) {
            const uint64_t np = (prefix << 8) | (uint64_t)t;
            ms.sc[0] = (uint32_t)np;
            ms.sc[1] = (uint32_t)(np >> 32);
            ms.sc[2] = need - above;
            ms.sc[3] = (mine == need - above) ? 1u : 0u;
        }
        consumer_sync();
        prefix = ((uint64_t)ms.sc[1] << 32) | ms.sc[0];
        need = ms.sc[2];
        const bool exact = ms.sc[3] != 0u;
        if (exact || shift == 0) break;
    }
    // everything with (key >> shift) >= prefix is selected: exactly kk keys
    for (uint32_t base = 0; base < total; base += kRowsPerBlock * kMergeTile) {
        uint64_t keys[kMergeTile];
#pragma unroll
        for (int j = 0; j < kMergeTile; ++j) {
            const uint32_t i = base + j * kRowsPerBlock + t;
            keys[j] = (i < total) ? __ldcg(all + i) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kMergeTile; ++j) {
            const uint64_t key = keys[j];
            const bool sel = key != 0ull && (key >> shift) >= prefix;
            const uint32_t ballot = __ballot_sync(0xffffffffu, sel);
            if (ballot) {
                const uint32_t leader = __ffs(ballot) - 1;
                uint32_t pos = 0;
                if (lane == leader) pos = atomicAdd(&ms.sc[4], (uint32_t)__popc(ballot));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (sel) st.buf[pos + __popc(ballot & ((1u << lane) - 1u))] = key;
            }
        }
    }
    consumer_sync();
    st.count = kk;
    if (t == 0) *st.cnt_smem = kk;
    consumer_sync();
    topk_prune(st, t);
}

// Decode the merged keys in st.buf[0..count) into the caller's output buffers.
struct TopKOutputs {
    uint64_t *out_keys;   // [k] sorted local keys, 0 padded (may be null)
    ShardHit *out_hits;   // [k] hits with global rows, {0,0,0} padded (may be null)
    uint64_t *out_rows;   // [k] (may be null)
    float *out_scores;    // [k] (may be null)
    uint32_t *out_count;  // (may be null)
    uint64_t row_base;
    uint32_t accumulate_count;
};
__device__ __forceinline__ void write_outputs(const TopKState &st, uint32_t t, uint32_t k,
                                              const TopKOutputs o) {
    for (uint32_t i = t; i < k; i += kRowsPerBlock) {
        const uint64_t key = (i < st.count) ? st.buf[i] : 0ull;
        if (o.out_keys) o.out_keys[i] = key;
        const uint64_t grow = o.row_base + key_local_row(key);
        const uint32_t sb = key_score_bits(key);
        if (o.out_hits) {
            ShardHit h;
            // an empty slot is {0,0,0}; a real NaN hit has ord 0 too but score_bits != 0
            h.global_row = key ? grow : 0ull;
            h.ord = key ? (uint32_t)(key >> 32) : 0u;
            h.score_bits = key ? sb : 0u;
            o.out_hits[i] = h;
        }
        if (i < st.count) {
            if (o.out_rows) o.out_rows[i] = grow;
            if (o.out_scores) o.out_scores[i] = __uint_as_float(sb);
        }
    }
    if (t == 0 && o.out_count) {
        if (o.accumulate_count) atomicAdd(o.out_count, st.count);
        else *o.out_count = st.count;
    }
}


// ---------------------------------------------------------------------------------------
// fused cross-shard exchange + merge (replaces ncclAllGather + merge_shards_kernel on the
// single-query path; semantics of ResultMerger::merge_top_k, distributed.rs:413-433).
// Called by the 256 consumer threads of the LAST CTA with this shard's sorted top-k keys in
// st.buf[0..count).  Writes this shard's hits into every rank's mailbox over NVLink, raises
// this rank's sequence flag on every peer, waits for all peers' flags, then merges the
// n_ranks*k gathered hits (concatenation position breaks ties == stable sort in shard order).
// Returns false if a peer did not show up within ~4 s of spinning (count is then poisoned).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool exchange_and_merge(TopKState &st, uint32_t t, const PeerXchg &x,
                                                   uint64_t row_base, uint64_t *scratch_keys,
                                                   const MergeScratch ms, uint64_t *out_rows,
                                                   float *out_scores, uint32_t *out_count) {
    const uint32_t k = st.k;
    const uint32_t slot = x.seq & 1u;
    // 1. my hits -> every rank's mailbox[slot][my rank]
    for (uint32_t r = 0; r < x.n_ranks; ++r) {
        ShardHit *dst = x.mailbox[r] + ((size_t)slot * x.n_ranks + x.rank) * x.kcap;
        for (uint32_t i = t; i < k; i += kRowsPerBlock) {
            const uint64_t key = (i < st.count) ? st.buf[i] : 0ull;
            uint4 w;
            const uint64_t grow = key ? row_base + key_local_row(key) : 0ull;
            w.x = (uint32_t)grow;
            w.y = (uint32_t)(grow >> 32);
            w.z = key ? (uint32_t)(key >> 32) : 0u;   // ord
            w.w = key ? key_score_bits(key) : 0u;     // exact score bits
            *reinterpret_cast<uint4 *>(dst + i) = w;
        }
    }
    __threadfence_system();
    consumer_sync();
    if (t < x.n_ranks) st_release_sys(x.flags[t] + x.rank, x.seq);
    // 2. wait until every rank's hits for this sequence number have landed here
    bool ok = true;
    if (t < x.n_ranks) {
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(x.flags[x.rank] + t) - x.seq) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 8000000000ll) {
                ok = false;
                break;
            }
        }
    }
    const uint32_t bad = consumer_sync_popc(!ok);
    if (bad) {
        if (t == 0 && out_count) *out_count = 0xffffffffu;
        return false;
    }
    // 3. merge keys: (ord, position in the concatenation)
    const ShardHit *mine = x.mailbox[x.rank] + (size_t)slot * x.n_ranks * x.kcap;
    const uint32_t total = x.n_ranks * k;
    for (uint32_t pos = t; pos < total; pos += kRowsPerBlock) {
        const uint32_t r = pos / k, i = pos - r * k;
        const uint4 w = __ldcg(reinterpret_cast<const uint4 *>(mine + (size_t)r * x.kcap + i));
        const bool valid = (w.z != 0u) || (w.w != 0u);
        scratch_keys[pos] = valid ? (((uint64_t)w.z << 32) | (uint64_t)(0xffffffffu - pos)) : 0ull;
    }
    __threadfence();
    if (t == 0) *st.cnt_smem = 0u;
    st.count = 0;
    st.cap = kCandCap;
    consumer_sync();
    merge_published(st, t, scratch_keys, total, k, ms);
    for (uint32_t i = t; i < st.count; i += kRowsPerBlock) {
        const uint32_t pos = 0xffffffffu - (uint32_t)st.buf[i];
        const uint32_t r = pos / k, j = pos - r * k;
        const uint4 w = __ldcg(reinterpret_cast<const uint4 *>(mine + (size_t)r * x.kcap + j));
        if (out_rows) out_rows[i] = ((uint64_t)w.y << 32) | w.x;
        if (out_scores) out_scores[i] = __uint_as_float(w.w);
    }
    if (t == 0 && out_count) *out_count = st.count;
    return true;
}

// ---------------------------------------------------------------------------------------
// per-row accumulation
// ---------------------------------------------------------------------------------------
template <int METRIC>
struct RowAcc {
    float d[8];  // dot lanes   (cosine, dot)  | d[0] = running L2 sum (euclidean)
    float s[8];  // sumsq lanes (cosine)
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            d[j] = 0.0f;
            s[j] = 0.0f;
        }
    }
    // one float4 of the row (elements 4u..4u+3 of the current 32-float chunk), with the
    // matching float4 of the query.  `half` = u & 1 selects lanes 0-3 or 4-7.
    template <int HALF>
    __device__ __forceinline__ void step(const float4 v, const float4 q) {
        if (METRIC == kEuclidean) {
            float t0 = __fsub_rn(q.x, v.x);
            d[0] = __fadd_rn(d[0], __fmul_rn(t0, t0));
            float t1 = __fsub_rn(q.y, v.y);
            d[0] = __fadd_rn(d[0], __fmul_rn(t1, t1));
            float t2 = __fsub_rn(q.z, v.z);
            d[0] = __fadd_rn(d[0], __fmul_rn(t2, t2));
            float t3 = __fsub_rn(q.w, v.w);
            d[0] = __fadd_rn(d[0], __fmul_rn(t3, t3));
        } else {
            constexpr int L = HALF * 4;
            d[L + 0] = __fadd_rn(d[L + 0], __fmul_rn(q.x, v.x));
            d[L + 1] = __fadd_rn(d[L + 1], __fmul_rn(q.y, v.y));
            d[L + 2] = __fadd_rn(d[L + 2], __fmul_rn(q.z, v.z));
            d[L + 3] = __fadd_rn(d[L + 3], __fmul_rn(q.w, v.w));
            if (METRIC == kCosine) {
                s[L + 0] = __fadd_rn(s[L + 0], __fmul_rn(v.x, v.x));
                s[L + 1] = __fadd_rn(s[L + 1], __fmul_rn(v.y, v.y));
                s[L + 2] = __fadd_rn(s[L + 2], __fmul_rn(v.z, v.z));
                s[L + 3] = __fadd_rn(s[L + 3], __fmul_rn(v.w, v.w));
            }
        }
    }
};

__device__ __forceinline__ float fold_lanes(const float *l) {
    float r = 0.0f;  // arr.iter().sum(): left fold from 0.0 (hnsw.rs:184)
#pragma unroll
    for (int j = 0; j < 8; ++j) r = __fadd_rn(r, l[j]);
    return r;
}

struct ScanParams {
    const float *query;      // [dim] device
    uint64_t *cand;          // [grid, k] per-CTA best keys (workspace)
    uint32_t *done_counter;  // [0] ticket for "last CTA merges", [1] dynamic row-block cursor;
                             // both left at 0 on exit
    uint64_t *out_keys;      // [k] merged local keys, descending, 0 padded (may be null)
    ShardHit *out_hits;      // [k] merged hits with global rows (may be null)
    uint64_t *out_rows;      // [k] global rows (may be null)
    float *out_scores;       // [k] (may be null)
    uint32_t *out_count;     // [1] (may be null)
    uint64_t row_base;       // global index of local row 0
    uint32_t n_rows;         // local rows
    uint32_t dim;
    uint32_t k;
    uint32_t n_stages;
    uint32_t q_floats;       // dim rounded up to a multiple of 32
    uint32_t evict_first;    // 1: stream the corpus through L2 with evict_first
    // k > kMaxFastK is served by repeated passes: pass p only admits keys strictly below the
    // last key of pass p-1 (read from device memory, so passes chain without a host sync).
    const uint64_t *key_ceiling;  // null = no ceiling; *key_ceiling == 0 = nothing left
    uint32_t accumulate_count;    // 1: atomicAdd into *out_count instead of storing
    PeerXchg xchg;                // n_ranks > 1: fused cross-shard exchange in the last CTA
    // optional pre-filter (search_with_pre_filter, vector_engine/src/lib.rs:3514-3557): bit r
    // set = row r takes part.  Padded to whole row blocks (8 words each).  Row blocks whose 8
    // words are all zero are never loaded.
    const uint32_t *row_mask;
    // Pipelined launches (nm_index_set_pipelining; launched with the programmatic stream
    // serialization attribute): pdl_seq >= 1 numbers the pipelined scans of one workspace,
    // *pdl_done is the highest sequence number whose last CTA has finished.  `cand` and
    // `done_counter` alternate between two sets by sequence parity, and a scan only starts
    // once scan pdl_seq - 2 — the previous user of its set — is done.
    uint32_t pdl_seq;
    uint32_t *pdl_done;
    // Conditional launch: when set, the kernel runs only if *gate != 0 (the exact redo of a
    // query the tensor-core pre-filter flagged, decided on the device: asynchronous searches
    // never read flags back to the host).
    const uint32_t *gate;
};

// Uniform "is this conditional launch needed?": true when any of gate[0..n) is non-zero, or
// when there is no gate.  Contains a barrier: call with all threads of the CTA.
__device__ __forceinline__ bool gate_open(const uint32_t *gate, uint32_t n) {
    if (!gate) return true;
    uint32_t any = 0u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) any |= __ldcg(gate + i);
    return __syncthreads_or((int)any) != 0;
}

// Read element `col` (0..31) of row t in a swizzled stage.
__device__ __forceinline__ float stage_elem(const uint8_t *stage, uint32_t t, uint32_t col) {
    uint32_t unit = (col >> 2) ^ (t & 7u);
    return *reinterpret_cast<const float *>(stage + t * 128u + (unit << 4) + ((col & 3u) << 2));
}

template <int METRIC>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve-up: stages | candidate buffer | query | barriers + scalars
    // (offset arithmetic, not an integer round-trip, so the compiler keeps the shared state space)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *stages = smem;
    uint64_t *cand_buf = reinterpret_cast<uint64_t *>(stages + p.n_stages * kStageBytes);
    float *q_s = reinterpret_cast<float *>(cand_buf + kCandCap);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(q_s + p.q_floats);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *thr_s = empty_bar + kMaxStages;
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(thr_s + 1);
    float *qmag_s = reinterpret_cast<float *>(cnt_s + 1);
    uint32_t *ticket_s = cnt_s + 2;
    uint32_t *rb_ring = cnt_s + 3;  // row block carried by each stage (kMaxStages entries)

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    if (!gate_open(p.gate, 1u)) return;
    const uint32_t n_stages = p.n_stages;
    const uint32_t n_rb = (p.n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    const uint32_t n_kc_full = p.dim / kChunkFloats;
    const uint32_t rem_cols = p.dim % kChunkFloats;
    const uint32_t n_kc = n_kc_full + (rem_cols ? 1u : 0u);

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        *thr_s = 0ull;
        *cnt_s = 0u;
        fence_mbar_init();
        if (p.pdl_seq) {
            // the scan two launches back used the same candidate lists and counters
            while ((int32_t)(ld_acquire_gpu(p.pdl_done) + 2u - p.pdl_seq) < 0) __nanosleep(200);
        }
    }
    __syncthreads();
    // from here on the next pipelined scan may be scheduled onto SMs as they become free
    if (p.pdl_seq) pdl_launch_dependents();

    if (warp == kConsumerWarps) {
        // ===================== TMA producer =====================
        // Row blocks are handed out dynamically: the first one is blockIdx.x, every further
        // one comes from a global cursor, so SMs that see more bandwidth simply take more
        // blocks and all CTAs finish together.  The block id travels to the consumers in
        // rb_ring[stage] (published by the mbarrier arrive); 0xffffffff ends the stream.
        if (tid == kRowsPerBlock) {
            const uint64_t policy = p.evict_first ? policy_evict_first() : policy_evict_normal();
            uint32_t stage = 0, phase = 0;
            uint32_t rb = blockIdx.x;
            for (;;) {
                if (p.row_mask) {
                    // pre-filter: skip row blocks without a single eligible row
                    while (rb < n_rb) {
                        const uint4 *m = reinterpret_cast<const uint4 *>(p.row_mask + rb * 8u);
                        const uint4 a = __ldg(m), b = __ldg(m + 1);
                        if ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0u) break;
                        rb = atomicAdd(p.done_counter + 1, 1u) + gridDim.x;
                    }
                }
                if (rb >= n_rb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    rb_ring[stage] = 0xffffffffu;
                    mbar_arrive(&full_bar[stage]);
                    break;
                }
                const uint32_t next = atomicAdd(p.done_counter + 1, 1u) + gridDim.x;
                for (uint32_t kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (kc == 0) rb_ring[stage] = rb;
                    mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d(stages + stage * kStageBytes, &tmap, (int32_t)(kc * kChunkFloats),
                                (int32_t)(rb * kRowsPerBlock), &full_bar[stage], policy);
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                rb = next;
            }
        }
        return;
    }

    // ===================== consumers: one row per thread =====================
    const uint32_t t = tid;
    const uint32_t lane = t & 31u;
    // query -> shared (zero padded to a multiple of 32 floats)
    for (uint32_t i = t; i < p.q_floats; i += kRowsPerBlock)
        q_s[i] = (i < p.dim) ? __ldcg(p.query + i) : 0.0f;  // L2: never a stale L1 line
    consumer_sync();
    float qmag = 0.0f;
    if (METRIC == kCosine) {
        // |q| with the same lane tree (simd::magnitude, hnsw.rs:198-229): lanes 0..7 of warp 0
        // each own one f32x8 lane, lane 0 folds them in order and adds the tail.
        if (warp == 0) {
            float acc = 0.0f;
            const uint32_t chunks = p.dim / 8u;
            if (lane < 8u) {
                for (uint32_t c = 0; c < chunks; ++c) {
                    float v = q_s[c * 8u + lane];
                    acc = __fadd_rn(acc, __fmul_rn(v, v));
                }
            }
            float r = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) r = __fadd_rn(r, __shfl_sync(0xffffffffu, acc, j));
            if (lane == 0) {
                for (uint32_t i = chunks * 8u; i < p.dim; ++i) {
                    float v = q_s[i];
                    r = __fadd_rn(r, __fmul_rn(v, v));
                }
                *qmag_s = __fsqrt_rn(r);
            }
        }
        consumer_sync();
        qmag = *qmag_s;
    }

    const uint64_t ceiling = p.key_ceiling ? __ldg(p.key_ceiling) : ~0ull;

    TopKState st;
    st.buf = cand_buf;
    st.cnt_smem = cnt_s;
    st.thr_smem = thr_s;
    st.count = 0;
    st.k = p.k;
    // capacity: >= k + one row block of offers, power of two, at most kCandCap
    st.cap = 512u;
    while (st.cap < p.k + (uint32_t)kRowsPerBlock) st.cap <<= 1;

    const uint32_t swz = t & 7u;
    uint32_t stage = 0, phase = 0;
    RowAcc<METRIC> acc;

    for (;;) {
        // the first stage of a row block also carries the block id
        mbar_wait(&full_bar[stage], phase);
        const uint32_t rb = rb_ring[stage];
        if (rb == 0xffffffffu) break;
        acc.reset();
        for (uint32_t kc = 0; kc < n_kc_full; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            const uint8_t *srow = stages + stage * kStageBytes + t * 128u;
            const float4 *qv = reinterpret_cast<const float4 *>(q_s) + kc * 8u;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                float4 v0 = *reinterpret_cast<const float4 *>(srow + (((uint32_t)u ^ swz) << 4));
                float4 v1 =
                    *reinterpret_cast<const float4 *>(srow + (((uint32_t)(u + 1) ^ swz) << 4));
                float4 q0 = qv[u], q1 = qv[u + 1];
                acc.template step<0>(v0, q0);
                acc.template step<1>(v1, q1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == n_stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        float dot, ssq = 0.0f;
        if (rem_cols) {
            // last, partial chunk: whole f32x8 groups first, then fold, then the scalar tail
            mbar_wait(&full_bar[stage], phase);
            const uint8_t *sbase = stages + stage * kStageBytes;
            const float *qt = q_s + n_kc_full * kChunkFloats;
            const uint32_t groups = rem_cols / 8u;
            if (METRIC == kEuclidean) {
                for (uint32_t c = 0; c < rem_cols; ++c) {
                    float df = __fsub_rn(qt[c], stage_elem(sbase, t, c));
                    acc.d[0] = __fadd_rn(acc.d[0], __fmul_rn(df, df));
                }
                dot = acc.d[0];
            } else {
                for (uint32_t g = 0; g < groups; ++g) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = stage_elem(sbase, t, g * 8u + j);
                        acc.d[j] = __fadd_rn(acc.d[j], __fmul_rn(qt[g * 8u + j], v));
                        if (METRIC == kCosine) acc.s[j] = __fadd_rn(acc.s[j], __fmul_rn(v, v));
                    }
                }
                dot = fold_lanes(acc.d);
                if (METRIC == kCosine) ssq = fold_lanes(acc.s);
                for (uint32_t c = groups * 8u; c < rem_cols; ++c) {
                    float v = stage_elem(sbase, t, c);
                    dot = __fadd_rn(dot, __fmul_rn(qt[c], v));
                    if (METRIC == kCosine) ssq = __fadd_rn(ssq, __fmul_rn(v, v));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == n_stages) {
                stage = 0;
                phase ^= 1u;
            }
        } else {
            if (METRIC == kEuclidean) {
                dot = acc.d[0];
            } else {
                dot = fold_lanes(acc.d);
                if (METRIC == kCosine) ssq = fold_lanes(acc.s);
            }
        }
        float score;
        if (METRIC == kCosine) {
            float rmag = __fsqrt_rn(ssq);
            score = (qmag == 0.0f || rmag == 0.0f) ? 0.0f
                                                   : __fdiv_rn(dot, __fmul_rn(qmag, rmag));
        } else if (METRIC == kEuclidean) {
            score = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fsqrt_rn(dot)));
        } else {
            score = dot;
        }
        const uint32_t row = rb * kRowsPerBlock + t;
        uint64_t key = (row < p.n_rows) ? make_key(__float_as_uint(score), row) : 0ull;
        if (key >= ceiling) key = 0ull;
        if (p.row_mask && key && !((__ldg(p.row_mask + (row >> 5)) >> (row & 31u)) & 1u)) key = 0ull;
        topk_offer(st, key, t);
    }

    // ---- publish this CTA's k best, last CTA merges everything ----
    topk_prune(st, t);
    uint64_t *my_cand = p.cand + (uint64_t)blockIdx.x * p.k;
    for (uint32_t i = t; i < p.k; i += kRowsPerBlock) my_cand[i] = (i < st.count) ? st.buf[i] : 0ull;
    __threadfence();
    consumer_sync();
    if (t == 0) *ticket_s = atomicAdd(p.done_counter, 1u);
    consumer_sync();
    if (*ticket_s != gridDim.x - 1) return;
    __threadfence();

    // all lists (this CTA's own included) are in p.cand: select the k best of grid*k keys
    MergeScratch ms;
    ms.hist = reinterpret_cast<uint32_t *>(stages);  // the stage ring is idle now
    ms.sc = ms.hist + 256;
    if (t == 0) *st.cnt_smem = 0u;
    st.count = 0;
    st.cap = kCandCap;
    consumer_sync();
    merge_published(st, t, p.cand, gridDim.x * p.k, p.k, ms);
    // Pipelined: everything above overlapped with the previous scan's tail.  Its outputs, its
    // turn in the peer exchange and its completion come first.
    if (p.pdl_seq) pdl_wait_prior_grids();
    if (p.xchg.n_ranks > 1) {
        // p.cand has room for gridDim.x * k >= n_ranks * k merge keys only if the grid is
        // at least n_ranks CTAs; the host guarantees a scratch of max(grid, n_ranks) * k
        exchange_and_merge(st, t, p.xchg, p.row_base, p.cand, ms, p.out_rows, p.out_scores,
                           p.out_count);
        consumer_sync();
        if (t == 0) {
            p.done_counter[0] = 0u;
            p.done_counter[1] = 0u;
            if (p.pdl_seq) {
                __threadfence();
                st_release_gpu(p.pdl_done, p.pdl_seq);
            }
        }
        return;
    }
    TopKOutputs o;
    o.out_keys = p.out_keys;
    o.out_hits = p.out_hits;
    o.out_rows = p.out_rows;
    o.out_scores = p.out_scores;
    o.out_count = p.out_count;
    o.row_base = p.row_base;
    o.accumulate_count = p.accumulate_count;
    write_outputs(st, t, p.k, o);
    consumer_sync();
    if (t == 0) {
        // every CTA took its ticket after its producer's last cursor fetch: safe to reset
        p.done_counter[0] = 0u;
        p.done_counter[1] = 0u;
        if (p.pdl_seq) {
            __threadfence();
            st_release_gpu(p.pdl_done, p.pdl_seq);
        }
    }
}

// A rank whose shard is empty still has to take part in the exchange.
__global__ void __launch_bounds__(kRowsPerBlock)
exchange_empty_shard_kernel(PeerXchg x, uint32_t k, uint64_t *scratch_keys, uint64_t *out_rows,
                            float *out_scores, uint32_t *out_count) {
    __shared__ __align__(16) uint64_t buf[kCandCap];
    __shared__ uint64_t thr_s;
    __shared__ uint32_t cnt_s;
    __shared__ uint32_t hist[256 + 16];
    const uint32_t t = threadIdx.x;
    if (t == 0) {
        thr_s = 0ull;
        cnt_s = 0u;
    }
    __syncthreads();
    TopKState st;
    st.buf = buf;
    st.cnt_smem = &cnt_s;
    st.thr_smem = &thr_s;
    st.count = 0;
    st.k = k;
    st.cap = kCandCap;
    MergeScratch ms;
    ms.hist = hist;
    ms.sc = hist + 256;
    exchange_and_merge(st, t, x, 0ull, scratch_keys, ms, out_rows, out_scores, out_count);
}

// ---------------------------------------------------------------------------------------
// cross-shard merge (ResultMerger::merge_top_k, query_router/src/distributed.rs:413-433):
// concat shard lists in shard order, stable sort by score desc, truncate k.  hits is
// [n_shards, k] as gathered; one CTA per query.
// ---------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;
__global__ void __launch_bounds__(kMergeThreads)
merge_shards_kernel(const ShardHit *hits, uint32_t n_shards, uint32_t k, uint32_t hits_stride_q,
                    uint32_t n_sort, uint64_t *out_rows, float *out_scores,
                    uint32_t *out_counts) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(merge_smem);
    const uint32_t q = blockIdx.x;
    const uint32_t t = threadIdx.x;
    const uint32_t total = n_shards * k;
    // gathered layout: [shard][query][k]
    auto hit_at = [&](uint32_t pos) -> const ShardHit & {
        uint32_t s = pos / k, i = pos % k;
        return hits[(uint64_t)s * hits_stride_q + (uint64_t)q * k + i];
    };
    for (uint32_t i = t; i < n_sort; i += kMergeThreads) {
        uint64_t key = 0ull;
        if (i < total) {
            const ShardHit &h = hit_at(i);
            bool valid = (h.ord != 0u) || (h.score_bits != 0u);
            // position in the concatenation breaks ties == stable sort; +1 keeps NaN hits > 0
            if (valid) key = ((uint64_t)h.ord << 32) | (uint64_t)(0xffffffffu - i);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= n_sort; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = t; i < (n_sort >> 1); i += kMergeThreads) {
                uint32_t lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                bool swap = desc ? (a < b) : (a > b);
                if (swap) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    uint32_t cnt = 0;
    for (uint32_t i = t; i < k; i += kMergeThreads) {
        uint64_t key = keys[i];
        if (key != 0ull) {
            const ShardHit &h = hit_at(0xffffffffu - (uint32_t)key);
            out_rows[(uint64_t)q * k + i] = h.global_row;
            out_scores[(uint64_t)q * k + i] = __uint_as_float(h.score_bits);
        }
    }
    if (t == 0) {
        // keys are sorted: count = first zero within k
        uint32_t lim = k < n_sort ? k : n_sort;
        while (cnt < lim && keys[cnt] != 0ull) ++cnt;
        out_counts[q] = cnt;
    }
}

// The same merge for gathers that do not fit the shared-memory sort (n_shards * k > 4096, e.g.
// k = 100 000 over 8 shards): every shard list is already sorted, so each hit computes its final
// rank directly — its position in its own list plus, for every other list, the number of hits
// that precede it in the total order (binary search; on equal scores the earlier shard wins,
// which is exactly the stable sort of the concatenation).  No size limit, no scratch.
__global__ void __launch_bounds__(kMergeThreads)
merge_shards_rank_kernel(const ShardHit *hits, uint32_t n_shards, uint32_t k, uint32_t hits_stride_q,
                         uint64_t *out_rows, float *out_scores, uint32_t *out_counts) {
    const uint32_t q = blockIdx.y;
    const uint32_t pos = blockIdx.x * kMergeThreads + threadIdx.x;
    auto list = [&](uint32_t s) { return hits + (uint64_t)s * hits_stride_q + (uint64_t)q * k; };
    auto valid_count = [&](const ShardHit *l) {  // valid hits come first, empty slots {0,0,0} last
        uint32_t lo = 0, hi = k;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const ShardHit h = l[mid];
            if (h.ord != 0u || h.score_bits != 0u) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    if (pos == 0) {
        uint64_t total = 0;
        for (uint32_t s = 0; s < n_shards; ++s) total += valid_count(list(s));
        out_counts[q] = (uint32_t)(total < k ? total : k);
    }
    if (pos >= n_shards * k) return;
    const uint32_t s = pos / k, i = pos - s * k;
    const ShardHit h = list(s)[i];
    if (h.ord == 0u && h.score_bits == 0u) return;
    uint64_t rank = i;
    for (uint32_t t = 0; t < n_shards && rank < k; ++t) {
        if (t == s) continue;
        const ShardHit *l = list(t);
        const uint32_t nv = valid_count(l);
        // hits of list t that precede h: ord > h.ord, or ord == h.ord when t is an earlier shard
        uint32_t lo = 0, hi = nv;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t o = l[mid].ord;
            const bool before = (o > h.ord) || (o == h.ord && t < s);
            if (before) lo = mid + 1;
            else hi = mid;
        }
        rank += lo;
    }
    if (rank < k) {
        out_rows[(uint64_t)q * k + rank] = h.global_row;
        out_scores[(uint64_t)q * k + rank] = __uint_as_float(h.score_bits);
    }
}

// ---------------------------------------------------------------------------------------
// synthetic corpus: bit-identical to oracle nmo_fill_synthetic
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t flat) {
    uint32_t u24 = (uint32_t)(splitmix64(splitmix64(seed) ^ flat) >> 40);
    return (float)u24 * 1.1920928955078125e-07f - 1.0f;  // 2^-23, both ops exact
}

__global__ void fill_synthetic_kernel(float *rows, uint64_t n, uint32_t dim, uint32_t pitch,
                                      uint64_t seed, uint64_t global_row0) {
    const uint64_t total = n * (uint64_t)pitch;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / pitch;
        uint32_t c = (uint32_t)(i % pitch);
        rows[i] = (c < dim) ? synth_value(seed, (global_row0 + r) * dim + c) : 0.0f;
    }
}
#endif  // __CUDACC__

}  // namespace nm
