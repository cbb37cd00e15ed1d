// filter_kernels.cuh — device-side evaluation of a metadata filter into the row bitmask of a
// pre-filtered scan ("filter first, then search the subset": search_with_pre_filter,
// vector_engine/src/lib.rs:3514-3557; evaluate_filter :3592-3640; the type rules of
// compare_tensor_value_to_filter :3658-3684).
//
// Metadata lives on the device as typed COLUMNS next to the mirror (one per field):
//   tags[row]  u8   0 missing | 1 null | 2 bool | 3 int | 4 float | 5 string
//   vals[row]  u64  bool 0/1 | i64 | f64 bits | string = code in the column's dictionary
// A FilterCondition tree arrives as a POSTFIX program (nm_filter_op, include/neumann_b200.h)
// whose leaves name a column.  Anything that needs the bytes of a string (=, <, CONTAINS,
// STARTS_WITH, IN over strings) was evaluated by the host ONCE PER DISTINCT STRING of the column
// into a bit table; the leaf only looks its row's code up.  Numbers, bools and nulls are compared
// here with the reference's rules: Int/Int as i64, Float/Float as f64 (NaN on either side is
// incomparable), mixed Int/Float through f64, anything else incomparable; an incomparable or
// missing value makes EVERY comparison false — `!=` included (lib.rs:3642-3656).
//
// One ballot per warp and 32 rows: the output is the u32 mask the scan kernels read, padded with
// zeros to whole 256-row blocks (8 words each).  Algorithmic traffic: 9 bytes per row and leaf
// (10M rows x 1 leaf = 90 MB; measured figures in DESIGN.md §4.8).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_types.hpp"

namespace nm {

#ifdef __CUDACC__
// Ordering of a row's value against the literal: 0 less, 1 equal, 2 greater, 3 incomparable.
__device__ __forceinline__ uint32_t filter_ord_i64(long long a, long long b) {
    return a < b ? 0u : (a == b ? 1u : 2u);
}
__device__ __forceinline__ uint32_t filter_ord_f64(double a, double b) {
    if (a != a || b != b) return 3u;  // partial_cmp -> None
    return a < b ? 0u : (a > b ? 2u : 1u);
}

// Which orderings satisfy a comparison, as a bit per ordering (bit 3 — incomparable — never set:
// every comparison is false then, `!=` included).
__device__ __forceinline__ uint32_t filter_cmp_accepts(uint32_t cmp) {
    switch (cmp) {
    case kFilterEq: return 0x2u;
    case kFilterNe: return 0x5u;
    case kFilterLt: return 0x1u;
    case kFilterLe: return 0x3u;
    case kFilterGt: return 0x4u;
    default: return 0x6u;  // kFilterGe
    }
}

// One leaf over the U rows a thread holds: bit u of the result = the leaf's verdict for row u.
// Everything that depends on the op alone (kind, literal type, comparison) is decided ONCE, outside
// the per-row work — the kernel is instruction-bound otherwise (112 warp instructions per row and
// leaf measured with the dispatch inside; ncu r02_filter_mask_full).
template <int U>
__device__ __forceinline__ uint32_t filter_leaf_rows(const FilterOpDev &op, const uint32_t (&tag)[U],
                                                     const uint64_t (&val)[U]) {
    uint32_t res = 0;
    if (op.kind == kFilterExists) {
#pragma unroll
        for (int u = 0; u < U; ++u) res |= (tag[u] != kTagMissing ? 1u : 0u) << u;
        return res;
    }
    if (op.kind == kFilterStrTable) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = tag[u] == kTagString && val[u] < op.table_bits;
            const uint32_t w = ok ? __ldg(op.table + (val[u] >> 5)) : 0u;
            res |= ((w >> ((uint32_t)val[u] & 31u)) & 1u) << u;
        }
        return res;
    }
    // kFilterCmp against a non-string literal
    const uint32_t accepts = filter_cmp_accepts(op.cmp);
    if (op.lit_tag == kTagInt) {
        // Int against Int without a branch; a Float value in the column (compared through f64,
        // lib.rs:3670-3677) sends the whole warp through the slower loop below.
        const long long b = (long long)op.lit;
        bool any_float = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t c = tag[u] == kTagInt ? filter_ord_i64((long long)val[u], b) : 3u;
            res |= ((accepts >> c) & 1u) << u;
            any_float |= tag[u] == kTagFloat;
        }
        if (__any_sync(0xffffffffu, any_float)) {
            const double bf = (double)b;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (tag[u] == kTagFloat) {
                    const uint32_t c = filter_ord_f64(__longlong_as_double((long long)val[u]), bf);
                    res |= ((accepts >> c) & 1u) << u;
                }
        }
    } else if (op.lit_tag == kTagFloat) {
        const double b = __longlong_as_double((long long)op.lit);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t c = 3u;
            if (tag[u] == kTagFloat) c = filter_ord_f64(__longlong_as_double((long long)val[u]), b);
            else if (tag[u] == kTagInt) c = filter_ord_f64((double)(long long)val[u], b);
            res |= ((accepts >> c) & 1u) << u;
        }
    } else if (op.lit_tag == kTagBool) {
        const uint32_t b = op.lit != 0 ? 1u : 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t c = tag[u] == kTagBool ? (val[u] != 0 ? 1u : 0u) + 1u - b : 3u;
            res |= ((accepts >> c) & 1u) << u;
        }
    } else if (op.lit_tag == kTagNull) {
#pragma unroll
        for (int u = 0; u < U; ++u) res |= ((accepts >> (tag[u] == kTagNull ? 1u : 3u)) & 1u) << u;
    }
    return res;
}

// ops: postfix program in global memory (n_ops <= kFilterMaxOps, stack depth <= 64 checked by
// the host).  mask: [n_words] u32, n_words = ceil(n_rows / 256) * 8.
// A warp owns one 256-row block of the mask at a time: lane l holds rows base + 32u + l, u < 8, so
// that the 16 column loads of a leaf are in flight together at constant offsets from one address,
// every load instruction reads 32 consecutive entries, and ballot u is mask word u of the block.
// The evaluation stack of row u is the bits of stack[u] (bit 0 = top).
template <typename StackT>  // uint32_t when the program's stack never holds more than 32 entries
__global__ void __launch_bounds__(256, 5)
filter_mask_kernel(const FilterOpDev *__restrict__ ops, uint32_t n_ops, uint64_t n_rows,
                   uint32_t *__restrict__ mask, uint64_t n_words) {
    __shared__ FilterOpDev s_ops[kFilterMaxOps];
    for (uint32_t i = threadIdx.x; i < n_ops; i += blockDim.x) s_ops[i] = ops[i];
    __syncthreads();
    constexpr int U = 8;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = n_words / U;
    for (uint64_t tile = (uint64_t)blockIdx.x * 8 + warp; tile < n_tiles; tile += (uint64_t)gridDim.x * 8) {
        const uint64_t base = tile * (32ull * U) + lane;
        const bool full = (tile + 1) * (32ull * U) <= n_rows;  // warp-uniform
        StackT stack[U];
#pragma unroll
        for (int u = 0; u < U; ++u) stack[u] = 0;
        for (uint32_t i = 0; i < n_ops; ++i) {
            const FilterOpDev &op = s_ops[i];
            if (op.kind == kFilterAnd || op.kind == kFilterOr) {
                const bool is_and = op.kind == kFilterAnd;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const StackT a = stack[u] & 1u, b = (stack[u] >> 1) & 1u;
                    stack[u] = ((stack[u] >> 2) << 1) | (is_and ? (a & b) : (a | b));
                }
                continue;
            }
            uint32_t res;
            if (op.kind == kFilterTrue) {
                res = (1u << U) - 1u;
            } else if (op.kind == kFilterFalse || op.tags == nullptr) {
                res = 0;  // no such column on this shard: every row is "missing"
            } else {
                uint32_t tag[U];
                uint64_t val[U];
                const bool need_val = op.kind != kFilterExists;
                const uint8_t *tp = op.tags + base;
                const uint64_t *vp = op.vals + base;
                if (full) {
#pragma unroll
                    for (int u = 0; u < U; ++u) tag[u] = __ldg(tp + 32 * u);
#pragma unroll
                    for (int u = 0; u < U; ++u) val[u] = need_val ? __ldg(vp + 32 * u) : 0ull;
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool in = base + 32u * u < n_rows;
                        tag[u] = in ? (uint32_t)__ldg(tp + 32 * u) : (uint32_t)kTagMissing;
                        val[u] = need_val && in ? __ldg(vp + 32 * u) : 0ull;
                    }
                }
                res = filter_leaf_rows<U>(op, tag, val);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) stack[u] = (stack[u] << 1) | (StackT)((res >> u) & 1u);
        }
        uint32_t word = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool pass = (full || base + 32u * u < n_rows) && (stack[u] & 1u) != 0;
            const uint32_t w = __ballot_sync(0xffffffffu, pass);
            if (lane == (uint32_t)u) word = w;
        }
        if (lane < (uint32_t)U) mask[tile * U + lane] = word;
    }
}

// Move one row's column entry (swap-remove keeps the columns in step with the mirror).
__global__ void column_move_kernel(uint8_t *tags, uint64_t *vals, uint64_t dst, uint64_t src) {
    tags[dst] = tags[src];
    vals[dst] = vals[src];
}
#endif  // __CUDACC__

}  // namespace nm
