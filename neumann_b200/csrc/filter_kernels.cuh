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
// One thread per row, one ballot per warp: the output is the u32 mask the scan kernels read,
// padded with zeros to whole 256-row blocks (8 words each).  HBM-bound on 9 bytes per row and
// referenced column: 10M rows x 2 leaves = 180 MB ~ 30 us.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_types.hpp"

namespace nm {

#ifdef __CUDACC__
__device__ __forceinline__ bool filter_cmp_result(int c, uint32_t cmp) {
    // c: -1 / 0 / 1, or 2 = incomparable
    if (c == 2) return false;
    switch (cmp) {
    case kFilterEq: return c == 0;
    case kFilterNe: return c != 0;
    case kFilterLt: return c < 0;
    case kFilterLe: return c <= 0;
    case kFilterGt: return c > 0;
    default: return c >= 0;
    }
}

__device__ __forceinline__ int filter_ord_f64(double a, double b) {
    if (a != a || b != b) return 2;  // partial_cmp -> None
    return a < b ? -1 : (a > b ? 1 : 0);
}

__device__ __forceinline__ bool filter_leaf(const FilterOpDev &op, uint64_t row) {
    const uint32_t tag = op.tags ? op.tags[row] : (uint32_t)kTagMissing;
    if (op.kind == kFilterExists) return tag != kTagMissing;
    if (tag == kTagMissing) return false;
    const uint64_t v = op.vals[row];
    if (op.kind == kFilterStrTable) {
        if (tag != kTagString || v >= op.table_bits) return false;
        return (op.table[v >> 5] >> (v & 31u)) & 1u;
    }
    // kFilterCmp against a non-string literal
    int c = 2;
    if (op.lit_tag == kTagInt) {
        const long long b = (long long)op.lit;
        if (tag == kTagInt) {
            const long long a = (long long)v;
            c = a < b ? -1 : (a > b ? 1 : 0);
        } else if (tag == kTagFloat) {
            c = filter_ord_f64(__longlong_as_double((long long)v), (double)b);
        }
    } else if (op.lit_tag == kTagFloat) {
        const double b = __longlong_as_double((long long)op.lit);
        if (tag == kTagFloat) c = filter_ord_f64(__longlong_as_double((long long)v), b);
        else if (tag == kTagInt) c = filter_ord_f64((double)(long long)v, b);
    } else if (op.lit_tag == kTagBool) {
        if (tag == kTagBool) c = (int)(v != 0) - (int)(op.lit != 0);
    } else if (op.lit_tag == kTagNull) {
        if (tag == kTagNull) c = 0;
    }
    return filter_cmp_result(c, op.cmp);
}

// ops: postfix program in global memory (n_ops <= kFilterMaxOps, stack depth <= 64 checked by
// the host).  mask: [n_words] u32, n_words = ceil(n_rows / 256) * 8.
__global__ void __launch_bounds__(256)
filter_mask_kernel(const FilterOpDev *__restrict__ ops, uint32_t n_ops, uint64_t n_rows,
                   uint32_t *__restrict__ mask, uint64_t n_words) {
    __shared__ FilterOpDev s_ops[kFilterMaxOps];
    for (uint32_t i = threadIdx.x; i < n_ops; i += blockDim.x) s_ops[i] = ops[i];
    __syncthreads();
    const uint64_t n_padded = n_words * 32ull;
    for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_padded;
         row += (uint64_t)gridDim.x * blockDim.x) {
        bool pass = false;
        if (row < n_rows) {
            uint64_t stack = 0;  // bit i = i-th entry from the top
            for (uint32_t i = 0; i < n_ops; ++i) {
                const FilterOpDev &op = s_ops[i];
                if (op.kind == kFilterAnd) {
                    const uint64_t r = (stack & 1ull) & ((stack >> 1) & 1ull);
                    stack = ((stack >> 2) << 1) | r;
                } else if (op.kind == kFilterOr) {
                    const uint64_t r = (stack & 1ull) | ((stack >> 1) & 1ull);
                    stack = ((stack >> 2) << 1) | r;
                } else {
                    bool v;
                    if (op.kind == kFilterTrue) v = true;
                    else if (op.kind == kFilterFalse) v = false;
                    else v = filter_leaf(op, row);
                    stack = (stack << 1) | (v ? 1ull : 0ull);
                }
            }
            pass = (stack & 1ull) != 0;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, pass);
        if ((threadIdx.x & 31u) == 0) mask[row >> 5] = word;
    }
}

// Move one row's column entry (swap-remove keeps the columns in step with the mirror).
__global__ void column_move_kernel(uint8_t *tags, uint64_t *vals, uint64_t dst, uint64_t src) {
    tags[dst] = tags[src];
    vals[dst] = vals[src];
}
#endif  // __CUDACC__

}  // namespace nm
