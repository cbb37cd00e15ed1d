// nm_types.hpp — plain types and constants shared by the kernels (scan/batch/prefilter
// headers, included by exactly one translation unit: nm_launch.cu) and the host-side
// translation units of the library.
#pragma once
#include <stdint.h>

namespace nm {

constexpr int kRowsPerBlock = 256;                  // rows per row block == consumer threads
constexpr int kConsumerWarps = kRowsPerBlock / 32;  // 8
constexpr int kScanThreads = kRowsPerBlock + 32;    // + 1 producer warp
constexpr int kChunkFloats = 32;                    // floats per box row (128 B)
constexpr int kStageBytes = kRowsPerBlock * 128;    // 32 KiB per stage
constexpr int kMaxStages = 6;
constexpr int kCandCap = 2048;                      // candidate buffer entries (u64)
constexpr int kMaxFastK = 1024;                     // kCandCap - kMaxFastK >= kRowsPerBlock
constexpr uint32_t kMaxLocalRows = 0x7ffffffeu;     // local row ids are 31 bit


enum Metric : int { kCosine = 0, kEuclidean = 1, kDot = 2 };

// Candidate record exchanged between shards (one ncclAllGather of these).
struct alignas(16) ShardHit {
    uint64_t global_row;
    uint32_t ord;         // score_to_ord(score_bits); 0 with valid==0 marks an empty slot
    uint32_t score_bits;  // exact score bits (keeps -0.0)
};

// Peer-memory exchange of per-shard hits (one process per GPU, buffers mapped with CUDA IPC).
// mailbox[r] / flags[r] are rank r's buffers as seen from THIS process (own rank = local).
//   mailbox layout: [2 slots][n_ranks writers][kcap] ShardHit     flags: [n_ranks] u32 sequence
constexpr int kMaxRanks = 8;
struct PeerXchg {
    ShardHit *mailbox[kMaxRanks];
    uint32_t *flags[kMaxRanks];
    uint32_t n_ranks;  // <= 1: exchange disabled
    uint32_t rank;
    uint32_t seq;      // launch sequence number, identical on every rank, starts at 1
    uint32_t kcap;     // ShardHit slots per (slot, writer)
};

struct alignas(16) RowMeta {
    float scale;      // s_r (0 for an all-zero row)
    uint32_t x1;      // sum |xt_i|
    float rmag;       // reference-arithmetic |x| (lane tree + sqrt): exact denominator input
    uint32_t flags;   // bit 0: row has a non-finite element
};

// ---- device-side metadata filter (filter_kernels.cuh) ----
enum FilterKind : uint8_t {
    kFilterTrue = 0, kFilterFalse = 1, kFilterAnd = 2, kFilterOr = 3, kFilterExists = 4,
    kFilterCmp = 5, kFilterStrTable = 6
};
enum FilterCmp : uint8_t { kFilterEq = 0, kFilterNe, kFilterLt, kFilterLe, kFilterGt, kFilterGe };
enum ValueTag : uint8_t { kTagMissing = 0, kTagNull, kTagBool, kTagInt, kTagFloat, kTagString };
constexpr uint32_t kFilterMaxOps = 128;
struct FilterOpDev {         // one postfix op with this shard's column pointers resolved
    const uint8_t *tags;     // null = the shard has no such column: every row is "missing"
    const uint64_t *vals;
    const uint32_t *table;   // kFilterStrTable: bit c = the predicate holds for dictionary code c
    uint64_t lit;            // kFilterCmp: i64 / f64 bits / bool
    uint32_t table_bits;
    uint8_t kind, cmp, lit_tag, pad;
};

constexpr uint32_t kTcTileRows = 128;  // corpus rows per tensor-core tile (tc_prefilter_kernels.cuh)

constexpr uint32_t kKeptCap = 1u << 20;  // kept (row, ub) entries per query before fallback

struct KeptEntry {
    uint32_t row;
    uint32_t ub_ord;
};


}  // namespace nm
